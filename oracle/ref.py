"""ctypes binding of oracle/_ref/libmoped_ref.so — the reference's own stage classes compiled
unmodified (see oracle/ref_harness.cpp, oracle/Makefile).

TEST INFRASTRUCTURE: imported only by tests/, bench.py's cpu_baseline / --impl reference legs,
tests/golden/make_golden.py and __graft_entry__.smoke(). The product (moped_b200/) never imports it.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libmoped_ref.so")                    # the reference's own flags (-ffast-math)
STRICT_PATH = os.path.join(_HERE, "_ref", "libmoped_ref_strict.so")          # the same sources, strict IEEE arithmetic

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def available() -> bool:
    return os.path.exists(LIB_PATH)


def strict_available() -> bool:
    return os.path.exists(STRICT_PATH)


def _load(path=None):
    lib = C.CDLL(path or LIB_PATH)
    sig = {
        "ref_create": (C.c_void_p, [C.c_int]),
        "ref_destroy": (None, [C.c_void_p]),
        "ref_max_threads": (C.c_int, []),
        "ref_set_models": (None, [C.c_void_p, C.c_int, _i32p, _f32p, _f32p, C.c_int]),
        "ref_get_model_desc": (None, [C.c_void_p, _f32p, C.c_int]),
        "ref_set_images": (None, [C.c_void_p, C.c_int, _f32p, _f32p]),
        "ref_set_features": (None, [C.c_void_p, C.c_int, C.c_int, _f32p, _f32p, _i32p]),
        "ref_get_features_desc": (None, [C.c_void_p, _f32p, C.c_int]),
        "ref_clear_frame": (None, [C.c_void_p]),
        "ref_norm_rows": (None, [_f32p, C.c_int, C.c_int]),
        "ref_build_match": (C.c_double, [C.c_void_p, C.c_float, C.c_float]),
        "ref_run_match": (C.c_double, [C.c_void_p, C.c_float, C.c_float]),
        "ref_ann_search": (None, [C.c_void_p, _f32p, C.c_int, C.c_float, _i32p, _f32p]),
        "ref_match_total": (C.c_int, [C.c_void_p]),
        "ref_match_models": (C.c_int, [C.c_void_p]),
        "ref_get_matches": (None, [C.c_void_p, _i32p, _i32p, _f32p, _f32p]),
        "ref_set_matches": (None, [C.c_void_p, C.c_int, _i32p, _i32p, _f32p, _f32p]),
        "ref_run_cluster": (C.c_double, [C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_int]),
        "ref_cluster_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
        "ref_get_clusters": (None, [C.c_void_p, _i32p, _i32p, _i32p]),
        "ref_set_clusters": (None, [C.c_void_p, C.c_int, _i32p, _i32p, _i32p]),
        "ref_run_pose": (C.c_double, [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_uint64]),
        "ref_object_count": (C.c_int, [C.c_void_p]),
        "ref_get_objects": (None, [C.c_void_p, _i32p, _f32p, _f32p]),
        "ref_set_objects": (None, [C.c_void_p, C.c_int, _i32p, _f32p]),
        "ref_run_filter": (C.c_double, [C.c_void_p, C.c_int, C.c_float, C.c_float]),
        "ref_draw_samples": (C.c_int, [C.c_void_p, C.c_int, _i32p, C.c_int, C.c_int, C.c_uint64, C.c_int, _i32p, _f32p]),
        "ref_hypothesis": (C.c_int, [C.c_void_p, C.c_int, _i32p, C.c_int, _i32p, C.c_int, _f32p, C.c_int, C.c_float, C.c_int,
                                     _f32p, _f32p, _f32p, _u8p]),
        "ref_hypotheses_batch": (C.c_double, [C.c_void_p, C.c_int, _i32p, C.c_int, _i32p, C.c_int, _f32p, C.c_int, C.c_int, C.c_float, C.c_int,
                                              _i32p, _f32p]),
        "ref_sift": (C.c_int, [_u8p, C.c_int, C.c_int, C.c_int]),
        "ref_sift_get": (None, [_f32p, _f32p]),
        "ref_add_model_xml": (C.c_int, [C.c_void_p, C.c_char_p]),
        "ref_model_count": (C.c_int, [C.c_void_p]),
        "ref_model_name": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p, C.c_int]),
        "ref_model_bbox": (None, [C.c_void_p, C.c_int, _f32p]),
        "ref_model_points": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p, C.POINTER(C.c_long)]),
        "ref_get_model_points": (None, [C.c_void_p, C.c_int, C.c_char_p, _f32p, _i32p, _f32p]),
        "ref_ransac": (C.c_int, [C.c_void_p, C.c_int, _i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_uint64, _f32p]),
        "ref_project": (None, [C.c_void_p, _f32p, _f32p, _i32p, C.c_int, _f32p]),
        "ref_run_pipeline": (C.c_int, [C.c_void_p, _f32p, C.c_uint64, _f64p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


_lib = None
_lib_fast = None
_lib_strict = None


def lib():
    global _lib, _lib_fast
    if _lib is None:
        _lib = _lib_fast = _load()
    return _lib


def use_strict(on: bool):
    """Route every call of this module to the strict-IEEE build of the reference (tests only) or back to the build with the
    reference's own flags. Contexts (Ref objects) belong to the build that created them."""
    global _lib, _lib_fast, _lib_strict
    lib()
    if on:
        if _lib_strict is None:
            _lib_strict = _load(STRICT_PATH)
        _lib = _lib_strict
    else:
        _lib = _lib_fast


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


# Default stage parameters, moped2/libmoped/src/config.hpp:83,101,110,115,118,120
DEFAULT_PARAMS = dict(
    match=(5.0, 0.8), cluster=(200.0, 20.0, 7, 100),
    pose=(600, 200, 4, 5, 6, 10.0), filter=(5, 4096.0, 2.0),
    pose2=(100, 500, 4, 6, 8, 5.0), filter2=(7, 4096.0, 3.0))


def pipeline_param_vector(p=DEFAULT_PARAMS, quality=None):
    m = list(p["match"])
    if quality is not None:
        m[0] = quality
    return np.array(m + list(p["cluster"]) + list(p["pose"]) + list(p["filter"]) + list(p["pose2"]) + list(p["filter2"]),
                    dtype=np.float32)


def sift(gray_u8, double_size=True):
    """The reference's FEAT_SIFT_CPU (libsiftfast) on one grayscale image: (xy[n,2], desc[n,128])."""
    g = np.ascontiguousarray(gray_u8, dtype=np.uint8)
    n = lib().ref_sift(g, g.shape[0], g.shape[1], 1 if double_size else 0)
    xy = np.zeros((n, 2), np.float32)
    desc = np.zeros((n, 128), np.float32)
    if n:
        lib().ref_sift_get(xy, desc)
    return xy, desc


class Ref:
    """One reference pipeline context (models + cameras + one frame's FrameData)."""

    def __init__(self, n_threads: int = 1):
        self.L = lib()
        self.h = self.L.ref_create(int(n_threads))
        self.n_threads = min(int(n_threads), self.L.ref_max_threads())
        self.D = 128
        self.n_models = 0

    def close(self):
        if self.h:
            self.L.ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- inputs
    def set_models(self, n_pts, xyz, desc):
        n_pts, xyz, desc = _i32(n_pts), _f32(xyz), _f32(desc)
        self.D = desc.shape[1]
        self.n_models = len(n_pts)
        self.N = desc.shape[0]
        self.L.ref_set_models(self.h, len(n_pts), n_pts, xyz, desc, self.D)

    # ---- model files through the reference's sXML reader + addModel loop
    def add_model_xml(self, path):
        st = self.L.ref_add_model_xml(self.h, os.fsencode(path))
        # row count of the SIFT-128 database the matcher will build (model_desc() sizes its output from it)
        self.D = 128
        self.N = sum(self.L.ref_model_points(self.h, i, b"SIFT", None) for i in range(self.L.ref_model_count(self.h)))
        return st

    def model_names(self):
        out = []
        for i in range(self.L.ref_model_count(self.h)):
            buf = C.create_string_buffer(4096)
            self.L.ref_model_name(self.h, i, buf, 4096)
            out.append(buf.value.decode())
        return out

    def model_bbox(self, i):
        b = np.zeros(6, np.float32)
        self.L.ref_model_bbox(self.h, i, b)
        return b

    def model_points(self, i, desc_type="SIFT"):
        """(xyz[n,3], desc_len[n], desc_values[sum]) of model i as the reference parsed them."""
        nv = C.c_long()
        n = self.L.ref_model_points(self.h, i, desc_type.encode(), C.byref(nv))
        xyz = np.zeros((n, 3), np.float32)
        ln = np.zeros(n, np.int32)
        vals = np.zeros(nv.value, np.float32)
        if n:
            self.L.ref_get_model_points(self.h, i, desc_type.encode(), xyz, ln, vals)
        return xyz, ln, vals

    def model_desc(self):
        out = np.empty((self.N, self.D), np.float32)
        self.L.ref_get_model_desc(self.h, out, self.D)
        return out

    def set_images(self, K, cam_pose):
        K, cam_pose = _f32(K).reshape(-1, 4), _f32(cam_pose).reshape(-1, 7)
        self.L.ref_set_images(self.h, len(K), K, cam_pose)

    def set_features(self, desc, xy, image_idx):
        desc, xy, image_idx = _f32(desc), _f32(xy), _i32(image_idx)
        self.Q = len(desc)
        self.L.ref_set_features(self.h, len(desc), desc.shape[1], desc, xy, image_idx)

    def features_desc(self):
        out = np.empty((self.Q, self.D), np.float32)
        self.L.ref_get_features_desc(self.h, out, self.D)
        return out

    def clear_frame(self):
        self.L.ref_clear_frame(self.h)

    @staticmethod
    def norm_rows(desc):
        d = _f32(desc).copy()
        lib().ref_norm_rows(d, d.shape[0], d.shape[1])
        return d

    # ---- MATCH
    def build_match(self, quality=0.0, ratio=0.8):
        return self.L.ref_build_match(self.h, quality, ratio)

    def run_match(self, quality=5.0, ratio=0.8):
        return self.L.ref_run_match(self.h, quality, ratio)

    def ann_search(self, q, eps=0.0):
        q = _f32(q)
        idx = np.empty((len(q), 2), np.int32)
        dist = np.empty((len(q), 2), np.float32)
        self.L.ref_ann_search(self.h, q, len(q), eps, idx, dist)
        return idx, dist

    def get_matches(self):
        nm = self.L.ref_match_models(self.h)
        t = self.L.ref_match_total(self.h)
        off = np.zeros(max(nm, self.n_models) + 1, np.int32)
        img = np.empty(t, np.int32)
        xy = np.empty((t, 2), np.float32)
        xyz = np.empty((t, 3), np.float32)
        if nm:
            self.L.ref_get_matches(self.h, off, img, xy, xyz)
            off[nm:] = t
        return dict(offsets=off, image=img, xy=xy, xyz=xyz)

    def set_matches(self, m):
        off = _i32(m["offsets"])
        self.L.ref_set_matches(self.h, len(off) - 1, off, _i32(m["image"]), _f32(m["xy"]), _f32(m["xyz"]))

    # ---- CLUSTER
    def run_cluster(self, radius=200.0, merge=20.0, minpts=7, maxiter=100):
        return self.L.ref_run_cluster(self.h, radius, merge, minpts, maxiter)

    def get_clusters(self):
        tot = C.c_int(0)
        n = self.L.ref_cluster_count(self.h, C.byref(tot))
        model = np.empty(n, np.int32)
        off = np.zeros(n + 1, np.int32)
        mem = np.empty(tot.value, np.int32)
        self.L.ref_get_clusters(self.h, model, off, mem)
        return dict(model=model, offsets=off, members=mem)

    def set_clusters(self, c):
        self.L.ref_set_clusters(self.h, len(c["model"]), _i32(c["model"]), _i32(c["offsets"]), _i32(c["members"]))

    # ---- POSE
    def run_pose(self, step="POSE", params=(600, 200, 4, 5, 6, 10.0), seed=1):
        a = params
        return self.L.ref_run_pose(self.h, step.encode(), a[0], a[1], a[2], a[3], a[4], a[5], seed)

    def get_objects(self):
        n = self.L.ref_object_count(self.h)
        model = np.empty(n, np.int32)
        pose = np.empty((n, 7), np.float32)
        score = np.empty(n, np.float32)
        self.L.ref_get_objects(self.h, model, pose, score)
        return dict(model=model, pose=pose, score=score)

    def set_objects(self, model, pose):
        model, pose = _i32(model), _f32(pose).reshape(-1, 7)
        self.L.ref_set_objects(self.h, len(model), model, pose)

    def run_filter(self, params=(5, 4096.0, 2.0)):
        return self.L.ref_run_filter(self.h, params[0], params[1], params[2])

    # ---- per hypothesis
    def draw_samples(self, model, members, n_pts_align, seed, n_hyp):
        members = _i32(members)
        pos = np.empty((n_hyp, n_pts_align), np.int32)
        quat = np.zeros((n_hyp, 4), np.float32)
        ok = self.L.ref_draw_samples(self.h, model, members, len(members), n_pts_align, seed, n_hyp, pos, quat)
        return ok, pos, quat

    def hypothesis(self, model, members, sample_pos, init_quat, max_lm, err_thr, min_npts):
        members, sample_pos, init_quat = _i32(members), _i32(sample_pos), _f32(init_quat)
        pose_lm = np.zeros(7, np.float32)
        pose_refit = np.zeros(7, np.float32)
        err = np.zeros(2, np.float32)
        mask = np.zeros(len(members), np.uint8)
        r = self.L.ref_hypothesis(self.h, model, members, len(members), sample_pos, len(sample_pos), init_quat,
                                  max_lm, err_thr, min_npts, pose_lm, pose_refit, err, mask)
        return r, pose_lm, pose_refit, err, mask

    def hypotheses_batch(self, model, members, sample_pos, init_quat, max_lm, err_thr, min_npts):
        """Many explicit hypotheses of one cluster on the OpenMP team. Returns (seconds, n_inliers, pose)."""
        members, sample_pos, init_quat = _i32(members), _i32(sample_pos), _f32(init_quat)
        n_hyp, n_samples = sample_pos.shape
        n_inl = np.zeros(n_hyp, np.int32)
        pose = np.zeros((n_hyp, 7), np.float32)
        sec = self.L.ref_hypotheses_batch(self.h, model, members, len(members), sample_pos, n_samples, init_quat, n_hyp,
                                          max_lm, err_thr, min_npts, n_inl, pose)
        return sec, n_inl, pose

    def ransac(self, model, members, params, seed):
        members = _i32(members)
        pose = np.zeros(7, np.float32)
        f = self.L.ref_ransac(self.h, model, members, len(members), params[0], params[1], params[3], params[4], params[5], seed, pose)
        return f, pose

    def project(self, pose7, xyz, image):
        xyz, image = _f32(xyz), _i32(image)
        uv = np.empty((len(xyz), 2), np.float32)
        self.L.ref_project(self.h, _f32(pose7), xyz, image, len(xyz), uv)
        return uv

    def run_pipeline(self, params=DEFAULT_PARAMS, quality=None, seed=1):
        times = np.zeros(6, np.float64)
        n = self.L.ref_run_pipeline(self.h, pipeline_param_vector(params, quality), seed, times)
        return n, times
