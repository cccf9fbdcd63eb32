"""ctypes binding of oracle/_ref/libmoped3d_ref.so — moped3d's depth-aware pose stage compiled unmodified from
/root/reference (oracle/ref3d_harness.cpp, `make -f oracle/Makefile ref`). TEST INFRASTRUCTURE: pins the restatement
(oracle/moped_oracle.c: mo_*_depth) for SURVEY.md §8f row 4; never imported by the product."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libmoped3d_ref.so")                  # the reference's own flags (-ffast-math)
STRICT_PATH = os.path.join(_HERE, "_ref", "libmoped3d_ref_strict.so")        # the same sources, strict IEEE arithmetic
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_libs = {}
_strict = False


def available() -> bool:
    return os.path.exists(LIB_PATH) and os.path.exists(STRICT_PATH)


def use_strict(on: bool):
    """Select which build of the reference the calls below go to."""
    global _strict
    _strict = bool(on)


def lib():
    path = STRICT_PATH if _strict else LIB_PATH
    if path not in _libs:
        L = C.CDLL(path)
        L.ref3d_cauchy_weight.restype = C.c_float
        L.ref3d_cauchy_weight.argtypes = [C.c_float]
        L.ref3d_lm_func.restype = None
        L.ref3d_lm_func.argtypes = [_f32p, C.c_int, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_float, _f32p]
        L.ref3d_hypothesis.restype = C.c_int
        L.ref3d_hypothesis.argtypes = [C.c_int, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_float, _i32p, C.c_int, _f32p, C.c_int, C.c_float,
                                       C.c_int, _f32p, _f32p, _f32p, _f32p, _u8p]
        L.ref3d_ransac.restype = C.c_int
        L.ref3d_ransac.argtypes = [C.c_int, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                   C.c_uint64, _f32p]
        for sfx in ("_v1",):
            getattr(L, "ref3d_cauchy_weight" + sfx).restype = C.c_float
            getattr(L, "ref3d_cauchy_weight" + sfx).argtypes = [C.c_float]
            getattr(L, "ref3d_lm_func" + sfx).restype = None
            getattr(L, "ref3d_lm_func" + sfx).argtypes = L.ref3d_lm_func.argtypes
            getattr(L, "ref3d_hypothesis" + sfx).restype = C.c_int
            getattr(L, "ref3d_hypothesis" + sfx).argtypes = L.ref3d_hypothesis.argtypes
            getattr(L, "ref3d_ransac" + sfx).restype = C.c_int
            getattr(L, "ref3d_ransac" + sfx).argtypes = L.ref3d_ransac.argtypes
        L.ref3d_cluster_linkage.restype = C.c_int
        L.ref3d_cluster_linkage.argtypes = [C.c_int, _f32p, _f32p, _f32p, C.c_int, C.c_int, _f32p, _f32p, C.c_float, C.c_int, C.c_int, C.c_float,
                                            C.c_float, C.c_int, C.c_float, C.c_float, _i32p, _i32p]
        L.ref3d_linkage_agglomerate.restype = C.c_int
        L.ref3d_linkage_agglomerate.argtypes = [_f32p, C.c_int, C.c_float, C.c_int, C.c_int, _i32p, _i32p]
        L.ref3d_filter_depth.restype = C.c_int
        L.ref3d_filter_depth.argtypes = [C.c_int, _i32p, _f32p, _i32p, _f32p, _f32p, C.c_int, _i32p, _f32p, C.c_int, C.c_float, C.c_float, C.c_float,
                                         C.c_float, C.c_int, C.c_float, C.c_uint64, _f32p, _f32p, _f32p, _f32p, C.c_int, C.c_int, _f32p, _f32p,
                                         _u8p, _f32p, _i32p, _i32p]
        L.ref3d_bench_log.restype = C.c_int
        L.ref3d_bench_log.argtypes = [C.c_char_p, C.c_int, _i32p, _f32p, _i32p, _f32p, C.c_int, _i32p, _i32p, _i32p, C.c_int, _i32p, _f32p, _f32p,
                                      _f32p, _f32p]
        _libs[path] = L
    return _libs[path]


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def lm_func(pose7, cl, K, cam_pose, alpha, variant=0):
    n = len(cl["xy"])
    out = np.zeros((3 if variant else 2) * n, np.float32)
    (lib().ref3d_lm_func_v1 if variant else lib().ref3d_lm_func)(_f(pose7), n, _f(cl["xy"]), _f(cl["xyz"]), _f(cl["world"]), _f(cl["fill"]), _f(K), _f(cam_pose), alpha, out)
    return out


def hypothesis(cl, K, cam_pose, alpha, sample_pos, init_quat, max_lm, err_thr, min_npts, variant=0):
    n = len(cl["xy"])
    init, lm, refit = (np.zeros(7, np.float32) for _ in range(3))
    err = np.zeros(2, np.float32)
    mask = np.zeros(n, np.uint8)
    sp = np.ascontiguousarray(sample_pos, dtype=np.int32)
    r = (lib().ref3d_hypothesis_v1 if variant else lib().ref3d_hypothesis)(n, _f(cl["xy"]), _f(cl["xyz"]), _f(cl["world"]), _f(cl["fill"]), _f(K), _f(cam_pose), alpha, sp, len(sp),
                               _f(init_quat), max_lm, err_thr, min_npts, init, lm, refit, err, mask)
    return dict(n_inliers=r, pose_init=init, pose_lm=lm, pose_refit=refit, lm_err=err, mask=mask)


def ransac(cl, K, cam_pose, alpha, params, seed, variant=0):
    """params = (MaxRANSACTests, MaxLMTests, NPtsAlign, MinNPtsObject, ErrorThreshold)"""
    n = len(cl["xy"])
    pose = np.zeros(7, np.float32)
    found = (lib().ref3d_ransac_v1 if variant else lib().ref3d_ransac)(n, _f(cl["xy"]), _f(cl["xyz"]), _f(cl["world"]), _f(cl["fill"]), _f(K), _f(cam_pose), alpha,
                               int(params[0]), int(params[1]), int(params[2]), int(params[3]), float(params[4]), int(seed), pose)
    return bool(found), pose


def cluster_linkage(xy, xyz, world, depth, distance, cutoff=0.1, min_pts=7, use3d_filter=2, weight_gamma=1.0, alpha=0.0, linkage_type=1,
                    sigma2d=-1.0, sigma3d=-1.0):
    """CLUSTER_LINKAGE_CPU::process on one model's matches (constructor defaults = moped3d/libmoped/src/config.hpp:45)."""
    xy, xyz, world, depth, distance = _f(xy), _f(xyz), _f(world), _f(depth), _f(distance)
    n = len(xy)
    off = np.zeros(n + 2, np.int32)
    mem = np.zeros(n + 1, np.int32)
    c = lib().ref3d_cluster_linkage(n, xy, xyz, world, depth.shape[1], depth.shape[0], depth, distance, cutoff, min_pts, use3d_filter, weight_gamma,
                                    alpha, linkage_type, sigma2d, sigma3d, off, mem)
    return off[:c + 1].copy(), mem[:off[c]].copy()


def linkage_agglomerate(K, cutoff=0.1, min_pts=7, linkage_type=1):
    K = _f(K)
    n = len(K)
    off = np.zeros(n + 2, np.int32)
    mem = np.zeros(n + 1, np.int32)
    c = lib().ref3d_linkage_agglomerate(K, n, cutoff, min_pts, linkage_type, off, mem)
    return off[:c + 1].copy(), mem[:off[c]].copy()


def bench_log(out_dir, fr):
    """MopedBench over one frame state (the dict layout of moped_b200.bench_log.MopedBenchLog) -> text of outputMopedBench.txt."""
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    n_models = len(fr["n_matches"])
    rec = np.zeros((len(fr["match_xy"]), 10), np.float32)
    if len(rec):
        rec[:, 0:2] = fr["match_xy"]; rec[:, 2:5] = fr["match_world"]; rec[:, 5] = fr["match_depth"]; rec[:, 6] = fr["match_fill"]
        rec[:, 7] = fr["match_valid"]; rec[:, 8] = fr["match_image"]
    xyz = np.concatenate(fr["model_xyz"]).astype(np.float32) if n_models else np.zeros((0, 3), np.float32)
    r = lib().ref3d_bench_log(out_dir.encode(), n_models, i32([len(x) for x in fr["model_xyz"]]), _f(xyz), i32(fr["n_matches"]), _f(rec),
                              len(fr["cluster_model"]), i32(fr["cluster_model"]), i32(fr["cluster_offsets"]), i32(fr["cluster_members"]),
                              len(fr["obj_model"]), i32(fr["obj_model"]), _f(fr["obj_pose"]), _f(fr["obj_score"]), _f(fr["K"]), _f(fr["cam_pose"]))
    if r != 0:
        raise RuntimeError("ref3d_bench_log failed")
    with open(os.path.join(out_dir, "outputMopedBench.txt")) as f:
        return f.read()


def filter_depth(n_model_pts, model_xyz, match_offsets, match_xy, match_xyz, obj_model, obj_pose, params, test_sample_size, seed, K4, cam_pose7,
                 depth_K4, depth_pose7, depth, fill_distance):
    """FILTER_PROJECTION_DEPTH_CPU::process compiled from the reference. params = (MinPoints, FeatureDistance, PlausibleSqDistance, MinScore,
    DepthFraction, MinKeypointFraction)."""
    f32 = lambda a: np.ascontiguousarray(a, np.float32)      # noqa: E731
    i32 = lambda a: np.ascontiguousarray(a, np.int32)        # noqa: E731
    mo = i32(match_offsets)
    om, op = i32(obj_model), f32(obj_pose).reshape(-1, 7)
    n = len(om)
    keep = np.zeros(n + 1, np.uint8)
    score = np.zeros(n + 1, np.float32)
    co = np.zeros(n + 2, np.int32)
    mem = np.zeros(int(mo[-1]) + 1, np.int32)
    d, f = f32(depth), f32(fill_distance)
    ns = lib().ref3d_filter_depth(len(mo) - 1, i32(n_model_pts), f32(model_xyz), mo, f32(match_xy), f32(match_xyz), n, om, op, int(params[0]),
                                  params[1], params[2], params[3], params[4], int(test_sample_size), params[5], int(seed), f32(K4), f32(cam_pose7),
                                  f32(depth_K4), f32(depth_pose7), d.shape[1], d.shape[0], d, f, keep, score, co, mem)
    return dict(keep=keep[:n].astype(bool), score=score[:n].copy(), offsets=co[:ns + 1].copy(), members=mem[:co[ns]].copy())
