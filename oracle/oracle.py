"""ctypes binding of oracle/libmoped_oracle.so — the plain-C restatement (oracle/moped_oracle.c).

TEST INFRASTRUCTURE: imported only by tests/, bench.py's cpu_baseline leg and
__graft_entry__.smoke(). The product (moped_b200/) never imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmoped_oracle.so")

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_u64p = C.POINTER(C.c_uint64)


class Camera(C.Structure):
    _fields_ = [("K", C.c_float * 4), ("TM", C.c_float * 12)]


def build():
    subprocess.check_call(["make", "-s", "-f", os.path.join(_HERE, "Makefile"), "oracle"])


def _load():
    if not os.path.exists(LIB_PATH):
        build()
    lib = C.CDLL(LIB_PATH)
    camp = C.POINTER(Camera)
    sig = {
        "mo_norm_rows": (None, [_f32p, C.c_int, C.c_int]),
        "mo_match_2nn": (None, [_f32p, C.c_int, C.c_int, _f32p, C.c_int, _i32p, _f32p]),
        "mo_match_emit": (C.c_int, [_i32p, _f32p, C.c_int, C.c_float, _i32p, C.c_int, _i32p, _i32p, _i32p]),
        "mo_meanshift": (C.c_int, [_f32p, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, _i32p, _i32p]),
        "mo_cluster": (C.c_int, [_i32p, _i32p, _f32p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, _i32p, _i32p, _i32p]),
        "mo_camera_init": (None, [camp, _f32p, _f32p]),
        "mo_rand": (C.c_int, [_u64p]),
        "mo_rand_sample": (C.c_int, [_u64p, _f32p, _i32p, C.c_void_p, C.c_int, C.c_int, _i32p]),
        "mo_init_pose": (None, [_u64p, _f32p]),
        "mo_lm_func": (None, [_f32p, _f32p, C.c_int, _f32p, _f32p, _i32p, camp]),
        "mo_levmar_dif": (C.c_int, [_f32p, C.c_int, C.c_int, _f32p, _f32p, _i32p, camp, _f32p]),
        "mo_optimize_camera": (C.c_float, [_f32p, C.c_int, C.c_int, _f32p, _f32p, _i32p, camp]),
        "mo_project": (None, [_f32p, _f32p, camp, _f32p]),
        "mo_test_all_points": (C.c_int, [_f32p, C.c_int, _f32p, _f32p, _i32p, camp, C.c_float, _u8p]),
        "mo_hypothesis": (C.c_int, [C.c_int, _f32p, _f32p, _i32p, camp, _i32p, C.c_int, _f32p, C.c_int, C.c_float, C.c_int,
                                    _f32p, _f32p, _f32p, _u8p]),
        "mo_ransac": (C.c_int, [_u64p, C.c_int, _f32p, _f32p, _i32p, C.c_void_p, camp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                _f32p, C.POINTER(C.c_int)]),
        "mo_cauchy_weight": (C.c_float, [C.c_float]),
        "mo_lm_func_depth": (None, [_f32p, _f32p, C.c_int, _f32p, _f32p, _f32p, _i32p, camp, C.c_float]),
        "mo_init_translation_depth": (None, [_f32p, _i32p, C.c_int, _f32p]),
        "mo_hypothesis_depth": (C.c_int, [C.c_int, _f32p, _f32p, _f32p, _f32p, _i32p, camp, C.c_float, _i32p, C.c_int, _f32p, C.c_int, C.c_float,
                                          C.c_int, _f32p, _f32p, _f32p, _u8p]),
        "mo_ransac_depth": (C.c_int, [_u64p, C.c_int, _f32p, _f32p, _f32p, _f32p, _i32p, C.c_void_p, camp, C.c_float, C.c_int, C.c_int, C.c_int,
                                      C.c_int, C.c_float, _f32p, C.POINTER(C.c_int)]),
        "mo_linkage_similarity": (None, [C.c_int, _f32p, _f32p, _f32p, C.c_int, C.c_int, _f32p, _f32p, C.c_int, C.c_float, C.c_float, _f32p]),
        "mo_linkage_agglomerate": (C.c_int, [_f32p, C.c_int, C.c_float, C.c_int, C.c_int, _i32p, _i32p]),
        "mo_cluster_linkage": (C.c_int, [C.c_int, _f32p, _f32p, _f32p, C.c_int, C.c_int, _f32p, _f32p, C.c_float, C.c_int, C.c_int, C.c_int,
                                         C.c_float, C.c_float, _i32p, _i32p]),
        "mo_set_lm_finite_check": (None, [C.c_int]),
        "mo_cauchy_weight_v1": (C.c_float, [C.c_float]),
        "mo_lm_func_depth_v1": (None, [_f32p, _f32p, C.c_int, _f32p, _f32p, _f32p, _f32p, _i32p, camp, C.c_float]),
        "mo_hypothesis_depth_v1": (C.c_int, [C.c_int, _f32p, _f32p, _f32p, _f32p, _i32p, camp, C.c_float, _i32p, C.c_int, _f32p, C.c_int, C.c_float,
                                             C.c_int, _f32p, _f32p, _f32p, _u8p]),
        "mo_ransac_depth_v1": (C.c_int, [_u64p, C.c_int, _f32p, _f32p, _f32p, _f32p, _i32p, C.c_void_p, camp, C.c_float, C.c_int, C.c_int, C.c_int,
                                         C.c_int, C.c_float, _f32p, C.POINTER(C.c_int)]),
        "mo_sift": (C.c_int, [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _f32p, _f32p]),
        "mo_sift_debug": (C.c_int, [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int)]),
        "mo_sift_gauss_kernel": (C.c_int, [C.c_float, _f32p]),
        "mo_sift_set_conv_fma": (None, [C.c_int]),
        "mo_filter": (C.c_int, [C.c_int, _i32p, _i32p, _f32p, _f32p, camp, C.c_int, _i32p, _f32p, C.c_int, C.c_float, C.c_float,
                                _u8p, _f32p, _i32p, _i32p]),
        "mo_filter_depth_select": (C.c_int, [_u64p, C.c_int, C.c_int, _i32p]),
        "mo_filter_depth": (C.c_int, [C.c_int, _i32p, _i32p, _f32p, _f32p, camp, C.c_int, _i32p, _f32p, C.c_int, C.c_float, C.c_float, C.c_float,
                                      C.c_float, C.c_float, _i32p, _f32p, camp, C.c_int, C.c_int, _f32p, _f32p, _u8p, _f32p, _i32p, _i32p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def cameras(K, cam_pose):
    K, cam_pose = _f32(K).reshape(-1, 4), _f32(cam_pose).reshape(-1, 7)
    arr = (Camera * len(K))()
    for i in range(len(K)):
        lib().mo_camera_init(C.byref(arr[i]), K[i], cam_pose[i])
    return arr


def norm_rows(desc):
    d = _f32(desc).copy()
    lib().mo_norm_rows(d, d.shape[0], d.shape[1])
    return d


def match_2nn(db, q):
    db, q = _f32(db), _f32(q)
    idx = np.empty((len(q), 2), np.int32)
    dist = np.empty((len(q), 2), np.float32)
    lib().mo_match_2nn(db, db.shape[0], db.shape[1], q, len(q), idx, dist)
    return idx, dist


def match_emit(idx, dist, ratio, model_of_row, n_models):
    idx, dist = _i32(idx), _f32(dist)
    Q = len(idx)
    mq = np.empty(Q, np.int32)
    mr = np.empty(Q, np.int32)
    off = np.zeros(n_models + 1, np.int32)
    n = lib().mo_match_emit(idx, dist, Q, ratio, _i32(model_of_row), n_models, mq, mr, off)
    return mq[:n], mr[:n], off


def match(db_desc, db_xyz, model_of_row, n_models, q_desc, q_xy, q_image, ratio):
    """MATCH step on already-normalised descriptors -> matches dict like ref.Ref.get_matches()."""
    idx, dist = match_2nn(db_desc, q_desc)
    mq, mr, off = match_emit(idx, dist, ratio, model_of_row, n_models)
    return dict(offsets=off, image=_i32(q_image)[mq], xy=_f32(q_xy)[mq], xyz=_f32(db_xyz)[mr], query=mq, row=mr), idx, dist


def meanshift(xy, radius=200.0, merge=20.0, minpts=7, maxiter=100):
    xy = _f32(xy).reshape(-1, 2)
    n = len(xy)
    off = np.zeros(n + 2, np.int32)
    mem = np.zeros(n + 1, np.int32)
    c = lib().mo_meanshift(xy, n, radius, merge, minpts, maxiter, off, mem)
    return off[:c + 1].copy(), mem[:off[c]].copy()


def cluster(matches, n_images, radius=200.0, merge=20.0, minpts=7, maxiter=100):
    off = _i32(matches["offsets"])
    M = int(off[-1])
    cm = np.zeros(M + 2, np.int32)
    co = np.zeros(M + 2, np.int32)
    mem = np.zeros(M + 1, np.int32)
    c = lib().mo_cluster(off, _i32(matches["image"]), _f32(matches["xy"]), len(off) - 1, n_images, radius, merge, minpts, maxiter, cm, co, mem)
    return dict(model=cm[:c].copy(), offsets=co[:c + 1].copy(), members=mem[:co[c]].copy())


def draw_samples(xy, image, tie_ids, n_pts_align, seed, n_hyp):
    xy, image = _f32(xy), _i32(image)
    tie = _i32(tie_ids) if tie_ids is not None else None
    st = C.c_uint64(seed)
    pos = np.full((n_hyp, n_pts_align), -1, np.int32)
    quat = np.zeros((n_hyp, 4), np.float32)
    ok = 0
    L = lib()
    for k in range(n_hyp):
        p = np.zeros(n_pts_align, np.int32)
        if not L.mo_rand_sample(C.byref(st), xy, image, tie.ctypes.data if tie is not None else None, len(xy), n_pts_align, p):
            continue
        pose = np.zeros(7, np.float32)
        L.mo_init_pose(C.byref(st), pose)
        pos[k] = p
        quat[k] = pose[:4]
        ok += 1
    return ok, pos, quat


def hypothesis(xy, xyz, image, cams, sample_pos, init_quat, max_lm, err_thr, min_npts):
    xy, xyz, image = _f32(xy), _f32(xyz), _i32(image)
    pose_lm = np.zeros(7, np.float32)
    pose_refit = np.zeros(7, np.float32)
    err = np.zeros(2, np.float32)
    mask = np.zeros(len(xy), np.uint8)
    sp = _i32(sample_pos)
    r = lib().mo_hypothesis(len(xy), xy, xyz, image, cams, sp, len(sp), _f32(init_quat), max_lm, err_thr, min_npts,
                            pose_lm, pose_refit, err, mask)
    return r, pose_lm, pose_refit, err, mask


def ransac(xy, xyz, image, tie_ids, cams, params, seed):
    xy, xyz, image = _f32(xy), _f32(xyz), _i32(image)
    tie = _i32(tie_ids) if tie_ids is not None else None
    st = C.c_uint64(seed)
    pose = np.zeros(7, np.float32)
    it = C.c_int(0)
    f = lib().mo_ransac(C.byref(st), len(xy), xy, xyz, image, tie.ctypes.data if tie is not None else None, cams,
                        params[0], params[1], params[3], params[4], params[5], pose, C.byref(it))
    return f, pose, it.value


def project(pose7, xyz, image, cams):
    xyz, image = _f32(xyz), _i32(image)
    uv = np.empty((len(xyz), 2), np.float32)
    p = _f32(pose7)
    for i in range(len(xyz)):
        lib().mo_project(p, xyz[i], C.byref(cams[int(image[i])]), uv[i])
    return uv


def filter_objects(matches, cams, obj_model, obj_pose, params=(5, 4096.0, 2.0)):
    off = _i32(matches["offsets"])
    M = int(off[-1])
    obj_model, obj_pose = _i32(obj_model), _f32(obj_pose).reshape(-1, 7)
    n = len(obj_model)
    keep = np.zeros(n + 1, np.uint8)
    score = np.zeros(n + 1, np.float32)
    co = np.zeros(n + 2, np.int32)
    mem = np.zeros(M + 1, np.int32)
    ns = lib().mo_filter(len(off) - 1, off, _i32(matches["image"]), _f32(matches["xy"]), _f32(matches["xyz"]), cams, n,
                         obj_model, obj_pose, params[0], params[1], params[2], keep, score, co, mem)
    return dict(keep=keep[:n].astype(bool), score=score[:n].copy(), offsets=co[:ns + 1].copy(), members=mem[:co[ns]].copy())


def filter_depth_test_points(model_offsets, model_xyz, sample_size, seed):
    """selectTestPoints (FILTER_PROJECTION_DEPTH_CPU.hpp:94-118): per model all keypoints, or `sample_size` of them drawn by randSample from
    the seedable stream (the models are visited in order, the stream runs on). Returns (test_offsets[n_models+1], test_xyz)."""
    mo = _i32(model_offsets)
    xyz = _f32(model_xyz).reshape(-1, 3)
    state = C.c_uint64(int(seed))
    offs, out = [0], []
    for m in range(len(mo) - 1):
        n = int(mo[m + 1] - mo[m])
        idx = np.zeros(max(n, 1), np.int32)
        k = lib().mo_filter_depth_select(C.byref(state), n, int(sample_size), idx)
        out.append(xyz[mo[m] + idx[:k]])
        offs.append(offs[-1] + k)
    return np.array(offs, np.int32), (np.concatenate(out) if out else np.zeros((0, 3), np.float32)).astype(np.float32)


def filter_depth(matches, cams, obj_model, obj_pose, params, test_offsets, test_xyz, depth_cam, depth, fill_distance):
    """params = (MinPoints, FeatureDistance, PlausibleSqDistance, MinScore, DepthFraction, MinKeypointFraction); depth / fill_distance: H x W."""
    off = _i32(matches["offsets"])
    M = int(off[-1])
    obj_model, obj_pose = _i32(obj_model), _f32(obj_pose).reshape(-1, 7)
    n = len(obj_model)
    keep = np.zeros(n + 1, np.uint8)
    score = np.zeros(n + 1, np.float32)
    co = np.zeros(n + 2, np.int32)
    mem = np.zeros(M + 1, np.int32)
    d, f = _f32(depth), _f32(fill_distance)
    ns = lib().mo_filter_depth(len(off) - 1, off, _i32(matches["image"]), _f32(matches["xy"]), _f32(matches["xyz"]), cams, n, obj_model, obj_pose,
                               int(params[0]), params[1], params[2], params[3], params[4], params[5], _i32(test_offsets), _f32(test_xyz), depth_cam,
                               d.shape[1], d.shape[0], d, f, keep, score, co, mem)
    return dict(keep=keep[:n].astype(bool), score=score[:n].copy(), offsets=co[:ns + 1].copy(), members=mem[:co[ns]].copy())


SIFT_TRACE = np.dtype([("octave", np.int32), ("index", np.int32), ("scan_row", np.int32), ("scan_col", np.int32), ("row", np.int32),
                       ("col", np.int32), ("X", np.float32, 3), ("fsize", np.float32), ("first_kp", np.int32)])


def sift(gray_u8, double_size=True, max_kp=65536):
    """FEAT_SIFT_CPU / libsiftfast restated (moped_sift_oracle.c): xy[n,2]=(col,row), scale_ori[n,2], desc[n,128], in the
    order FEAT_SIFT_CPU emits them with one OpenMP thread."""
    g = np.ascontiguousarray(gray_u8, dtype=np.uint8)
    xy = np.zeros((max_kp, 2), np.float32)
    so = np.zeros((max_kp, 2), np.float32)
    desc = np.zeros((max_kp, 128), np.float32)
    n = lib().mo_sift(g, g.shape[0], g.shape[1], 1 if double_size else 0, max_kp, xy, so, desc)
    n = min(n, max_kp)
    return xy[:n].copy(), so[:n].copy(), desc[:n].copy()


def sift_octave_dims(height, width, double_size=True):
    r, c = (2 * height - 2, 2 * width - 2) if double_size else (height, width)
    dims = []
    while r > 12 and c > 12:
        dims.append((r, c))
        r, c = r >> 1, c >> 1
    return dims


def sift_debug(gray_u8, double_size=True, octave=0, max_trace=65536):
    """Gaussian (6) and DoG (5) images of one octave plus the accepted-extremum trace (creation order)."""
    g = np.ascontiguousarray(gray_u8, dtype=np.uint8)
    r, c = sift_octave_dims(g.shape[0], g.shape[1], double_size)[octave]
    gauss = np.zeros((6, r, c), np.float32)
    dog = np.zeros((5, r, c), np.float32)
    trace = np.zeros(max_trace, SIFT_TRACE)
    nt = C.c_int(0)
    lib().mo_sift_debug(g, g.shape[0], g.shape[1], 1 if double_size else 0, octave, gauss.ctypes.data, dog.ctypes.data, max_trace,
                        trace.ctypes.data, C.byref(nt))
    return gauss, dog, trace[:min(nt.value, max_trace)].copy()


# ---- moped3d depth-aware pose stage (oracle only so far) ----------------------------------------
def cauchy_weights(fill, variant=0):
    fn = lib().mo_cauchy_weight_v1 if variant else lib().mo_cauchy_weight
    return np.array([fn(float(f)) for f in np.asarray(fill).ravel()], np.float32)


def lm_func_depth(pose7, cl, cams, alpha, variant=0):
    """variant 0 = ..._BACKPROJECTION_DEPTH_CPU (2 residuals per correspondence), 1 = ..._REPROJECTION_DEPTH_CPU (3)."""
    n = len(cl["xy"])
    img = np.zeros(n, np.int32)
    if variant:
        out = np.zeros(3 * n, np.float32)
        lib().mo_lm_func_depth_v1(_f32(pose7), out, n, _f32(cl["xy"]), _f32(cl["xyz"]), _f32(cl["world"]), cauchy_weights(cl["fill"], 1), img, cams, alpha)
    else:
        out = np.zeros(2 * n, np.float32)
        lib().mo_lm_func_depth(_f32(pose7), out, n, _f32(cl["xyz"]), _f32(cl["world"]), cauchy_weights(cl["fill"]), img, cams, alpha)
    return out


def hypothesis_depth(cl, cams, alpha, sample_pos, init_quat, max_lm, err_thr, min_npts, variant=0):
    n = len(cl["xy"])
    lm, refit = np.zeros(7, np.float32), np.zeros(7, np.float32)
    err = np.zeros(2, np.float32)
    mask = np.zeros(n, np.uint8)
    sp = _i32(sample_pos)
    fn = lib().mo_hypothesis_depth_v1 if variant else lib().mo_hypothesis_depth
    r = fn(n, _f32(cl["xy"]), _f32(cl["xyz"]), _f32(cl["world"]), cauchy_weights(cl["fill"], variant), np.zeros(n, np.int32), cams,
           alpha, sp, len(sp), _f32(init_quat), max_lm, err_thr, min_npts, lm, refit, err, mask)
    return dict(n_inliers=r, pose_lm=lm, pose_refit=refit, lm_err=err, mask=mask)


def ransac_depth(cl, cams, alpha, params, seed, variant=0):
    """params = (MaxRANSACTests, MaxLMTests, NPtsAlign, MinNPtsObject, ErrorThreshold); seed = the LCG state ref3d_srand gets"""
    n = len(cl["xy"])
    st = C.c_uint64(int(seed))
    pose = np.zeros(7, np.float32)
    it = C.c_int(0)
    fn = lib().mo_ransac_depth_v1 if variant else lib().mo_ransac_depth
    found = fn(C.byref(st), n, _f32(cl["xy"]), _f32(cl["xyz"]), _f32(cl["world"]), cauchy_weights(cl["fill"], variant),
               np.zeros(n, np.int32), None, cams, alpha, int(params[0]), int(params[1]), int(params[2]), int(params[3]),
               float(params[4]), pose, C.byref(it))
    return bool(found), pose, it.value


def cluster_linkage(xy, xyz, world, depth, distance, cutoff=0.1, min_pts=7, use3d_filter=2, linkage_type=1, sigma2d=-1.0, sigma3d=-1.0):
    """moped3d CLUSTER_LINKAGE_CPU on one model's matches -> (offsets, members). depth / distance: H x W float maps."""
    xy, xyz, world = _f32(xy), _f32(xyz), _f32(world)
    depth, distance = _f32(depth), _f32(distance)
    n = len(xy)
    off = np.zeros(n + 2, np.int32)
    mem = np.zeros(n + 1, np.int32)
    c = lib().mo_cluster_linkage(n, xy, xyz, world, depth.shape[1], depth.shape[0], depth, distance, cutoff, min_pts, use3d_filter, linkage_type,
                                 sigma2d, sigma3d, off, mem)
    return off[:c + 1].copy(), mem[:off[c]].copy()


def linkage_similarity(xy, xyz, world, depth, distance, use3d_filter=2, sigma2d=-1.0, sigma3d=-1.0):
    xy, xyz, world, depth, distance = _f32(xy), _f32(xyz), _f32(world), _f32(depth), _f32(distance)
    n = len(xy)
    K = np.zeros((n, n), np.float32)
    lib().mo_linkage_similarity(n, xy, xyz, world, depth.shape[1], depth.shape[0], depth, distance, use3d_filter, sigma2d, sigma3d, K)
    return K


def linkage_agglomerate(K, cutoff=0.1, min_pts=7, linkage_type=1):
    K = _f32(K)
    n = len(K)
    off = np.zeros(n + 2, np.int32)
    mem = np.zeros(n + 1, np.int32)
    c = lib().mo_linkage_agglomerate(K, n, cutoff, min_pts, linkage_type, off, mem)
    return off[:c + 1].copy(), mem[:off[c]].copy()
