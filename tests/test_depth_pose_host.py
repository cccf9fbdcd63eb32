"""The DEVICE SOURCE of the depth-aware pose stages (moped_b200/csrc/lm_exact.cuh + depth_pose.cuh: an LM for a team of
lanes that takes every sum in levmar's order) compiled by g++ with the lanes emulated by a loop (tests/cpp/depth_host.cpp)
and compared bit for bit with the oracle — and, where the compiled reference is present, with the strict-IEEE build of
moped3d's own stage classes. It checks the arithmetic and the split into phases of the code the kernels in pose_depth.cu
instantiate (team width 32), on a machine without a GPU; lanes of a phase are visited in ascending and descending order, so a
phase that depended on the order of its lanes would show. The CUDA path itself is checked by tests/test_gpu_depth_pose.py.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle, ref3d
from test_oracle3d_pose import ALPHA, CAM, K, make_cluster

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
PARAMS = (100, 8.0, 6)        # MaxLMTests, ErrorThreshold, MinNPtsObject (moped3d/libmoped/src/config.hpp:46)


@pytest.fixture(scope="module")
def host_lib():
    out_dir = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, "libdepth_host.so")
    src = os.path.join(ROOT, "tests", "cpp", "depth_host.cpp")
    deps = [src] + [os.path.join(ROOT, "moped_b200", "csrc", f) for f in ("lm_exact.cuh", "depth_pose.cuh", "simt_phases.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        # -ffp-contract=off = the .cu's -fmad=false; x86-64-v3 like the oracle (FMA available, so a contraction would show)
        subprocess.check_call(["g++", "-x", "c++", "-std=c++17", "-O2", "-march=x86-64-v3", "-ffp-contract=off", "-fPIC", "-shared", "-Wall",
                               "-o", out, src])
    L = C.CDLL(out)
    L.dh_hypothesis.restype = C.c_int
    L.dh_hypothesis.argtypes = [C.c_int] * 4 + [_f32p] * 4 + [_i32p, C.c_void_p, C.c_float, _i32p, C.c_int, _f32p, C.c_int, C.c_float, C.c_int,
                                C.c_int, _f32p, _f32p, _f32p, _u8p]
    return L


@pytest.fixture(scope="module")
def cams():
    return oracle.cameras(K[None], CAM[None])


def run_host(L, variant, width, order, cl, cams, pos, quat, finite_check=0):
    n = len(cl["xy"])
    lm, rf, er, mk = np.zeros(7, np.float32), np.zeros(7, np.float32), np.zeros(2, np.float32), np.zeros(n, np.uint8)
    pos = np.ascontiguousarray(pos, np.int32)
    r = L.dh_hypothesis(variant, width, order, n, cl["xy"], cl["xyz"], cl["world"], oracle.cauchy_weights(cl["fill"], variant), np.zeros(n, np.int32),
                        C.addressof(cams), ALPHA, pos, len(pos), np.ascontiguousarray(quat, np.float32), PARAMS[0], PARAMS[1], PARAMS[2],
                        finite_check, lm, rf, er, mk)
    return dict(n_inliers=r, pose_lm=lm, pose_refit=rf, lm_err=er, mask=mk)


def same(a, b):
    if a["n_inliers"] != b["n_inliers"] or not np.array_equal(a["mask"], b["mask"]) or not np.array_equal(a["lm_err"], b["lm_err"]):
        return False
    return a["n_inliers"] < 0 or (np.array_equal(a["pose_lm"], b["pose_lm"]) and np.array_equal(a["pose_refit"], b["pose_refit"]))


@pytest.mark.parametrize("variant", [0, 1])
def test_device_source_equals_oracle_bit_for_bit(host_lib, cams, variant):
    rng = np.random.default_rng(11 + variant)
    accepted = refits = 0
    for seed, n, n_align in [(100, 40, 5), (101, 40, 5), (102, 40, 6), (103, 9, 5), (104, 150, 5), (105, 150, 6), (106, 64, 5), (107, 33, 5)]:
        cl = make_cluster(seed, n=n, outliers=0.3 if n < 100 else 0.2)
        for h in range(12):
            pos = rng.choice(cl["good"], n_align, replace=False) if h % 3 else rng.choice(n, n_align, replace=False)
            quat = (rng.integers(0, 256, 4) / 256.0).astype(np.float32)
            o = oracle.hypothesis_depth(cl, cams, ALPHA, pos, quat, PARAMS[0], PARAMS[1], PARAMS[2], variant=variant)
            for width, order in ((1, 0), (32, 0), (32, 1), (8, 1)):          # 8 = four teams per warp (not used by the kernels yet)
                d = run_host(host_lib, variant, width, order, cl, cams, pos, quat)
                assert same(d, o), (variant, seed, h, width, order, d, o)
            accepted += o["n_inliers"] > PARAMS[2]
            refits += o["lm_err"][1] >= 0
    assert accepted >= 20 and refits >= 20          # the refit path (up to ~120 inliers = 360 residual rows) was exercised


@pytest.mark.skipif(not ref3d.available(), reason="oracle/_ref/libmoped3d_ref_strict.so not built (needs /root/reference at build time)")
@pytest.mark.parametrize("variant", [0, 1])
def test_device_source_equals_the_strict_reference_build(host_lib, cams, variant):
    """Directly against moped3d's own stage class compiled with strict IEEE arithmetic (levmar's non-finite stop kept)."""
    rng = np.random.default_rng(5 + variant)
    ref3d.use_strict(True)
    try:
        n_acc = 0
        for seed in range(6):
            cl = make_cluster(300 + seed)
            for h in range(10):
                pos = rng.choice(cl["good"], 5, replace=False) if h % 2 else rng.choice(len(cl["xy"]), 5, replace=False)
                quat = (rng.integers(0, 256, 4) / 256.0).astype(np.float32)
                r = ref3d.hypothesis(cl, K, CAM, ALPHA, pos, quat, PARAMS[0], PARAMS[1], PARAMS[2], variant=variant)
                d = run_host(host_lib, variant, 32, 0, cl, cams, pos, quat, finite_check=1)
                assert same(d, r), (variant, seed, h, d, r)
                n_acc += r["n_inliers"] > PARAMS[2]
        assert n_acc >= 10
    finally:
        ref3d.use_strict(False)


def test_degenerate_inputs(host_lib, cams):
    """Points behind the camera (the -z + 10 branch), an all-zero quaternion (NaN pose: LM runs to its iteration limit or stops,
    exactly as the oracle does), duplicate sample positions."""
    cl = make_cluster(7)
    cl["xyz"][:5] *= 40.0                                   # model points far outside: some land behind the camera during LM
    for variant in (0, 1):
        for pos, quat in [([0, 1, 2, 3, 4], [0.5, 0.5, 0.5, 0.5]), ([0, 0, 1, 1, 2], [0.25, 0.0, 0.0, 0.75]), ([5, 6, 7, 8, 9], [0, 0, 0, 0])]:
            o = oracle.hypothesis_depth(cl, cams, ALPHA, np.array(pos, np.int32), np.array(quat, np.float32), PARAMS[0], PARAMS[1], PARAMS[2], variant=variant)
            for width, order in ((1, 0), (32, 1)):
                d = run_host(host_lib, variant, width, order, cl, cams, pos, quat)
                assert d["n_inliers"] == o["n_inliers"] and np.array_equal(d["mask"], o["mask"])
                assert np.array_equal(d["lm_err"], o["lm_err"], equal_nan=True)
                if o["n_inliers"] >= 0:
                    assert np.array_equal(d["pose_lm"], o["pose_lm"], equal_nan=True) and np.array_equal(d["pose_refit"], o["pose_refit"], equal_nan=True)


@pytest.mark.parametrize("finite_check", [0, 1])
def test_moped2_reprojection_variant_equals_oracle_bit_for_bit(host_lib, finite_check):
    """variant 2 = the moped2 stage's residual (POSE_RANSAC_LM_DIFF_REPROJECTION_CPU) under the same order-preserving LM — the
    "pose_exact_order" mode of mc_pose_hypotheses / mc_pose_ransac — against mo_hypothesis, which tests/test_oracle_vs_strict_ref.py
    pins bit for bit to the strict build of that stage (with levmar's non-finite stop on, hence both settings here)."""
    from moped_b200 import synth
    cams2 = oracle.cameras(synth.K_DEFAULT, synth.CAM_IDENTITY)
    rng = np.random.default_rng(77)
    oracle.lib().mo_set_lm_finite_check(finite_check)
    try:
        accepted = 0
        for pts, frac, n_align, (max_lm, thr, min_npts) in ((80, 0.5, 5, (200, 10.0, 6)), (40, 0.1, 6, (500, 5.0, 8)), (150, 0.3, 5, (200, 10.0, 6))):
            cl = synth.make_ransac_clusters(3, pts, frac, seed=900 + pts)
            for k in range(3):
                s = slice(cl["offsets"][k], cl["offsets"][k + 1])
                xy, xyz, img = cl["xy"][s], cl["xyz"][s], cl["image"][s]
                n = len(xy)
                zeros3, zeros1 = np.zeros((n, 3), np.float32), np.zeros(n, np.float32)
                for h in range(10):
                    pos = rng.choice(n, n_align, replace=False).astype(np.int32)
                    quat = (rng.integers(0, 256, 4) / 256.0).astype(np.float32)
                    on, olm, orefit, oerr, omask = oracle.hypothesis(xy, xyz, img, cams2, pos, quat, max_lm, thr, min_npts)
                    for width, order in ((1, 0), (32, 0), (32, 1)):
                        lm, rf, er, mk = np.zeros(7, np.float32), np.zeros(7, np.float32), np.zeros(2, np.float32), np.zeros(n, np.uint8)
                        r = host_lib.dh_hypothesis(2, width, order, n, xy, xyz, zeros3, zeros1, img, C.addressof(cams2), 0.0, pos, n_align, quat,
                                                   max_lm, thr, min_npts, finite_check, lm, rf, er, mk)
                        assert r == on and np.array_equal(mk, omask) and np.array_equal(er, oerr), (pts, k, h, width, order, r, on, er, oerr)
                        if on >= 0:
                            assert np.array_equal(lm, olm) and np.array_equal(rf, orefit), (pts, k, h, width, order)
                    accepted += on > min_npts
        assert accepted >= 15
    finally:
        oracle.lib().mo_set_lm_finite_check(0)


def test_large_cluster_refit_and_eight_samples(host_lib, cams):
    """700 correspondences, ~630 inliers: the refit runs on 1256 (variant 0) / 1902 (variant 1) residual rows — no cap on the
    consistent set (the reference has none); NPtsAlign = 8 is the kernels' upper limit."""
    rng = np.random.default_rng(0)
    for variant in (0, 1):
        cl = make_cluster(900 + variant, n=700, outliers=0.1)
        for n_align in (5, 8):
            pos = rng.choice(cl["good"], n_align, replace=False)
            quat = np.array([0.5, 0.25, 0.75, 0.5], np.float32)
            o = oracle.hypothesis_depth(cl, cams, ALPHA, pos, quat, PARAMS[0], PARAMS[1], PARAMS[2], variant=variant)
            d = run_host(host_lib, variant, 32, 1, cl, cams, pos, quat)
            assert same(d, o), (variant, n_align)
        assert o["n_inliers"] >= 0
