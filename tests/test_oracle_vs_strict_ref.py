"""The C restatements against a STRICT-IEEE build of the reference's own sources (oracle/_ref/libmoped_ref_strict.so: the
same files as oracle/_ref/libmoped_ref.so, compiled without -ffast-math and without contraction; oracle/Makefile).

Why: the reference's own flags (-ffast-math) let its compiler re-associate and contract, so against THAT build the
floating-point stages can only be compared to a tolerance (tests/test_oracle_vs_ref.py, tests/test_sift_oracle.py). Against the
strict build the restatements must be exact — this separates "is the restatement right" (bit-exact here) from "how far does
the reference move under its own fast-math" (the tolerances of the other tests):
  * norm(), Levenberg-Marquardt hypotheses (accept decision, inlier mask, both poses, ||e||^2) and whole RANSAC runs on the
    shared seedable stream: bit for bit;
  * feature extraction: the whole keypoint list (order, coord2D, descriptors) bit for bit, with the oracle's convolution taps
    switched from FMA (the variant the CUDA kernels implement) to multiply-then-add (what the strict build computes).
"""
import os

import numpy as np
import pytest

from conftest import cluster_points


@pytest.fixture(scope="module")
def strict_ref():
    from oracle import ref
    if not (ref.available() and ref.strict_available()):
        pytest.skip("oracle/_ref/libmoped_ref_strict.so not built (needs /root/reference at build time)")
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle import oracle
    ref.use_strict(True)
    oracle.lib().mo_set_lm_finite_check(1)      # a strict build keeps levmar's stop=7 (non-finite ||e||^2); -ffast-math folds it away
    yield ref
    oracle.lib().mo_set_lm_finite_check(0)
    ref.use_strict(False)


def test_norm_rows_bit_exact(strict_ref, oracle_mod):
    rng = np.random.default_rng(5)
    d = rng.gamma(0.5, size=(300, 128)).astype(np.float32)
    assert np.array_equal(strict_ref.Ref.norm_rows(d), oracle_mod.norm_rows(d))


def test_lm_hypotheses_and_ransac_bit_exact(strict_ref, oracle_mod):
    from moped_b200 import synth
    db = synth.make_db(10, 400, seed=11)
    fr = synth.make_frame(db, 800, n_visible=3, pts_visible=50, seed=11)
    r = strict_ref.Ref(1)
    r.set_models(db["n_pts"], db["xyz"], db["desc"])
    r.set_images(synth.K_DEFAULT, synth.CAM_IDENTITY)
    r.set_features(fr["desc"], fr["xy"], fr["image_idx"])
    r.clear_frame(); r.run_match(0.0, 0.8); r.run_cluster()
    m, c = r.get_matches(), r.get_clusters()
    assert len(c["model"]) >= 3
    cams = oracle_mod.cameras(synth.K_DEFAULT, synth.CAM_IDENTITY)
    xy, xyz, img, tie, co = cluster_points(m, c)
    n_hyp = n_acc = 0
    for k in range(len(c["model"])):
        s = slice(co[k], co[k + 1])
        mem = c["members"][c["offsets"][k]:c["offsets"][k + 1]]
        model = int(c["model"][k])
        for n_align, params in ((5, (600, 200, 4, 5, 6, 10.0)), (6, (100, 500, 4, 6, 8, 5.0))):
            ok, pos, quat = r.draw_samples(model, mem, n_align, 500 + k, 24)
            assert ok
            for h in range(len(pos)):
                rn, rlm, rrefit, rerr, rmask = r.hypothesis(model, mem, pos[h], quat[h], params[1], params[5], params[4])
                on, olm, orefit, oerr, omask = oracle_mod.hypothesis(xy[s], xyz[s], img[s], cams, pos[h], quat[h], params[1], params[5], params[4])
                assert rn == on and np.array_equal(rmask, omask) and np.array_equal(rerr, oerr), (k, h)
                if rn >= 0:
                    assert np.array_equal(rlm, olm) and np.array_equal(rrefit, orefit), (k, h)
                n_hyp += 1
                n_acc += rn > params[4]
            f, pose = r.ransac(model, mem, params, 40 + k)
            f2, pose2, _ = oracle_mod.ransac(xy[s], xyz[s], img[s], tie[s], cams, params, 40 + k)
            assert f == f2
            if f:
                assert np.array_equal(pose, pose2)
    assert n_hyp >= 100 and n_acc >= 10


def test_sift_bit_exact_with_multiply_add_taps(strict_ref, oracle_mod):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "sift_golden.npz"))
    oracle_mod.lib().mo_sift_set_conv_fma(0)
    try:
        for name in ("bag0_crop_double", "bag1_crop_single", "ex2_odd_double", "ex2_odd_single"):
            im, dbl = g[f"{name}/image"], bool(g[f"{name}/double"])
            rxy, rdesc = strict_ref.sift(im, dbl)
            xy, so, desc = oracle_mod.sift(im, dbl)
            assert len(xy) == len(rxy) > 50, (name, len(xy), len(rxy))
            assert np.array_equal(xy, rxy), name
            assert np.array_equal(desc, rdesc), (name, np.abs(desc - rdesc).max())
    finally:
        oracle_mod.lib().mo_sift_set_conv_fma(1)
