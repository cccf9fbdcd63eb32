"""EXACT-ORDER mode of the POSE / POSE2 steps (mc_set_option "pose_exact_order", 1) inside the device-resident frame pipeline: the
staged RANSAC driver with the order-preserving LM (pose_exact.cu). With MATCH exact by certificate, CLUSTER order-preserving and
FILTER compiled without fused multiply-add, a whole frame equals the oracle's stage chain (tests/oracle_chain.py: the C restatement,
bit-identical to the strict-IEEE build of the reference's own stage classes) — the same objects, in the same order, with the same pose
BITS. north_star's RANSAC/LM gate (1e-3 rad, 1e-4 m per accepted hypothesis on identical sample sets) is met with zero difference."""
import numpy as np
import pytest

import oracle_chain
from conftest import cluster_points, golden_matches

pytestmark = pytest.mark.gpu


@pytest.fixture()
def exact_ctx(gpu_ctx):
    gpu_ctx.set_option("pose_exact_order", 1)
    yield gpu_ctx
    gpu_ctx.set_option("pose_exact_order", 0)


def _same_objects(out, want):
    assert np.array_equal(out["model"], want["model"]), (out["model"], want["model"])
    assert np.array_equal(out["pose"], want["pose"]), np.abs(out["pose"] - want["pose"]).max()
    assert np.array_equal(out["score"], want["score"]), np.abs(out["score"] - want["score"]).max()


def test_staged_exact_ransac_equals_the_oracle_on_golden_clusters(exact_ctx, golden, oracle_mod):
    """mc_pose_ransac in exact-order mode runs the staged driver (first hypotheses of all tasks, escalating levels, warp refit): for POSE
    and POSE2 parameters the same test wins as in the oracle's sequential loop and the refitted pose has the same bits — also with the
    first-round width of the batch setting (pose_warps_per_task 1 and 4) and for clusters whose refit has more than 512 inliers."""
    m = golden_matches(golden)
    c = dict(model=golden["cluster_model"], offsets=golden["cluster_offsets"], members=golden["cluster_members"])
    xy, xyz, img, tie, co = cluster_points(m, c)
    exact_ctx.set_cameras(golden["K"], golden["cam_pose"])
    cams = oracle_mod.cameras(golden["K"], golden["cam_pose"])
    for warps in (8, 4, 1):
        exact_ctx.set_tuning(0, warps, 0)
        for P in ((600, 200, 4, 5, 6, 10.0), (100, 500, 4, 6, 8, 5.0)):
            found, pose, nt = exact_ctx.pose_ransac(co, xy, xyz, img, P, seed=7)
            for task in range(len(found)):
                k = task // P[2]; s = slice(co[k], co[k + 1])
                f, p, it = oracle_mod.ransac(xy[s], xyz[s], img[s], None, cams, P, (7 + oracle_chain.GOLDEN * (task + 1)) & oracle_chain.M64)
                assert bool(found[task]) == bool(f) and nt[task] == it, (warps, P, task, found[task], f, nt[task], it)
                if f:
                    assert np.array_equal(pose[task], p), (warps, P, task)
            assert found.any()
    exact_ctx.set_tuning(0, 8, 0)
    # a cluster with ~700 inliers (the default kernels cap a refit at 512; this mode does not) and one that exhausts its tests
    from moped_b200 import synth
    big = synth.make_ransac_clusters(1, 800, 0.1, seed=99)
    junk = synth.make_ransac_clusters(1, 40, 1.0, seed=98)
    off = np.array([0, 800, 840], np.int32)
    bxy = np.concatenate([big["xy"], junk["xy"]]); bxyz = np.concatenate([big["xyz"], junk["xyz"]]); bim = np.zeros(840, np.int32)
    exact_ctx.set_cameras(synth.K_DEFAULT, synth.CAM_IDENTITY)
    cams2 = oracle_mod.cameras(synth.K_DEFAULT, synth.CAM_IDENTITY)
    P = (48, 200, 2, 5, 6, 10.0)
    found, pose, nt = exact_ctx.pose_ransac(off, bxy, bxyz, bim, P, seed=3)
    for task in range(4):
        k = task // 2; s = slice(off[k], off[k + 1])
        f, p, it = oracle_mod.ransac(bxy[s], bxyz[s], bim[s], None, cams2, P, (3 + oracle_chain.GOLDEN * (task + 1)) & oracle_chain.M64)
        assert bool(found[task]) == bool(f) and nt[task] == it, task
        if f:
            assert np.array_equal(pose[task], p), task
    assert found[:2].all() and not found[2:].any() and (nt[2:] == 48).all()


def test_process_frame_in_exact_order_mode_equals_the_oracle_chain(exact_ctx, small_case):
    from moped_b200 import synth
    c = small_case
    exact_ctx.db_upload(c["dbn"], c["db"]["xyz"], c["db"]["model_of_row"], c["n_obj"])
    exact_ctx.set_cameras(synth.K_DEFAULT, synth.CAM_IDENTITY)
    want = oracle_chain.frame(c["dbn"], c["db"]["xyz"], c["db"]["model_of_row"], c["n_obj"], c["qn"], c["fr"]["xy"], c["fr"]["image_idx"],
                              synth.K_DEFAULT, synth.CAM_IDENTITY)
    assert sorted(want["model"].tolist()) == sorted(c["fr"]["gt_model"].tolist())
    for warps in (8, 1):
        exact_ctx.set_tuning(0, warps, 0)
        _same_objects(exact_ctx.process_frame(c["qn"], c["fr"]["xy"], c["fr"]["image_idx"]), want)
    exact_ctx.set_tuning(0, 8, 0)


def test_filter_scores_are_bit_exact(gpu_ctx, golden, oracle_mod):
    """filter.cu is compiled without fused multiply-add: reprojection errors, and with them scores, ownership and pruning, are the oracle's."""
    m = golden_matches(golden)
    cams = oracle_mod.cameras(golden["K"], golden["cam_pose"])
    gpu_ctx.set_cameras(golden["K"], golden["cam_pose"])
    for prm in ((5, 4096.0, 2.0), (7, 4096.0, 3.0)):
        of = oracle_mod.filter_objects(m, cams, golden["filter_in_model"], golden["filter_in_pose"], prm)
        gf = gpu_ctx.filter(m, golden["filter_in_model"], golden["filter_in_pose"], prm)
        assert np.array_equal(of["score"], gf["score"]), np.abs(of["score"] - gf["score"]).max()


def test_frame_batch_in_exact_order_mode_equals_the_oracle_chain_per_frame(exact_ctx, oracle_mod):
    """mc_process_frames (one MATCH pass, stage chains of the frames on concurrent lanes, CUDA graphs) in exact-order mode: every frame of
    a batch — different feature counts, an empty frame — equals the oracle chain bit for bit."""
    from moped_b200 import synth
    db = synth.make_db(24, 600, seed=77)
    dbn = oracle_mod.norm_rows(db["desc"])
    frames = [synth.make_frame(db, q, n_visible=v, frame_id=40 + i, seed=77) for i, (q, v) in enumerate(((1500, 3), (700, 2), (2000, 5), (900, 0)))]
    qn = [oracle_mod.norm_rows(f["desc"]) for f in frames]
    sizes = [len(q) for q in qn]
    sizes.insert(2, 0)                                   # an empty frame in the middle
    fo = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    exact_ctx.db_upload(dbn, db["xyz"], db["model_of_row"], 24)
    exact_ctx.set_cameras(synth.K_DEFAULT, synth.CAM_IDENTITY)
    exact_ctx.set_tuning(4, 4, 0)
    try:
        for rep in range(2):                             # second call replays the captured graphs
            out = exact_ctx.process_frames(np.concatenate(qn), np.concatenate([f["xy"] for f in frames]), np.concatenate([f["image_idx"] for f in frames]),
                                           fo, None, 32)
            k = 0
            for slot, n in enumerate(sizes):
                if n == 0:
                    assert len(out[slot]["model"]) == 0
                    continue
                f = frames[k]
                want = oracle_chain.frame(dbn, db["xyz"], db["model_of_row"], 24, qn[k], f["xy"], f["image_idx"], synth.K_DEFAULT, synth.CAM_IDENTITY)
                _same_objects(out[slot], want)
                assert sorted(want["model"].tolist()) == sorted(f["gt_model"].tolist()), (slot, want["model"], f["gt_model"])
                k += 1
    finally:
        exact_ctx.set_tuning(8, 8, 0)
