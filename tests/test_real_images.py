"""REAL descriptors through the hot path (SURVEY.md §8d real-image config, BASELINE.json configs[0] substitute):
tests/golden/real_images.npz holds SIFT features the reference's own FEAT_SIFT_CPU/libsiftfast extracted from the
reference's shipped imagery (moped2/test_data/timing.bag frames, moped-example/test JPEGs), planar models built like
Moped::createPlanarModelsFromImages and loaded through the reference's sXML reader, and the outputs of the reference's
CPU stages in exact-matching mode (tests/golden/make_real_golden.py). Real descriptors bring what the synthetic ones do
not: exact duplicates (a model's own image: distance 0), repeated keypoint coordinates with different orientations,
clustered neighbours."""
import os

import numpy as np
import pytest

from conftest import ROOT, quat_angle


@pytest.fixture(scope="module")
def real():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "real_images.npz")))


def _frames(g):
    fo = g["frame_offsets"]
    for f in range(len(fo) - 1):
        yield f, g[f"f{f}_q_desc"], g["q_xy"][fo[f]:fo[f + 1]]


def test_oracle_on_real_descriptors(real, oracle_mod):
    """The C restatement equals the compiled reference on real data: 2-NN rows and distances bit for bit, matches,
    clusters."""
    g = real
    n_models = len(g["n_pts"])
    for f, q, xy in _frames(g):
        idx, dist = oracle_mod.match_2nn(g["db_desc"], q)
        assert np.array_equal(idx, g[f"f{f}_ann_idx"]), f
        assert np.array_equal(dist, g[f"f{f}_ann_dist"]), f
        img = np.zeros(len(q), np.int32)
        m, _, _ = oracle_mod.match(g["db_desc"], g["db_xyz"], g["model_of_row"], n_models, q, xy, img, 0.8)
        assert np.array_equal(m["offsets"], g[f"f{f}_match_offsets"]) and np.array_equal(m["xy"], g[f"f{f}_match_xy"])
        assert np.array_equal(m["xyz"], g[f"f{f}_match_xyz"])
        c = oracle_mod.cluster(m, 1, 200.0, 20.0, 7, 100)
        assert np.array_equal(c["model"], g[f"f{f}_cluster_model"]) and np.array_equal(c["offsets"], g[f"f{f}_cluster_offsets"])
        assert np.array_equal(c["members"], g[f"f{f}_cluster_members"])
    # the data really has the hard cases
    assert (g["f3_ann_dist"][:, 0] == 0).sum() >= 500            # a model's own image: exact duplicates
    xy0 = g["q_xy"][g["frame_offsets"][0]:g["frame_offsets"][1]]
    assert len(np.unique(xy0, axis=0)) < len(xy0)                 # same keypoint, several orientations


def _same_objects(models_a, poses_a, g, f):
    """Same objects as the reference's CPU pipeline. RANSAC draws differ (the reference uses rand(), the CUDA path its
    own counter-based stream), so poses agree only as well as the reference agrees with ITSELF under other seeds: the
    fixture records that spread per object (8 seeds); allowed = 2 x spread + (1 mm, 5 mrad)."""
    models_b, poses_b = g[f"f{f}_obj_model"], g[f"f{f}_obj_pose"]
    assert sorted(models_a.tolist()) == sorted(models_b.tolist()), (models_a, models_b)
    used = set()
    for m, p in zip(models_a, poses_a):
        best = None
        for k, (m2, p2) in enumerate(zip(models_b, poses_b)):
            if m2 != m or k in used:
                continue
            d = (np.abs(p[4:] - p2[4:]).max(), quat_angle(p[:4], p2[:4]))
            if best is None or d < best[1]:
                best = (k, d)
        used.add(best[0])
        tol_t = 2 * g[f"f{f}_obj_spread_t"][best[0]] + 1e-3
        tol_r = 2 * g[f"f{f}_obj_spread_r"][best[0]] + 5e-3
        assert best[1][0] < tol_t and best[1][1] < tol_r, (f, m, p, best, tol_t, tol_r)


@pytest.mark.gpu
def test_cuda_path_on_real_descriptors(real, gpu_ctx, oracle_mod):
    from moped_b200 import capi
    g = real
    n_models = len(g["n_pts"])
    gpu_ctx.db_upload(g["db_desc"], g["db_xyz"], g["model_of_row"], n_models)
    gpu_ctx.set_cameras(g["K"], g["cam_pose"])
    gpu_ctx.set_tuning(8, 8, 1)
    singles = []
    for f, q, xy in _frames(g):
        img = np.zeros(len(q), np.int32)
        for mode in (capi.MATCH_TENSOR, capi.MATCH_EXACT):
            rows, dist, acc, stats = gpu_ctx.match(q, 0.8, mode)
            assert np.array_equal(rows, g[f"f{f}_ann_idx"]), (f, mode)          # bit-exact against the reference's exact mode
            assert np.array_equal(dist, g[f"f{f}_ann_dist"]), (f, mode)
            with np.errstate(divide="ignore", invalid="ignore"):
                assert np.array_equal(acc, g[f"f{f}_ann_dist"][:, 0] / g[f"f{f}_ann_dist"][:, 1] < np.float32(0.8))
        m = dict(offsets=g[f"f{f}_match_offsets"], image=g[f"f{f}_match_image"], xy=g[f"f{f}_match_xy"], xyz=g[f"f{f}_match_xyz"])
        c = gpu_ctx.cluster(m, 1, 200.0, 20.0, 7, 100)
        assert np.array_equal(c["model"], g[f"f{f}_cluster_model"]) and np.array_equal(c["offsets"], g[f"f{f}_cluster_offsets"])
        assert np.array_equal(c["members"], g[f"f{f}_cluster_members"])
        out = gpu_ctx.process_frame(q, xy, img, max_objects=64)
        _same_objects(out["model"], out["pose"], g, f)
        singles.append(out)
    # the planar pose of a model's own image is known in closed form: R = I, t = (-cx, -cy, f) * scale
    own = singles[3]
    k = list(own["model"]).index(0)
    K, s = g["K"][0], float(g["scale"])
    assert np.abs(own["pose"][k][4:] - np.array([-K[2] * s, -K[3] * s, K[0] * s])).max() < 2e-3
    assert quat_angle(own["pose"][k][:4], np.array([0, 0, 0, 1.0])) < 5e-3
    # and as one batch
    fo = g["frame_offsets"]
    q_all = np.concatenate([q for _, q, _ in _frames(g)])
    batch = gpu_ctx.process_frames(q_all, g["q_xy"], np.zeros(len(q_all), np.int32), fo, max_objects=64)
    for b, s1 in zip(batch, singles):
        assert np.array_equal(b["model"], s1["model"]) and np.array_equal(b["pose"], s1["pose"])


@pytest.mark.gpu
def test_images_in_objects_out(real, gpu_ctx):
    """mc_process_images: FEAT chained to MATCH..FILTER2 on the device, on a frame the reference ships (timing.bag frame 4
    = frame 1 of the fixture, stored as an image in tests/golden/sift_golden.npz). Same objects as the reference's CPU
    pipeline run on the reference's own features, poses within the reference's seed-to-seed spread; and the same objects,
    pose for pose, as feeding the device-extracted features through the host entry points by hand."""
    g = real
    sg = np.load(os.path.join(ROOT, "tests", "golden", "sift_golden.npz"))
    im = sg["bag4_full_double/image"]
    assert g["frame_names"][1] == "bag4" and len(sg["bag4_full_double/xy"]) == len(g["f1_q_desc"])
    n_models = len(g["n_pts"])
    gpu_ctx.db_upload(g["db_desc"], g["db_xyz"], g["model_of_row"], n_models)
    gpu_ctx.set_cameras(g["K"], g["cam_pose"])
    gpu_ctx.set_tuning(8, 8, 1)
    frames = np.stack([im, im, np.ascontiguousarray(im[:, ::-1])])        # the third frame (mirrored) shows no known object pose-consistently
    out = gpu_ctx.process_images(frames, True, max_keypoints=2048, max_objects=64, want_times=True)
    assert out[0]["n_features"] in (582, 583) and out[0]["stage_ms"][0] > 0
    _same_objects(out[0]["model"], out[0]["pose"], g, 1)
    assert np.array_equal(out[0]["model"], out[1]["model"]) and np.array_equal(out[0]["pose"], out[1]["pose"])
    # by hand: features to the host, the MATCH stage's normalisation on the host, then the frame entry point
    xy, so, desc = gpu_ctx.sift(im, True)
    ss = np.zeros(len(desc), np.float32)
    for k in range(128):
        ss = ss + desc[:, k] * desc[:, k]                                  # sequential fp32 sum like MATCH_CUDA::normalise
    q = desc * (1.0 / np.sqrt(ss).astype(np.float64)).astype(np.float32)[:, None]
    hand = gpu_ctx.process_frames(q, xy, np.zeros(len(q), np.int32), np.array([0, len(q)], np.int32), max_objects=64)[0]
    assert np.array_equal(hand["model"], out[0]["model"]) and np.array_equal(hand["pose"], out[0]["pose"])
    with pytest.raises(Exception, match="max_keypoints"):
        gpu_ctx.process_images(frames, True, max_keypoints=100)
