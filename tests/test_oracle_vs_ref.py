"""The C restatement against the compiled reference itself (oracle/_ref) on seeded inputs, including the
edge cases the domain has. Skipped where oracle/_ref is not built (it needs /root/reference)."""
import numpy as np
import pytest

from conftest import cluster_points


def _setup(ref_mod, db, K=None, cam=None, threads=1):
    from moped_b200 import synth
    r = ref_mod.Ref(threads)
    r.set_models(db["n_pts"], db["xyz"], db["desc"])
    r.set_images(synth.K_DEFAULT if K is None else K, synth.CAM_IDENTITY if cam is None else cam)
    return r


@pytest.mark.parametrize("n_obj,pts,q,ragged", [(5, 300, 400, False), (12, 100, 257, True), (3, 50, 64, False)])
def test_match_exact_mode(ref_mod, oracle_mod, n_obj, pts, q, ragged):
    from moped_b200 import synth
    db = synth.make_db(n_obj, pts, seed=77 + n_obj, ragged=ragged)
    fr = synth.make_frame(db, q, n_visible=min(3, n_obj), pts_visible=30, seed=77)
    r = _setup(ref_mod, db)
    r.set_features(fr["desc"], fr["xy"], fr["image_idx"])
    r.clear_frame(); r.run_match(0.0, 0.8)
    dbn, qn = r.model_desc(), r.features_desc()
    ridx, rdist = r.ann_search(qn, 0.0)
    oidx, odist = oracle_mod.match_2nn(dbn, qn)
    assert np.array_equal(ridx, oidx) and np.array_equal(rdist, odist)
    rm = r.get_matches()
    om, _, _ = oracle_mod.match(dbn, db["xyz"], db["model_of_row"], n_obj, qn, fr["xy"], fr["image_idx"], 0.8)
    for k in ("offsets", "image", "xy", "xyz"):
        assert np.array_equal(rm[k], om[k]), k


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_meanshift_random_groups(ref_mod, oracle_mod, seed):
    """Random point clouds per (model, image), sizes around MinPts and well above, two images."""
    rng = np.random.default_rng(seed)
    n_models = 6
    sizes = rng.integers(0, 90, size=n_models)
    sizes[0] = 0; sizes[1] = 6
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    M = int(off[-1])
    centres = rng.uniform(50, 600, size=(n_models, 3, 2))
    xy = np.concatenate([centres[m][rng.integers(0, 3, size=sizes[m])] + rng.normal(0, 15, size=(sizes[m], 2)) for m in range(n_models)]).astype(np.float32)
    img = rng.integers(0, 2, size=M).astype(np.int32)
    m = dict(offsets=off, image=img, xy=xy, xyz=rng.uniform(-0.1, 0.1, size=(M, 3)).astype(np.float32))
    from moped_b200 import synth
    db = synth.make_db(n_models, 10, seed=5)
    r = _setup(ref_mod, db, K=np.stack([synth.K_DEFAULT] * 2), cam=np.stack([synth.CAM_IDENTITY] * 2))
    r.set_matches(m)
    for prm in ((200.0, 20.0, 7, 100), (60.0, 20.0, 7, 100), (40.0, 30.0, 3, 2)):
        r.run_cluster(*prm)
        rc = r.get_clusters()
        oc = oracle_mod.cluster(m, 2, *prm)
        for k in ("model", "offsets", "members"):
            assert np.array_equal(rc[k], oc[k]), (prm, k)


def test_pose_and_filter_chain(ref_mod, oracle_mod):
    from moped_b200 import synth
    db = synth.make_db(10, 400, seed=11)
    fr = synth.make_frame(db, 800, n_visible=3, pts_visible=50, seed=11)
    r = _setup(ref_mod, db)
    r.set_features(fr["desc"], fr["xy"], fr["image_idx"])
    r.clear_frame(); r.run_match(0.0, 0.8); r.run_cluster()
    m, c = r.get_matches(), r.get_clusters()
    assert len(c["model"]) >= 3
    cams = oracle_mod.cameras(synth.K_DEFAULT, synth.CAM_IDENTITY)
    xy, xyz, img, tie, co = cluster_points(m, c)
    # sampling streams are identical
    for k in range(len(c["model"])):
        s = slice(co[k], co[k + 1])
        mem = c["members"][c["offsets"][k]:c["offsets"][k + 1]]
        ok, pos, quat = r.draw_samples(int(c["model"][k]), mem, 6, 300 + k, 8)
        ok2, pos2, quat2 = oracle_mod.draw_samples(xy[s], img[s], tie[s], 6, 300 + k, 8)
        assert ok == ok2 and np.array_equal(pos, pos2) and np.array_equal(quat, quat2)
        f, pose = r.ransac(int(c["model"][k]), mem, (100, 500, 4, 6, 8, 5.0), 40 + k)
        f2, pose2, _ = oracle_mod.ransac(xy[s], xyz[s], img[s], tie[s], cams, (100, 500, 4, 6, 8, 5.0), 40 + k)
        assert f == f2
        if f:
            assert np.abs(pose[4:] - pose2[4:]).max() < 1e-3
    # filter on the reference's own POSE output
    r.run_pose("POSE", (600, 200, 4, 5, 6, 10.0), seed=9)
    ob = r.get_objects()
    r.run_filter((5, 4096.0, 2.0))
    ob2, c2 = r.get_objects(), r.get_clusters()
    f = oracle_mod.filter_objects(m, cams, ob["model"], ob["pose"], (5, 4096.0, 2.0))
    assert np.array_equal(ob["model"][f["keep"]], ob2["model"])
    assert np.array_equal(f["offsets"], c2["offsets"]) and np.array_equal(f["members"], c2["members"])


def test_sample_failure_when_too_few_distinct_points(ref_mod, oracle_mod):
    """randSample fails when the cluster has fewer distinct (image, coord2D) than NPtsAlign."""
    xy = np.array([[10, 10]] * 4 + [[20, 20]] * 3, np.float32)
    img = np.zeros(7, np.int32)
    ok, pos, quat = oracle_mod.draw_samples(xy, img, None, 5, 3, 2)
    assert ok == 0


def test_hypotheses_batch_equals_single_calls(ref_mod):
    """bench.py's configs[3] CPU baseline (ref_hypotheses_batch, OpenMP over hypotheses) is the same computation as
    the per-hypothesis entry the parity tests use."""
    from moped_b200 import synth
    cl = synth.make_ransac_clusters(2, 80, 0.5)
    hy = synth.make_hypotheses(cl, 24, 5)
    assert hy["sample_pos"].shape == (48, 5) and all(len(set(r)) == 5 for r in hy["sample_pos"].tolist())
    r = ref_mod.Ref(2)
    r.set_models(np.diff(cl["offsets"]).astype(np.int32), cl["xyz"], np.full((len(cl["xyz"]), 128), 0.1, np.float32))
    r.set_images(synth.K_DEFAULT, synth.CAM_IDENTITY)
    r.set_matches(dict(offsets=cl["offsets"], image=cl["image"], xy=cl["xy"], xyz=cl["xyz"]))
    members = np.arange(80, dtype=np.int32)
    for c in range(2):
        sel = slice(24 * c, 24 * (c + 1))
        sec, n_in, pose = r.hypotheses_batch(c, members, hy["sample_pos"][sel], hy["init_quat"][sel], 200, 10.0, 6)
        assert sec > 0
        for k in range(24):
            ni, _, prf, _, _ = r.hypothesis(c, members, hy["sample_pos"][24 * c + k], hy["init_quat"][24 * c + k], 200, 10.0, 6)
            assert ni == n_in[k]
            if ni >= 0:
                assert np.array_equal(prf, pose[k])
