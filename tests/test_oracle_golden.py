"""The C restatement (oracle/moped_oracle.c) against the golden vectors produced by the reference's own
stage classes (tests/golden/make_golden.py). CPU only."""
import numpy as np

from conftest import cluster_points, golden_matches, quat_angle


def test_match_2nn_bit_exact(golden, oracle_mod):
    idx, dist = oracle_mod.match_2nn(golden["db_desc"], golden["q_desc"])
    assert np.array_equal(idx, golden["ann_idx"])
    assert np.array_equal(dist, golden["ann_dist"])            # same summation order, same bits


def test_match_emit(golden, oracle_mod):
    m, _, _ = oracle_mod.match(golden["db_desc"], golden["db_xyz"], golden["model_of_row"], len(golden["n_pts"]), golden["q_desc"],
                               golden["q_xy"], golden["q_image"], 0.8)
    for k in ("offsets", "image", "xy", "xyz"):
        assert np.array_equal(m[k], golden["match_" + k]), k


def test_norm_rows_close(golden, oracle_mod):
    # the reference normalises with RSQRTSS + Newton under -ffast-math: equal to 1 ulp, idempotent input
    again = oracle_mod.norm_rows(golden["q_desc"])
    assert np.abs(again - golden["q_desc"]).max() <= 2e-7


def test_cluster(golden, oracle_mod):
    m = golden_matches(golden)
    for tag, prm in (("", (200.0, 20.0, 7, 100)), ("hi_", (150.0, 20.0, 7, 100)), ("lo_", (60.0, 20.0, 7, 100)), ("it1_", (200.0, 20.0, 7, 1))):
        c = oracle_mod.cluster(m, 2, *prm)
        for k in ("model", "offsets", "members"):
            assert np.array_equal(c[k], golden[f"cluster_{tag}{k}"]), (tag, k)


def test_draw_samples(golden, oracle_mod):
    m = golden_matches(golden)
    c = dict(model=golden["cluster_model"], offsets=golden["cluster_offsets"], members=golden["cluster_members"])
    xy, xyz, img, tie, co = cluster_points(m, c)
    H = int(golden["hyp_H"])
    for k in range(len(c["model"])):
        s = slice(co[k], co[k + 1])
        ok, pos, quat = oracle_mod.draw_samples(xy[s], img[s], tie[s], 5, int(golden["hyp_seed0"]) + k, H)
        assert ok == H
        assert np.array_equal(pos, golden["hyp_pos"][k * H:(k + 1) * H])
        assert np.array_equal(quat, golden["hyp_quat"][k * H:(k + 1) * H])


def test_hypotheses(golden, oracle_mod):
    """Same (samples, init quat) -> same accept decision and pose within tolerance. LM in fp32 on a quartic
    cost is chaotic for badly conditioned 5-point fits, so the agreement is statistical: the two CPU builds
    (reference with -ffast-math, restatement without) already differ on a few percent of the hypotheses."""
    m = golden_matches(golden)
    c = dict(model=golden["cluster_model"], offsets=golden["cluster_offsets"], members=golden["cluster_members"])
    xy, xyz, img, tie, co = cluster_points(m, c)
    cams = oracle_mod.cameras(golden["K"], golden["cam_pose"])
    n = len(golden["hyp_cluster"])
    agree, both, dt, dr, mask_equal = 0, 0, [], [], 0
    mo = 0
    for h in range(n):
        k = int(golden["hyp_cluster"][h]); s = slice(co[k], co[k + 1]); sz = co[k + 1] - co[k]
        r, plm, prf, err, mask = oracle_mod.hypothesis(xy[s], xyz[s], img[s], cams, golden["hyp_pos"][h], golden["hyp_quat"][h], 200, 10.0, 6)
        gr = int(golden["hyp_n_inl"][h])
        agree += (r > 6) == (gr > 6)
        if r > 6 and gr > 6:
            both += 1
            dt.append(np.abs(prf[4:] - golden["hyp_pose_refit"][h][4:]).max())
            dr.append(quat_angle(prf[:4], golden["hyp_pose_refit"][h][:4]))
            mask_equal += np.array_equal(mask, golden["hyp_mask"][mo:mo + sz])
        mo += sz
    dt, dr = np.array(dt), np.array(dr)
    assert agree >= 0.95 * n, (agree, n)
    assert both >= 0.7 * n
    assert np.median(dt) < 1e-4 and np.percentile(dt, 90) < 5e-4 and dt.max() < 5e-3, (np.median(dt), dt.max())
    assert np.median(dr) < 1e-3 and dr.max() < 2e-2
    assert mask_equal >= 0.9 * both


def test_ransac(golden, oracle_mod):
    m = golden_matches(golden)
    c = dict(model=golden["cluster_model"], offsets=golden["cluster_offsets"], members=golden["cluster_members"])
    xy, xyz, img, tie, co = cluster_points(m, c)
    cams = oracle_mod.cameras(golden["K"], golden["cam_pose"])
    for k in range(len(c["model"])):
        s = slice(co[k], co[k + 1])
        f, pose, it = oracle_mod.ransac(xy[s], xyz[s], img[s], tie[s], cams, (600, 200, 4, 5, 6, 10.0), int(golden["ransac_seed0"]) + k)
        assert f == int(golden["ransac_found"][k])
        assert np.abs(pose[4:] - golden["ransac_pose"][k][4:]).max() < 1e-3
        assert quat_angle(pose[:4], golden["ransac_pose"][k][:4]) < 5e-3


def test_project(golden, oracle_mod):
    cams = oracle_mod.cameras(golden["K"], golden["cam_pose"])
    uv = oracle_mod.project(golden["project_pose"], golden["match_xyz"][:64], golden["match_image"][:64], cams)
    assert np.abs(uv - golden["project_uv"]).max() < 2e-3


def test_filter(golden, oracle_mod):
    m = golden_matches(golden)
    cams = oracle_mod.cameras(golden["K"], golden["cam_pose"])
    f = oracle_mod.filter_objects(m, cams, golden["filter_in_model"], golden["filter_in_pose"], (5, 4096.0, 2.0))
    keep = f["keep"]
    assert np.array_equal(golden["filter_in_model"][keep], golden["filter_out_model"])
    assert np.array_equal(golden["filter_in_pose"][keep], golden["filter_out_pose"])
    assert np.abs(f["score"][keep] - golden["filter_out_score"]).max() < 1e-3
    assert np.array_equal(f["offsets"], golden["filter_cluster_offsets"])
    assert np.array_equal(f["members"], golden["filter_cluster_members"])
