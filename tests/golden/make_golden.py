"""Generates tests/golden/moped_golden.npz from the reference's OWN stage classes (oracle/_ref, compiled from
/root/reference by oracle/Makefile). Run in the build container only (needs /root/reference for the build):

    make -f oracle/Makefile ref && python tests/golden/make_golden.py

The fixture holds the inputs (as normalised by the reference) and the reference's outputs for every stage of
the hot path; tests/test_oracle_golden.py pins the C restatement to it and the -m gpu tests pin the CUDA path."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from moped_b200 import synth
from oracle import ref

SEED = 4242
n_obj, pts = 6, 200
db = synth.make_db(n_obj, pts, seed=SEED)
fr = synth.make_frame(db, 320, n_visible=3, pts_visible=45, seed=SEED)
# a second camera / image so that the per-image split of CLUSTER and the (image, coord2D) keys are exercised
fr2 = synth.make_frame(db, 120, n_visible=2, pts_visible=40, frame_id=1, seed=SEED, image_idx=1)
desc = np.concatenate([fr["desc"], fr2["desc"]]); xy = np.concatenate([fr["xy"], fr2["xy"]]); img = np.concatenate([fr["image_idx"], fr2["image_idx"]])
# duplicate (image, coord2D) keys: a feature detected twice at the same pixel with different orientations
xy[5] = xy[3]; xy[200] = xy[3]
K = np.stack([synth.K_DEFAULT, np.array([790.0, 805.0, 310.0, 250.0], np.float32)])
cam = np.stack([synth.CAM_IDENTITY, np.array([0.0, 0.0087265, 0.0, 0.99996192, 0.02, -0.01, 0.0], np.float32)])

r = ref.Ref(1)
r.set_models(db["n_pts"], db["xyz"], db["desc"]); r.set_images(K, cam); r.set_features(desc, xy, img)
r.clear_frame(); r.run_match(0.0, 0.8)
dbn, qn = r.model_desc(), r.features_desc()          # what the kd-tree and the queries really held
ann_idx, ann_dist = r.ann_search(qn, 0.0)
ann5_idx, _ = r.ann_search(qn, 5.0)
m = r.get_matches()
r.run_cluster(200.0, 20.0, 7, 100)
c = r.get_clusters()
out = dict(n_pts=db["n_pts"], db_xyz=db["xyz"], db_desc=dbn, model_of_row=db["model_of_row"], q_desc=qn, q_xy=xy, q_image=img, K=K, cam_pose=cam,
           ann_idx=ann_idx, ann_dist=ann_dist, ann5_idx=ann5_idx,
           match_offsets=m["offsets"], match_image=m["image"], match_xy=m["xy"], match_xyz=m["xyz"],
           cluster_model=c["model"], cluster_offsets=c["offsets"], cluster_members=c["members"])
# other CLUSTER parameter sets seen in the tree
for tag, prm in (("hi", (150.0, 20.0, 7, 100)), ("lo", (60.0, 20.0, 7, 100)), ("it1", (200.0, 20.0, 7, 1))):
    r.run_cluster(*prm); cc = r.get_clusters()
    out[f"cluster_{tag}_model"], out[f"cluster_{tag}_offsets"], out[f"cluster_{tag}_members"] = cc["model"], cc["offsets"], cc["members"]
r.run_cluster(200.0, 20.0, 7, 100)
# hypotheses: (sample set, init quat) drawn by the reference's own randSample/initPose, then its RANSAC body
H = 24
hyp = dict(cluster=[], pos=[], quat=[], n_inl=[], pose_lm=[], pose_refit=[], err=[], mask=[])
for k in range(len(c["model"])):
    mem = c["members"][c["offsets"][k]:c["offsets"][k + 1]]; mdl = int(c["model"][k])
    ok, pos, quat = r.draw_samples(mdl, mem, 5, 9000 + k, H)
    assert ok == H
    for h in range(H):
        n, plm, prf, err, mask = r.hypothesis(mdl, mem, pos[h], quat[h], 200, 10.0, 6)
        hyp["cluster"].append(k); hyp["pos"].append(pos[h]); hyp["quat"].append(quat[h]); hyp["n_inl"].append(n)
        hyp["pose_lm"].append(plm); hyp["pose_refit"].append(prf); hyp["err"].append(err); hyp["mask"].append(mask)
out.update(hyp_cluster=np.array(hyp["cluster"], np.int32), hyp_pos=np.array(hyp["pos"], np.int32), hyp_quat=np.array(hyp["quat"], np.float32),
           hyp_n_inl=np.array(hyp["n_inl"], np.int32), hyp_pose_lm=np.array(hyp["pose_lm"], np.float32), hyp_pose_refit=np.array(hyp["pose_refit"], np.float32),
           hyp_err=np.array(hyp["err"], np.float32), hyp_mask=np.concatenate(hyp["mask"]).astype(np.uint8), hyp_seed0=np.int64(9000), hyp_H=np.int32(H))
# whole RANSAC per cluster with the seeded stand-in RNG
rs = [r.ransac(int(c["model"][k]), c["members"][c["offsets"][k]:c["offsets"][k + 1]], (600, 200, 4, 5, 6, 10.0), 500 + k) for k in range(len(c["model"]))]
out.update(ransac_found=np.array([a[0] for a in rs], np.int32), ransac_pose=np.array([a[1] for a in rs], np.float32), ransac_seed0=np.int64(500))
# objects: the RANSAC poses, a perturbed duplicate of each, and a bogus one -> FILTER
om = np.concatenate([c["model"], c["model"], [0]]).astype(np.int32)
op = np.concatenate([out["ransac_pose"], out["ransac_pose"] + np.array([0, 0, 0, 0, 0.004, -0.003, 0.01], np.float32), [[0, 0, 0, 1, 0, 0, 2.0]]]).astype(np.float32)
r.set_objects(om, op)
r.run_filter((5, 4096.0, 2.0))
o2, c2 = r.get_objects(), r.get_clusters()
out.update(filter_in_model=om, filter_in_pose=op, filter_out_model=o2["model"], filter_out_pose=o2["pose"], filter_out_score=o2["score"],
           filter_cluster_model=c2["model"], filter_cluster_offsets=c2["offsets"], filter_cluster_members=c2["members"])
# project()
pp = out["ransac_pose"][0]
out["project_uv"] = r.project(pp, m["xyz"][:64], m["image"][:64]); out["project_pose"] = pp
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "moped_golden.npz"), **out)
print("clusters", len(c["model"]), "matches", m["offsets"][-1], "hyps", len(hyp["cluster"]), "accepted", int((out["hyp_n_inl"] > 6).sum()),
      "ransac found", out["ransac_found"], "filter survivors", len(o2["model"]), "ann eps5 != eps0:", int((ann5_idx[:, 0] != ann_idx[:, 0]).sum()))
