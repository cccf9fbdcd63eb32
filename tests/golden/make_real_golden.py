"""Generates tests/golden/real_images.npz: the REAL-IMAGE fixture of the hot path (BASELINE.json
configs[0] substitute). Run in the build container only (needs /root/reference and cv2 for JPEG decoding):

    make -f oracle/Makefile ref && python tests/golden/make_real_golden.py

The reference ships no model set (moped2/download_models.sh fetches it) and no expected results, so configs[0] cannot
be run literally. What it does ship is imagery: five 640x480 grey frames inside moped2/test_data/timing.bag and eleven
JPEGs under moped-example/test/. This script
  1. decodes them (the bag's CompressedImage payloads are plain JPEGs),
  2. extracts SIFT features with the reference's OWN step 1 (FEAT_SIFT_CPU over the vendored libsiftfast, ScaleOrigin
     "-1" like config.hpp:72) — real descriptors, real keypoint coordinates,
  3. builds planar models from some of the images exactly as Moped::createPlanarModelsFromImages does
     (moped.cpp:196-236: p3d = (u*scale, v*scale, 0), one point per feature), writes them as .moped.xml files in the
     modelling tools' format and loads them back through the reference's sXML reader (so the fixture's database is what
     libmoped would hold after Moped::addModel),
  4. runs the reference's CPU stages (exact matching, Quality=0) on the remaining frames and stores inputs + outputs.
tests/test_real_images.py pins the C restatement (CPU) and the CUDA path (-m gpu) to it."""
import os
import sys
import tempfile

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from moped_b200 import model_xml      # noqa: E402
from oracle import ref                # noqa: E402

REF = "/root/reference"
SCALE = np.float32(0.001)             # metres per pixel of the planar models
K = np.array([[811.4, 811.5, 307.4, 248.9]], np.float32)       # moped2/startmoped.launch intrinsics
CAM = np.array([[0, 0, 0, 1, 0, 0, 0]], np.float32)


def bag_frames(path):
    b = open(path, "rb").read()
    out, i = [], 0
    while True:
        s = b.find(b"\xff\xd8\xff", i)
        if s < 0:
            return out
        e = b.find(b"\xff\xd9", s)
        out.append(cv2.imdecode(np.frombuffer(b[s:e + 2], np.uint8), cv2.IMREAD_GRAYSCALE))
        i = e + 2


bag = bag_frames(os.path.join(REF, "moped2/test_data/timing.bag"))
assert len(bag) == 5 and bag[0].shape == (480, 640)
ex = [cv2.resize(cv2.imread(os.path.join(REF, "moped-example/test/image_%04d.jpg" % k), cv2.IMREAD_GRAYSCALE), (640, 360), interpolation=cv2.INTER_AREA)
      for k in (1, 2, 6)]

feats = {}
for name, im, dbl in [("bag0", bag[0], True), ("bag1", bag[1], True), ("bag2", bag[2], True), ("bag3", bag[3], True), ("bag4", bag[4], True),
                      ("ex1", ex[0], False), ("ex2", ex[1], False), ("ex6", ex[2], False)]:
    feats[name] = ref.sift(im, dbl)
    print(name, im.shape, len(feats[name][0]), "features")

# planar models: the whole of bag frame 0, the bottle region of bag frame 2, two example images (distractor objects)
models = []
xy, d = feats["bag0"]
models.append(("scene_bag0", xy, d))
xy, d = feats["bag2"]
sel = (xy[:, 0] > 270) & (xy[:, 0] < 390) & (xy[:, 1] > 50) & (xy[:, 1] < 290)
models.append(("bottle_bag2", xy[sel], d[sel]))
for nm in ("ex6",):
    models.append(("planar_" + nm, feats[nm][0], feats[nm][1]))

tmp = tempfile.mkdtemp()
r = ref.Ref(1)
for nm, xy, d in models:
    p = os.path.join(tmp, nm + ".moped.xml")
    xyz = np.concatenate([xy * SCALE, np.zeros((len(xy), 1), np.float32)], axis=1).astype(np.float32)
    model_xml.write_model_xml(p, nm, xyz, d, exact=True)
    assert r.add_model_xml(p) == 1
names = r.model_names()
n_pts, db_xyz, db_desc = [], [], []
for i in range(len(names)):
    x, ln, v = r.model_points(i, "SIFT")
    assert (ln == 128).all()
    n_pts.append(len(x)); db_xyz.append(x); db_desc.append(v.reshape(-1, 128))
n_pts = np.array(n_pts, np.int32); db_xyz = np.concatenate(db_xyz); db_desc = np.concatenate(db_desc)
model_of_row = np.repeat(np.arange(len(names), dtype=np.int32), n_pts)
print("models:", names, n_pts.tolist())

frames = ["bag1", "bag4", "ex2", "bag0"]                  # bag0 = a model's own image (exact planar pose)
out = dict(model_names=np.array(names), n_pts=n_pts, db_xyz=db_xyz, model_of_row=model_of_row, K=K, cam_pose=CAM,
           frame_names=np.array(frames), scale=SCALE)
r.set_images(K, CAM)
fo = [0]
qx, qd = [], []
for f, nm in enumerate(frames):
    xy, d = feats[nm]
    img = np.zeros(len(xy), np.int32)
    r.set_features(d, xy, img)
    r.clear_frame(); r.run_match(0.0, 0.8)
    if f == 0:
        out["db_desc"] = r.model_desc()                      # as normalised by the reference's MATCH
    qn = r.features_desc()
    idx, dist = r.ann_search(qn, 0.0)
    idx5, _ = r.ann_search(qn, 5.0)
    m = r.get_matches()
    r.run_cluster(200.0, 20.0, 7, 100)
    c = r.get_clusters()
    r.clear_frame()
    r.set_features(d, xy, img)
    n_obj, _ = r.run_pipeline(quality=0.0, seed=1)
    o = r.get_objects()
    om, op, osc = o["model"], o["pose"], o["score"]
    # RANSAC draws differ between runs (and between the reference's rand() and the CUDA stream): record how far the
    # reference's own poses move over other seeds, per object, as the yardstick of the pose comparison
    spread_t, spread_r = np.zeros(len(om), np.float32), np.zeros(len(om), np.float32)
    for seed in range(2, 10):
        r.clear_frame(); r.set_features(d, xy, img)
        r.run_pipeline(quality=0.0, seed=seed)
        o2 = r.get_objects()
        assert sorted(o2["model"].tolist()) == sorted(om.tolist()), (nm, seed)
        for k in range(len(om)):
            j = list(o2["model"]).index(om[k])
            qa, qb = op[k][:4] / np.linalg.norm(op[k][:4]), o2["pose"][j][:4] / np.linalg.norm(o2["pose"][j][:4])
            spread_t[k] = max(spread_t[k], np.abs(op[k][4:] - o2["pose"][j][4:]).max())
            spread_r[k] = max(spread_r[k], 2 * np.arccos(min(1.0, abs(float(np.dot(qa, qb))))))
    print("  pose spread over 8 other seeds: t", spread_t, "r", spread_r)
    print(nm, "features", len(xy), "matches", int(m["offsets"][-1]), "clusters", len(c["model"]), "objects", [(names[k], np.round(p, 3).tolist()) for k, p in zip(om, op)])
    out.update({f"f{f}_q_desc": qn, f"f{f}_ann_idx": idx, f"f{f}_ann_dist": dist, f"f{f}_ann5_idx": idx5,
                f"f{f}_match_offsets": m["offsets"], f"f{f}_match_xy": m["xy"], f"f{f}_match_xyz": m["xyz"], f"f{f}_match_image": m["image"],
                f"f{f}_cluster_model": c["model"], f"f{f}_cluster_offsets": c["offsets"], f"f{f}_cluster_members": c["members"],
                f"f{f}_obj_model": om, f"f{f}_obj_pose": op, f"f{f}_obj_score": osc,
                f"f{f}_obj_spread_t": spread_t, f"f{f}_obj_spread_r": spread_r})
    qx.append(xy); qd.append(d); fo.append(fo[-1] + len(xy))
# descriptors are stored as the reference's MATCH normalised them (f*_q_desc, db_desc): the fixture pins MATCH..FILTER2
out.update(q_xy=np.concatenate(qx), frame_offsets=np.array(fo, np.int32))
path = os.path.join(ROOT, "tests", "golden", "real_images.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path) / 1e6, "MB")
