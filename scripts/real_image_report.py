"""Per-stage times of the reference CPU stages (oracle/_ref, all host cores) and of libmoped_cuda on the real-image
fixture (tests/golden/real_images.npz: BASELINE.json configs[0] substitute). Run on the GPU box:
    python scripts/real_image_report.py > gpurun_out/real_images_r1.md"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from moped_b200 import capi
from oracle import ref

g = dict(np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "real_images.npz")))
n_models = len(g["n_pts"])
cores = os.cpu_count() or 1
ctx = capi.Context(0)
ctx.db_upload(g["db_desc"], g["db_xyz"], g["model_of_row"], n_models)
ctx.set_cameras(g["K"], g["cam_pose"])
r = ref.Ref(cores)
r.set_models(g["n_pts"], g["db_xyz"], g["db_desc"])
r.set_images(g["K"], g["cam_pose"])
r.build_match(5.0, 0.8)
names = ["match", "cluster", "pose", "filter", "pose2", "filter2"]
print(f"# Real-image fixture: reference CPU stages ({cores} host cores, shipped defaults: ANN eps=5) vs libmoped_cuda (1 x B200)\n")
print("Database: " + ", ".join(f"{n} ({k} pts)" for n, k in zip(g["model_names"], g["n_pts"])) + "; per frame ms, median of 5 runs after a warm-up.\n")
print("| frame | features | objects ref / cuda | " + " | ".join(f"{n} ref / cuda" for n in names) + " | total ref / cuda |")
print("|---|---:|---|" + "---:|" * 7)
fo = g["frame_offsets"]
for f in range(len(fo) - 1):
    q, xy = g[f"f{f}_q_desc"], g["q_xy"][fo[f]:fo[f + 1]]
    img = np.zeros(len(q), np.int32)
    rt, n_ref = [], 0
    for it in range(6):
        r.clear_frame(); r.set_features(q, xy, img)
        n_ref, t = r.run_pipeline(seed=1 + it)
        if it: rt.append(np.array(t) * 1e3)
    rt = np.median(np.stack(rt), axis=0)
    ct = []
    for it in range(6):
        out = ctx.process_frame(q, xy, img, max_objects=64, want_times=True)
        if it: ct.append(np.array(out["stage_ms"]))
    ct = np.median(np.stack(ct), axis=0)
    print(f"| {g['frame_names'][f]} | {len(q)} | {n_ref} / {len(out['model'])} | " + " | ".join(f"{a:.3f} / {b:.3f}" for a, b in zip(rt, ct)) + f" | {rt.sum():.3f} / {ct.sum():.3f} |")
print("\nThe database is tiny (1275 descriptors), so this is the regime where the CPU kd-tree is at its best and the GPU path is pure launch latency; "
      "it is a parity fixture (tests/test_real_images.py), not a throughput configuration.")
