"""Sweep of the scheduling knobs of the batch path on a B200 (frames/s of mc_process_frames_dev at 1 M descriptors)."""
import os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from moped_b200 import capi, synth

def norm(x):
    n = np.sqrt((x * x).sum(axis=1, dtype=np.float32)).astype(np.float32)
    return (x * (np.float32(1.0) / n)[:, None]).astype(np.float32)

n_obj = int(os.environ.get("OBJ", 1000))
db = synth.make_db(n_obj, 1000)
dbn = norm(db["desc"])
ctx = capi.Context(0)
ctx.db_upload(dbn, db["xyz"], db["model_of_row"], n_obj)
ctx.set_cameras(synth.K_DEFAULT, synth.CAM_IDENTITY)
if os.environ.get("FUSED"):
    ctx.set_option("ransac_fused", 1)
dev = torch.device("cuda", 0)
NF = int(os.environ.get("NF", 64))
frames = [synth.make_frame(db, 2000, n_visible=8, frame_id=i) for i in range(NF)]
q = torch.from_numpy(np.concatenate([norm(f["desc"]) for f in frames])).to(dev)
xy = torch.from_numpy(np.concatenate([f["xy"] for f in frames])).to(dev)
img = torch.from_numpy(np.concatenate([f["image_idx"] for f in frames])).to(dev)
for B in [int(x) for x in os.environ.get("BATCHES", "64").split(",")]:
    fo = (np.arange(B + 1) * 2000).astype(np.int32)
    cfgs = [tuple(int(v) for v in c.split(":")) for c in os.environ.get("CONFIGS", "32:1:1,32:2:1,32:4:1,64:1:1,32:1:2,32:1:4,16:1:1").split(",")]
    for lanes, warps, chunks in cfgs:
        if lanes > B and lanes != 1:
            continue
        ctx.set_tuning(lanes, warps, chunks)
        ms = np.zeros(2, np.float32)
        for _ in range(2):
            out = ctx.process_frames_dev(q.data_ptr(), xy.data_ptr(), img.data_ptr(), fo, times=ms)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 5
        for _ in range(n):
            out = ctx.process_frames_dev(q.data_ptr(), xy.data_ptr(), img.data_ptr(), fo)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n
        nobj = sum(len(o["model"]) for o in out)
        print(f"B={B:3d} lanes={lanes:2d} warps={warps} chunks={chunks} : {dt*1e3:8.3f} ms/batch  {B/dt:8.1f} frames/s  match {ms[0]:.3f} ms rest {ms[1]:.3f} ms  objects {nobj}", flush=True)
