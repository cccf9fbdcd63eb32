import os, sys
os.environ["MOPED_CUDA_TRACE"] = "1"
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from moped_b200 import capi, synth
def norm(x):
    n = np.sqrt((x * x).sum(axis=1, dtype=np.float32)).astype(np.float32)
    return (x * (np.float32(1.0) / n)[:, None]).astype(np.float32)
n_obj = 1000
db = synth.make_db(n_obj, 1000)
ctx = capi.Context(0)
ctx.db_upload(norm(db["desc"]), db["xyz"], db["model_of_row"], n_obj)
ctx.set_cameras(synth.K_DEFAULT, synth.CAM_IDENTITY)
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
lanes, warps, chunks = [int(x) for x in (sys.argv[2:5] if len(sys.argv) > 4 else (16, 2, 1))]
frames = [synth.make_frame(db, 2000, n_visible=8, frame_id=i) for i in range(B)]
q = torch.from_numpy(np.concatenate([norm(f["desc"]) for f in frames])).to(dev)
xy = torch.from_numpy(np.concatenate([f["xy"] for f in frames])).to(dev)
img = torch.from_numpy(np.concatenate([f["image_idx"] for f in frames])).to(dev)
fo = (np.arange(B + 1) * 2000).astype(np.int32)
ctx.set_tuning(lanes, warps, chunks)
for i in range(3):
    print(f"--- run {i}", file=sys.stderr)
    out = ctx.process_frames_dev(q.data_ptr(), xy.data_ptr(), img.data_ptr(), fo)
print([o["info"].tolist() for o in out])
