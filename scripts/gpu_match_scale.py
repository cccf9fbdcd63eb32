"""MATCH at growing query counts (frame batches): coarse kernel time, certificate statistics, total."""
import os, sys, time
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
ctx.set_profiling(True)
dev = torch.device("cuda", 0)
NF = int(os.environ.get("NF", 32))
frames = [synth.make_frame(db, 2000, n_visible=8, frame_id=i) for i in range(NF)]
qh = np.concatenate([norm(f["desc"]) for f in frames])
q = torch.from_numpy(qh).to(dev)
for B in (1, 2, 4, 8, 16, 32):
    if B > NF: break
    Q = 2000 * B
    rows, dist, acc, stats = ctx.match(qh[:Q])
    nn_row = torch.empty((Q, 2), dtype=torch.int32, device=dev)
    nn_dist = torch.empty((Q, 2), dtype=torch.float32, device=dev)
    a = torch.empty((Q,), dtype=torch.uint8, device=dev)
    for _ in range(2):
        ctx.match_dev(q.data_ptr(), Q, 0.8, 0, nn_row.data_ptr(), nn_dist.data_ptr(), a.data_ptr())
    ctx.synchronize()
    t0 = time.perf_counter()
    n = 5
    for _ in range(n):
        ctx.match_dev(q.data_ptr(), Q, 0.8, 0, nn_row.data_ptr(), nn_dist.data_ptr(), a.data_ptr())
    ctx.synchronize()
    dt = (time.perf_counter() - t0) / n * 1e3
    k = ctx.coarse_kernel_ms()
    tf = 2.0 * Q * len(dbn) * 128 / (k * 1e-3) / 1e12
    print(f"Q={Q:6d} total {dt:8.3f} ms  coarse {k:8.3f} ms ({tf:7.1f} TFLOP/s)  certified {stats[0]} fallback {stats[1]} n_cand {stats[2]} splits {stats[3]}  accepted {int(acc.sum())}", flush=True)
