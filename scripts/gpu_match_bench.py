"""Scratch: matching at scale. NOBJ objects x 1000 descriptors, Q queries; tensor mode vs exact mode on the GPU."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moped_b200 import synth, capi
n_obj = int(os.environ.get("NOBJ", "1000")); Q = int(os.environ.get("Q", "2000")); reps = int(os.environ.get("REPS", "10"))
db = synth.make_db(n_obj, 1000)
def nrm(x):
    n = np.sqrt((x * x).sum(axis=1, dtype=np.float32)).astype(np.float32); return (x / n[:, None]).astype(np.float32)
dbn = nrm(db["desc"])
frames = [synth.make_frame(db, Q, n_visible=8, frame_id=i) for i in range(4)]
qn = [nrm(f["desc"]) for f in frames]
ctx = capi.Context(0); ctx.db_upload(dbn, db["xyz"], db["model_of_row"], n_obj); ctx.set_profiling(True)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
dev = torch.device("cuda", 0)
d_q = [torch.from_numpy(q).to(dev) for q in qn]
nn_row = torch.empty((Q, 2), dtype=torch.int32, device=dev); nn_dist = torch.empty((Q, 2), dtype=torch.float32, device=dev); acc = torch.empty((Q,), dtype=torch.uint8, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
res = {}
for mode, name in ((capi.MATCH_EXACT, "exact"), (capi.MATCH_TENSOR, "tensor")):
    tms, kms = [], []
    for i in range(reps + 2):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream); ctx.match_dev(d_q[i % 4].data_ptr(), Q, 0.8, mode, nn_row.data_ptr(), nn_dist.data_ptr(), acc.data_ptr()); e1.record(stream)
        torch.cuda.synchronize()
        if i >= 2:
            tms.append(e0.elapsed_time(e1))
            if mode == capi.MATCH_TENSOR: kms.append(ctx.coarse_kernel_ms())
    print(f"[{name}] match ms mean {np.mean(tms):.3f} min {np.min(tms):.3f}" + (f"  coarse kernel ms mean {np.mean(kms):.3f} min {np.min(kms):.3f} -> {2*Q*len(dbn)*128/np.mean(kms)/1e9:.0f} TFLOP/s" if kms else ""), flush=True)
    for k in range(4):
        ctx.match_dev(d_q[k].data_ptr(), Q, 0.8, mode, nn_row.data_ptr(), nn_dist.data_ptr(), acc.data_ptr()); torch.cuda.synchronize()
        res[(name, k)] = (nn_row.cpu().numpy().copy(), nn_dist.cpu().numpy().copy(), acc.cpu().numpy().copy())
for k in range(4):
    a, b = res[("exact", k)], res[("tensor", k)]
    r, d, ac, st = ctx.match(qn[k], 0.8, capi.MATCH_TENSOR)
    print(f"frame {k}: rows equal {np.array_equal(a[0], b[0])} dist equal {np.array_equal(a[1], b[1])} acc equal {np.array_equal(a[2], b[2])} accepted {int(a[2].sum())} stats(cert,fallback,ncand,splits) {st}", flush=True)
