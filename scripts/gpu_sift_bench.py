"""Times mc_sift_extract_dev on a B200: batches of 640x480 frames (the reference's shipped frame, shifted per slot)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from moped_b200 import capi

g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "sift_golden.npz"))
base = g["bag4_full_double/image"]
ctx = capi.Context(0)
st = torch.cuda.Stream()
ctx.set_stream(st.cuda_stream)
for B in [int(a) for a in (sys.argv[1:] or ["1", "8", "64"])]:
    frames = np.stack([np.roll(base, 5 * i, axis=1) for i in range(B)])
    max_kp = 4096
    with torch.cuda.stream(st):
        d_gray = torch.from_numpy(frames).cuda()
        d_xy = torch.zeros(B, max_kp, 2, device="cuda"); d_so = torch.zeros(B, max_kp, 2, device="cuda")
        d_desc = torch.zeros(B, max_kp, 128, device="cuda"); d_cnt = torch.zeros(B, dtype=torch.int32, device="cuda")
        for _ in range(3):
            ctx.sift_dev(d_gray.data_ptr(), B, 480, 640, True, max_kp, d_xy.data_ptr(), d_so.data_ptr(), d_desc.data_ptr(), d_cnt.data_ptr())
        st.synchronize()
        l0 = ctx.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        e0.record(st)
        for _ in range(n):
            ctx.sift_dev(d_gray.data_ptr(), B, 480, 640, True, max_kp, d_xy.data_ptr(), d_so.data_ptr(), d_desc.data_ptr(), d_cnt.data_ptr())
        e1.record(st)
        st.synchronize()
        ms = e0.elapsed_time(e1) / n
    print(json.dumps(dict(B=B, ms_per_batch=ms, frames_per_s=B / ms * 1e3, keypoints=int(d_cnt.sum()), launches_per_batch=(ctx.launches - l0) // n)))
