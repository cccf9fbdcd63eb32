"""Does the stage chain of one batch run beside the persistent matching kernel of the next one? (GPU box)
Times, on two streams of two contexts: MATCH alone, CLUSTER..FILTER2 alone, both enqueued together — per reserved-SM setting."""
import json
import os
import sys

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from moped_b200 import capi, synth  # noqa: E402

B, Q, OBJ = int(os.environ.get("PROBE_FRAMES", 64)), 2000, int(os.environ.get("PROBE_OBJECTS", 1000))
SB = int(os.environ.get("PROBE_STAGE_FRAMES", B))      # frames whose stages run here (a rank of an N-GPU job matches all B frames, runs B/N)


def nrm(x):
    n = np.sqrt((x * x).sum(1, dtype=np.float32))
    return (x / n[:, None]).astype(np.float32)


db = synth.make_db(OBJ, 1000)
dbn = nrm(db["desc"])
frames = [synth.make_frame(db, Q, n_visible=8, frame_id=i) for i in range(B)]
dev = torch.device("cuda", 0)
q = torch.from_numpy(np.concatenate([nrm(f["desc"]) for f in frames])).to(dev)
xy = torch.from_numpy(np.concatenate([f["xy"] for f in frames])).to(dev)
img = torch.from_numpy(np.concatenate([f["image_idx"] for f in frames])).to(dev)
QT = B * Q
fo = (np.arange(B + 1) * Q).astype(np.int32)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
ONE = bool(int(os.environ.get("PROBE_ONE_CONTEXT", "1")))      # MATCH and the stages on ONE context (the partition belongs to a context)
cm = capi.Context(0)
cs = cm if ONE else capi.Context(0)
cm.set_stream(s1.cuda_stream)
cm.db_upload(dbn, db["xyz"], db["model_of_row"], OBJ)
if not ONE:
    cs.set_stream(s2.cuda_stream)
    cs.db_upload(dbn[:256], db["xyz"][:256], db["model_of_row"][:256], OBJ)
    cs.db_set_global_tables(db["xyz"], db["model_of_row"], OBJ)
for c in (cm, cs):
    c.set_cameras(synth.K_DEFAULT, synth.CAM_IDENTITY)
cs.set_tuning(32, 4, 1)
if ONE:
    cm.set_option("defer_lane_join", 1)
params = cm.default_params()
MO = 64
row = torch.empty((QT, 2), dtype=torch.int32, device=dev)
dist = torch.empty((QT, 2), dtype=torch.float32, device=dev)
acc = torch.empty((QT,), dtype=torch.uint8, device=dev)
row2, dist2, acc2 = torch.empty_like(row), torch.empty_like(dist), torch.empty_like(acc)
o_info = torch.zeros((B, 4), dtype=torch.int32, device=dev)
o_model = torch.zeros((B, MO), dtype=torch.int32, device=dev)
o_pose = torch.zeros((B, MO, 7), dtype=torch.float32, device=dev)
o_score = torch.zeros((B, MO), dtype=torch.float32, device=dev)


def match(r, d, a):
    cm.match_dev(q.data_ptr(), QT, params.match_ratio, params.match_mode, r.data_ptr(), d.data_ptr(), a.data_ptr())


def stages():
    cs.process_frames_matched_dev(row.data_ptr(), acc.data_ptr(), xy.data_ptr(), img.data_ptr(), fo, 0, SB, params, MO,
                                  o_info.data_ptr(), o_model.data_ptr(), o_pose.data_ptr(), o_score.data_ptr())


def timed(fn_a, fn_b, n=5):
    """(ms until fn_a's work is done, ms until fn_b's work is done) with both enqueued at once, fn_b (the stages) first.
    One context: the lanes fork from s1 when the stages are enqueued (deferred join), MATCH follows on s1, the join comes last."""
    out = []
    for _ in range(n):
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); ea = torch.cuda.Event(enable_timing=True); eb = torch.cuda.Event(enable_timing=True)
        e0.record(s1)
        if ONE:
            with torch.cuda.stream(s1):
                if fn_b:
                    fn_b()
                if fn_a:
                    fn_a()
                ea.record(s1)
                cm.join_lanes()
                eb.record(s1)
        else:
            s2.wait_event(e0)
            if fn_b:
                with torch.cuda.stream(s2):
                    fn_b()
            if fn_a:
                with torch.cuda.stream(s1):
                    fn_a()
            ea.record(s1); eb.record(s2)
        torch.cuda.synchronize()
        out.append((e0.elapsed_time(ea), e0.elapsed_time(eb)))
    return np.median(np.array(out), axis=0).tolist()


with torch.cuda.stream(s1):
    match(row, dist, acc)
torch.cuda.synchronize()
ref_objects = None
for part, graphs in [(int(x.split(":")[0]), int(x.split(":")[1])) for x in os.environ.get("PROBE_CONFIGS", "0:1,16:1,16:0,24:1,32:1").split(",")]:
    try:
        cm.set_option("stage_sm_partition", part)
    except Exception as e:
        print(json.dumps({"stage_sm_partition": part, "error": str(e)}), flush=True)
        continue
    cm.set_option("frame_graphs", graphs)
    sms = cm.sm_partition()
    for _ in range(3):
        with torch.cuda.stream(s1):
            stages()
            if ONE:
                cm.join_lanes()
        torch.cuda.synchronize()
    objs = (o_info[:, 0].cpu().numpy().copy(), o_model.cpu().numpy().copy(), o_pose.cpu().numpy().copy())
    same = True if ref_objects is None else all(np.array_equal(a, b) for a, b in zip(ref_objects, objs))
    ref_objects = ref_objects or objs
    m = timed(lambda: match(row2, dist2, acc2), None)
    s = timed(None, stages)
    both = timed(lambda: match(row2, dist2, acc2), stages)
    print(json.dumps({"frames": B, "stage_frames": SB, "db_objects": OBJ, "stage_sm_partition": part, "frame_graphs": graphs, "sms": sms, "same_objects": same, "match_alone_ms": m[0], "stages_alone_ms": s[1], "together_match_ms": both[0],
                      "together_stages_ms": both[1], "objects": int(o_info[:, 0].sum().item())}), flush=True)
