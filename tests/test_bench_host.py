"""Host-side logic of bench.py that the driver's runs depend on (no GPU): the schedule bench.py picks from the frames per GPU, the
sub-result runner of the default line (child JSON lines condensed, a failing child reported without losing the main line), and the
file-descriptor redirection that keeps NCCL's banner off stdout."""
import argparse
import json
import os
import subprocess
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _args(**kw):
    d = dict(frames=64, pipeline=-1, stage_sms=-1, lanes=0, pose_warps=0, objects=1000, pts=1000, features=2000, pose_mode="default")
    d.update(kw)
    return argparse.Namespace(**d)


def test_schedule_by_frames_per_gpu():
    a = bench.resolve_auto(_args(), 1)                      # 64 frames on one GPU: one call per step, a lane per frame
    assert (a.pipeline, a.stage_sms, a.lanes, a.pose_warps) == (0, 0, 64, 4)
    a = bench.resolve_auto(_args(), 2)
    assert (a.pipeline, a.lanes) == (0, 32)
    a = bench.resolve_auto(_args(), 4)                      # <= 16 frames per GPU: MATCH of step i+1 beside the stages of step i
    assert (a.pipeline, a.stage_sms, a.lanes) == (1, 16, 16)
    a = bench.resolve_auto(_args(), 8)
    assert (a.pipeline, a.stage_sms, a.lanes) == (1, 16, 16)
    a = bench.resolve_auto(_args(frames=1), 1)              # configs[1]: a single frame
    assert (a.pipeline, a.stage_sms, a.lanes) == (1, 16, 16)
    a = bench.resolve_auto(_args(lanes=8, pose_warps=2, pipeline=0, stage_sms=0), 1)      # explicit flags win
    assert (a.pipeline, a.lanes, a.pose_warps) == (0, 8, 2)


def test_both_arms_describe_the_same_configuration():
    a, b = bench.resolve_auto(_args(coarse_kind=1, reserve_sms=0, batch_graph=1, merge_levels=1, chunks=1), 1), \
        bench.resolve_auto(_args(coarse_kind=1, reserve_sms=0, batch_graph=1, merge_levels=1, chunks=1), 1)
    assert bench.workload_config(a, 1) == bench.workload_config(b, 1)
    c = bench.workload_config(a, 1)
    assert c["db_descriptors"] == 1000000 and c["features_per_frame"] == 2000 and c["frames_per_step"] == 64
    assert bench.default_metric_config(a) and not bench.default_metric_config(_args(features=4000))


def test_other_config_lines_condense_children_and_survive_failures(monkeypatch):
    calls = []

    def fake_run(cmd, **kw):
        calls.append(cmd)
        if "sift" in cmd:                                                         # one child fails
            return types.SimpleNamespace(returncode=3, stdout="", stderr="boom\nlast line")
        line = {"metric": "frames_per_s", "value": 12.5, "unit": "frames/s", "ms_per_step": 80.0, "steps": 5, "warmup": 3, "gpu_launches": 7, "dtype": "f32",
                "e2e": {"value": 11.0}, "clocks": {"sm_mhz": 1965.0, "reasons": []}, "single_frame": {"latency_ms": 1.5},
                "roofline": {"kernel": "k", "bound": "tensor", "achieved": 1.0, "peak": 2.0, "unit": "TFLOP/s", "frac": 0.5}}
        return types.SimpleNamespace(returncode=0, stdout="NCCL version 2.28\n" + json.dumps(line) + "\n", stderr="")

    monkeypatch.setattr(subprocess, "run", fake_run)
    res = bench.other_config_lines()
    assert len(res) == 6 and len(calls) == 6
    assert all("--other-configs" in c and c[c.index("--other-configs") + 1] == "0" for c in calls)      # children never recurse
    ok = [v for v in res.values() if "error" not in v]
    bad = [v for v in res.values() if "error" in v]
    assert len(ok) == 5 and len(bad) == 1 and "rc 3" in bad[0]["error"]
    assert all(v["value"] == 12.5 and v["e2e"] == 11.0 and v["roofline"]["frac"] == 0.5 and v["single_frame_latency_ms"] == 1.5 for v in ok)
    json.dumps(res)                                                               # must fit into the main JSON line


def test_stdout_is_restored_after_the_quiet_nccl_init(capfd):
    class FakeDist:
        def init_process_group(self, *a, **k):
            os.write(1, b"NCCL version banner\n")                                 # what NCCL does: straight to file descriptor 1

        def all_reduce(self, t):
            pass

    fake_torch = types.SimpleNamespace(zeros=lambda *a, **k: 0, cuda=types.SimpleNamespace(synchronize=lambda *a, **k: None))
    saved = sys.modules.get("torch")
    sys.modules["torch"] = fake_torch
    try:
        bench.init_nccl_quietly(FakeDist(), None)
    finally:
        if saved is not None:
            sys.modules["torch"] = saved
        else:
            del sys.modules["torch"]
    os.write(1, b"{\"json\": 1}\n")
    out, err = capfd.readouterr()
    assert "banner" not in out and "banner" in err and out.strip() == '{"json": 1}'
