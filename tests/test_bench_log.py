"""moped_b200/bench_log.py (writer of moped3d's MopedBench text log) against MopedBench itself — MopedBench.cpp and
Benchmark.cpp compiled as they are into oracle/_ref/libmoped3d_ref*.so and driven over the same frame state. Every line must be
identical to the strict-IEEE build's except the numbers after TIME: (wall-clock of the reference run); against the build with the
reference's own -ffast-math flags numeric fields may differ in the sixth significant digit."""
import re

import numpy as np
import pytest

from moped_b200.bench_log import MopedBenchLog
from oracle import ref3d

pytestmark = pytest.mark.skipif(not ref3d.available(), reason="oracle/_ref/libmoped3d_ref*.so not built (needs /root/reference at build time)")


def make_frame(seed, n_models=3):
    rng = np.random.default_rng(seed)
    K = np.array([525.0, 525.0, 319.5, 239.5], np.float32)
    cam = np.array([0.02, -0.01, 0.03, 0.999, 0.01, -0.02, 0.03], np.float32)
    cam[:4] /= np.linalg.norm(cam[:4])
    model_xyz = [rng.uniform(-0.1, 0.1, (int(rng.integers(20, 60)), 3)).astype(np.float32) for _ in range(n_models)]
    n_matches = rng.integers(0, 25, n_models)
    n_matches[1] = 0                                           # a model without matches
    M = int(n_matches.sum())
    fr = dict(n_matches=n_matches, match_xy=rng.uniform(0, 640, (M, 2)).astype(np.float32), match_world=rng.normal(0, 1, (M, 3)).astype(np.float32),
              match_depth=rng.uniform(0.4, 3, M).astype(np.float32), match_fill=np.where(rng.random(M) < 0.5, 0, rng.uniform(0, 30, M)).astype(np.float32),
              match_valid=(rng.random(M) < 0.8), match_image=np.zeros(M, np.int32), model_xyz=model_xyz,
              model_names=[f"obj{m}" for m in range(n_models)], K=K, cam_pose=cam)
    cm, co, mem = [], [0], []
    for m in range(n_models):
        left = list(rng.permutation(int(n_matches[m])))
        while len(left) >= 4 and rng.random() < 0.8:
            k = int(rng.integers(3, len(left) + 1))
            mem += left[:k]; left = left[k:]
            cm.append(m); co.append(len(mem))
    fr.update(cluster_model=np.array(cm, np.int32), cluster_offsets=np.array(co, np.int32), cluster_members=np.array(mem, np.int32))
    n_obj = 3
    q = rng.normal(size=(n_obj, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    t = np.stack([rng.uniform(-0.3, 0.3, n_obj), rng.uniform(-0.2, 0.2, n_obj), rng.uniform(0.6, 1.5, n_obj)], 1)
    fr.update(obj_model=np.array([0, 2, 0], np.int32), obj_pose=np.concatenate([q, t], 1).astype(np.float32), obj_score=rng.uniform(2, 40, n_obj).astype(np.float32))
    return fr


def ours(fr):
    log = MopedBenchLog()
    for step in ("CLUSTER", "POSE", "FILTER", "FILTER2"):
        log.step(step, fr, 0.0)
    log.all_done(fr)
    return log.text()


def strip_times(text):
    return [re.sub(r"^(TIME:[A-Z0-9]+:).*$", r"\1", ln) for ln in text.splitlines()]


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_identical_to_the_strict_build(seed, tmp_path):
    fr = make_frame(seed)
    ref3d.use_strict(True)
    try:
        ref_text = ref3d.bench_log(str(tmp_path), fr)
    finally:
        ref3d.use_strict(False)
    a, b = strip_times(ours(fr)), strip_times(ref_text)
    assert len(a) == len(b) and len(a) > 20
    for x, y in zip(a, b):
        assert x == y
    kinds = {ln.split(":")[0] + ":" + ln.split(":")[1].split(" ")[0] for ln in b}
    assert {"PRE:CLUSTER", "POST:CLUSTER", "PRE:POSE", "POST:POSE", "PRE:FILTER", "POST:FILTER", "POST:FILTER2", "TIME:CLUSTER", "OBJ: obj0"} <= kinds | {b[-1][:9]}


def test_same_lines_as_the_fast_math_build(tmp_path):
    fr = make_frame(7)
    a, b = strip_times(ours(fr)), strip_times(ref3d.bench_log(str(tmp_path), fr))
    assert len(a) == len(b)
    num = re.compile(r"-?\d+\.?\d*(?:e[-+]?\d+)?")
    for x, y in zip(a, b):
        assert num.sub("#", x) == num.sub("#", y)                     # same structure, same hull vertex count and order
        for u, v in zip(num.findall(x), num.findall(y)):
            assert abs(float(u) - float(v)) <= 2e-5 * max(1.0, abs(float(v)))
