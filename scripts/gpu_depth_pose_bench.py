"""First timing of the depth-aware pose stages (pose_depth.cu) on one B200 next to the compiled reference on the host.

  explicit hypotheses: 64 clusters x HYP hypotheses (the shape of BASELINE configs[3]) through mc_pose_depth_hypotheses
                       (host buffers in, results out), variants 0 / 1 / 2 (2 = the moped2 residual in exact-order mode), and the
                       default (re-associating) moped2 kernels on the same clusters for comparison;
  RANSAC:              64 clusters x 4 tries through mc_pose_depth_ransac (moped3d POSE parameters 192, 100, 4, 5, 6, 8, 0.5);
  reference:           ref3d.hypothesis / ref3d.ransac (moped3d's own class, its -ffast-math flags) on ONE host core, bounded sample.

Prints one JSON object per line; scripts/gpu_next_round_first.sh stores them under gpurun_out/."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np

from moped_b200 import capi
from oracle import oracle, ref3d
from test_oracle3d_pose import ALPHA, CAM, K, make_cluster

HYP = int(os.environ.get("DEPTH_BENCH_HYP", "256"))
N_CLUSTERS = 64
LM, THR, MIN_NPTS = 100, 8.0, 6


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    return (time.perf_counter() - t0) / reps, out


def main():
    ctx = capi.Context(0)
    ctx.set_cameras(K[None], CAM[None])
    clusters = [make_cluster(1000 + i, n=80, outliers=0.5) for i in range(N_CLUSTERS)]
    off = np.concatenate([[0], np.cumsum([len(c["xy"]) for c in clusters])]).astype(np.int32)
    cat = lambda k: np.concatenate([c[k] for c in clusters]).astype(np.float32)
    xy, xyz, world = cat("xy"), cat("xyz"), cat("world")
    img = np.zeros(off[-1], np.int32)
    rng = np.random.default_rng(1)
    hyp_cluster = np.repeat(np.arange(N_CLUSTERS, dtype=np.int32), HYP)
    sample_pos = np.stack([rng.choice(80, 5, replace=False) for _ in range(N_CLUSTERS * HYP)]).astype(np.int32)
    init_quat = (rng.integers(0, 256, (N_CLUSTERS * HYP, 4)) / 256.0).astype(np.float32)
    P = (192, LM, 4, 5, MIN_NPTS, THR)
    for variant in (0, 1, 2):
        cw = np.concatenate([oracle.cauchy_weights(c["fill"], min(variant, 1)) for c in clusters])
        dt, out = timed(lambda: ctx.pose_depth_hypotheses(variant, off, xy, xyz, world, cw, img, hyp_cluster, sample_pos, init_quat, P, ALPHA,
                                                          want_mask=False))
        print(json.dumps(dict(what="explicit hypotheses, order-preserving LM", variant=variant, hypotheses=len(hyp_cluster), s_per_call=dt,
                              hypotheses_per_s=len(hyp_cluster) / dt, accepted=int((out[0] > MIN_NPTS).sum()), lm_failed=int((out[0] < 0).sum()))))
    ctx.set_option("depth_team_lanes", 8)
    for variant in (0, 2):
        cw = np.concatenate([oracle.cauchy_weights(c["fill"], min(variant, 1)) for c in clusters])
        dt, out = timed(lambda: ctx.pose_depth_hypotheses(variant, off, xy, xyz, world, cw, img, hyp_cluster, sample_pos, init_quat, P, ALPHA,
                                                          want_mask=False))
        print(json.dumps(dict(what="explicit hypotheses, order-preserving LM, teams of 8 lanes (four hypotheses per warp)", variant=variant,
                              hypotheses=len(hyp_cluster), s_per_call=dt, hypotheses_per_s=len(hyp_cluster) / dt)))
    ctx.set_option("depth_team_lanes", 32)
    dt, out = timed(lambda: ctx.pose_hypotheses(off, xy, xyz, img, hyp_cluster, sample_pos, init_quat, P, want_mask=False))
    print(json.dumps(dict(what="explicit hypotheses, default moped2 kernels (re-associating), same clusters", hypotheses=len(hyp_cluster),
                          s_per_call=dt, hypotheses_per_s=len(hyp_cluster) / dt, accepted=int((out[0] > MIN_NPTS).sum()))))
    for variant in (0, 1):
        cw = np.concatenate([oracle.cauchy_weights(c["fill"], variant) for c in clusters])
        dt, out = timed(lambda: ctx.pose_depth_ransac(variant, off, xy, xyz, world, cw, img, P, ALPHA, seed=3))
        print(json.dumps(dict(what="RANSAC, 64 clusters x 4 tries", variant=variant, s_per_call=dt, tasks_per_s=len(out[0]) / dt,
                              found=int(out[0].sum()), tests=int(out[2].sum()))))
    if ref3d.available():
        for variant in (0, 1):
            t0 = time.perf_counter(); n = 0
            while time.perf_counter() - t0 < 5.0:
                h = n % len(hyp_cluster)
                ref3d.hypothesis(clusters[hyp_cluster[h]], K, CAM, ALPHA, sample_pos[h], init_quat[h], LM, THR, MIN_NPTS, variant=variant)
                n += 1
            dt = time.perf_counter() - t0
            print(json.dumps(dict(what="reference (moped3d class, -ffast-math), explicit hypotheses, ONE host core", variant=variant,
                                  hypotheses=n, hypotheses_per_s=n / dt)))
            t0 = time.perf_counter(); n = 0
            for c in clusters[:16]:
                for t in range(4):
                    ref3d.ransac(c, K, CAM, ALPHA, (192, LM, 5, MIN_NPTS, THR), 100 + n, variant=variant); n += 1
            dt = time.perf_counter() - t0
            print(json.dumps(dict(what="reference RANSAC, 16 clusters x 4 tries, ONE host core", variant=variant, tasks_per_s=n / dt)))


if __name__ == "__main__":
    main()
