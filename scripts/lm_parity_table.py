#!/usr/bin/env python
"""Per-hypothesis LM / RANSAC parity table (task 1d): identical exported (sample set, initial quaternion) pairs through

  S  the STRICT-IEEE build of the reference's own POSE stage   (oracle/_ref/libmoped_ref_strict.so; the C restatement
     oracle/moped_oracle.c is bit-identical to it, tests/test_oracle_vs_strict_ref.py — the oracle is used when the strict .so is absent)
  F  the reference built with ITS OWN flags (-ffast-math)        (oracle/_ref/libmoped_ref.so)
  D  libmoped_cuda's default pose kernels (butterfly sums, FMA)  (mc_pose_hypotheses)
  X  libmoped_cuda's exact-order mode                            (mc_set_option pose_exact_order=1)

and prints, for the pairs D-S, D-F, F-S (the reference's own spread between its two builds) and X-S, per ACCEPTED hypothesis the
translation / rotation difference of the refitted pose and whether the inlier sets are identical. north_star's gate is 1e-4 m and
1e-3 rad per accepted hypothesis with identical inlier sets except at threshold ties.

    python scripts/lm_parity_table.py [--clusters 16] [--hyp 32] [--out gpurun_out/lm_parity]     (needs a B200)
Writes <out>.md (summary) and <out>.csv (one row per hypothesis).
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from moped_b200 import capi, synth  # noqa: E402
from oracle import oracle, ref  # noqa: E402

P = (600, 200, 1, 5, 6, 10.0)        # POSE defaults (moped2/libmoped/src/config.hpp:110)


def quat_angle(q1, q2):
    """rotation angle between two unit quaternions (x, y, z, w): 2 asin |vec(q1^-1 q2)|, in float64 (well conditioned near zero,
    where 2 acos |q1.q2| is quantised to ~7e-4 rad by fp32 inputs)"""
    a = np.asarray(q1, np.float64); b = np.asarray(q2, np.float64)
    a = a / np.linalg.norm(a); b = b / np.linalg.norm(b)
    if np.dot(a, b) < 0:
        b = -b
    v = a[3] * b[:3] - b[3] * a[:3] - np.cross(a[:3], b[:3])
    return 2.0 * np.arcsin(min(1.0, float(np.linalg.norm(v))))


def run_ref(cl, hy, strict):
    ref.use_strict(strict)
    try:
        r = ref.Ref(1)
        n_pts = np.diff(cl["offsets"]).astype(np.int32)
        r.set_models(n_pts, cl["xyz"], np.zeros((len(cl["xyz"]), 128), np.float32) + np.float32(0.1))
        r.set_images(synth.K_DEFAULT, synth.CAM_IDENTITY)
        r.set_matches(dict(offsets=cl["offsets"], image=cl["image"], xy=cl["xy"], xyz=cl["xyz"]))
        out = []
        for h in range(len(hy["hyp_cluster"])):
            c = int(hy["hyp_cluster"][h])
            members = np.arange(cl["offsets"][c + 1] - cl["offsets"][c], dtype=np.int32)
            n, plm, prf, err, mask = r.hypothesis(c, members, hy["sample_pos"][h], hy["init_quat"][h], P[1], P[5], P[4])
            out.append((n, prf.copy(), mask.astype(bool)))
        r.close()
        return out
    finally:
        ref.use_strict(False)


def run_oracle(cl, hy):
    cams = oracle.cameras(synth.K_DEFAULT, synth.CAM_IDENTITY)
    out = []
    for h in range(len(hy["hyp_cluster"])):
        c = int(hy["hyp_cluster"][h]); s = slice(cl["offsets"][c], cl["offsets"][c + 1])
        n, plm, prf, err, mask = oracle.hypothesis(cl["xy"][s], cl["xyz"][s], cl["image"][s], cams, hy["sample_pos"][h], hy["init_quat"][h], P[1], P[5], P[4])
        out.append((n, prf.copy(), mask.astype(bool)))
    return out


def run_gpu(ctx, cl, hy, exact):
    ctx.set_option("pose_exact_order", 1 if exact else 0)
    try:
        n_in, plm, prf, err, masks = ctx.pose_hypotheses(cl["offsets"], cl["xy"], cl["xyz"], cl["image"], hy["hyp_cluster"], hy["sample_pos"],
                                                         hy["init_quat"], P, want_mask=True)
    finally:
        ctx.set_option("pose_exact_order", 0)
    return [(int(n_in[h]), prf[h].copy(), np.asarray(masks[h]).astype(bool)) for h in range(len(n_in))]


def compare(a, b):
    """per hypothesis: (accepted_a, accepted_b, |dt|, drot, same_mask) — differences only where both accept"""
    rows = []
    for (na, pa, ma), (nb, pb, mb) in zip(a, b):
        aa, ab = na > P[4], nb > P[4]
        if aa and ab:
            rows.append((aa, ab, float(np.abs(pa[4:] - pb[4:]).max()), quat_angle(pa[:4], pb[:4]), bool(np.array_equal(ma, mb))))
        else:
            rows.append((aa, ab, np.nan, np.nan, False))
    return rows


def summarise(name, rows):
    acc_a = np.array([r[0] for r in rows]); acc_b = np.array([r[1] for r in rows])
    both = acc_a & acc_b
    dt = np.array([r[2] for r in rows])[both]; dr = np.array([r[3] for r in rows])[both]
    same = np.array([r[4] for r in rows])[both]
    pct = lambda v, q: float(np.percentile(v, q)) if len(v) else float("nan")
    return (f"| {name} | {len(rows)} | {int((acc_a == acc_b).sum())} | {int(both.sum())} | {pct(dt, 50):.2e} | {pct(dt, 90):.2e} | {pct(dt, 99):.2e} | "
            f"{(dt.max() if len(dt) else float('nan')):.2e} | {int((dt <= 1e-4).sum())} | {pct(dr, 50):.2e} | {pct(dr, 90):.2e} | {pct(dr, 99):.2e} | "
            f"{(dr.max() if len(dr) else float('nan')):.2e} | {int((dr <= 1e-3).sum())} | {int(same.sum())} |")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clusters", type=int, default=16)
    ap.add_argument("--hyp", type=int, default=32)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "lm_parity"))
    args = ap.parse_args()
    cl = synth.make_ransac_clusters(args.clusters, 80, 0.5, seed=777)
    hy = synth.make_hypotheses(cl, args.hyp, 5, seed=778)
    # half of the hypotheses from inlier-only samples, so that enough of them are accepted
    rng = np.random.default_rng(9)
    cams = oracle.cameras(synth.K_DEFAULT, synth.CAM_IDENTITY)
    for h in range(0, len(hy["hyp_cluster"]), 2):
        c = int(hy["hyp_cluster"][h]); s = slice(cl["offsets"][c], cl["offsets"][c + 1])
        uv = oracle.project(cl["gt_pose"][c], cl["xyz"][s], cl["image"][s], cams)
        good = np.nonzero(((uv - cl["xy"][s]) ** 2).sum(1) < 4.0)[0]
        if len(good) >= 5:
            hy["sample_pos"][h] = rng.choice(good, 5, replace=False)
    have_strict = ref.available() and ref.strict_available()
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    S = run_ref(cl, hy, True) if have_strict else run_oracle(cl, hy)
    O = run_oracle(cl, hy)
    F = run_ref(cl, hy, False) if ref.available() else None
    ctx = capi.Context(0)
    ctx.set_cameras(synth.K_DEFAULT, synth.CAM_IDENTITY)
    D = run_gpu(ctx, cl, hy, False)
    X = run_gpu(ctx, cl, hy, True)
    ctx.close()
    lines = ["| pair | hypotheses | same accept decision | accepted by both | dt p50 [m] | dt p90 | dt p99 | dt max | dt <= 1e-4 m | drot p50 [rad] | drot p90 | drot p99 | drot max | drot <= 1e-3 | identical inlier sets |",
             "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
    pairs = [("X (CUDA exact-order) vs S (strict reference)", X, S), ("oracle (C restatement) vs S", O, S), ("D (CUDA default) vs S", D, S)]
    if F is not None:
        pairs += [("D (CUDA default) vs F (reference, -ffast-math)", D, F), ("F vs S (the reference against itself)", F, S)]
    table = {}
    for name, a, b in pairs:
        table[name] = compare(a, b)
        lines.append(summarise(name, table[name]))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out + ".md", "w") as f:
        f.write(f"# LM parity per hypothesis — {args.clusters} clusters x {args.hyp} explicit hypotheses (80 points, 50 % outliers, POSE defaults {P})\n\n")
        f.write("S = " + ("strict-IEEE build of the reference" if have_strict else "C restatement (strict .so absent)") + "; dt = max |delta translation| of the refitted pose, "
                "drot = rotation angle between the refitted quaternions; differences over hypotheses accepted (> MinNPtsObject inliers) by both sides.\n\n")
        f.write("\n".join(lines) + "\n")
    with open(args.out + ".csv", "w") as f:
        f.write("hyp,cluster," + ",".join(f"{k}_{c}" for k in ("XS", "OS", "DS", "DF", "FS")[:len(pairs)] for c in ("accA", "accB", "dt", "drot", "same_mask")) + "\n")
        for h in range(len(S)):
            f.write(f"{h},{int(hy['hyp_cluster'][h])}," + ",".join(f"{int(r[h][0])},{int(r[h][1])},{r[h][2]:.3e},{r[h][3]:.3e},{int(r[h][4])}" for r in table.values()) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
