#!/usr/bin/env python
"""Host experiment behind the LDL^T solve of the default LM kernels (pose.cu: solve7). TEST / EVIDENCE TOOLING, CPU only.

Builds two patched copies of the oracle (oracle/moped_oracle.c) in a temporary directory:
  probe   the unchanged LM with a probe after every solve of the damped normal equations: an unpivoted LDL^T factorisation of the same
          system -> histogram of the smallest pivot ratio d_j / a_jj, how often a pivot is non-positive, how far the LDL^T solution is
          from the pivoting LU's;
  ldl     the LM with the rule the CUDA kernels use: LDL^T when every pivot keeps more than 1e-3 of its diagonal entry, levmar's
          pivoting LU otherwise;
and runs explicit hypotheses of the bench's RANSAC clusters through the unchanged oracle and the `ldl` copy: accept decisions, inlier
counts, refitted poses (the comparison the GPU parity table makes between the CUDA kernels and the reference).

    python scripts/lm_ldl_experiment.py > profiles/lm_ldl_experiment_r2b.txt
"""
from __future__ import annotations

import ctypes as C
import importlib.util
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from moped_b200 import synth  # noqa: E402
from oracle import oracle as o_ref  # noqa: E402

SOLVE_CALL = "\t\tint solved = lu_solve(jtj, jte, Dp, m); ++nlss;"
LM_HEADER = "/* Levenberg-Marquardt with forward-difference Jacobian and Broyden rank-1 updates, target vector 0,"

PROBE = r'''
#include <stdio.h>
double lmx_stats[16];
long lmx_hist[12];
void lmx_probe(const float *A, const float *b, const float *x_lu, int solved, float mu, int k) {
	float L[7][7], d[7], y[7], x[7], minratio = 1e30f;
	int ok = 1;
	for (int j = 0; j < 7; j++) {
		float s = A[j * 7 + j];
		for (int k2 = 0; k2 < j; k2++) s -= L[j][k2] * L[j][k2] * d[k2];
		d[j] = s;
		float r = s / A[j * 7 + j];
		if (!(r > 0)) { ok = 0; r = 0; }
		if (r < minratio) minratio = r;
		if (!ok) break;
		float inv = 1.0f / s;
		for (int i = j + 1; i < 7; i++) { float t = A[i * 7 + j]; for (int k2 = 0; k2 < j; k2++) t -= L[i][k2] * L[j][k2] * d[k2]; L[i][j] = t * inv; }
	}
	lmx_stats[0] += 1;
	if (!solved) lmx_stats[1] += 1;
	if (!ok) { lmx_stats[2] += 1; lmx_hist[11]++; return; }
	int bin = (int)floor(-log10f(minratio));
	if (bin < 0) bin = 0;
	if (bin > 10) bin = 10;
	lmx_hist[bin]++;
	for (int i = 0; i < 7; i++) { float s = b[i]; for (int k2 = 0; k2 < i; k2++) s -= L[i][k2] * y[k2]; y[i] = s; }
	for (int i = 6; i >= 0; i--) { float s = y[i] / d[i]; for (int k2 = i + 1; k2 < 7; k2++) s -= L[k2][i] * x[k2]; x[i] = s; }
	double num = 0, den = 0;
	for (int i = 0; i < 7; i++) { num += (double)(x[i] - x_lu[i]) * (x[i] - x_lu[i]); den += (double)x_lu[i] * x_lu[i]; }
	double rel = sqrt(num / (den + 1e-300));
	if (minratio > 1e-3f) { lmx_stats[3] += 1; lmx_stats[4] += rel; } else { lmx_stats[6] += 1; lmx_stats[7] += rel; }
}
'''

LDL = r'''
long lmx_fast, lmx_slow;
static int ldl_or_lu(const float *A, const float *b, float *x, int m) {
	float L[7][7], d[7], y[7];
	int ok = 1;
	for (int j = 0; j < 7 && ok; j++) {
		float s = A[j * 7 + j];
		for (int k2 = 0; k2 < j; k2++) s -= L[j][k2] * L[j][k2] * d[k2];
		d[j] = s;
		if (!(s > 1e-3f * A[j * 7 + j])) { ok = 0; break; }
		float inv = 1.0f / s;
		for (int i = j + 1; i < 7; i++) { float t = A[i * 7 + j]; for (int k2 = 0; k2 < j; k2++) t -= L[i][k2] * L[j][k2] * d[k2]; L[i][j] = t * inv; }
	}
	if (!ok) { lmx_slow++; return lu_solve(A, b, x, m); }
	lmx_fast++;
	for (int i = 0; i < 7; i++) { float s = b[i]; for (int k2 = 0; k2 < i; k2++) s -= L[i][k2] * y[k2]; y[i] = s; }
	for (int i = 6; i >= 0; i--) { float s = y[i] / d[i]; for (int k2 = i + 1; k2 < 7; k2++) s -= L[k2][i] * x[k2]; x[i] = s; }
	return 1;
}

'''


def build(tmp, name, src):
    d = os.path.join(tmp, name)
    os.makedirs(d)
    open(os.path.join(d, "moped_oracle.c"), "w").write(src)
    shutil.copy(os.path.join(ROOT, "oracle", "oracle.py"), d)
    subprocess.check_call(["gcc", "-O2", "-march=native", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-w", "-I" + os.path.join(ROOT, "oracle"),
                           "-o", os.path.join(d, "libmoped_oracle.so"), os.path.join(d, "moped_oracle.c"),
                           os.path.join(ROOT, "oracle", "moped_sift_oracle.c"), os.path.join(ROOT, "oracle", "moped_linkage_oracle.c"), "-lm"])
    spec = importlib.util.spec_from_file_location("oracle_" + name, os.path.join(d, "oracle.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod, C.CDLL(os.path.join(d, "libmoped_oracle.so"))


def quat_angle(a, b):
    return 2 * np.arccos(min(1.0, abs(float(np.dot(a, b))) / (np.linalg.norm(a) * np.linalg.norm(b))))


def main():
    src = open(os.path.join(ROOT, "oracle", "moped_oracle.c")).read()
    assert SOLVE_CALL in src and LM_HEADER in src, "oracle/moped_oracle.c changed: adapt the patch points"
    probe_src = src.replace(SOLVE_CALL, SOLVE_CALL + "\n\t\t{ extern void lmx_probe(const float *, const float *, const float *, int, float, int); "
                            "lmx_probe(jtj, jte, Dp, solved, mu, k); }") + PROBE
    ldl_src = src.replace(SOLVE_CALL, "\t\tint solved = ldl_or_lu(jtj, jte, Dp, m); ++nlss;").replace(LM_HEADER, LDL + LM_HEADER)
    with tempfile.TemporaryDirectory() as tmp:
        o_probe, lib_probe = build(tmp, "probe", probe_src)
        o_ldl, lib_ldl = build(tmp, "ldl", ldl_src)
        stats = (C.c_double * 16).in_dll(lib_probe, "lmx_stats")
        hist = (C.c_long * 12).in_dll(lib_probe, "lmx_hist")
        cl = synth.make_ransac_clusters(8, 80, 0.5)
        hy = synth.make_hypotheses(cl, 64, 5)
        cams = o_probe.cameras(synth.K_DEFAULT, synth.CAM_IDENTITY)
        for h in range(len(hy["hyp_cluster"])):
            c = hy["hyp_cluster"][h]
            s = slice(cl["offsets"][c], cl["offsets"][c + 1])
            o_probe.hypothesis(cl["xy"][s], cl["xyz"][s], cl["image"][s], cams, hy["sample_pos"][h], hy["init_quat"][h], 200, 10.0, 6)
        n = stats[0]
        print(f"probe: {len(hy['hyp_cluster'])} hypotheses (8 clusters x 64, 80 points, 50 % outliers), {int(n)} solves, LU reported singular {int(stats[1])}")
        print("  smallest LDL^T pivot ratio d_j / a_jj, histogram over solves (bin = -log10, last = non-positive pivot):", list(hist))
        print(f"  pivots all above 1e-3: {int(stats[3])} solves ({100 * stats[3] / n:.2f} %), mean relative difference LDL^T vs LU solution {stats[4] / max(stats[3], 1):.2e}")
        print(f"  otherwise: {int(stats[6] + stats[2])} solves ({100 * (stats[6] + stats[2]) / n:.2f} %)")
        print()
        print("LM with the CUDA kernels' rule (LDL^T above 1e-3, pivoting LU otherwise) against the unchanged oracle, per explicit hypothesis:")
        tot_h = 0
        for (ncl, npts, outl, per, na, P) in [(16, 80, 0.5, 64, 5, (600, 200, 1, 5, 6, 10.0)), (8, 60, 0.3, 64, 6, (100, 500, 1, 6, 8, 5.0)),
                                              (8, 200, 0.4, 32, 5, (600, 200, 1, 5, 6, 10.0))]:
            cl = synth.make_ransac_clusters(ncl, npts, outl, seed=50 + na)
            hy = synth.make_hypotheses(cl, per, na, seed=7)
            cr, cq = o_ref.cameras(synth.K_DEFAULT, synth.CAM_IDENTITY), o_ldl.cameras(synth.K_DEFAULT, synth.CAM_IDENTITY)
            R, L = [], []
            for h in range(len(hy["hyp_cluster"])):
                c = hy["hyp_cluster"][h]
                s = slice(cl["offsets"][c], cl["offsets"][c + 1])
                a = (cl["xy"][s], cl["xyz"][s], cl["image"][s])
                R.append(o_ref.hypothesis(*a, cr, hy["sample_pos"][h], hy["init_quat"][h], P[1], P[5], P[4]))
                L.append(o_ldl.hypothesis(*a, cq, hy["sample_pos"][h], hy["init_quat"][h], P[1], P[5], P[4]))
            rin, lin = np.array([r[0] for r in R]), np.array([r[0] for r in L])
            both = (rin > P[4]) & (lin > P[4])
            idx = np.nonzero(both)[0]
            dt = np.array([np.abs(R[i][2][4:] - L[i][2][4:]).max() for i in idx])
            dr = np.array([quat_angle(R[i][2][:4], L[i][2][:4]) for i in idx])
            tot_h += len(R)
            print(f"  {ncl} clusters x {per} hypotheses, {na}-point samples, {npts} points, {int(100 * outl)} % outliers: same accept decision "
                  f"{((rin > P[4]) == (lin > P[4])).mean():.4f}, accepted by both {int(both.sum())}, same inlier count {(rin == lin).mean():.3f}, "
                  f"dt p50 / p90 / max {np.median(dt):.2e} / {np.quantile(dt, .9):.2e} / {dt.max():.2e} m, "
                  f"drot p50 / p90 / max {np.median(dr):.2e} / {np.quantile(dr, .9):.2e} / {dr.max():.2e} rad, LM_ERROR {int((rin < 0).sum())} vs {int((lin < 0).sum())}")
        fast, slow = C.c_long.in_dll(lib_ldl, "lmx_fast").value, C.c_long.in_dll(lib_ldl, "lmx_slow").value
        print(f"  {tot_h} hypotheses, {fast + slow} solves: LDL^T {fast} ({100 * fast / (fast + slow):.2f} %), pivoting LU {slow}")


if __name__ == "__main__":
    main()
