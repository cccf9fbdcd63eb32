// lm_exact.cuh — slevmar_dif for a TEAM of W lanes with every floating-point sum taken in levmar's own order.
//
// Why: moped3d's depth-aware pose stages minimise squared metre-scale distances (~1e-6) in fp32; that LM is so
// ill-conditioned that two builds of the UNMODIFIED reference (its own -ffast-math flags vs strict IEEE) already differ
// from each other by milliradians. A tolerance against such a target says little, so this LM does not
// re-associate anything: products and sums are rounded separately (the translation unit is compiled with -fmad=false)
// and every reduction runs in the order of libs.tgz!levmar-2.4 (lm_core.c:427-836, misc_core.c:135-168,712-790,
// Axb_core.c:888-1035), i.e. the result is the strict-IEEE build's, bit for bit, however many lanes work on it.
//
// Provenance: this file is DERIVED FROM levmar 2.4 (Manolis Lourakis; GPL; vendored by the reference as libs.tgz!levmar-2.4): the LM driver,
// the Jacobian approximation and the LU solve restate lm_core.c / misc_core.c / Axb_core.c operation by operation — bit-exactness
// with the reference leaves no freedom there. The lane-team decomposition below is this repository's.
//
// How the work is split without touching the order:
//   * residuals, the finite-difference Jacobian and the Broyden rank-1 update are independent per correspondence /
//     per residual row  -> lanes stride over them;
//   * J^T J (28 entries) and J^T e (7) are each ONE sequential chain over the residual rows (descending, like
//     levmar's loop) but the 35 chains are independent -> one chain per lane;
//   * ||e||^2 is levmar's four interleaved accumulators -> one accumulator per lane (4 lanes), summed s0+s1+s2+s3;
//   * the 7x7 Crout LU, the damping logic and the stop tests are scalar -> every lane runs them redundantly on the
//     same values (no communication).
// State that crosses lanes lives in a scratch block (`Work`, shared or global memory); the code is a sequence of phases
// (simt_phases.cuh: `team.each(f)` = __syncwarp(); f(lane); __syncwarp(); on the device, a loop over the lanes on the host),
// which is what lets tests/cpp/depth_host.cpp compile this very source with g++ and check it against the oracle (and through it
// against the compiled reference) on the CPU. Nothing here is a CPU path of the product.
#pragma once

#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stddef.h>

#include "simt_phases.cuh"

#define LMX_FN PHX_FN
#define LMX_MEM PHX_MEM

namespace lmx {

constexpr int M = 7;                 // parameters: raw quaternion (x,y,z,w) + translation

template <int W> using Team = phx::WarpTeam<W>;      // W lanes working on one LM problem (simt_phases.cuh)

// scratch of one team for a problem with up to n residuals
struct Work {
	float *e, *hx, *wrk, *wrk2, *jac;    // n, n, n, n, n*M
	float *jtj, *jte, *acc;              // M*M (lower triangle used), M, 4
};
LMX_FN size_t work_floats(int n) { return (size_t)n * (4 + M) + 64; }
LMX_FN Work work_carve(float *base, int n) {
	Work w;
	w.jtj = base; w.jte = base + 49; w.acc = base + 56;
	float *v = base + 64;
	w.e = v; w.hx = v + n; w.wrk = v + 2 * (size_t)n; w.wrk2 = v + 3 * (size_t)n; w.jac = v + 4 * (size_t)n;
	return w;
}

// Pt<4>::norm (moped.hpp:122): fp32 sum of squares, fp32 sqrt, the reciprocal taken in double and narrowed
LMX_FN void quat_norm(float *q) {
	float d = 0.f;
	for (int x = 0; x < 4; x++) d += q[x] * q[x];
	d = (float)(1. / (double)sqrtf(d));
	for (int x = 0; x < 4; x++) q[x] *= d;
}
// TransformMatrix::init (moped.hpp:175-182)
LMX_FN void tm_init(float *T, const float *q, const float *t) {
	T[0] = 1 - 2 * q[1] * q[1] - 2 * q[2] * q[2]; T[1] = 2 * q[0] * q[1] - 2 * q[3] * q[2]; T[2] = 2 * q[0] * q[2] + 2 * q[3] * q[1]; T[3] = t[0];
	T[4] = 2 * q[0] * q[1] + 2 * q[3] * q[2]; T[5] = 1 - 2 * q[0] * q[0] - 2 * q[2] * q[2]; T[6] = 2 * q[1] * q[2] - 2 * q[3] * q[0]; T[7] = t[1];
	T[8] = 2 * q[0] * q[2] - 2 * q[3] * q[1]; T[9] = 2 * q[1] * q[2] + 2 * q[3] * q[0]; T[10] = 1 - 2 * q[0] * q[0] - 2 * q[1] * q[1]; T[11] = t[2];
}
// what every residual function of the path does first: normalise a COPY of the quaternion, build the 3x4 matrix
LMX_FN void pose_matrix(const float *p7, float *T) {
	float q[4] = { p7[0], p7[1], p7[2], p7[3] };
	quat_norm(q);
	tm_init(T, q, p7 + 4);
}

// sAx_eq_b_LU_noLapack (Axb_core.c:888-1035) for the damped normal equations (JtJ + mu I) x = Jte; JtJ's lower triangle
// is read from L (row-major M x M). Crout LU, implicit row scaling, partial pivoting; forward substitution with the
// permutation vector and levmar's skip of leading zeros. Returns false if singular.
// Every array index is a compile-time constant (loops fully unrolled, the data-dependent pivot row handled by selects), so the
// 7 x 7 matrix lives in registers: a dynamically indexed row would put the whole matrix into local memory. The arithmetic — which
// products, in which order, rounded where — is levmar's, including what it does on a column of NaNs (`maxi` keeps the previous
// column's row, which may lie ABOVE the diagonal; a first column of NaNs counts as singular, see below).
LMX_FN bool lu_solve(const float *L, float mu, const float *B, float *x) {
	float a[M][M], work[M];
	int idx[M], maxi = -1;
#pragma unroll
	for (int i = 0; i < M; i++)
#pragma unroll
		for (int j = 0; j < M; j++) a[i][j] = i >= j ? L[i * M + j] : L[j * M + i];
#pragma unroll
	for (int i = 0; i < M; i++) { a[i][i] += mu; x[i] = B[i]; }
	bool singular = false;
#pragma unroll
	for (int i = 0; i < M; i++) {
		float mx = 0.f;
#pragma unroll
		for (int j = 0; j < M; j++) { const float t = fabsf(a[i][j]); if (t > mx) mx = t; }
		if (mx == 0.f) singular = true;
		work[i] = 1.0f / mx;
	}
	if (singular) return false;
#pragma unroll
	for (int j = 0; j < M; j++) {
#pragma unroll
		for (int i = 0; i < j; i++) {
			float sum = a[i][j];
#pragma unroll
			for (int k = 0; k < i; k++) sum -= a[i][k] * a[k][j];
			a[i][j] = sum;
		}
		float mx = 0.f;
#pragma unroll
		for (int i = j; i < M; i++) {
			float sum = a[i][j];
#pragma unroll
			for (int k = 0; k < j; k++) sum -= a[i][k] * a[k][j];
			a[i][j] = sum;
			const float t = work[i] * fabsf(sum);
			if (t >= mx) { mx = t; maxi = i; }
		}
		// levmar starts with maxi = -1 and a column of NaNs never sets it (t >= mx is false): the reference then swaps with the
		// row BEFORE its matrix (undefined behaviour). No such access here: the system counts as singular.
		if (maxi < 0) return false;
		if (j != maxi) {
			// rows j and maxi trade places (maxi is any row: below the diagonal normally, possibly above after a NaN column)
#pragma unroll
			for (int i = 0; i < M; i++) {
				if (i == j) continue;
				const bool sw = maxi == i;
#pragma unroll
				for (int k = 0; k < M; k++) {
					const float u = a[i][k], v = a[j][k];
					a[i][k] = sw ? v : u; a[j][k] = sw ? u : v;
				}
				work[i] = sw ? work[j] : work[i];
			}
		}
		idx[j] = maxi;
		if (a[j][j] == 0.f) a[j][j] = FLT_EPSILON;
		if (j != M - 1) {
			const float t = 1.0f / a[j][j];
#pragma unroll
			for (int i = j + 1; i < M; i++) a[i][j] *= t;
		}
	}
	int k = 0;
#pragma unroll
	for (int i = 0; i < M; i++) {
		// sum = x[idx[i]]; x[idx[i]] = x[i];  with the data-dependent index resolved by selects
		const int pj = idx[i];
		float sum = x[0];
#pragma unroll
		for (int r = 1; r < M; r++) sum = pj == r ? x[r] : sum;
		const float xi = x[i];
#pragma unroll
		for (int r = 0; r < M; r++) x[r] = pj == r ? xi : x[r];
		if (k != 0) {
#pragma unroll
			for (int j = 0; j < i; j++) if (j >= k - 1) sum -= a[i][j] * x[j];
		} else if (sum != 0.f) k = i + 1;
		x[i] = sum;
	}
#pragma unroll
	for (int i = M - 1; i >= 0; i--) {
		float sum = x[i];
#pragma unroll
		for (int j = i + 1; j < M; j++) sum -= a[i][j] * x[j];
		x[i] = sum / a[i][i];
	}
	return true;
}

// e = 0 - y and ||e||^2 in slevmar_L2nrmxmy's order (misc_core.c:712-790, x = 0): blocks of 8 walked downwards, element
// i-a of a block into accumulator a%4; the tail (ascending) into accumulator 0; s0+s1+s2+s3.
template <int W>
LMX_FN float l2_neg(const Team<W> &team, const Work &w, float *e, const float *y, int n) {
	team.each([&](int lane) {
		for (int i = lane; i < n; i += W) e[i] = 0.f - y[i];
	});
	team.each([&](int lane) {
		const int blockn = (n >> 3) << 3;
		for (int a = lane; a < 4; a += W) {
			float s = 0.f;
			for (int i = blockn - 1; i > 0; i -= 8) {
				s += e[i - a] * e[i - a];
				s += e[i - a - 4] * e[i - a - 4];
			}
			if (a == 0) for (int i = blockn; i < n; i++) s += e[i] * e[i];
			w.acc[a] = s;
		}
	});
	return w.acc[0] + w.acc[1] + w.acc[2] + w.acc[3];
}

// Residual model: Fn::R residuals per correspondence; fn.point(T, k, r) writes the R residuals of correspondence k under
// the 3x4 pose matrix T.
template <int W, class Fn>
LMX_FN void eval(const Team<W> &team, const Fn &fn, const float *p, int n_pts, float *out) {
	float T[12];
	pose_matrix(p, T);
	team.each([&](int lane) {
		for (int k = lane; k < n_pts; k += W) fn.point(T, k, out + Fn::R * k);
	});
}

// slevmar_dif(func, p, x = 0, m = 7, n = R * n_pts, itmax, opts = NULL, ...) — lm_core.c:427-836 with the defaults of
// lm.h:83-85. `finite_check` keeps levmar's stop = 7 on a non-finite ||e||^2 (the reference's -ffast-math build folds it
// away; a strict build keeps it). Returns the iteration count or -1 (stop 4 / 7); *err_out = ||e||^2 at the solution.
// `stop_flag` (nullable): a word that may drop below `my_index` while the LM runs — a RANSAC test with a lower index has succeeded,
// so this one cannot be chosen any more and gives up (returns -1; its result is never used). Read by member 0 and broadcast.
template <int W, class Fn>
LMX_FN int levmar_dif(const Team<W> &team, const Fn &fn, float *p, int n_pts, int itmax, const Work &w, bool finite_check, float *err_out,
                      const volatile int *stop_flag = nullptr, int my_index = 0) {
	const int n = Fn::R * n_pts;
	const float tau = 1E-03f, eps1 = 1E-17f, eps2 = 1E-17f, eps2_sq = 1E-17f * 1E-17f, eps3 = 1E-17f, delta = 1E-06f;
	float Dp[M], diag[M], pDp[M];
	float mu = 0.f, jte_inf = 0.f, p_L2 = 0.f, Dp_L2, p_eL2, pDp_eL2, tmp;
	int nu = 20, nu2, stop = 0, K = 10, updjac = 0, updp = 1, newjac = 0, k;

	eval(team, fn, p, n_pts, w.hx);
	p_eL2 = l2_neg(team, w, w.e, w.hx, n);
	if (finite_check && !isfinite(p_eL2)) stop = 7;

	for (k = 0; k < itmax && !stop; ++k) {
		if (stop_flag && team.from_first(*stop_flag) < my_index) { stop = 4; break; }
		if (p_eL2 <= eps3) { stop = 6; break; }

		if ((updp && nu > 16) || updjac == K) {
			// forward differences (slevmar_fdif_forw_jac_approx, misc_core.c:135-168): one column per parameter
			float Tj[M][12], dinv[M];
			for (int j = 0; j < M; j++) {
				float d = 1E-04f * p[j];
				d = fabsf(d);
				if (d < delta) d = delta;
				const float save = p[j];
				p[j] += d;
				pose_matrix(p, Tj[j]);
				p[j] = save;
				dinv[j] = 1.0f / d;
			}
			team.each([&](int lane) {
				for (int pt = lane; pt < n_pts; pt += W)
					for (int j = 0; j < M; j++) {
						float r[Fn::R];
						fn.point(Tj[j], pt, r);
						for (int c = 0; c < Fn::R; c++) {
							const int i = Fn::R * pt + c;
							w.jac[i * M + j] = (r[c] - w.hx[i]) * dinv[j];
						}
					}
			});
			nu = 2; updjac = 0; updp = 0; newjac = 1;
		}

		if (newjac) {
			newjac = 0;
			// J^T J (lower triangle) and J^T e: 35 independent chains over the rows l = n-1 .. 0 (lm_core.c:583-624)
			team.each([&](int lane) {
				for (int t = lane; t < 35; t += W) {
					float s = 0.f;
					if (t < 28) {
						int i = 0;
						while ((i + 1) * (i + 2) / 2 <= t) i++;
						const int j = t - i * (i + 1) / 2;
						for (int l = n - 1; l >= 0; l--) s += w.jac[l * M + j] * w.jac[l * M + i];
						w.jtj[i * M + j] = s;
					} else {
						const int i = t - 28;
						for (int l = n - 1; l >= 0; l--) s += w.jac[l * M + i] * w.e[l];
						w.jte[i] = s;
					}
				}
			});
			p_L2 = jte_inf = 0.f;
			for (int i = 0; i < M; i++) {
				tmp = fabsf(w.jte[i]);
				if (jte_inf < tmp) jte_inf = tmp;
				diag[i] = w.jtj[i * M + i];
				p_L2 += p[i] * p[i];
			}
		}

		if (jte_inf <= eps1) { stop = 1; break; }

		if (k == 0) {
			tmp = -FLT_MAX;
			for (int i = 0; i < M; i++) if (diag[i] > tmp) tmp = diag[i];
			mu = tau * tmp;
		}

		float jte_l[M];
		for (int i = 0; i < M; i++) jte_l[i] = w.jte[i];
		const bool solved = lu_solve(w.jtj, mu, jte_l, Dp);
		if (solved) {
			Dp_L2 = 0.f;
			for (int i = 0; i < M; i++) { tmp = Dp[i]; pDp[i] = p[i] + tmp; Dp_L2 += tmp * tmp; }
			if (Dp_L2 <= eps2_sq * p_L2) { stop = 2; break; }
			if (Dp_L2 >= (p_L2 + eps2) / (1E-12f * 1E-12f)) { stop = 4; break; }

			eval(team, fn, pDp, n_pts, w.wrk);
			pDp_eL2 = l2_neg(team, w, w.wrk2, w.wrk, n);
			if (finite_check && !isfinite(pDp_eL2)) { stop = 7; break; }
			const float dF = p_eL2 - pDp_eL2;
			if (updp || dF > 0) {
				// Broyden rank-1 update of J, row by row (lm_core.c:742-752)
				team.each([&](int lane) {
					for (int i = lane; i < n; i += W) {
						float t = 0.f;
						for (int l = 0; l < M; l++) t += w.jac[i * M + l] * Dp[l];
						t = (w.wrk[i] - w.hx[i] - t) / Dp_L2;
						for (int j = 0; j < M; j++) w.jac[i * M + j] += t * Dp[j];
					}
				});
				++updjac; newjac = 1;
			}
			float dL = 0.f;
			for (int i = 0; i < M; i++) dL += Dp[i] * (mu * Dp[i] + jte_l[i]);
			if (dL > 0.f && dF > 0.f) {
				tmp = 2.0f * dF / dL - 1.0f;
				tmp = 1.0f - tmp * tmp * tmp;
				mu = mu * ((tmp >= 0.3333333334f) ? tmp : 0.3333333334f);
				nu = 2;
				for (int i = 0; i < M; i++) p[i] = pDp[i];
				team.each([&](int lane) {
					for (int i = lane; i < n; i += W) { w.e[i] = w.wrk2[i]; w.hx[i] = w.wrk[i]; }
				});
				p_eL2 = pDp_eL2;
				updp = 1;
				continue;
			}
		}
		mu *= nu;
		nu2 = nu << 1;
		if (nu2 <= nu) { stop = 5; break; }
		nu = nu2;
	}
	*err_out = p_eL2;
	return (stop != 4 && stop != 7) ? k : -1;
}

} // namespace lmx
