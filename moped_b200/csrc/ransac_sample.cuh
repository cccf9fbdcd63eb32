// ransac_sample.cuh — randSample + initPose of the RANSAC stages on the seedable stream shared with the oracle
// (POSE_RANSAC_LM_DIFF_REPROJECTION_CPU.hpp:76-98,182-186; the moped3d depth variants use the same two functions).
// Included by pose.cu and pose_depth.cu.
#pragma once
#include "common.cuh"

#include <math_constants.h>

namespace mc {

constexpr int kMaxAlign = 8;        // NPtsAlign <= 8 (5 and 6 in every reference config)

// ---- the seedable LCG shared with the oracle (oracle/moped_oracle.c: mo_rand) with O(log k) jump-ahead ----
__device__ __forceinline__ void lcg_jump(uint64_t k, uint64_t &A, uint64_t &C) {
	uint64_t a = 6364136223846793005ULL, c = 1442695040888963407ULL;
	A = 1; C = 0;
	while (k) {
		if (k & 1) { C = C * a + c; A = A * a; }
		c = c * a + c; a = a * a;
		k >>= 1;
	}
}
__device__ __forceinline__ int lcg_out(uint64_t s) { return (int)((s >> 33) & 0x7fffffffULL); }

// randSample + initPose for hypothesis h of a task (POSE_..._CPU.hpp:76-98,182-186): the task's LCG stream
// hands draw h*(n+4)+i to cluster point i (key = (float)rand()) and the next four draws to the quaternion
// (w first: g++ evaluates the arguments right to left). Points are taken in ascending (key, match index)
// order, skipping repeated (image, coord2D). Lane group cooperative. Returns false if fewer than n_align
// distinct points exist.
template <int G>
__device__ bool draw_sample(uint64_t seed, int h, int n, int n_align, const float *xy, const int32_t *image, const int32_t *tie,
                            unsigned mask, int lig, int (&pos)[kMaxAlign], float (&quat)[4]) {
	uint64_t A0, C0, AG, CG;
	lcg_jump((uint64_t)h * (uint64_t)(n + 4) + (uint64_t)lig + 1, A0, C0);   // state after (index+1) steps
	lcg_jump((uint64_t)G, AG, CG);
	const uint64_t s_first = seed * A0 + C0;
	float last_key = -1.f; int last_tie = -1;
	int taken = 0;
	for (int round = 0; round < n && taken < n_align; round++) {
		// smallest (key, tie) strictly above the last popped one
		float bk = CUDART_INF_F; int bt = 0x7fffffff, bp = -1;
		uint64_t s = s_first;
		for (int i = lig; i < n; i += G) {
			const float key = (float)lcg_out(s);
			const int t = tie ? tie[i] : i;
			const bool above = key > last_key || (key == last_key && t > last_tie);
			if (above && (key < bk || (key == bk && t < bt))) { bk = key; bt = t; bp = i; }
			s = s * AG + CG;
		}
#pragma unroll
		for (int o = G / 2; o; o >>= 1) {
			const float ok = __shfl_xor_sync(mask, bk, o);
			const int ot = __shfl_xor_sync(mask, bt, o), op = __shfl_xor_sync(mask, bp, o);
			if (ok < bk || (ok == bk && ot < bt)) { bk = ok; bt = ot; bp = op; }
		}
		if (bp < 0) break;
		last_key = bk; last_tie = bt;
		bool dup = false;
#pragma unroll
		for (int j = 0; j < kMaxAlign; j++)
			if (j < taken) {
				const int sidx = pos[j];
				dup = dup || (image[sidx] == image[bp] && xy[2 * sidx] == xy[2 * bp] && xy[2 * sidx + 1] == xy[2 * bp + 1]);
			}
		if (!dup) {
#pragma unroll
			for (int j = 0; j < kMaxAlign; j++) if (j == taken) pos[j] = bp;
			taken++;
		}
	}
	uint64_t Aq, Cq;
	lcg_jump((uint64_t)h * (uint64_t)(n + 4) + (uint64_t)n + 1, Aq, Cq);
	uint64_t s = seed * Aq + Cq;
#pragma unroll
	for (int j = 3; j >= 0; j--) {
		quat[j] = (float)((lcg_out(s) & 255) / 256.);
		s = s * 6364136223846793005ULL + 1442695040888963407ULL;
	}
	return taken == n_align;
}

} // namespace mc
