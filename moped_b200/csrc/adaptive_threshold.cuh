// adaptive_threshold.cuh — the per-feature acceptance rule of moped3d's depth-adaptive MATCH step, written once for the device
// kernel (adaptive.cu) and for host-compiled checks (oracle/ref3d_match_dropin.cpp builds this header with g++ -ffp-contract=off).
//
// What it computes (moped3d/libmoped/src/match/MATCH_ADAPTIVE_FLANN_CPU.hpp:182-205,360-372,419-470): every model has a ratio
// curve over depth — a ramp from `ratio_low` (depth 0) to `ratio_high` (depth_peak), a plateau up to depth_fade, a ramp down to
// zero at twice depth_fade, zero beyond `maximum_depth`. A feature's threshold is the curve at the depth under its pixel blended
// with the curve at a default depth; the blend weight is a Cauchy kernel of the depth map's fill distance at that pixel (1 where
// the depth was measured, towards 0 where it was filled in from far away). The feature is accepted when the ratio of its two
// nearest squared distances is below that threshold.
//
// Rounding contract: the reference evaluates this in fp32 with a few sub-expressions promoted to double by its literals; the same
// promotions are spelled out here, and every product and sum is a separate rounding (no fused multiply-add on either side), so
// host and device produce the bits of a strict-IEEE build of the reference. Plain C++98.
#pragma once
#include "../../include/moped_cuda.h"

#ifdef __CUDA_ARCH__
#define MC_AT_HD __host__ __device__ __forceinline__
#define MC_AT_FMUL(a, b) __fmul_rn((a), (b))
#define MC_AT_FADD(a, b) __fadd_rn((a), (b))
#define MC_AT_FSUB(a, b) __fsub_rn((a), (b))
#define MC_AT_FDIV(a, b) __fdiv_rn((a), (b))
#define MC_AT_DMUL(a, b) __dmul_rn((a), (b))
#define MC_AT_DADD(a, b) __dadd_rn((a), (b))
#define MC_AT_DDIV(a, b) __ddiv_rn((a), (b))
#else
#ifdef __CUDACC__
#define MC_AT_HD __host__ __device__ inline
#else
#define MC_AT_HD inline
#endif
#define MC_AT_FMUL(a, b) ((float)(a) * (float)(b))
#define MC_AT_FADD(a, b) ((float)(a) + (float)(b))
#define MC_AT_FSUB(a, b) ((float)(a) - (float)(b))
#define MC_AT_FDIV(a, b) ((float)(a) / (float)(b))
#define MC_AT_DMUL(a, b) ((double)(a) * (double)(b))
#define MC_AT_DADD(a, b) ((double)(a) + (double)(b))
#define MC_AT_DDIV(a, b) ((double)(a) / (double)(b))
#endif

namespace mc {

struct AdaptiveParams {        // constants of the stage (MATCH_ADAPTIVE_FLANN_CPU.hpp:103-108: 4.0 m, 1.0 m, 0.1)
	float maximum_depth, default_depth, cauchy_scale;
};

// value of a model's ratio curve at `depth`
MC_AT_HD float ratio_curve(const mc_adaptive_model &c, float depth, float maximum_depth) {
	if (depth > maximum_depth) return 0.f;
	if (depth < c.depth_peak) {
		const float along = MC_AT_FDIV(depth, c.depth_peak);
		return MC_AT_FADD(c.ratio_low, MC_AT_FMUL(along, MC_AT_FSUB(c.ratio_high, c.ratio_low)));
	}
	if (depth < c.depth_fade) return c.ratio_high;
	const float twice = MC_AT_FMUL(c.depth_fade, 2.f);
	if (depth < twice) return MC_AT_FMUL(MC_AT_FDIV(MC_AT_FSUB(twice, depth), c.depth_fade), c.ratio_high);
	return 0.f;
}

// threshold of a feature lying over (depth, fill_distance) whose nearest row belongs to the model with curve `c`
MC_AT_HD float adaptive_threshold(const mc_adaptive_model &c, float depth, float fill_distance, const AdaptiveParams &P) {
	const float t = MC_AT_FDIV(fill_distance, P.cauchy_scale);
	const float w = (float)MC_AT_DDIV(1.0, MC_AT_DADD(1.0, (double)MC_AT_FMUL(t, t)));                 // Cauchy weight, evaluated in double like the reference's literals make it
	const float here = ratio_curve(c, depth, P.maximum_depth), fallback = ratio_curve(c, P.default_depth, P.maximum_depth);
	return (float)MC_AT_DADD((double)MC_AT_FMUL(w, here), MC_AT_DMUL(MC_AT_DADD(1.0, -(double)w), (double)fallback));
}

// pixel under a feature: truncation towards zero, then the reference's clamp to [0, width] x [0, height] (sic, :423-424); the
// linear index is kept inside the map (the reference reads past its buffer for x == width on the last row)
MC_AT_HD int adaptive_pixel(float fx, float fy, int width, int height) {
	int x = (int)fx, y = (int)fy;
	x = x < 0 ? 0 : (x > width ? width : x);
	y = y < 0 ? 0 : (y > height ? height : y);
	const long long idx = (long long)y * width + x, last = (long long)width * height - 1;
	return (int)(idx > last ? last : idx);
}

// the whole decision for one feature; d0 / d1 = squared distances of its two nearest rows
MC_AT_HD bool adaptive_accept(const mc_adaptive_model &c, float depth, float fill_distance, float d0, float d1, const AdaptiveParams &P) {
	if (depth > P.maximum_depth) return false;                     // the reference does not even search these (:431-434)
	return MC_AT_FDIV(d0, d1) < adaptive_threshold(c, depth, fill_distance, P);
}

} // namespace mc
