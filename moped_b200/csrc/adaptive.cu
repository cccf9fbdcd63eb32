// adaptive.cu — moped3d's MATCH step on the device: the exact two nearest rows of every feature (match.cu) followed by the
// depth-adaptive ratio test as a finalize kernel, plus the host-side derivation of a model's ratio curve.
//
// Replaces MATCH_ADAPTIVE_FLANN_CPU::Update / process (moped3d/libmoped/src/match/MATCH_ADAPTIVE_FLANN_CPU.hpp:101-174,419-490).
// The reference looks a depth up per feature, evaluates a per-model piecewise-linear threshold on the host and then tests
// dx[0]/dx[1] against it; here the curves of all models are a small table in device memory (16 bytes per model), the depth and
// fill-distance planes are uploaded once per frame, and one thread per feature reads its pixel, its nearest row's model and that
// model's curve (adaptive_threshold.cuh) — the decision leaves the device as one byte per feature.
#include "common.cuh"
#include "adaptive_threshold.cuh"

#include <cmath>

namespace mc {

__global__ void k_match_finalize_adaptive(const int32_t *__restrict__ nn_row, const float *__restrict__ nn_dist, int Q, const float *__restrict__ q_xy,
                                          const float *__restrict__ depth, const float *__restrict__ fill, int width, int height,
                                          const int32_t *__restrict__ model_of_row, int64_t table_base, const mc_adaptive_model *__restrict__ curves,
                                          int n_models, AdaptiveParams P, uint8_t *__restrict__ accepted) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= Q) return;
	bool ok = false;
	const int r0 = nn_row[2 * i], r1 = nn_row[2 * i + 1];
	if (r0 >= 0 && r1 >= 0) {
		const int px = adaptive_pixel(q_xy[2 * i], q_xy[2 * i + 1], width, height);
		const int m = model_of_row[r0 - table_base];
		if (m >= 0 && m < n_models) ok = adaptive_accept(curves[m], depth[px], fill[px], nn_dist[2 * i], nn_dist[2 * i + 1], P);
	}
	accepted[i] = ok ? 1 : 0;
}

// ---- ratio curve of one model (host, once per model change) ---------------------------------------------------------------------
// The reference asks: at which depth does the largest face of the model's bounding box, centred on the optical axis and parallel
// to the image plane, project to a square of `target` pixels a side? It answers by doubling an upper bound from 2 m and bisecting
// (at most 100 evaluations in total, stop within 1 %). The projected side is evaluated like the reference does — pinhole
// projection of the face's corners, extent of the projections, square root of the area — but only for the two distinct values each
// image coordinate takes (the four corners share them pairwise), which yields the same floats.
namespace {

struct Face { float half_a, half_b; };      // half extents of the largest face, a along image x, b along image y

Face largest_face(const float *lo, const float *hi) {
	const float ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
	// centred coordinates: +-e/2; the face areas are products of full extents rebuilt from the halves, as the reference forms them
	const float hx = ex / 2, hy = ey / 2, hz = ez / 2;
	const float fx = hx - (-hx), fy = hy - (-hy), fz = hz - (-hz);
	const float area_xy = fx * fy, area_xz = fx * fz, area_yz = fy * fz;
	Face f;
	if (area_xy >= area_yz && area_xy >= area_xz) { f.half_a = hx; f.half_b = hy; }
	else if (area_xz >= area_yz && area_xz >= area_xy) { f.half_a = hx; f.half_b = hz; }
	else { f.half_a = hy; f.half_b = hz; }
	return f;
}

float projected_side(const Face &f, const float *K, float z) {
	const float u_lo = (K[0] * -f.half_a + K[2] * z) / z, u_hi = (K[0] * f.half_a + K[2] * z) / z;
	const float v_lo = (K[1] * -f.half_b + K[3] * z) / z, v_hi = (K[1] * f.half_b + K[3] * z) / z;
	const float du = std::fmax(u_lo, u_hi) - std::fmin(u_lo, u_hi), dv = std::fmax(v_lo, v_hi) - std::fmin(v_lo, v_hi);
	return std::sqrt(du * dv);
}

float depth_for_side(const Face &f, const float *K, float target) {
	const int budget = 100;
	const float tolerance = 0.01;                     // a double literal narrowed to float, as the reference's parameter is
	int used = 0;
	float near = 0.f, far = 2.f;
	while (used < budget) {                           // push the far bound out until the face looks smaller than the target
		used++;
		if (projected_side(f, K, far) > target) far *= 2;
		else break;
	}
	const float slack = target * tolerance;
	while (used < budget) {
		used++;
		const float mid = (near + far) / 2;
		const float side = projected_side(f, K, mid);
		if (std::fabs(side - target) < slack) return mid;
		if (side > target) near = mid;
		else far = mid;
	}
	return (near + far) / 2;
}

} // namespace
} // namespace mc

using namespace mc;

extern "C" {

void mc_adaptive_model_init(mc_adaptive_model *out, const float *bbox_min, const float *bbox_max, const float *K, int n_features,
                            float min_ratio_min, float min_ratio_max, float max_ratio_min, float max_ratio_max, float dimension_peak,
                            float dimension_fade) {
	if (!out || !bbox_min || !bbox_max || !K) return;
	const Face f = largest_face(bbox_min, bbox_max);
	out->depth_peak = depth_for_side(f, K, dimension_peak);
	out->depth_fade = depth_for_side(f, K, dimension_fade);
	// sparse models (few features) get the permissive ends of both ratio ranges, dense ones the strict ends: a logistic in the
	// feature count centred on 1750 features with scale 250 (MATCH_ADAPTIVE_FLANN_CPU.hpp:103-104,160-166), evaluated in double
	const float centre = 1750, scale = 250;
	const float count = (float)n_features;
	const float x = (centre - count) / scale;
	const float density = (float)(1.0 / (1.0 + std::exp(-1.0 * (double)x)));
	out->ratio_low = min_ratio_min + density * (min_ratio_max - min_ratio_min);
	out->ratio_high = max_ratio_min + density * (max_ratio_max - max_ratio_min);
}

mc_status mc_match_adaptive(mc_ctx *ctx, const float *q_desc, const float *q_xy, int Q, const float *depth, const float *fill_distance, int width,
                            int height, const mc_adaptive_model *models, int n_models, float maximum_depth, float default_depth, float cauchy_scale,
                            int32_t *nn_row, float *nn_dist, uint8_t *accepted) {
	if (!ctx || !q_desc || !q_xy || !depth || !fill_distance || !models || !nn_row || !nn_dist || !accepted || Q < 0 || width < 1 || height < 1 ||
	    n_models < 1) {
		if (ctx) ctx->err = "mc_match_adaptive: bad argument";
		return MC_ERR_ARG;
	}
	if (!ctx->d_db) { ctx->err = "mc_match_adaptive: no database uploaded"; return MC_ERR_STATE; }
	if (n_models < ctx->n_models) { ctx->err = "mc_match_adaptive: fewer ratio curves than models in the database"; return MC_ERR_ARG; }
	if (Q == 0) return MC_OK;
	MC_CUDA(cudaSetDevice(ctx->device));
	const size_t px = (size_t)width * height;
	MC_TRY(reserve(ctx, ctx->q_desc, sizeof(float) * (size_t)Q * ctx->D));
	MC_TRY(reserve(ctx, ctx->nn_row, sizeof(int32_t) * 2 * (size_t)Q));
	MC_TRY(reserve(ctx, ctx->nn_dist, sizeof(float) * 2 * (size_t)Q));
	MC_TRY(reserve(ctx, ctx->accepted, (size_t)Q));
	MC_TRY(reserve(ctx, ctx->q_xy, sizeof(float) * 2 * (size_t)Q));
	const size_t o_fill = (sizeof(float) * px + 255) & ~(size_t)255, o_curves = 2 * o_fill;
	MC_TRY(reserve(ctx, ctx->scratch[19], o_curves + sizeof(mc_adaptive_model) * (size_t)n_models));
	char *b = (char *)ctx->scratch[19].p;
	MC_CUDA(cudaMemcpyAsync(ctx->q_desc.p, q_desc, sizeof(float) * (size_t)Q * ctx->D, cudaMemcpyHostToDevice, ctx->stream));
	MC_CUDA(cudaMemcpyAsync(ctx->q_xy.p, q_xy, sizeof(float) * 2 * (size_t)Q, cudaMemcpyHostToDevice, ctx->stream));
	MC_CUDA(cudaMemcpyAsync(b, depth, sizeof(float) * px, cudaMemcpyHostToDevice, ctx->stream));
	MC_CUDA(cudaMemcpyAsync(b + o_fill, fill_distance, sizeof(float) * px, cudaMemcpyHostToDevice, ctx->stream));
	MC_CUDA(cudaMemcpyAsync(b + o_curves, models, sizeof(mc_adaptive_model) * (size_t)n_models, cudaMemcpyHostToDevice, ctx->stream));
	// the fixed-ratio flags of the plain matcher are overwritten by the adaptive kernel
	MC_TRY(match_device(ctx, (const float *)ctx->q_desc.p, Q, 1.0f, MC_MATCH_TENSOR, (int32_t *)ctx->nn_row.p, (float *)ctx->nn_dist.p, (uint8_t *)ctx->accepted.p));
	AdaptiveParams P;
	P.maximum_depth = maximum_depth; P.default_depth = default_depth; P.cauchy_scale = cauchy_scale;
	k_match_finalize_adaptive<<<(Q + 255) / 256, 256, 0, ctx->stream>>>((const int32_t *)ctx->nn_row.p, (const float *)ctx->nn_dist.p, Q, (const float *)ctx->q_xy.p,
	                                                                  (const float *)b, (const float *)(b + o_fill), width, height, ctx->d_model_of_row,
	                                                                  ctx->table_base, (const mc_adaptive_model *)(b + o_curves), n_models, P,
	                                                                  (uint8_t *)ctx->accepted.p);
	MC_LAUNCH_CHECK();
	MC_CUDA(cudaMemcpyAsync(nn_row, ctx->nn_row.p, sizeof(int32_t) * 2 * (size_t)Q, cudaMemcpyDeviceToHost, ctx->stream));
	MC_CUDA(cudaMemcpyAsync(nn_dist, ctx->nn_dist.p, sizeof(float) * 2 * (size_t)Q, cudaMemcpyDeviceToHost, ctx->stream));
	MC_CUDA(cudaMemcpyAsync(accepted, ctx->accepted.p, (size_t)Q, cudaMemcpyDeviceToHost, ctx->stream));
	MC_CUDA(cudaStreamSynchronize(ctx->stream));
	return MC_OK;
}

} // extern "C"
