// common.cuh — shared declarations of libmoped_cuda (B200 / sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/moped_cuda.h"

namespace mc {

// ---- sizes fixed by the tensor-core matcher -------------------------------------------------
constexpr int kD = 128;                 // descriptor length of the tcgen05 path (SIFT)
constexpr int kTileRows = 128;          // DB rows / query rows per operand tile image
constexpr int kTileBytes = kTileRows * kD * 2;   // 32 KiB: fp16, two 128B-swizzle K atoms of 64 elements
constexpr int kMTile = 256;             // queries per CTA (two 128-row halves)
#ifndef MC_TOPK
#define MC_TOPK 4
#endif
constexpr int kTopK = MC_TOPK;          // coarse candidates kept per (query, DB split) by the fp16 pass
#ifndef MC_TOPK8
#define MC_TOPK8 8
#endif
constexpr int kTopK8 = MC_TOPK8;               // ... by the 8-bit pass (its certificate charges a ~12x larger score error)
constexpr int kMaxSplits = 64;

struct Camera {            // FrameData::images[i]: K=(fx,fy,cx,cy), TM = 3x4 of cameraPose (moped.hpp:226-241)
	float K[4];
	float TM[12];
};

struct FrameDesc {         // what differs between the frames of a stream: read by the stage chain through ONE indirection,
	const int32_t *nn_row;      // so that the chain itself (a CUDA graph per lane) never changes
	const uint8_t *accepted;
	const float *q_xy;
	const int32_t *q_image;
	int32_t *out_info, *out_model;
	float *out_pose, *out_score;
	int Q, max_objects;
};

struct DevBuf {            // grow-only device scratch
	void *p = nullptr;
	size_t cap = 0;
};

} // namespace mc

struct mc_ctx {
	int device = 0;
	int num_sms = 148;
	cudaStream_t stream = nullptr;
	bool own_stream = false;
	std::string err;
	int64_t launches = 0;

	// model database (device resident)
	int64_t n_rows = 0, row_base = 0, n_tiles = 0;
	int D = 0, n_models = 0;
	float *d_db = nullptr;            // n_rows x D fp32, row-major (exact re-rank / exact scan)
	__half *d_db_img = nullptr;       // n_tiles x 32 KiB pre-swizzled fp16 operand tiles (tcgen05 B operand)
	uint8_t *d_db_img8 = nullptr;     // n_tiles x 16 KiB pre-swizzled u8 / s8 operand tiles, one scale for all rows
	float db_scale = 0.f;             // 8-bit image: q8 = round(db_scale * x); 0 = no usable image (empty / non-finite database)
	bool db_signed = false;           // 8-bit image holds s8 (some element is negative) instead of u8
	float db_err_max = 0.f;           // max over rows of |x - q8 / db_scale|_2
	float *d_xyz = nullptr;           // table_rows x 3 (coord3D of rows table_base .. table_base+table_rows)
	int32_t *d_model_of_row = nullptr;
	int64_t table_base = 0, table_rows = 0;   // = this shard's rows unless mc_db_set_global_tables was called
	float db_norm2_min = 1.f, db_norm2_max = 1.f;

	// cameras
	mc::Camera *d_cams = nullptr;
	int n_images = 0;

	// scratch
	mc::DevBuf q_desc, q_img, q_norm2, tau, cand_score, cand_row, flag_list, flag_count, nn_key;
	mc::DevBuf q_img8, q_signed, q_scale, q_err, flag_list2, tau2, cand_score2, cand_row2;   // 8-bit pass + fp16 second-chance pass
	int coarse_kind = 1;              // mc_set_option "match_coarse_kind": 1 = 8-bit pass first (default), 0 = fp16 pass only
	// spatial partition (sm_partition.cu, mc_set_option "stage_sm_partition"): green contexts, the stream the coarse kernel runs on
	void *green_match = nullptr, *green_stage = nullptr;
	cudaStream_t match_green_stream = nullptr;
	cudaEvent_t ev_green[2] = {nullptr, nullptr};
	int match_sms = 0, stage_sms = 0; // SMs of the two partitions (0 = no partition)
	int match_splits = 0;             // mc_set_option "match_splits": database splits per query tile in the coarse pass (0 = chosen from the grid)
	int match_stagger = 1;            // mc_set_option "match_stagger": CTAs of one DB split start at different tiles
	int match_reserve_sms = 0;        // mc_set_option "match_reserve_sms": SMs the persistent matching kernel leaves to concurrent work
	mc::DevBuf nn_row, nn_dist, accepted, q_xy, q_image;
	mc::DevBuf scratch[24];
	void *h_pinned = nullptr; size_t h_pinned_cap = 0;
	int last_stats[4] = {0, 0, 0, 0};
	int last_match_q = 0, last_match_tensor = 0;   // queries / mode of the last match_device pass (mc_match_last_stats)
	bool profile = false;             // record CUDA events around the dominant kernel (k_match_coarse)
	cudaEvent_t ev_coarse[2] = {nullptr, nullptr};
	bool ev_valid = false;

	// frame batches (mc_process_frames*): the stages after MATCH of different frames run concurrently on
	// "lanes" — child contexts with their own stream and scratch that borrow this context's database tables
	// and cameras. A lane never owns the database.
	mc_ctx *parent = nullptr;
	std::vector<mc_ctx *> lanes;
	int n_lanes_wanted = 8;           // mc_set_tuning
	int pose_warps = 8;               // warps per RANSAC task CTA (mc_set_tuning); results do not depend on it
	int64_t fit_thread_min = 16384;   // mc_set_option: explicit-hypothesis calls with at least this many use one thread per hypothesis
	bool fit_stream = true;           // mc_set_option "pose_fit_stream": the persistent phase-synchronous one-thread-per-hypothesis kernel (0: k_pose_fit_thread)
	int ransac_shard_rank = 0, ransac_shard_world = 1;   // set around the RANSAC calls of mc_process_frame_sharded_dev (pose_staged.cuh)
	bool ransac_merge_levels = true;  // mc_set_option "ransac_merge_levels" (pose_staged.cuh)
	bool ransac_fused = false;        // mc_set_option: the one-CTA-per-task RANSAC kernel instead of the staged kernels (A/B aid)
	int depth_team_lanes = 32;        // mc_set_option: lanes per explicit hypothesis in k_depth_hypotheses (32 or 8)
	bool linkage_cached = true;       // mc_set_option: cached-row-maximum agglomeration (linkage_cached.cuh) for average linkage (0: the O(n^2)-scan kernel)
	bool pose_exact_order = false;    // mc_set_option: mc_pose_hypotheses / mc_pose_ransac run the order-preserving LM (pose_depth.cu, variant 2)
	bool lm_finite_check = false;     // mc_set_option: depth pose stages keep levmar's stop on a non-finite ||e||^2 (a strict-IEEE build of the reference)
	int match_chunks = 1;             // MATCH launches per batch (mc_set_tuning): chunk c+1 matches while chunk c runs its lanes
	cudaEvent_t ev_fork = nullptr, ev_done = nullptr;
	std::vector<cudaEvent_t> ev_chunk;
	mc::DevBuf batch_out;             // per-frame result slots of the running batch
	// per-lane CUDA graph of the stage chain CLUSTER..FILTER2 (+ compaction and export): replayed once per frame
	mc::DevBuf frame_desc;            // device copy of the current frame's FrameDesc
	struct FrameGraph { uint64_t cfg, ptr_key; cudaGraphExec_t exec; int nodes; };
	std::vector<FrameGraph> fgraphs;  // one per configuration seen on this lane (sizes, parameters), most recent last
	struct BatchGraph { uint64_t key; cudaGraphExec_t exec; int nodes; };
	std::vector<BatchGraph> bgraphs;  // one graph per batch configuration: the stage chains of all frames of a call as parallel branches (pipeline.cu)
	bool batch_graph = true;          // mc_set_option "batch_graph"
	bool capturing = false;           // reserve() must not allocate while the lane's stream is being captured
	bool frame_graphs = true;         // mc_set_option "frame_graphs"
	bool defer_lane_join = false;     // mc_set_option "defer_lane_join": mc_process_frames_matched_dev returns without ordering the context's
	                                  // stream after the lanes; mc_join_lanes does that later (MATCH of the next chunk overlaps this chunk's stages)
	int lanes_pending = 0;            // lanes used since the last join
	int batch_stats[4] = {0, 0, 0, 0};   // {frames, accepted matches, objects, lanes used} of the last batch
	mc::DevBuf link_buf;                 // linkage clustering (linkage.cu): inputs, similarity matrices, agglomeration state
	void *sift_state = nullptr;          // feature extraction (sift.cu): scale-space plan and buffers, created on first use
};

namespace mc {

#define MC_CUDA(call)                                                                         \
	do {                                                                                      \
		cudaError_t e__ = (call);                                                             \
		if (e__ != cudaSuccess) {                                                             \
			ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__);                   \
			return MC_ERR_CUDA;                                                               \
		}                                                                                     \
	} while (0)

#define MC_LAUNCH_CHECK()                                                                     \
	do {                                                                                      \
		ctx->launches++;                                                                      \
		cudaError_t e__ = cudaGetLastError();                                                 \
		if (e__ != cudaSuccess) {                                                             \
			ctx->err = std::string("kernel launch: ") + cudaGetErrorString(e__);              \
			return MC_ERR_CUDA;                                                               \
		}                                                                                     \
	} while (0)

#define MC_TRY(expr)                                                                          \
	do {                                                                                      \
		mc_status s__ = (expr);                                                               \
		if (s__ != MC_OK) return s__;                                                         \
	} while (0)

inline mc_status reserve(mc_ctx *ctx, DevBuf &b, size_t bytes) {
	if (bytes <= b.cap && b.p) return MC_OK;
	if (ctx->capturing) { ctx->err = "scratch allocation requested during graph capture"; return MC_ERR_STATE; }
	if (b.p) { MC_CUDA(cudaFree(b.p)); b.p = nullptr; b.cap = 0; }
	size_t cap = bytes < 256 ? 256 : bytes + bytes / 4;
	MC_CUDA(cudaMalloc(&b.p, cap));
	b.cap = cap;
	return MC_OK;
}

inline mc_status pinned(mc_ctx *ctx, size_t bytes) {
	if (bytes <= ctx->h_pinned_cap) return MC_OK;
	if (ctx->h_pinned) { MC_CUDA(cudaFreeHost(ctx->h_pinned)); ctx->h_pinned = nullptr; ctx->h_pinned_cap = 0; }
	size_t cap = bytes + bytes / 4 + 4096;
	MC_CUDA(cudaMallocHost(&ctx->h_pinned, cap));
	ctx->h_pinned_cap = cap;
	return MC_OK;
}

// ---- stage entry points implemented in the .cu files (device pointers, async on ctx->stream) ----
mc_status db_build_images(mc_ctx *ctx);
mc_status match_configure_device(mc_ctx *ctx);
mc_status cluster_configure_device(mc_ctx *ctx);
mc_status sm_partition_create(mc_ctx *ctx, int stage_sms);
void sm_partition_destroy(mc_ctx *ctx);
mc_status sm_partition_stage_stream(mc_ctx *ctx, cudaStream_t *out, int priority);
void sift_free(mc_ctx *ctx);
mc_status sift_set_two_pass(mc_ctx *ctx, int on);
mc_status sift_set_gather(mc_ctx *ctx, int on);
mc_status sift_extract_device(mc_ctx *ctx, const uint8_t *d_gray, int B, int H, int W, int dbl, int max_kp,
                              float *d_xy, float *d_so, float *d_desc, int32_t *d_counts, int32_t *d_offsets, int match_normalise);
mc_status process_frames_host(mc_ctx *ctx, const float *d_q, const float *d_qxy, const int32_t *d_qimg, const int32_t *frame_offsets, int n_frames,
                              const mc_pipeline_params *P, int max_objects, int32_t *n_objects, int32_t *obj_model, float *obj_pose, float *obj_score,
                              int32_t *frame_info, float *stage_ms);
size_t frame_shard_slot_bytes(int Q, const mc_pipeline_params *P);
mc_status process_frame_sharded_device(mc_ctx *ctx, int phase, const int32_t *d_nn_row, const uint8_t *d_accepted, const float *d_qxy,
                                       const int32_t *d_qimg, int Q, const mc_pipeline_params *P, int shard_rank, int shard_world, void *d_exchange,
                                       int max_objects, int32_t *d_out_info, int32_t *d_out_model, float *d_out_pose, float *d_out_score);
mc_status match_device(mc_ctx *ctx, const float *d_q, int Q, float ratio, int mode,
                       int32_t *d_nn_row, float *d_nn_dist, uint8_t *d_accepted);
mc_status match_merge_device(mc_ctx *ctx, const int32_t *rows_all, const float *dist_all, size_t stride, int n_shards, int Q, float ratio,
                             int32_t *d_nn_row, float *d_nn_dist, uint8_t *d_accepted);

} // namespace mc
