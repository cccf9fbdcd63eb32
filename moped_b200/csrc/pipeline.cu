// pipeline.cu — device-resident chaining of the stages: the FrameData fields that
// the reference keeps in STL containers between `process()` calls (matches, clusters, objects;
// moped2/libmoped/src/util.hpp:68-110) live here as flat device arrays, so that a frame goes
// MATCH -> CLUSTER -> POSE -> FILTER -> POSE2 -> FILTER2 (moped2/libmoped/src/config.hpp:83-120) with one
// host->device copy of the queries and one device->host copy of the objects.
#include "common.cuh"

namespace mc {

mc_status cluster_device(mc_ctx *ctx, const int32_t *d_match_offsets, const int32_t *d_match_image, const float *d_match_xy,
                         int n_models, int n_images, int max_matches, float radius, float merge, int min_pts, int max_iter,
                         int32_t *d_out_n, int32_t *d_cluster_model, int32_t *d_cluster_offsets, int32_t *d_members);
mc_status pose_ransac_device(mc_ctx *ctx, const int32_t *d_cluster_offsets, const int32_t *d_n_clusters, int n_clusters_cap,
                             const float *d_xy, const float *d_xyz, const int32_t *d_image, const int32_t *d_tie,
                             const mc_pose_params *pp, uint8_t *d_found, float *d_pose, int32_t *d_n_tests);
mc_status pose_append_device(mc_ctx *ctx, const int32_t *d_cluster_model, const int32_t *d_n_clusters, int n_clusters_cap, int max_obj,
                             const uint8_t *d_found, const float *d_pose, int32_t *d_n_obj, int obj_cap, int32_t *d_obj_model, float *d_obj_pose);
mc_status filter_device(mc_ctx *ctx, const int32_t *d_match_offsets, const int32_t *d_match_image, const float *d_match_xy,
                        const float *d_match_xyz, int n_models, int max_matches, const int32_t *d_obj_model, const float *d_obj_pose,
                        const int32_t *d_n_obj, int n_obj_cap, int min_points, float feat_dist, float min_score,
                        uint8_t *d_keep, float *d_score, int32_t *d_out_n, int32_t *d_cluster_model, int32_t *d_cluster_offsets,
                        int32_t *d_members, int32_t *d_surv_model, float *d_surv_pose, float *d_surv_score);

constexpr int kCompactCap = 8192;    // accepted matches per frame handled by the single-CTA compaction

// Accepted queries -> matches bucketed by model, query order preserved inside a model
// (`matches[correspModel[nx[0]]].push_back(...)` in query order, MATCH_ANN_CPU.hpp:165-176; the order matters:
// mean-shift is order dependent). Single CTA: stable compaction of the accepted queries, then a bitonic sort
// of (model << 13 | rank) keys in shared memory.
__global__ void __launch_bounds__(1024)
k_match_compact(const int32_t *__restrict__ nn_row, const uint8_t *__restrict__ accepted, int Q, int64_t row_base,
                const int32_t *__restrict__ model_of_row, const float *__restrict__ db_xyz, const float *__restrict__ q_xy,
                const int32_t *__restrict__ q_image, int n_models, int32_t *__restrict__ acc_list,
                int32_t *__restrict__ match_offsets, int32_t *__restrict__ match_query, int32_t *__restrict__ match_row,
                int32_t *__restrict__ match_image, float *__restrict__ match_xy, float *__restrict__ match_xyz,
                int32_t *__restrict__ status) {
	__shared__ uint32_t keys[kCompactCap];
	__shared__ int s_warp[32];
	__shared__ int s_base;
	const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
	if (tid == 0) s_base = 0;
	__syncthreads();
	for (int q0 = 0; q0 < Q; q0 += 1024) {
		const int q = q0 + tid;
		const bool acc = q < Q && accepted[q];
		const unsigned bal = __ballot_sync(0xffffffffu, acc);
		if (lane == 0) s_warp[w] = __popc(bal);
		__syncthreads();
		int off = s_base;
		for (int i = 0; i < w; i++) off += s_warp[i];
		const int rank = off + __popc(bal & ((1u << lane) - 1));
		if (acc && rank < kCompactCap) {
			acc_list[rank] = q;
			keys[rank] = ((uint32_t)model_of_row[nn_row[2 * q] - row_base] << 13) | (uint32_t)rank;
		}
		__syncthreads();
		if (tid == 0) { int t = 0; for (int i = 0; i < 32; i++) t += s_warp[i]; s_base += t; }
		__syncthreads();
	}
	int M = s_base;
	if (M > kCompactCap) { if (tid == 0) status[0] = 1; M = kCompactCap; }
	int P = 1;
	while (P < M) P <<= 1;
	for (int i = M + tid; i < P; i += 1024) keys[i] = 0xFFFFFFFFu;
	__syncthreads();
	for (int k = 2; k <= P; k <<= 1)
		for (int j = k >> 1; j > 0; j >>= 1) {
			for (int i = tid; i < P; i += 1024) {
				const int ixj = i ^ j;
				if (ixj > i) {
					const uint32_t a = keys[i], b = keys[ixj];
					const bool up = (i & k) == 0;
					if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
				}
			}
			__syncthreads();
		}
	for (int i = tid; i < M; i += 1024) {
		const uint32_t key = keys[i];
		const int m = (int)(key >> 13), q = acc_list[key & 8191u];
		const int row = nn_row[2 * q];
		match_query[i] = q; match_row[i] = row;
		match_image[i] = q_image[q];
		match_xy[2 * i] = q_xy[2 * q]; match_xy[2 * i + 1] = q_xy[2 * q + 1];
		const int64_t lr = (int64_t)row - row_base;
		match_xyz[3 * i] = db_xyz[3 * lr]; match_xyz[3 * i + 1] = db_xyz[3 * lr + 1]; match_xyz[3 * i + 2] = db_xyz[3 * lr + 2];
		const int mprev = i > 0 ? (int)(keys[i - 1] >> 13) : -1;
		for (int mm = mprev + 1; mm <= m; mm++) match_offsets[mm] = i;
	}
	const int mlast = M > 0 ? (int)(keys[M - 1] >> 13) : -1;
	for (int mm = mlast + 1 + tid; mm <= n_models; mm += 1024) match_offsets[mm] = M;
}

// cluster members -> contiguous per-cluster point arrays (what preprocessAllMatches + `cl` build,
// POSE_..._CPU.hpp:213-237,288-290). tie = match index inside the model (the reference's pointer order).
__global__ void k_gather_points(const int32_t *__restrict__ n_p /* {#clusters, #members} */, const int32_t *__restrict__ cluster_model,
                                const int32_t *__restrict__ cluster_offsets, const int32_t *__restrict__ members,
                                const int32_t *__restrict__ match_offsets, const int32_t *__restrict__ match_image,
                                const float *__restrict__ match_xy, const float *__restrict__ match_xyz,
                                float *__restrict__ pt_xy, float *__restrict__ pt_xyz, int32_t *__restrict__ pt_image, int32_t *__restrict__ pt_tie) {
	const int n_clusters = n_p[0];
	for (int c = blockIdx.x; c < n_clusters; c += gridDim.x) {
		const int lo = match_offsets[cluster_model[c]];
		for (int t = cluster_offsets[c] + threadIdx.x; t < cluster_offsets[c + 1]; t += blockDim.x) {
			const int mi = members[t], j = lo + mi;
			pt_xy[2 * t] = match_xy[2 * j]; pt_xy[2 * t + 1] = match_xy[2 * j + 1];
			pt_xyz[3 * t] = match_xyz[3 * j]; pt_xyz[3 * t + 1] = match_xyz[3 * j + 1]; pt_xyz[3 * t + 2] = match_xyz[3 * j + 2];
			pt_image[t] = match_image[j];
			pt_tie[t] = mi;
		}
	}
}

__global__ void k_copy_objects(const int32_t *__restrict__ n_p, const int32_t *__restrict__ src_model, const float *__restrict__ src_pose,
                               int32_t *__restrict__ dst_n, int32_t *__restrict__ dst_model, float *__restrict__ dst_pose) {
	const int n = n_p[0];
	for (int i = threadIdx.x; i < n; i += blockDim.x) {
		dst_model[i] = src_model[i];
		for (int j = 0; j < 7; j++) dst_pose[7 * i + j] = src_pose[7 * i + j];
	}
	if (threadIdx.x == 0) *dst_n = n;
}

struct FrameBufs {
	int32_t *nn_row; float *nn_dist; uint8_t *accepted;
	int32_t *acc_list, *match_offsets, *match_query, *match_row, *match_image; float *match_xy, *match_xyz;
	int32_t *status, *cl_n, *cl_model, *cl_offsets, *cl_members;
	float *pt_xy, *pt_xyz; int32_t *pt_image, *pt_tie;
	uint8_t *found; float *task_pose; int32_t *n_tests;
	int32_t *n_obj, *obj_model; float *obj_pose, *obj_score; uint8_t *keep;
	int32_t *f_n, *f_model, *f_offsets, *f_members, *surv_model; float *surv_pose, *surv_score;
};

static mc_status carve(mc_ctx *ctx, FrameBufs &B, int Q, int n_models, int cl_cap, int task_cap, int obj_cap) {
	size_t off = 0;
	auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
	const size_t o_nn_row = take(8ull * Q), o_nn_dist = take(8ull * Q), o_acc = take(Q);
	const size_t o_acc_list = take(4ull * Q), o_moff = take(4ull * (n_models + 2)), o_mq = take(4ull * Q), o_mr = take(4ull * Q),
	             o_mi = take(4ull * Q), o_mxy = take(8ull * Q), o_mxyz = take(12ull * Q);
	const size_t o_status = take(64), o_cln = take(64), o_clm = take(4ull * (Q + 2)), o_clo = take(4ull * (Q + 2)), o_clmem = take(4ull * (Q + 2));
	const size_t o_pxy = take(8ull * Q), o_pxyz = take(12ull * Q), o_pim = take(4ull * Q), o_ptie = take(4ull * Q);
	const size_t o_found = take(task_cap), o_tpose = take(28ull * task_cap), o_ntests = take(4ull * task_cap);
	const size_t o_nobj = take(64), o_om = take(4ull * obj_cap), o_op = take(28ull * obj_cap), o_os = take(4ull * obj_cap), o_keep = take(obj_cap);
	const size_t o_fn = take(64), o_fm = take(4ull * (obj_cap + 2)), o_fo = take(4ull * (obj_cap + 2)), o_fmem = take(4ull * (Q + 2)),
	             o_sm = take(4ull * obj_cap), o_sp = take(28ull * obj_cap), o_ss = take(4ull * obj_cap);
	MC_TRY(reserve(ctx, ctx->scratch[10], off));
	char *b = (char *)ctx->scratch[10].p;
	B.nn_row = (int32_t *)(b + o_nn_row); B.nn_dist = (float *)(b + o_nn_dist); B.accepted = (uint8_t *)(b + o_acc);
	B.acc_list = (int32_t *)(b + o_acc_list); B.match_offsets = (int32_t *)(b + o_moff); B.match_query = (int32_t *)(b + o_mq);
	B.match_row = (int32_t *)(b + o_mr); B.match_image = (int32_t *)(b + o_mi); B.match_xy = (float *)(b + o_mxy); B.match_xyz = (float *)(b + o_mxyz);
	B.status = (int32_t *)(b + o_status); B.cl_n = (int32_t *)(b + o_cln); B.cl_model = (int32_t *)(b + o_clm); B.cl_offsets = (int32_t *)(b + o_clo);
	B.cl_members = (int32_t *)(b + o_clmem);
	B.pt_xy = (float *)(b + o_pxy); B.pt_xyz = (float *)(b + o_pxyz); B.pt_image = (int32_t *)(b + o_pim); B.pt_tie = (int32_t *)(b + o_ptie);
	B.found = (uint8_t *)(b + o_found); B.task_pose = (float *)(b + o_tpose); B.n_tests = (int32_t *)(b + o_ntests);
	B.n_obj = (int32_t *)(b + o_nobj); B.obj_model = (int32_t *)(b + o_om); B.obj_pose = (float *)(b + o_op); B.obj_score = (float *)(b + o_os);
	B.keep = (uint8_t *)(b + o_keep);
	B.f_n = (int32_t *)(b + o_fn); B.f_model = (int32_t *)(b + o_fm); B.f_offsets = (int32_t *)(b + o_fo); B.f_members = (int32_t *)(b + o_fmem);
	B.surv_model = (int32_t *)(b + o_sm); B.surv_pose = (float *)(b + o_sp); B.surv_score = (float *)(b + o_ss);
	return MC_OK;
}

// One frame, queries resident on the device. Results are copied to the host at the end (one sync).
mc_status process_frame_device(mc_ctx *ctx, const float *d_q, const float *d_qxy, const int32_t *d_qimg, int Q, const mc_pipeline_params *P,
                               int max_objects, int32_t *n_objects, int32_t *obj_model, float *obj_pose, float *obj_score, float *stage_ms,
                               const int32_t *d_nn_row_in, const uint8_t *d_accepted_in) {
	if (!ctx->d_db) { ctx->err = "process_frame: no database uploaded"; return MC_ERR_STATE; }
	if (!ctx->d_cams) { ctx->err = "process_frame: cameras not set"; return MC_ERR_STATE; }
	if (Q <= 0) { *n_objects = 0; return MC_OK; }
	const int min_pts = P->cluster_min_pts > 0 ? P->cluster_min_pts : 1;
	const int cl_cap = Q / min_pts + 1;
	const int max_obj_per = P->pose.max_objects_per_cluster > P->pose2.max_objects_per_cluster ? P->pose.max_objects_per_cluster : P->pose2.max_objects_per_cluster;
	const int obj_cap = 2 * cl_cap * max_obj_per + 8;
	const int task_cap = obj_cap * max_obj_per + 8;
	FrameBufs B;
	MC_TRY(carve(ctx, B, Q, ctx->n_models, cl_cap, task_cap, obj_cap));
	cudaEvent_t ev[7];
	if (stage_ms) for (int i = 0; i < 7; i++) MC_CUDA(cudaEventCreate(&ev[i]));
	auto mark = [&](int i) { if (stage_ms) cudaEventRecord(ev[i], ctx->stream); };
	MC_CUDA(cudaMemsetAsync(B.status, 0, 64, ctx->stream));
	MC_CUDA(cudaMemsetAsync(B.n_obj, 0, 64, ctx->stream));
	mark(0);
	// MATCH (skipped when the caller already holds merged nearest neighbours of an object-sharded database)
	const int32_t *nn_row = d_nn_row_in ? d_nn_row_in : B.nn_row;
	const uint8_t *accepted = d_accepted_in ? d_accepted_in : B.accepted;
	if (!d_nn_row_in) MC_TRY(match_device(ctx, d_q, Q, P->match_ratio, P->match_mode, B.nn_row, B.nn_dist, B.accepted));
	k_match_compact<<<1, 1024, 0, ctx->stream>>>(nn_row, accepted, Q, ctx->table_base, ctx->d_model_of_row, ctx->d_xyz, d_qxy, d_qimg, ctx->n_models,
	                                            B.acc_list, B.match_offsets, B.match_query, B.match_row, B.match_image, B.match_xy, B.match_xyz, B.status);
	MC_LAUNCH_CHECK();
	mark(1);
	// CLUSTER
	MC_TRY(cluster_device(ctx, B.match_offsets, B.match_image, B.match_xy, ctx->n_models, ctx->n_images, Q, P->cluster_radius, P->cluster_merge,
	                      P->cluster_min_pts, P->cluster_max_iterations, B.cl_n, B.cl_model, B.cl_offsets, B.cl_members));
	mark(2);
	// POSE
	k_gather_points<<<ctx->num_sms, 128, 0, ctx->stream>>>(B.cl_n, B.cl_model, B.cl_offsets, B.cl_members, B.match_offsets, B.match_image, B.match_xy,
	                                                     B.match_xyz, B.pt_xy, B.pt_xyz, B.pt_image, B.pt_tie);
	MC_LAUNCH_CHECK();
	MC_TRY(pose_ransac_device(ctx, B.cl_offsets, B.cl_n, cl_cap, B.pt_xy, B.pt_xyz, B.pt_image, B.pt_tie, &P->pose, B.found, B.task_pose, B.n_tests));
	MC_TRY(pose_append_device(ctx, B.cl_model, B.cl_n, cl_cap, P->pose.max_objects_per_cluster, B.found, B.task_pose, B.n_obj, obj_cap, B.obj_model, B.obj_pose));
	mark(3);
	// FILTER
	MC_TRY(filter_device(ctx, B.match_offsets, B.match_image, B.match_xy, B.match_xyz, ctx->n_models, Q, B.obj_model, B.obj_pose, B.n_obj, obj_cap,
	                     P->filter_min_points, P->filter_feature_distance, P->filter_min_score, B.keep, B.obj_score, B.f_n, B.f_model, B.f_offsets,
	                     B.f_members, B.surv_model, B.surv_pose, B.surv_score));
	mark(4);
	// POSE2: the surviving objects stay in the list, new hypotheses from the rebuilt clusters are appended (:295-303)
	k_copy_objects<<<1, 256, 0, ctx->stream>>>(B.f_n, B.surv_model, B.surv_pose, B.n_obj, B.obj_model, B.obj_pose);
	MC_LAUNCH_CHECK();
	k_gather_points<<<ctx->num_sms, 128, 0, ctx->stream>>>(B.f_n, B.f_model, B.f_offsets, B.f_members, B.match_offsets, B.match_image, B.match_xy,
	                                                     B.match_xyz, B.pt_xy, B.pt_xyz, B.pt_image, B.pt_tie);
	MC_LAUNCH_CHECK();
	const int cl2_cap = obj_cap / 2;
	MC_TRY(pose_ransac_device(ctx, B.f_offsets, B.f_n, cl2_cap, B.pt_xy, B.pt_xyz, B.pt_image, B.pt_tie, &P->pose2, B.found, B.task_pose, B.n_tests));
	MC_TRY(pose_append_device(ctx, B.f_model, B.f_n, cl2_cap, P->pose2.max_objects_per_cluster, B.found, B.task_pose, B.n_obj, obj_cap, B.obj_model, B.obj_pose));
	mark(5);
	// FILTER2
	MC_TRY(filter_device(ctx, B.match_offsets, B.match_image, B.match_xy, B.match_xyz, ctx->n_models, Q, B.obj_model, B.obj_pose, B.n_obj, obj_cap,
	                     P->filter2_min_points, P->filter2_feature_distance, P->filter2_min_score, B.keep, B.obj_score, B.f_n, B.f_model, B.f_offsets,
	                     B.f_members, B.surv_model, B.surv_pose, B.surv_score));
	mark(6);
	// results -> host
	const size_t bytes = 64 + (size_t)max_objects * (4 + 28 + 4);
	MC_TRY(pinned(ctx, bytes + 64));
	char *hp = (char *)ctx->h_pinned;
	int32_t *h_n = (int32_t *)hp; int32_t *h_status = (int32_t *)(hp + 32);
	int32_t *h_model = (int32_t *)(hp + 64); float *h_pose = (float *)(hp + 64 + 4ull * max_objects); float *h_score = (float *)(hp + 64 + 32ull * max_objects);
	const int ncopy = max_objects < obj_cap ? max_objects : obj_cap;
	MC_CUDA(cudaMemcpyAsync(h_n, B.f_n, 8, cudaMemcpyDeviceToHost, ctx->stream));
	MC_CUDA(cudaMemcpyAsync(h_status, B.status, 4, cudaMemcpyDeviceToHost, ctx->stream));
	MC_CUDA(cudaMemcpyAsync(h_model, B.surv_model, 4ull * ncopy, cudaMemcpyDeviceToHost, ctx->stream));
	MC_CUDA(cudaMemcpyAsync(h_pose, B.surv_pose, 28ull * ncopy, cudaMemcpyDeviceToHost, ctx->stream));
	MC_CUDA(cudaMemcpyAsync(h_score, B.surv_score, 4ull * ncopy, cudaMemcpyDeviceToHost, ctx->stream));
	MC_CUDA(cudaStreamSynchronize(ctx->stream));
	if (stage_ms) {
		for (int i = 0; i < 6; i++) { cudaEventElapsedTime(&stage_ms[i], ev[i], ev[i + 1]); }
		for (int i = 0; i < 7; i++) cudaEventDestroy(ev[i]);
	}
	if (*h_status) { ctx->err = "process_frame: more than 8192 accepted matches in one frame"; return MC_ERR_CAPACITY; }
	int n = h_n[0];
	if (n > ncopy) { ctx->err = "process_frame: object buffer too small"; *n_objects = n; return MC_ERR_CAPACITY; }
	*n_objects = n;
	memcpy(obj_model, h_model, 4ull * n);
	memcpy(obj_pose, h_pose, 28ull * n);
	memcpy(obj_score, h_score, 4ull * n);
	return MC_OK;
}

} // namespace mc
