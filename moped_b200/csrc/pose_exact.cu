// pose_exact.cu — the POSE / POSE2 steps in EXACT-ORDER mode (mc_set_option "pose_exact_order"): the staged RANSAC driver of
// pose_staged.cuh instantiated with the order-preserving Levenberg-Marquardt of lm_exact.cuh and the moped2 residual
// (depth_pose.cuh, variant 2 = lmFuncQuat of POSE_RANSAC_LM_DIFF_REPROJECTION_CPU.hpp:100-138).
//
// Compiled with -fmad=false -ftz=true like pose_depth.cu: every product and sum rounds separately and every reduction runs in
// levmar's order, so a RANSAC task returns what the strict-IEEE build of the reference returns on the same sample stream, bit for
// bit — the same winning test, the same refitted pose (tests/test_gpu_depth_pose.py, tests/test_gpu_stages.py) — while keeping the
// staged driver's shape (first hypotheses of all tasks packed four to a warp, escalating levels, one warp per refit).
//   sample fit   8-lane team, the LM work arrays (5 vectors + Jacobian of at most 16 residual rows) in SHARED memory
//   inlier count each lane a strided share of the cluster, integer group sum (no mask is kept: the refit rebuilds the set)
//   refit        a warp, inliers compacted in cluster order, LM work arrays in a per-task slice of global scratch (L1/L2 resident);
//                no cap on the number of inliers
#include "common.cuh"
#include "ransac_sample.cuh"
#include "depth_pose.cuh"
#include "pose_staged.cuh"

namespace mc {

static_assert(sizeof(lmx::Cam) == sizeof(Camera), "lmx::Cam must mirror mc::Camera");

struct ExactFit {
	// sample fit: Work for R * kMaxAlign = 16 residual rows + the selection list
	static constexpr int kFitRows = 2 * kMaxAlign;
	static constexpr int kFitSmemFloats = kFitRows * (4 + lmx::M) + 64 + kMaxAlign;
	static constexpr int kRefitListInts = 1;
	// refit of a cluster of n points: Work for 2n rows (22 n + 64 floats) + the inlier list (n ints), rounded up
	static constexpr int kRefitFloatsPerPoint = 2 * (4 + lmx::M) + 1, kRefitFloatsPerTask = 96;

	static __device__ __forceinline__ lmx::Cluster view(int n, const float *xy, const float *xyz, const int32_t *image, const Camera *cams) {
		lmx::Cluster c;
		c.n = n; c.xy = xy; c.xyz = xyz; c.world = nullptr; c.cauchy = nullptr; c.image = image;
		c.cams = reinterpret_cast<const lmx::Cam *>(cams); c.alpha = 0.f;
		return c;
	}

	// squared reprojection error < thr, the arithmetic of testAllPoints / project() (lmx::test_all_points)
	static __device__ __forceinline__ bool inlier(const float *T, const lmx::Cluster &c, int i, float thr) {
		const lmx::Cam &cam = c.cams[c.image[i]];
		float p3[3];
		lmx::to_camera(T, cam.TM, c.xyz + 3 * i, p3);
		float u = FLT_MAX, v = FLT_MAX;
		if (!(p3[2] < 0.001)) { u = p3[0] / p3[2] * cam.K[0] + cam.K[2]; v = p3[1] / p3[2] * cam.K[1] + cam.K[3]; }
		const float a = u - c.xy[2 * i], b = v - c.xy[2 * i + 1];
		return a * a + b * b < thr;
	}

	static __device__ int fit(const int (&pos)[kMaxAlign], int n_align, const float (&quat)[4], int n, const float *xy, const float *xyz,
	                          const int32_t *image, const Camera *cams, int max_lm, float thr, unsigned mask, int lig, float (&pose)[7], float &err,
	                          const volatile int *abort_if_below, int my_index, float *smem) {
		lmx::Team<8> team;
		team.init(threadIdx.x & 31);
		const lmx::Cluster c = view(n, xy, xyz, image, cams);
		int32_t *sel = reinterpret_cast<int32_t *>(smem + kFitSmemFloats - kMaxAlign);
		int mine = 0;
#pragma unroll
		for (int j = 0; j < kMaxAlign; j++) if (j == lig) mine = pos[j];
		team.each([&](int l) { if (l < n_align) sel[l] = mine; });
		pose[0] = quat[0]; pose[1] = quat[1]; pose[2] = quat[2]; pose[3] = quat[3];
		pose[4] = 0.f; pose[5] = 0.f; pose[6] = 0.5f;                                   // initPose (:182-186)
		err = lmx::optimize_camera<2, 8>(team, c, sel, n_align, pose, max_lm, smem, false, abort_if_below, my_index);
		if (err == -1.f) return -1;
		float T[12];
		lmx::tm_init(T, pose, pose + 4);
		int cnt = 0;
		for (int i = lig; i < n; i += 8) cnt += inlier(T, c, i, thr) ? 1 : 0;
#pragma unroll
		for (int o = 4; o; o >>= 1) cnt += __shfl_xor_sync(mask, cnt, o);
		return cnt;
	}

	static __device__ bool refit(float (&pose)[7], int n, const float *xy, const float *xyz, const int32_t *image, const Camera *cams, float thr,
	                             int max_lm, int lane, int *, float *scratch, float &err) {
		lmx::Team<32> team;
		team.init(lane);
		const lmx::Cluster c = view(n, xy, xyz, image, cams);
		int32_t *sel = reinterpret_cast<int32_t *>(scratch + lmx::work_floats(2 * n));
		float T[12];
		lmx::tm_init(T, pose, pose + 4);
		int n_inl = 0;                                         // `consistent` in cluster order (:172-178)
		for (int i0 = 0; i0 < n; i0 += 32) {
			const int i = i0 + lane;
			const bool in = i < n && inlier(T, c, i, thr);
			const unsigned m = __ballot_sync(0xffffffffu, in);
			if (in) sel[n_inl + __popc(m & ((1u << lane) - 1))] = i;
			n_inl += __popc(m);
		}
		__syncwarp();
		err = lmx::optimize_camera<2, 32>(team, c, sel, n_inl, pose, max_lm, scratch, false);
		return err != -1.f;
	}
};

mc_status pose_ransac_exact_device(mc_ctx *ctx, const int32_t *d_cluster_offsets, const int32_t *d_n_clusters, int n_clusters_cap, int n_points_cap,
                                   const float *d_xy, const float *d_xyz, const int32_t *d_image, const int32_t *d_tie,
                                   const mc_pose_params *pp, uint8_t *d_found, float *d_pose, int32_t *d_n_tests) {
	return ransac_staged_launch<ExactFit>(ctx, d_cluster_offsets, d_n_clusters, n_clusters_cap, n_points_cap, d_xy, d_xyz, d_image, d_tie, pp, d_found,
	                                      d_pose, d_n_tests);
}

} // namespace mc
