// depth_pose.cuh — one RANSAC test of moped3d's depth-aware pose stages for a team of W lanes:
//   variant 0  POSE_RANSAC_LM_DIFF_BACKPROJECTION_DEPTH_CPU (moped3d/libmoped/src/pose/POSE_RANSAC_LM_DIFF_BACKPROJECTION_DEPTH_CPU.hpp:57-470,
//              the stage of moped3d's shipped pipeline, config.hpp:46): two residuals per correspondence;
//   variant 1  POSE_RANSAC_LM_DIFF_REPROJECTION_DEPTH_CPU (…/POSE_RANSAC_LM_DIFF_REPROJECTION_DEPTH_CPU.hpp:57-435): three.
// initPose (translation = mean world3D of the samples), optimizeCamera (lm_exact.cuh), testAllPoints (2-D reprojection,
// moped.hpp project()), refit on the consistent set in cluster order. Same host/device source rules as lm_exact.cuh.
#pragma once

#include "lm_exact.cuh"

namespace lmx {

struct Cam {               // layout of mc::Camera / mo_camera: K = (fx, fy, cx, cy), TM = 3x4 of cameraPose
	float K[4];
	float TM[12];
};

// the correspondences of one cluster (struct of arrays, cluster order) and the subset an LM problem runs on
struct Cluster {
	int n;
	const float *xy, *xyz, *world, *cauchy;   // n x 2, n x 3, n x 3 (depthData.coord3D), n (getCauchyWeight(fillDistance))
	const int32_t *image;
	const Cam *cams;
	float alpha;
};

// TransformMatrix::transform then Image::TM.inverseTransform (moped.hpp:183-200): model point -> camera frame
LMX_FN void to_camera(const float *T, const float *TM, const float *o, float *d) {
	const float x = o[0] * T[0] + o[1] * T[1] + o[2] * T[2] + T[3];
	const float y = o[0] * T[4] + o[1] * T[5] + o[2] * T[6] + T[7];
	const float z = o[0] * T[8] + o[1] * T[9] + o[2] * T[10] + T[11];
	const float a = x - TM[3], b = y - TM[7], c = z - TM[11];
	d[0] = a * TM[0] + b * TM[4] + c * TM[8];
	d[1] = a * TM[1] + b * TM[5] + c * TM[9];
	d[2] = a * TM[2] + b * TM[6] + c * TM[10];
}

template <int V> struct DepthResiduals;

// lmFuncQuat of the back-projection variant (:100-176): squared distance from the transformed model point to its
// projection pHat on the feature's viewing ray, squared distance from pHat to world3D; weights 1-(1-Alpha)w and (1-Alpha)w
template <> struct DepthResiduals<0> {
	static constexpr int R = 2;
	Cluster c;
	const int32_t *sel;            // correspondence k of the problem = cluster point sel[k]
	LMX_MEM void point(const float *T, int k, float *res) const {
		const int i = sel[k];
		float p3[3];
		to_camera(T, c.cams[c.image[i]].TM, c.xyz + 3 * i, p3);
		if (p3[2] < 0) {
			res[0] = -p3[2] + 10;
			res[1] = -p3[2] + 10;
		} else {
			const float *w3 = c.world + 3 * i;
			const float vx = w3[0], vy = w3[1], vz = w3[2];
			const float norm = sqrtf(vx * vx + vy * vy + vz * vz);
			const float nx = vx / norm, ny = vy / norm, nz = vz / norm;
			const float dot = nx * p3[0] + ny * p3[1] + nz * p3[2];
			const float hx = nx * dot, hy = ny * dot, hz = nz * dot;
			const float ax = p3[0] - hx, ay = p3[1] - hy, az = p3[2] - hz;
			const float dxy = sqrtf(ax * ax + ay * ay + az * az);
			const float bx = w3[0] - hx, by = w3[1] - hy, bz = w3[2] - hz;
			const float dz = sqrtf(bx * bx + by * by + bz * bz);
			res[0] = dxy * dxy;
			res[1] = dz * dz;
		}
		const float wi = c.cauchy[i];
		const float weight3D = (1 - c.alpha) * wi;
		const float weight2D = 1 - weight3D;
		res[0] *= weight2D;
		res[1] *= weight3D;
	}
};

// lmFuncQuat of the reprojection+depth variant (:150-215): squared pixel differences and 50 x the squared distance between
// p3D and (p3D . world3D) p3D, the latter also for points behind the camera
template <> struct DepthResiduals<1> {
	static constexpr int R = 3;
	Cluster c;
	const int32_t *sel;
	LMX_MEM void point(const float *T, int k, float *res) const {
		const int i = sel[k];
		const Cam &cam = c.cams[c.image[i]];
		float p3[3];
		to_camera(T, cam.TM, c.xyz + 3 * i, p3);
		const float u = p3[0] / p3[2] * cam.K[0] + cam.K[2];
		const float v = p3[1] / p3[2] * cam.K[1] + cam.K[3];
		if (p3[2] < 0) {
			res[0] = -p3[2] + 10;
			res[1] = -p3[2] + 10;
		} else {
			const float dx = u - c.xy[2 * i], dy = v - c.xy[2 * i + 1];
			res[0] = dx * dx;
			res[1] = dy * dy;
		}
		const float *w3 = c.world + 3 * i;
		const float vecTP = p3[0] * w3[0] + p3[1] * w3[1] + p3[2] * w3[2];
		const float pwx = p3[0] * vecTP, pwy = p3[1] * vecTP, pwz = p3[2] * vecTP;
		const float dx = p3[0] - pwx, dy = p3[1] - pwy, dz = p3[2] - pwz;
		const float depthError = sqrtf(dx * dx + dy * dy + dz * dz);
		res[2] = depthError * depthError;
		res[2] *= 50;
		const float wi = c.cauchy[i];
		const float weight3D = (1 - c.alpha) * wi;
		res[0] *= (1 - weight3D);
		res[1] *= (1 - weight3D);
		res[2] *= weight3D;
	}
};

// variant 2: lmFuncQuat of the moped2 stage POSE_RANSAC_LM_DIFF_REPROJECTION_CPU (moped2/libmoped/src/pose/
// POSE_RANSAC_LM_DIFF_REPROJECTION_CPU.hpp:100-138): squared pixel differences, (-z + 10) twice behind the camera. No depth
// inputs (Cluster::world / cauchy unused). With it the same order-preserving LM serves the moped2 POSE / POSE2 steps
// ("pose_exact_order"): results equal the strict-IEEE build of that stage bit for bit.
template <> struct DepthResiduals<2> {
	static constexpr int R = 2;
	Cluster c;
	const int32_t *sel;
	LMX_MEM void point(const float *T, int k, float *res) const {
		const int i = sel[k];
		const Cam &cam = c.cams[c.image[i]];
		float p3[3];
		to_camera(T, cam.TM, c.xyz + 3 * i, p3);
		const float u = p3[0] / p3[2] * cam.K[0] + cam.K[2];
		const float v = p3[1] / p3[2] * cam.K[1] + cam.K[3];
		if (p3[2] < 0) {
			res[0] = -p3[2] + 10;
			res[1] = -p3[2] + 10;
		} else {
			const float a = u - c.xy[2 * i], b = v - c.xy[2 * i + 1];
			res[0] = a * a;
			res[1] = b * b;
		}
	}
};

// optimizeCamera (:222-250): LM from `pose` on the selected correspondences; on success the pose is replaced (quaternion
// re-normalised) and ||e||^2 returned, on LM_ERROR the pose is untouched and -1 returned.
template <int V, int W>
LMX_FN float optimize_camera(const Team<W> &team, const Cluster &c, const int32_t *sel, int n_sel, float *pose, int itmax, float *scratch,
                             bool finite_check, const volatile int *stop_flag = nullptr, int my_index = 0) {
	DepthResiduals<V> fn;
	fn.c = c; fn.sel = sel;
	const Work w = work_carve(scratch, DepthResiduals<V>::R * n_sel);
	float p[M], err;
	for (int i = 0; i < M; i++) p[i] = pose[i];
	const int r = levmar_dif(team, fn, p, n_sel, itmax, w, finite_check, &err, stop_flag, my_index);
	if (r < 0) return (float)r;
	for (int i = 0; i < M; i++) pose[i] = p[i];
	quat_norm(pose);
	return err;
}

// testAllPoints (:252-266) + project() (moped.hpp:330-354): mask[i] = squared reprojection error < thr. Returns the count.
template <int W>
LMX_FN int test_all_points(const Team<W> &team, const Cluster &c, const float *pose, float thr, uint8_t *mask, int32_t *count_slot) {
	float T[12];
	tm_init(T, pose, pose + 4);
	team.each([&](int lane) {
		for (int i = lane; i < c.n; i += W) {
			const Cam &cam = c.cams[c.image[i]];
			float p3[3];
			to_camera(T, cam.TM, c.xyz + 3 * i, p3);
			float u = FLT_MAX, v = FLT_MAX;
			if (!(p3[2] < 0.001)) { u = p3[0] / p3[2] * cam.K[0] + cam.K[2]; v = p3[1] / p3[2] * cam.K[1] + cam.K[3]; }
			const float a = u - c.xy[2 * i], b = v - c.xy[2 * i + 1];
			const float err = a * a + b * b;
			mask[i] = err < thr;
		}
	});
	// the consistent set in cluster order -> sel-style list behind the mask is built by the caller; here only the count
	team.each([&](int lane) {
		if (lane == 0) {
			int k = 0;
			for (int i = 0; i < c.n; i++) k += mask[i];
			*count_slot = k;
		}
	});
	return *count_slot;
}

// Scratch of one team for a cluster of n points: LM work for R*n residuals + the selection list + a count word.
LMX_FN size_t hypothesis_scratch_floats(int n, int R) { return work_floats(R * n) + (size_t)n + 16; }

// One RANSAC test on an explicit (sample positions, initial quaternion): the loop body of RANSAC() (:283-312).
// Outputs like the oracle's mo_hypothesis_depth*: returns the inlier count of the sample fit or -1 (LM_ERROR);
// pose_lm = pose after the sample fit, pose_refit = pose after the refit on the inliers (= pose_lm if there were not
// more than min_npts of them), lm_err2 = (||e||^2 of the fit or -1, ||e||^2 of the refit or -1 / -2 = not run).
// `mask` (n bytes, team-visible) receives the inlier flags.
template <int V, int W>
LMX_FN int hypothesis(const Team<W> &team, const Cluster &c, const int32_t *sample_pos, int n_samples, const float *init_quat, int max_lm,
                      float err_thr, int min_npts, float *scratch, uint8_t *mask, bool finite_check, float *pose_lm, float *pose_refit,
                      float *lm_err2) {
	constexpr int R = DepthResiduals<V>::R;
	int32_t *sel = (int32_t *)(scratch + work_floats(R * c.n));
	int32_t *count_slot = sel + c.n;
	float pose[M] = { init_quat[0], init_quat[1], init_quat[2], init_quat[3], 0, 0, 0 };
	if (V == 2) {   // moped2's initPose (POSE_RANSAC_LM_DIFF_REPROJECTION_CPU.hpp:182-186)
		pose[6] = 0.5f;
	} else {        // initPose (:262-276): Pt += over the samples, then / n
		float sx = 0.f, sy = 0.f, sz = 0.f;
		for (int j = 0; j < n_samples; j++) { const float *w = c.world + 3 * sample_pos[j]; sx += w[0]; sy += w[1]; sz += w[2]; }
		pose[4] = sx / n_samples; pose[5] = sy / n_samples; pose[6] = sz / n_samples;
	}
	team.each([&](int lane) {
		for (int i = lane; i < c.n; i += W) mask[i] = 0;
		for (int j = lane; j < n_samples; j += W) sel[j] = sample_pos[j];
	});
	lm_err2[1] = -2;
	int ret = -1;
	const float r = optimize_camera<V>(team, c, sel, n_samples, pose, max_lm, scratch, finite_check);
	lm_err2[0] = r;
	if (r != -1.f) {
		for (int i = 0; i < M; i++) pose_lm[i] = pose[i];
		ret = test_all_points(team, c, pose, err_thr, mask, count_slot);
		if (ret > min_npts) {
			team.each([&](int lane) {
				if (lane == 0) {
					int k = 0;
					for (int i = 0; i < c.n; i++) if (mask[i]) sel[k++] = i;
				}
			});
			lm_err2[1] = optimize_camera<V>(team, c, sel, ret, pose, max_lm, scratch, finite_check);
		}
		for (int i = 0; i < M; i++) pose_refit[i] = pose[i];
	}
	return ret;
}

} // namespace lmx
