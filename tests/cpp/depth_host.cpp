// tests/cpp/depth_host.cpp — TEST INFRASTRUCTURE: the device source of the order-preserving LM and of the depth-aware
// RANSAC test (moped_b200/csrc/lm_exact.cuh, depth_pose.cuh) compiled by g++ with the lanes of a team emulated by a loop,
// so that its arithmetic and its split into phases can be checked against the oracle on a machine without a GPU
// (tests/test_depth_pose_host.py). Built with -ffp-contract=off (the .cu is built with -fmad=false). Not part of the product:
// nothing under moped_b200/ builds or loads this file.
#include "../../moped_b200/csrc/depth_pose.cuh"

#include <vector>
#include <xmmintrin.h>

namespace {
template <int V, int W>
int run(const lmx::Cluster &c, const int32_t *sample_pos, int n_samples, const float *init_quat, int max_lm, float err_thr, int min_npts,
        bool finite_check, float *pose_lm, float *pose_refit, float *lm_err2, uint8_t *mask) {
	std::vector<float> scratch(lmx::hypothesis_scratch_floats(c.n, lmx::DepthResiduals<V>::R));
	lmx::Team<W> team;
	team.init(0);
	return lmx::hypothesis<V, W>(team, c, sample_pos, n_samples, init_quat, max_lm, err_thr, min_npts, scratch.data(), mask, finite_check, pose_lm,
	                             pose_refit, lm_err2);
}
}

extern "C" int dh_hypothesis(int variant, int width, int lane_order, int n, const float *xy, const float *xyz, const float *world,
                             const float *cauchy, const int32_t *image, const float *cams16, float alpha, const int32_t *sample_pos,
                             int n_samples, const float *init_quat, int max_lm, float err_thr, int min_npts, int finite_check, float *pose_lm,
                             float *pose_refit, float *lm_err2, uint8_t *mask) {
	lmx::Cluster c;
	c.n = n; c.xy = xy; c.xyz = xyz; c.world = world; c.cauchy = cauchy; c.image = image;
	c.cams = reinterpret_cast<const lmx::Cam *>(cams16); c.alpha = alpha;
	phx::g_host_order = lane_order;
	const unsigned csr = _mm_getcsr();
	_mm_setcsr(csr | 0x8040u);                       // FTZ | DAZ: the .cu is built with -ftz=true, the reference process runs that way
	int r = -2;
	const bool fc = finite_check != 0;
#define DH_RUN(V, W) r = run<V, W>(c, sample_pos, n_samples, init_quat, max_lm, err_thr, min_npts, fc, pose_lm, pose_refit, lm_err2, mask)
	if (variant == 0 && width == 1) DH_RUN(0, 1);
	else if (variant == 0 && width == 32) DH_RUN(0, 32);
	else if (variant == 1 && width == 1) DH_RUN(1, 1);
	else if (variant == 1 && width == 32) DH_RUN(1, 32);
	else if (variant == 2 && width == 1) DH_RUN(2, 1);
	else if (variant == 2 && width == 32) DH_RUN(2, 32);
	else if (variant == 0 && width == 8) DH_RUN(0, 8);
	else if (variant == 1 && width == 8) DH_RUN(1, 8);
	else if (variant == 2 && width == 8) DH_RUN(2, 8);
#undef DH_RUN
	_mm_setcsr(csr);
	return r;
}
