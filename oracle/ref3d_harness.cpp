/*
 * ref3d_harness.cpp — TEST INFRASTRUCTURE (never part of the product): flat-array glue around moped3d's depth-aware pose
 * stage, POSE_RANSAC_LM_DIFF_BACKPROJECTION_DEPTH_CPU (moped3d/libmoped/src/pose/…:57-437), compiled UNMODIFIED from
 * /root/reference together with the vendored levmar 2.4 into oracle/_ref/libmoped3d_ref.so by oracle/Makefile (target
 * ref3d). It pins the C restatement of the stage (oracle/moped_oracle.c: mo_*_depth) — the checker of SURVEY.md §8f row 4's
 * first component. A separate library because moped3d's MopedNS (FrameData::Match with depthData, typed Image) clashes with
 * moped2's.
 *
 * Same shims as ref_harness.cpp: `#define class struct` (private members per hypothesis), `#define rand moped3d_ref_rand`
 * (seedable LCG instead of libc's global rand()), FTZ|DAZ while inside reference code (the executable is -ffast-math).
 */
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <cfloat>
#include <stdint.h>
#include <xmmintrin.h>

#include <moped.hpp>
#include <util.hpp>
#include <lm.h>

static __thread uint64_t g_rng_state = 0x9E3779B97F4A7C15ULL;
extern "C" int moped3d_ref_rand(void) {
	g_rng_state = g_rng_state * 6364136223846793005ULL + 1442695040888963407ULL;
	return (int)((g_rng_state >> 33) & 0x7fffffffULL);
}
extern "C" void ref3d_srand(uint64_t seed) { g_rng_state = seed; }

#ifndef MAX_THREADS
#define MAX_THREADS 64
#endif

#define rand moped3d_ref_rand
#define class struct
#include <pose/POSE_RANSAC_LM_DIFF_BACKPROJECTION_DEPTH_CPU.hpp>
#include <pose/POSE_RANSAC_LM_DIFF_REPROJECTION_DEPTH_CPU.hpp>
#include <cluster/CLUSTER_LINKAGE_CPU.hpp>
#include <filter/FILTER_PROJECTION_DEPTH_CPU.hpp>
#include <iostream>
#undef class
#undef rand

/* moped3d's benchmark log (MopedBench.cpp / Benchmark.cpp, compiled as they are): the parity-log format of SURVEY.md 8f row 4 */
#include <unistd.h>
#include <Benchmark.cpp>
#include <MopedBench.cpp>

using namespace MopedNS;
typedef POSE_RANSAC_LM_DIFF_BACKPROJECTION_DEPTH_CPU POSE3D_T;      /* variant 0: the stage of moped3d's shipped pipeline (config.hpp:46,48) */
typedef POSE_RANSAC_LM_DIFF_REPROJECTION_DEPTH_CPU POSE3R_T;        /* variant 1: three residuals per correspondence, Cauchy scale 25 */

struct FtzGuard {
	unsigned saved;
	FtzGuard() { saved = _mm_getcsr(); _mm_setcsr(saved | 0x8040u); }
	~FtzGuard() { _mm_setcsr(saved); }
};

/* one camera (every correspondence of a call is seen by it) + the LmData records of n correspondences */
struct Cluster3D {
	Image image;
	vector<POSE3D_T::LmData> data;
	vector<POSE3D_T::LmData *> ptrs;
	Cluster3D(POSE3D_T &alg, int n, const float *xy, const float *xyz, const float *world, const float *fill, const float *K4, const float *cam_pose7)
	: image(IMAGE_TYPE_GRAY_IMAGE) {
		image.intrinsicLinearCalibration.init(K4[0], K4[1], K4[2], K4[3]);
		image.cameraPose.rotation.init(cam_pose7[0], cam_pose7[1], cam_pose7[2], cam_pose7[3]);
		image.cameraPose.translation.init(cam_pose7[4], cam_pose7[5], cam_pose7[6]);
		image.TM.init(image.cameraPose);
		data.resize(n);
		for (int i = 0; i < n; i++) {                       /* preprocessAllMatches, :330-355 */
			data[i].image = &image;
			data[i].coord2D.init(xy[2 * i], xy[2 * i + 1]);
			data[i].coord3D.init(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
			data[i].world3D.init(world[3 * i], world[3 * i + 1], world[3 * i + 2]);
			data[i].cauchyWeight = alg.getCauchyWeight(fill[i]);
			ptrs.push_back(&data[i]);
		}
	}
};

template <class ALG> struct ClusterT {
	Image image;
	vector<typename ALG::LmData> data;
	vector<typename ALG::LmData *> ptrs;
	ClusterT(ALG &alg, int n, const float *xy, const float *xyz, const float *world, const float *fill, const float *K4, const float *cam_pose7)
	: image(IMAGE_TYPE_GRAY_IMAGE) {
		image.intrinsicLinearCalibration.init(K4[0], K4[1], K4[2], K4[3]);
		image.cameraPose.rotation.init(cam_pose7[0], cam_pose7[1], cam_pose7[2], cam_pose7[3]);
		image.cameraPose.translation.init(cam_pose7[4], cam_pose7[5], cam_pose7[6]);
		image.TM.init(image.cameraPose);
		data.resize(n);
		for (int i = 0; i < n; i++) {
			data[i].image = &image;
			data[i].coord2D.init(xy[2 * i], xy[2 * i + 1]);
			data[i].coord3D.init(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
			data[i].world3D.init(world[3 * i], world[3 * i + 1], world[3 * i + 2]);
			data[i].cauchyWeight = alg.getCauchyWeight(fill[i]);
			ptrs.push_back(&data[i]);
		}
	}
};

template <class ALG>
static int hypothesis_t(int per_item, int n, const float *xy, const float *xyz, const float *world, const float *fill, const float *K4, const float *cam_pose7,
                        float alpha, const int *sample_pos, int n_samples, const float *init_quat, int maxLM, float errThr, int minNPts,
                        float *pose_init, float *pose_lm, float *pose_refit, float *lm_err, unsigned char *inlier_mask) {
	ALG alg(1, maxLM, 1, n_samples, minNPts, errThr, alpha);
	ClusterT<ALG> cl(alg, n, xy, xyz, world, fill, K4, cam_pose7);
	vector<typename ALG::LmData *> samples;
	for (int j = 0; j < n_samples; j++) samples.push_back(cl.ptrs[sample_pos[j]]);
	Pose pose;
	alg.initPose(pose, samples);
	pose.rotation.init(init_quat[0], init_quat[1], init_quat[2], init_quat[3]);
	for (int j = 0; j < 7; j++) pose_init[j] = pose[j];
	Float r = alg.optimizeCamera(pose, samples, maxLM);
	lm_err[0] = r; lm_err[1] = -2;
	for (int i = 0; i < n; i++) inlier_mask[i] = 0;
	if ((int)r == -1) return -1;
	for (int j = 0; j < 7; j++) pose_lm[j] = pose[j];
	vector<typename ALG::LmData *> consistent;
	alg.testAllPoints(consistent, pose, cl.ptrs, errThr);
	for (size_t k = 0; k < consistent.size(); k++) inlier_mask[consistent[k] - &cl.data[0]] = 1;
	if ((int)consistent.size() > minNPts) lm_err[1] = alg.optimizeCamera(pose, consistent, maxLM);
	for (int j = 0; j < 7; j++) pose_refit[j] = pose[j];
	return (int)consistent.size();
}

extern "C" {

/* the second depth pose variant, POSE_RANSAC_LM_DIFF_REPROJECTION_DEPTH_CPU (…:57-435): lmFuncQuat (:106-213, 3 residuals per
 * correspondence), one explicit hypothesis, whole RANSAC */
float ref3d_cauchy_weight_v1(float fill_distance) {
	POSE3R_T alg(1, 1, 1, 5, 6, 8, 0.5);
	return alg.getCauchyWeight(fill_distance);
}
void ref3d_lm_func_v1(const float *pose7, int n, const float *xy, const float *xyz, const float *world, const float *fill, const float *K4,
                      const float *cam_pose7, float alpha, float *errors) {
	FtzGuard g;
	POSE3R_T alg(1, 1, 1, 5, 6, 8, alpha);
	ClusterT<POSE3R_T> cl(alg, n, xy, xyz, world, fill, K4, cam_pose7);
	float p[7]; memcpy(p, pose7, sizeof p);
	POSE3R_T::lmFuncQuat(p, errors, 7, 3 * n, (void *)&cl.ptrs);
}
int ref3d_hypothesis_v1(int n, const float *xy, const float *xyz, const float *world, const float *fill, const float *K4, const float *cam_pose7,
                        float alpha, const int *sample_pos, int n_samples, const float *init_quat, int maxLM, float errThr, int minNPts,
                        float *pose_init, float *pose_lm, float *pose_refit, float *lm_err, unsigned char *inlier_mask) {
	FtzGuard g;
	return hypothesis_t<POSE3R_T>(3, n, xy, xyz, world, fill, K4, cam_pose7, alpha, sample_pos, n_samples, init_quat, maxLM, errThr, minNPts,
	                              pose_init, pose_lm, pose_refit, lm_err, inlier_mask);
}
int ref3d_ransac_v1(int n, const float *xy, const float *xyz, const float *world, const float *fill, const float *K4, const float *cam_pose7,
                    float alpha, int maxRansac, int maxLM, int nPtsAlign, int minNPts, float errThr, uint64_t seed, float *pose_out) {
	FtzGuard g;
	POSE3R_T alg(maxRansac, maxLM, 1, nPtsAlign, minNPts, errThr, alpha);
	ClusterT<POSE3R_T> cl(alg, n, xy, xyz, world, fill, K4, cam_pose7);
	ref3d_srand(seed);
	Pose pose;
	bool found = alg.RANSAC(pose, cl.ptrs);
	for (int j = 0; j < 7; j++) pose_out[j] = pose[j];
	return found ? 1 : 0;
}

/* getCauchyWeight (:187-190) */
float ref3d_cauchy_weight(float fill_distance) {
	POSE3D_T alg(1, 1, 1, 5, 6, 8, 0.5);
	return alg.getCauchyWeight(fill_distance);
}

/* lmFuncQuat (:108-183): 2 residuals per correspondence */
void ref3d_lm_func(const float *pose7, int n, const float *xy, const float *xyz, const float *world, const float *fill, const float *K4,
                   const float *cam_pose7, float alpha, float *errors) {
	FtzGuard g;
	POSE3D_T alg(1, 1, 1, 5, 6, 8, alpha);
	Cluster3D cl(alg, n, xy, xyz, world, fill, K4, cam_pose7);
	float p[7]; memcpy(p, pose7, sizeof p);
	POSE3D_T::lmFuncQuat(p, errors, 7, 2 * n, (void *)&cl.ptrs);
}

/* The body of one RANSAC iteration (:283-312) for an EXPLICIT (sample positions, initial quaternion) pair; the initial
 * translation is the reference's own initPose (:262-276: mean world3D of the samples). Returns -1 when LM failed on the
 * samples, else #inliers. Outputs like ref_hypothesis of the moped2 harness. */
int ref3d_hypothesis(int n, const float *xy, const float *xyz, const float *world, const float *fill, const float *K4, const float *cam_pose7,
                     float alpha, const int *sample_pos, int n_samples, const float *init_quat, int maxLM, float errThr, int minNPts,
                     float *pose_init, float *pose_lm, float *pose_refit, float *lm_err, unsigned char *inlier_mask) {
	FtzGuard g;
	POSE3D_T alg(1, maxLM, 1, n_samples, minNPts, errThr, alpha);
	Cluster3D cl(alg, n, xy, xyz, world, fill, K4, cam_pose7);
	vector<POSE3D_T::LmData *> samples;
	for (int j = 0; j < n_samples; j++) samples.push_back(cl.ptrs[sample_pos[j]]);
	Pose pose;
	alg.initPose(pose, samples);
	pose.rotation.init(init_quat[0], init_quat[1], init_quat[2], init_quat[3]);
	for (int j = 0; j < 7; j++) pose_init[j] = pose[j];
	Float r = alg.optimizeCamera(pose, samples, maxLM);
	lm_err[0] = r; lm_err[1] = -2;
	for (int i = 0; i < n; i++) inlier_mask[i] = 0;
	if ((int)r == -1) return -1;                           /* `int LMIterations = LMInfo; if (LMIterations == -1) continue;` :292-297 */
	for (int j = 0; j < 7; j++) pose_lm[j] = pose[j];
	vector<POSE3D_T::LmData *> consistent;
	alg.testAllPoints(consistent, pose, cl.ptrs, errThr);
	for (size_t k = 0; k < consistent.size(); k++) inlier_mask[consistent[k] - &cl.data[0]] = 1;
	if ((int)consistent.size() > minNPts) lm_err[1] = alg.optimizeCamera(pose, consistent, maxLM);
	for (int j = 0; j < 7; j++) pose_refit[j] = pose[j];
	return (int)consistent.size();
}

/* Whole RANSAC() (:278-314) on one cluster with the seeded RNG; returns found (0/1). */
int ref3d_ransac(int n, const float *xy, const float *xyz, const float *world, const float *fill, const float *K4, const float *cam_pose7,
                 float alpha, int maxRansac, int maxLM, int nPtsAlign, int minNPts, float errThr, uint64_t seed, float *pose_out) {
	FtzGuard g;
	POSE3D_T alg(maxRansac, maxLM, 1, nPtsAlign, minNPts, errThr, alpha);
	Cluster3D cl(alg, n, xy, xyz, world, fill, K4, cam_pose7);
	ref3d_srand(seed);
	Pose pose;
	bool found = alg.RANSAC(pose, cl.ptrs);
	for (int j = 0; j < 7; j++) pose_out[j] = pose[j];
	return found ? 1 : 0;
}

/* CLUSTER_LINKAGE_CPU::process (moped3d/libmoped/src/cluster/CLUSTER_LINKAGE_CPU.hpp:577-704) on the matches of ONE model:
 * coord2D, model coord3D and depthData.coord3D (world) per match, a depth map (depth per pixel; stored in slot 2 of the 4 floats
 * per pixel the class reads, moped.hpp:281-287) and its fill-distance map. Output: clusters as CSR over match indices, in the
 * order and with the member order the class produces. */
int ref3d_cluster_linkage(int n, const float *xy, const float *xyz, const float *world, int W, int H, const float *depth, const float *distance,
                          float cutoff, int minPts, int use3DFilter, float weightGamma, float alpha, int linkageType, float sigma2D, float sigma3D,
                          int *cluster_offsets, int *members) {
	FtzGuard g;
	omp_set_num_threads(1);
	CLUSTER_LINKAGE_CPU alg(cutoff, minPts, use3DFilter, weightGamma, alpha, linkageType, sigma2D, sigma3D);
	string step("CLUSTER"); alg.setStepNameAndAlg(step, 0);
	vector<SP_Model> models(1, SP_Model(new Model));
	models[0]->name = "m0";
	alg.modelsUpdated(models);
	FrameData fd;
	SP_Image dm(new Image(IMAGE_TYPE_DEPTH_MAP));
	dm->name = "cam/depth"; dm->width = W; dm->height = H;
	dm->data.assign((size_t)W * H * 4 * sizeof(Float), 0);
	for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) dm->setDepth(x, y, depth[(size_t)y * W + x]);
	SP_Image pm(new Image(IMAGE_TYPE_PROB_MAP));
	pm->name = dm->name + ".distance"; pm->width = W; pm->height = H;
	pm->data.assign((size_t)W * H * sizeof(Float), 0);
	for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) pm->setProb(x, y, distance[(size_t)y * W + x]);
	fd.images.push_back(dm); fd.images.push_back(pm);
	fd.matches.resize(1);
	fd.matches[0].resize(n);
	for (int i = 0; i < n; i++) {
		FrameData::Match &m = fd.matches[0][i];
		m.imageIdx = 0;
		m.coord2D.init(xy[2 * i], xy[2 * i + 1]);
		m.coord3D.init(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
		m.depthData.depthValid = true;
		m.depthData.coord3D.init(world[3 * i], world[3 * i + 1], world[3 * i + 2]);
		m.depthData.depth = world[3 * i + 2];
		m.depthData.fillDistance = 0;
	}
	alg.process(fd);
	int nc = 0, k = 0;
	cluster_offsets[0] = 0;
	for (size_t c = 0; c < fd.clusters[0].size(); c++) {
		for (list<int>::iterator it = fd.clusters[0][c].begin(); it != fd.clusters[0][c].end(); ++it) members[k++] = *it;
		cluster_offsets[++nc] = k;
	}
	return nc;
}

/* hierarchicalCluster (:414-531) alone on a caller-supplied similarity matrix (n x n, row-major) */
int ref3d_linkage_agglomerate(const float *K, int n, float cutoff, int minPts, int linkageType, int *cluster_offsets, int *members) {
	FtzGuard g;
	CLUSTER_LINKAGE_CPU alg(cutoff, minPts, 0, 1, 0, linkageType, -1, -1);
	SP_Image Ki = alg.getSimilarityMatrix(n);
	for (int y = 0; y < n; y++) for (int x = 0; x < n; x++) Ki->setProb(x, y, K[(size_t)y * n + x]);
	vector<FrameData::Cluster> cl = alg.hierarchicalCluster(Ki);
	int nc = 0, k = 0;
	cluster_offsets[0] = 0;
	for (size_t c = 0; c < cl.size(); c++) {
		for (list<int>::iterator it = cl[c].begin(); it != cl[c].end(); ++it) members[k++] = *it;
		cluster_offsets[++nc] = k;
	}
	return nc;
}

/* FILTER_PROJECTION_DEPTH_CPU::process (filter/FILTER_PROJECTION_DEPTH_CPU.hpp:145-329) on one frame state: models (keypoints = the
 * population the class draws its test points from, TestSampleSize of them with the harness's seedable rand() when a model has more),
 * matches per model (coord2D, coord3D, all seen by the grey camera image 0), objects in list order, a depth map (4 floats per pixel,
 * depth in the third: Image::getDepth) with its own intrinsics / pose and the fill-distance map "<name>.distance".
 * Outputs: keep / score per input object, clusters of the survivors (model-major, list order). Returns #survivors. */
int ref3d_filter_depth(int n_models, const int *n_model_pts, const float *model_xyz, const int *match_offsets, const float *match_xy,
                       const float *match_xyz, int n_obj, const int *obj_model, const float *obj_pose7,
                       int minPoints, float featureDistance, float plausibleSqDistance, float minScore, float depthFraction, int testSampleSize,
                       float minKeypointFraction, uint64_t seed, const float *K4, const float *cam_pose7, const float *depthK4,
                       const float *depth_pose7, int width, int height, const float *depth, const float *fill_distance,
                       unsigned char *keep, float *score, int *cluster_offsets, int *members) {
	FtzGuard g;
	ref3d_srand(seed);
	vector<SP_Model> models;
	size_t row = 0;
	for (int m = 0; m < n_models; m++) {
		SP_Model mod(new Model); mod->name = "obj" + toString(m);
		vector<Model::IP> &ips = mod->IPs["SIFT"];
		ips.resize(n_model_pts[m]);
		for (int i = 0; i < n_model_pts[m]; i++, row++) ips[i].coord3D.init(model_xyz[3 * row], model_xyz[3 * row + 1], model_xyz[3 * row + 2]);
		models.push_back(mod);
	}
	FrameData fd;
	list<SP_Object> objects;
	fd.objects = &objects;
	SP_Image im(new Image(IMAGE_TYPE_GRAY_IMAGE));
	im->name = "cam"; im->width = width; im->height = height;
	im->intrinsicLinearCalibration.init(K4[0], K4[1], K4[2], K4[3]);
	im->cameraPose.rotation.init(cam_pose7[0], cam_pose7[1], cam_pose7[2], cam_pose7[3]);
	im->cameraPose.translation.init(cam_pose7[4], cam_pose7[5], cam_pose7[6]);
	im->TM.init(im->cameraPose);
	fd.images.push_back(im);
	SP_Image dm(new Image(IMAGE_TYPE_DEPTH_MAP));
	dm->name = "depth"; dm->width = width; dm->height = height;
	dm->intrinsicLinearCalibration.init(depthK4[0], depthK4[1], depthK4[2], depthK4[3]);
	dm->cameraPose.rotation.init(depth_pose7[0], depth_pose7[1], depth_pose7[2], depth_pose7[3]);
	dm->cameraPose.translation.init(depth_pose7[4], depth_pose7[5], depth_pose7[6]);
	dm->TM.init(dm->cameraPose);
	dm->data.resize((size_t)width * height * 4 * sizeof(Float));
	for (int y = 0; y < height; y++) for (int x = 0; x < width; x++) dm->setDepth(x, y, depth[(size_t)y * width + x]);
	fd.images.push_back(dm);
	SP_Image fm(new Image(IMAGE_TYPE_PROB_MAP));
	fm->name = "depth.distance"; fm->width = width; fm->height = height;
	fm->data.resize((size_t)width * height * sizeof(Float));
	for (int y = 0; y < height; y++) for (int x = 0; x < width; x++) fm->setProb(x, y, fill_distance[(size_t)y * width + x]);
	fd.images.push_back(fm);
	fd.matches.resize(n_models);
	for (int m = 0; m < n_models; m++) {
		fd.matches[m].resize(match_offsets[m + 1] - match_offsets[m]);
		for (int j = match_offsets[m]; j < match_offsets[m + 1]; j++) {
			FrameData::Match &ma = fd.matches[m][j - match_offsets[m]];
			ma.coord2D.init(match_xy[2 * j], match_xy[2 * j + 1]);
			ma.coord3D.init(match_xyz[3 * j], match_xyz[3 * j + 1], match_xyz[3 * j + 2]);
			ma.imageIdx = 0;
		}
	}
	vector<Object *> raw;
	for (int o = 0; o < n_obj; o++) {
		SP_Object ob(new Object);
		ob->model = models[obj_model[o]];
		ob->pose.rotation.init(obj_pose7[7 * o], obj_pose7[7 * o + 1], obj_pose7[7 * o + 2], obj_pose7[7 * o + 3]);
		ob->pose.translation.init(obj_pose7[7 * o + 4], obj_pose7[7 * o + 5], obj_pose7[7 * o + 6]);
		ob->score = 0;
		objects.push_back(ob);
		raw.push_back(ob.get());
	}
	vector<SP_Object> hold(objects.begin(), objects.end());      /* erased objects stay alive: their scores are read below */
	FILTER_PROJECTION_DEPTH_CPU alg(minPoints, featureDistance, plausibleSqDistance, minScore, depthFraction, testSampleSize, minKeypointFraction);
	alg.modelsUpdated(models);
	std::streambuf *saved = std::cerr.rdbuf(NULL);                /* the class narrates every object on stderr */
	alg.process(fd);
	std::cerr.rdbuf(saved);
	std::cerr.clear();
	for (int o = 0; o < n_obj; o++) { keep[o] = 0; score[o] = raw[o]->score; }
	for (list<SP_Object>::iterator it = objects.begin(); it != objects.end(); ++it)
		for (int o = 0; o < n_obj; o++) if (raw[o] == it->get()) keep[o] = 1;
	int ns = 0, t = 0;
	cluster_offsets[0] = 0;
	for (int m = 0; m < n_models; m++)
		for (size_t c = 0; c < fd.clusters[m].size(); c++) {
			for (list<int>::iterator it = fd.clusters[m][c].begin(); it != fd.clusters[m][c].end(); ++it) members[t++] = *it;
			cluster_offsets[++ns] = t;
		}
	return ns;
}

/* Runs MopedBench (MopedBench.cpp:185-227) over one frame state the way Moped::processImages drives it — init(), then
 * beforeAlgorithm/afterAlgorithm for CLUSTER, POSE, FILTER, FILTER2, then allDone() — and leaves its text log in
 * <out_dir>/outputMopedBench.txt. The frame state: models (n_model_pts[m] SIFT points each, xyz concatenated), matches per model
 * (10 floats each: x y  wx wy wz  depth fillDistance depthValid  imageIdx  pad), clusters (CSR over all clusters, model id per cluster),
 * objects (model id, pose7 = quaternion xyzw + translation, score), one grey camera image (K, pose). */
int ref3d_bench_log(const char *out_dir, int n_models, const int *n_model_pts, const float *model_xyz, const int *n_matches, const float *match_rec,
                    int n_clusters, const int *cluster_model, const int *cluster_offsets, const int *members,
                    int n_obj, const int *obj_model, const float *obj_pose7, const float *obj_score, const float *K4, const float *cam_pose7) {
	char cwd[4096];
	if (!getcwd(cwd, sizeof cwd) || chdir(out_dir) != 0) return -1;
	{
		vector<SP_Model> models;
		size_t row = 0;
		for (int m = 0; m < n_models; m++) {
			SP_Model mod(new Model); mod->name = "obj" + toString(m);
			vector<Model::IP> &ips = mod->IPs["SIFT"];
			ips.resize(n_model_pts[m]);
			for (int i = 0; i < n_model_pts[m]; i++, row++) ips[i].coord3D.init(model_xyz[3 * row], model_xyz[3 * row + 1], model_xyz[3 * row + 2]);
			models.push_back(mod);
		}
		FrameData fd;
		list<SP_Object> objects;
		fd.objects = &objects;
		SP_Image im(new Image(IMAGE_TYPE_GRAY_IMAGE));
		im->width = 640; im->height = 480;
		im->intrinsicLinearCalibration.init(K4[0], K4[1], K4[2], K4[3]);
		im->cameraPose.rotation.init(cam_pose7[0], cam_pose7[1], cam_pose7[2], cam_pose7[3]);
		im->cameraPose.translation.init(cam_pose7[4], cam_pose7[5], cam_pose7[6]);
		im->TM.init(im->cameraPose);
		fd.images.push_back(im);
		fd.matches.resize(n_models);
		fd.clusters.resize(n_models);
		size_t r = 0;
		for (int m = 0; m < n_models; m++) {
			fd.matches[m].resize(n_matches[m]);
			for (int i = 0; i < n_matches[m]; i++, r++) {
				const float *q = match_rec + 10 * r;
				FrameData::Match &ma = fd.matches[m][i];
				ma.coord2D.init(q[0], q[1]);
				ma.coord3D.init(0, 0, 0);
				ma.depthData.coord3D.init(q[2], q[3], q[4]);
				ma.depthData.depth = q[5];
				ma.depthData.fillDistance = q[6];
				ma.depthData.depthValid = q[7] != 0;
				ma.imageIdx = (int)q[8];
			}
		}
		for (int c = 0; c < n_clusters; c++) {
			FrameData::Cluster cl;
			for (int k = cluster_offsets[c]; k < cluster_offsets[c + 1]; k++) cl.push_back(members[k]);
			fd.clusters[cluster_model[c]].push_back(cl);
		}
		for (int o = 0; o < n_obj; o++) {
			SP_Object ob(new Object);
			ob->model = models[obj_model[o]];
			ob->pose.rotation.init(obj_pose7[7 * o], obj_pose7[7 * o + 1], obj_pose7[7 * o + 2], obj_pose7[7 * o + 3]);
			ob->pose.translation.init(obj_pose7[7 * o + 4], obj_pose7[7 * o + 5], obj_pose7[7 * o + 6]);
			ob->score = obj_score[o];
			objects.push_back(ob);
		}
		MopedBench mb;
		mb.init();
		const char *steps[4] = { "CLUSTER", "POSE", "FILTER", "FILTER2" };
		for (int s = 0; s < 4; s++) {
			string step(steps[s]);
			mb.beforeAlgorithm(step, fd);
			mb.afterAlgorithm(step, fd);
		}
		mb.allDone(fd);
	}   /* ~MopedBench closes the log */
	return chdir(cwd);
}

} /* extern "C" */
