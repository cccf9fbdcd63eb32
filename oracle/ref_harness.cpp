/*
 * ref_harness.cpp — TEST INFRASTRUCTURE (never linked into the product).
 *
 * Thin C-ABI around the reference's OWN hot-path stage classes, compiled unmodified from
 * /root/reference by oracle/Makefile into oracle/_ref/libmoped_ref.so:
 *   MATCH_ANN_CPU                         moped2/libmoped/src/match/MATCH_ANN_CPU.hpp:52-178
 *   CLUSTER_MEAN_SHIFT_CPU                moped2/libmoped/src/cluster/CLUSTER_MEAN_SHIFT_CPU.hpp:50-199
 *   POSE_RANSAC_LM_DIFF_REPROJECTION_CPU  moped2/libmoped/src/pose/POSE_RANSAC_LM_DIFF_REPROJECTION_CPU.hpp:57-307
 *   FILTER_PROJECTION_CPU                 moped2/libmoped/src/filter/FILTER_PROJECTION_CPU.hpp:50-162
 * plus the vendored ANN 1.1.1 and levmar 2.4 from moped2/libmoped/libs/libs.tgz.
 *
 * Nothing of the reference is restated here: this file only builds FrameData from flat arrays,
 * calls alg->modelsUpdated()/alg->process() exactly as MopedPimpl::processImages does
 * (moped2/libmoped/src/moped.cpp:166-194), and copies results back out. It is the arbiter the
 * C restatement (oracle/moped_oracle.c) and the CUDA path are checked against, and the
 * `cpu_baseline.kind == "reference"` timer of bench.py.
 *
 * Two preprocessor shims, neither of which edits reference source:
 *   - `#define class struct` around the four stage headers so the harness can call their private
 *     members per hypothesis (randSample / initPose / optimizeCamera / testAllPoints).
 *   - `#define rand moped_ref_rand` so RANSAC draws from a seedable per-thread LCG instead of the
 *     process-global, lock-protected libc rand() (reference RANSAC is otherwise non-reproducible).
 */
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <stdint.h>
#include <xmmintrin.h>
#include <pmmintrin.h>

#include <moped.hpp>
#include <util.hpp>
#include <ANN.h>
#include <lm.h>
#include <sXML.hpp>

/* ---- seedable stand-in for libc rand() (31-bit output like glibc; RAND_MAX = 2^31-1) ---- */
static __thread uint64_t g_rng_state = 0x9E3779B97F4A7C15ULL;
extern "C" int moped_ref_rand(void) {
	g_rng_state = g_rng_state * 6364136223846793005ULL + 1442695040888963407ULL;
	return (int)((g_rng_state >> 33) & 0x7fffffffULL);
}
extern "C" void ref_srand(uint64_t seed) { g_rng_state = seed; }

#ifndef MAX_THREADS
#define MAX_THREADS 64
#endif

#define rand moped_ref_rand
#define class struct
#include <match/MATCH_ANN_CPU.hpp>
#include <cluster/CLUSTER_MEAN_SHIFT_CPU.hpp>
#include <pose/POSE_RANSAC_LM_DIFF_REPROJECTION_CPU.hpp>
#include <filter/FILTER_PROJECTION_CPU.hpp>
#include <feat/FEAT_SIFT_CPU.hpp>
#undef class
#undef rand

using namespace MopedNS;

typedef POSE_RANSAC_LM_DIFF_REPROJECTION_CPU POSE_T;

/* The reference executable is linked with -ffast-math, which makes crt set FTZ|DAZ at start-up
 * (SURVEY.md Appendix C). GCC >= 13 no longer does that for shared objects, so do it explicitly,
 * for the calling thread and for every OpenMP worker. */
static void set_ftz_daz_workers() {
	#pragma omp parallel
	{
		if (omp_get_thread_num() != 0) {
			_MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON);
			_MM_SET_DENORMALS_ZERO_MODE(_MM_DENORMALS_ZERO_ON);
		}
	}
}
/* The calling (Python) thread gets FTZ|DAZ only while it is inside reference code. */
struct FtzGuard {
	unsigned saved;
	FtzGuard() { saved = _mm_getcsr(); _mm_setcsr(saved | 0x8040u); }
	~FtzGuard() { _mm_setcsr(saved); }
};

static double now_s() {
	struct timespec t; clock_gettime(CLOCK_REALTIME, &t);
	return t.tv_sec + 1e-9 * t.tv_nsec;
}

struct RefCtx {
	vector<SP_Model> models;
	vector<SP_Image> images;
	list<SP_Object> objects;
	FrameData fd;
	MATCH_ANN_CPU *match;
	float matchQuality, matchRatio;
	int nThreads;
	RefCtx() : match(NULL), matchQuality(-1), matchRatio(-1), nThreads(1) {}
};

static string step_name(const char *s) { return string(s); }

extern "C" {

void *ref_create(int n_threads) {
	RefCtx *c = new RefCtx;
	c->nThreads = n_threads < 1 ? 1 : (n_threads > MAX_THREADS ? MAX_THREADS : n_threads);
	omp_set_num_threads(c->nThreads);
	set_ftz_daz_workers();
	c->fd.objects = &c->objects;
	return c;
}

void ref_destroy(void *h) {
	RefCtx *c = (RefCtx *)h;
	if (c->match) delete c->match;
	delete c;
}

int ref_max_threads(void) { return MAX_THREADS; }

/* Model database: n_models objects, n_pts[m] points each, rows concatenated in model order. */
void ref_set_models(void *h, int n_models, const int *n_pts, const float *xyz, const float *desc, int D) {
	RefCtx *c = (RefCtx *)h;
	c->models.clear();
	long row = 0;
	for (int m = 0; m < n_models; m++) {
		SP_Model mod(new Model);
		char nm[32]; snprintf(nm, sizeof nm, "obj%06d", m);
		mod->name = nm;
		vector<Model::IP> &ips = mod->IPs["SIFT"];
		ips.resize(n_pts[m]);
		for (int i = 0; i < n_pts[m]; i++, row++) {
			ips[i].coord3D.init(xyz[3 * row], xyz[3 * row + 1], xyz[3 * row + 2]);
			ips[i].descriptor.assign(desc + row * D, desc + (row + 1) * D);
		}
		c->models.push_back(mod);
	}
	if (c->match) { delete c->match; c->match = NULL; }
}

/* Model load through the reference's OWN sXML reader (include/sXML.hpp) followed by the loop of
 * MopedPimpl::addModel(sXML&) and addModel(SP_Model&) (moped.cpp:100-149). MopedPimpl itself cannot be compiled here
 * (moped.cpp includes config.hpp, hence every FEAT_* stage and OpenCV), so those fifteen lines of glue are repeated
 * verbatim in behaviour: last <Points> child, every child of it is a point, `istringstream >>` for p3d and desc,
 * replace-by-name. Returns 1 if a model was added/replaced, 0 if the file has no Points (addModel returns ""),
 * -1 if sXML::fromFile fails. */
int ref_add_model_xml(void *h, const char *path) {
	RefCtx *c = (RefCtx *)h;
	sXML sxml;
	string fileName(path);
	if (!sxml.fromFile(fileName)) return -1;
	SP_Model m(new Model);
	m->name = sxml["name"];
	m->boundingBox[0].init(10E10, 10E10, 10E10);
	m->boundingBox[1].init(-10E10, -10E10, -10E10);
	sXML *points = NULL;
	foreach (pts, sxml.children)
		if (pts.name == "Points") points = &pts;
	if (points == NULL) return 0;
	foreach (pt, points->children) {
		Model::IP ip;
		ip.coord3D.init(0, 0, 0);                 /* the reference leaves it uninitialised; files under test always give p3d */
		std::istringstream iss(pt["p3d"]);
		iss >> ip.coord3D;
		m->boundingBox[0].min(ip.coord3D);
		m->boundingBox[1].max(ip.coord3D);
		std::istringstream jss(pt["desc"]);
		Float f;
		while (jss >> f) ip.descriptor.push_back(f);
		m->IPs[pt["desc_type"]].push_back(ip);
	}
	int found = false;
	foreach (mm, c->models)
		if ((found = (mm->name == m->name))) mm = m;
	if (!found) c->models.push_back(m);
	if (c->match) { delete c->match; c->match = NULL; }
	return 1;
}

int ref_model_count(void *h) { return (int)((RefCtx *)h)->models.size(); }

int ref_model_name(void *h, int i, char *buf, int cap) {
	RefCtx *c = (RefCtx *)h;
	snprintf(buf, cap, "%s", c->models[i]->name.c_str());
	return (int)c->models[i]->name.size();
}

void ref_model_bbox(void *h, int i, float *bbox6) {
	RefCtx *c = (RefCtx *)h;
	for (int k = 0; k < 3; k++) { bbox6[k] = c->models[i]->boundingBox[0][k]; bbox6[3 + k] = c->models[i]->boundingBox[1][k]; }
}

/* #points of model i under desc_type, and the total number of descriptor values they hold */
int ref_model_points(void *h, int i, const char *desc_type, long *n_values) {
	RefCtx *c = (RefCtx *)h;
	map<string, vector<Model::IP> >::iterator it = c->models[i]->IPs.find(desc_type);
	if (it == c->models[i]->IPs.end()) { if (n_values) *n_values = 0; return 0; }
	long v = 0;
	for (size_t k = 0; k < it->second.size(); k++) v += (long)it->second[k].descriptor.size();
	if (n_values) *n_values = v;
	return (int)it->second.size();
}

void ref_get_model_points(void *h, int i, const char *desc_type, float *xyz, int *desc_len, float *desc_values) {
	RefCtx *c = (RefCtx *)h;
	vector<Model::IP> &ips = c->models[i]->IPs[desc_type];
	long v = 0;
	for (size_t k = 0; k < ips.size(); k++) {
		for (int j = 0; j < 3; j++) xyz[3 * k + j] = ips[k].coord3D[j];
		desc_len[k] = (int)ips[k].descriptor.size();
		for (size_t j = 0; j < ips[k].descriptor.size(); j++) desc_values[v++] = ips[k].descriptor[j];
	}
}

/* ---- feature extraction with the reference's own step 1 (FEAT_SIFT_CPU over the vendored libsiftfast 1.1), used
 * only to turn the reference's shipped test images into real descriptors for the parity fixtures of the hot path
 * (SURVEY.md 8d "real-image config"). ScaleOrigin "-1" doubles the image like config.hpp:72; "0" does not. ---- */
static vector<FrameData::DetectedFeature> g_sift_out;

int ref_sift(const unsigned char *gray, int height, int width, int double_size) {
	FEAT_SIFT_CPU alg(double_size ? "-1" : "0");
	map<string, string> cfg;
	alg.getConfig(cfg);
	alg.setConfig(cfg);                       /* sets libsiftfast's DoubleImSize from ScaleOrigin (FEAT_SIFT_CPU.hpp:69-76) */
	SP_Image im(new Image);
	im->name = "img"; im->width = width; im->height = height;
	im->data.assign(gray, gray + (size_t)width * height);
	FrameData fd;
	fd.images.push_back(im);
	alg.process(fd);
	g_sift_out.clear();
	for (map<string, vector<FrameData::DetectedFeature> >::iterator it = fd.detectedFeatures.begin(); it != fd.detectedFeatures.end(); ++it)
		g_sift_out.insert(g_sift_out.end(), it->second.begin(), it->second.end());
	return (int)g_sift_out.size();
}

void ref_sift_get(float *xy, float *desc) {
	for (size_t i = 0; i < g_sift_out.size(); i++) {
		xy[2 * i] = g_sift_out[i].coord2D[0]; xy[2 * i + 1] = g_sift_out[i].coord2D[1];
		memcpy(desc + 128 * i, &g_sift_out[i].descriptor[0], 128 * sizeof(float));
	}
}

/* Descriptors as the reference holds them now (MATCH normalises the Model in place). */
void ref_get_model_desc(void *h, float *out, int D) {
	RefCtx *c = (RefCtx *)h;
	long row = 0;
	for (size_t m = 0; m < c->models.size(); m++) {
		vector<Model::IP> &ips = c->models[m]->IPs["SIFT"];
		for (size_t i = 0; i < ips.size(); i++, row++)
			memcpy(out + row * D, &ips[i].descriptor[0], D * sizeof(float));
	}
}

/* Cameras: K = (fx, fy, cx, cy); cam_pose = quat(x,y,z,w) + t, as Image::cameraPose. */
void ref_set_images(void *h, int n_images, const float *K, const float *cam_pose) {
	RefCtx *c = (RefCtx *)h;
	c->images.clear();
	for (int i = 0; i < n_images; i++) {
		SP_Image im(new Image);
		im->width = 640; im->height = 480;
		im->intrinsicLinearCalibration.init(K[4 * i], K[4 * i + 1], K[4 * i + 2], K[4 * i + 3]);
		im->intrinsicNonlinearCalibration.init(0.f, 0.f, 0.f, 0.f);
		im->cameraPose.rotation.init(cam_pose[7 * i], cam_pose[7 * i + 1], cam_pose[7 * i + 2], cam_pose[7 * i + 3]);
		im->cameraPose.translation.init(cam_pose[7 * i + 4], cam_pose[7 * i + 5], cam_pose[7 * i + 6]);
		im->TM.init(im->cameraPose);            /* moped.cpp:168-169 */
		c->images.push_back(im);
	}
	c->fd.images = c->images;
}

void ref_set_features(void *h, int Q, int D, const float *desc, const float *xy, const int *image_idx) {
	RefCtx *c = (RefCtx *)h;
	vector<FrameData::DetectedFeature> &f = c->fd.detectedFeatures["SIFT"];
	f.resize(Q);
	for (int i = 0; i < Q; i++) {
		f[i].imageIdx = image_idx[i];
		f[i].coord2D.init(xy[2 * i], xy[2 * i + 1]);
		f[i].descriptor.assign(desc + (long)i * D, desc + (long)(i + 1) * D);
	}
}

void ref_get_features_desc(void *h, float *out, int D) {
	RefCtx *c = (RefCtx *)h;
	vector<FrameData::DetectedFeature> &f = c->fd.detectedFeatures["SIFT"];
	for (size_t i = 0; i < f.size(); i++) memcpy(out + i * D, &f[i].descriptor[0], D * sizeof(float));
}

void ref_clear_frame(void *h) {
	RefCtx *c = (RefCtx *)h;
	c->fd.matches.clear(); c->fd.clusters.clear(); c->fd.oldClusters.clear();
	c->objects.clear(); c->fd.oldObjects.clear();
}

/* L2-normalise rows with the matcher's own norm() (MATCH_ANN_CPU.hpp:54-57). */
void ref_norm_rows(float *desc, int n, int D) {
	FtzGuard ftz_guard;
	vector<float> v(D);
	for (int i = 0; i < n; i++) {
		v.assign(desc + (long)i * D, desc + (long)(i + 1) * D);
		MATCH_ANN_CPU::norm(v);
		memcpy(desc + (long)i * D, &v[0], D * sizeof(float));
	}
}

/* (Re)build the matcher (kd-tree) if models or parameters changed; returns build seconds. */
static double ensure_match(RefCtx *c, float quality, float ratio) {
	FtzGuard ftz_guard;
	if (c->match && c->matchQuality == quality && c->matchRatio == ratio) return 0.;
	if (c->match) delete c->match;
	int D = c->models.empty() || c->models[0]->IPs["SIFT"].empty() ? 128 : (int)c->models[0]->IPs["SIFT"][0].descriptor.size();
	c->match = new MATCH_ANN_CPU(D, "SIFT", quality, ratio);
	string sn("MATCH_SIFT"); c->match->setStepNameAndAlg(sn, 0);
	c->match->modelsUpdated(c->models);
	double t0 = now_s();
	c->match->Update();
	c->matchQuality = quality; c->matchRatio = ratio;
	return now_s() - t0;
}

double ref_build_match(void *h, float quality, float ratio) { return ensure_match((RefCtx *)h, quality, ratio); }

/* MATCH step. Returns process() wall seconds (index build excluded, see ensure_match). */
double ref_run_match(void *h, float quality, float ratio) {
	FtzGuard ftz_guard;
	RefCtx *c = (RefCtx *)h;
	ensure_match(c, quality, ratio);
	c->fd.matches.clear();
	double t0 = now_s();
	c->match->process(c->fd);
	return now_s() - t0;
}

/* Raw 2-NN of already-normalised queries through the matcher's kd-tree (MATCH_ANN_CPU.hpp:162). */
void ref_ann_search(void *h, const float *q, int Q, float eps, int *idx, float *dist) {
	FtzGuard ftz_guard;
	RefCtx *c = (RefCtx *)h;
	ensure_match(c, c->matchQuality < 0 ? 0.f : c->matchQuality, c->matchRatio < 0 ? 0.8f : c->matchRatio);
	int D = c->match->DescriptorSize;
	ANNpoint pt = annAllocPt(D);
	ANNidx nx[2]; ANNdist ds[2];
	for (int i = 0; i < Q; i++) {
		for (int j = 0; j < D; j++) pt[j] = q[(long)i * D + j];
		c->match->kdtree->annkSearch(pt, 2, nx, ds, eps);
		idx[2 * i] = nx[0]; idx[2 * i + 1] = nx[1];
		dist[2 * i] = ds[0]; dist[2 * i + 1] = ds[1];
	}
	annDeallocPt(pt);
}

/* ---- matches ---- */
int ref_match_total(void *h) {
	RefCtx *c = (RefCtx *)h; int t = 0;
	for (size_t m = 0; m < c->fd.matches.size(); m++) t += c->fd.matches[m].size();
	return t;
}
int ref_match_models(void *h) { return (int)((RefCtx *)h)->fd.matches.size(); }

void ref_get_matches(void *h, int *offsets, int *image, float *xy, float *xyz) {
	RefCtx *c = (RefCtx *)h; int t = 0;
	for (size_t m = 0; m < c->fd.matches.size(); m++) {
		offsets[m] = t;
		for (size_t i = 0; i < c->fd.matches[m].size(); i++, t++) {
			const FrameData::Match &ma = c->fd.matches[m][i];
			image[t] = ma.imageIdx;
			xy[2 * t] = ma.coord2D[0]; xy[2 * t + 1] = ma.coord2D[1];
			xyz[3 * t] = ma.coord3D[0]; xyz[3 * t + 1] = ma.coord3D[1]; xyz[3 * t + 2] = ma.coord3D[2];
		}
	}
	offsets[c->fd.matches.size()] = t;
}

void ref_set_matches(void *h, int n_models, const int *offsets, const int *image, const float *xy, const float *xyz) {
	RefCtx *c = (RefCtx *)h;
	c->fd.matches.clear(); c->fd.matches.resize(n_models);
	for (int m = 0; m < n_models; m++) {
		c->fd.matches[m].resize(offsets[m + 1] - offsets[m]);
		for (int t = offsets[m]; t < offsets[m + 1]; t++) {
			FrameData::Match &ma = c->fd.matches[m][t - offsets[m]];
			ma.imageIdx = image[t];
			ma.coord2D.init(xy[2 * t], xy[2 * t + 1]);
			ma.coord3D.init(xyz[3 * t], xyz[3 * t + 1], xyz[3 * t + 2]);
		}
	}
}

/* ---- CLUSTER step ---- */
double ref_run_cluster(void *h, float radius, float merge, int minpts, int maxiter) {
	FtzGuard ftz_guard;
	RefCtx *c = (RefCtx *)h;
	CLUSTER_MEAN_SHIFT_CPU alg(radius, merge, minpts, maxiter);
	string sn("CLUSTER"); alg.setStepNameAndAlg(sn, 0);
	alg.modelsUpdated(c->models);
	c->fd.clusters.clear();
	double t0 = now_s();
	alg.process(c->fd);
	return now_s() - t0;
}

int ref_cluster_count(void *h, int *total_members) {
	RefCtx *c = (RefCtx *)h; int n = 0, t = 0;
	for (size_t m = 0; m < c->fd.clusters.size(); m++)
		for (size_t k = 0; k < c->fd.clusters[m].size(); k++) { n++; t += c->fd.clusters[m][k].size(); }
	if (total_members) *total_members = t;
	return n;
}

void ref_get_clusters(void *h, int *model, int *offsets, int *members) {
	RefCtx *c = (RefCtx *)h; int n = 0, t = 0;
	for (size_t m = 0; m < c->fd.clusters.size(); m++)
		for (size_t k = 0; k < c->fd.clusters[m].size(); k++) {
			model[n] = (int)m; offsets[n] = t; n++;
			foreach( p, c->fd.clusters[m][k] ) members[t++] = p;
		}
	offsets[n] = t;
}

void ref_set_clusters(void *h, int n_clusters, const int *model, const int *offsets, const int *members) {
	RefCtx *c = (RefCtx *)h;
	c->fd.clusters.clear(); c->fd.clusters.resize(c->models.size());
	for (int k = 0; k < n_clusters; k++) {
		FrameData::Cluster cl;
		for (int t = offsets[k]; t < offsets[k + 1]; t++) cl.push_back(members[t]);
		c->fd.clusters[model[k]].push_back(cl);
	}
}

/* ---- POSE / POSE2 step (full RANSAC, reference control flow) ---- */
double ref_run_pose(void *h, const char *step, int maxRansac, int maxLM, int maxObj, int nPtsAlign, int minNPts, float errThr, uint64_t seed) {
	FtzGuard ftz_guard;
	RefCtx *c = (RefCtx *)h;
	POSE_T alg(maxRansac, maxLM, maxObj, nPtsAlign, minNPts, errThr);
	string sn = step_name(step); alg.setStepNameAndAlg(sn, 0);
	alg.modelsUpdated(c->models);
	#pragma omp parallel
	{ ref_srand(seed + 0x632BE59BD9B4E019ULL * (uint64_t)(omp_get_thread_num() + 1)); }
	ref_srand(seed + 0x632BE59BD9B4E019ULL);
	double t0 = now_s();
	alg.process(c->fd);
	return now_s() - t0;
}

int ref_object_count(void *h) { return (int)((RefCtx *)h)->objects.size(); }

static int model_index(RefCtx *c, const Object &o) {
	for (size_t m = 0; m < c->models.size(); m++) if (c->models[m].get() == o.model.get()) return (int)m;
	return -1;
}

void ref_get_objects(void *h, int *model, float *pose, float *score) {
	RefCtx *c = (RefCtx *)h; int n = 0;
	foreach( o, c->objects ) {
		model[n] = model_index(c, *o);
		for (int j = 0; j < 7; j++) pose[7 * n + j] = o->pose[j];
		score[n] = o->score;
		n++;
	}
}

void ref_set_objects(void *h, int n, const int *model, const float *pose) {
	RefCtx *c = (RefCtx *)h;
	c->objects.clear();
	for (int i = 0; i < n; i++) {
		SP_Object o(new Object);
		o->model = c->models[model[i]];
		for (int j = 0; j < 7; j++) o->pose[j] = pose[7 * i + j];
		o->score = 0;
		c->objects.push_back(o);
	}
}

/* ---- FILTER / FILTER2 step ---- */
double ref_run_filter(void *h, int minPoints, float featDist, float minScore) {
	FtzGuard ftz_guard;
	RefCtx *c = (RefCtx *)h;
	FILTER_PROJECTION_CPU alg(minPoints, featDist, minScore);
	string sn("FILTER"); alg.setStepNameAndAlg(sn, 0);
	alg.modelsUpdated(c->models);
	double t0 = now_s();
	alg.process(c->fd);
	return now_s() - t0;
}

/* ---- per-hypothesis access to the POSE class's private members ---- */
struct HypCtx {
	vector< vector<POSE_T::LmData> > lmData;
	vector<POSE_T::LmData *> cl;
};

static void build_cluster(RefCtx *c, POSE_T &alg, HypCtx &hc, int model, const int *members, int n) {
	alg.preprocessAllMatches(hc.lmData, c->fd.matches, c->fd.images);
	hc.cl.clear();
	for (int i = 0; i < n; i++) hc.cl.push_back(&hc.lmData[model][members[i]]);
}

/* Draw n_hyp (sample set, init quaternion) pairs in the order RANSAC() would
 * (randSample :76-98 then initPose :182-186), from the seeded stand-in RNG.
 * sample_pos = positions inside `members`; returns #hypotheses for which randSample succeeded
 * (it fails for all or none: it depends only on the number of distinct points). */
int ref_draw_samples(void *h, int model, const int *members, int n, int nPtsAlign, uint64_t seed, int n_hyp, int *sample_pos, float *init_quat) {
	FtzGuard ftz_guard;
	RefCtx *c = (RefCtx *)h;
	POSE_T alg(1, 1, 1, nPtsAlign, 0, 0);
	HypCtx hc; build_cluster(c, alg, hc, model, members, n);
	ref_srand(seed);
	int ok = 0;
	for (int k = 0; k < n_hyp; k++) {
		vector<POSE_T::LmData *> samples;
		if (!alg.randSample(samples, hc.cl, nPtsAlign)) { for (int j = 0; j < nPtsAlign; j++) sample_pos[k * nPtsAlign + j] = -1; continue; }
		Pose pose; alg.initPose(pose, samples);
		for (int j = 0; j < nPtsAlign; j++) {
			int pos = -1;
			for (int i = 0; i < n; i++) if (hc.cl[i] == samples[j]) { pos = i; break; }
			sample_pos[k * nPtsAlign + j] = pos;
		}
		for (int j = 0; j < 4; j++) init_quat[4 * k + j] = pose.rotation[j];
		ok++;
	}
	return ok;
}

/* One RANSAC iteration body on an explicit (sample set, init quat): optimizeCamera :140-164,
 * testAllPoints :166-180, and the refit on inliers (:204-208) when #inliers > minNPts.
 * Returns -1 when LM failed on the samples (iteration skipped, :199), else #inliers.
 * pose_lm = pose after the sample fit; pose_refit = pose after the inlier refit (or copy of pose_lm).
 * lm_err[0] = info[1] of the sample fit, lm_err[1] = of the refit (or -2 if no refit). */
int ref_hypothesis(void *h, int model, const int *members, int n, const int *sample_pos, int n_samples, const float *init_quat,
                   int maxLM, float errThr, int minNPts, float *pose_lm, float *pose_refit, float *lm_err, unsigned char *inlier_mask) {
	FtzGuard ftz_guard;
	RefCtx *c = (RefCtx *)h;
	POSE_T alg(1, maxLM, 1, n_samples, minNPts, errThr);
	HypCtx hc; build_cluster(c, alg, hc, model, members, n);
	vector<POSE_T::LmData *> samples;
	for (int j = 0; j < n_samples; j++) samples.push_back(hc.cl[sample_pos[j]]);
	Pose pose;
	pose.rotation.init(init_quat[0], init_quat[1], init_quat[2], init_quat[3]);
	pose.translation.init(0., 0., 0.5);
	Float r = alg.optimizeCamera(pose, samples, maxLM);
	lm_err[0] = r; lm_err[1] = -2;
	for (int i = 0; i < n; i++) inlier_mask[i] = 0;
	if ((int)r == -1) return -1;                      /* RANSAC(): `int LMIterations = optimizeCamera(...)` */
	for (int j = 0; j < 7; j++) pose_lm[j] = pose[j];
	vector<POSE_T::LmData *> consistent;
	alg.testAllPoints(consistent, pose, hc.cl, errThr);
	for (size_t k = 0; k < consistent.size(); k++)
		for (int i = 0; i < n; i++) if (hc.cl[i] == consistent[k]) inlier_mask[i] = 1;
	if ((int)consistent.size() > minNPts) lm_err[1] = alg.optimizeCamera(pose, consistent, maxLM);
	for (int j = 0; j < 7; j++) pose_refit[j] = pose[j];
	return (int)consistent.size();
}

/* The same iteration body for MANY explicit hypotheses of one cluster (BASELINE.json configs[3]), spread over the
 * OpenMP team the way process() spreads its tasks (`#pragma omp parallel for`, POSE_..._CPU.hpp:282). Only
 * #inliers and the final pose are kept. Returns the wall-clock seconds of the loop. */
double ref_hypotheses_batch(void *h, int model, const int *members, int n, const int *sample_pos, int n_samples, const float *init_quat,
                            int n_hyp, int maxLM, float errThr, int minNPts, int *n_inliers, float *pose_out) {
	FtzGuard ftz_guard;
	RefCtx *c = (RefCtx *)h;
	POSE_T alg(1, maxLM, 1, n_samples, minNPts, errThr);
	HypCtx hc; build_cluster(c, alg, hc, model, members, n);
	double t0 = now_s();
	#pragma omp parallel for schedule(dynamic, 8)
	for (int k = 0; k < n_hyp; k++) {
		vector<POSE_T::LmData *> samples;
		for (int j = 0; j < n_samples; j++) samples.push_back(hc.cl[sample_pos[k * n_samples + j]]);
		Pose pose;
		pose.rotation.init(init_quat[4 * k], init_quat[4 * k + 1], init_quat[4 * k + 2], init_quat[4 * k + 3]);
		pose.translation.init(0., 0., 0.5);
		Float r = alg.optimizeCamera(pose, samples, maxLM);
		n_inliers[k] = -1;
		for (int j = 0; j < 7; j++) pose_out[7 * k + j] = 0;
		if ((int)r == -1) continue;
		vector<POSE_T::LmData *> consistent;
		alg.testAllPoints(consistent, pose, hc.cl, errThr);
		if ((int)consistent.size() > minNPts) alg.optimizeCamera(pose, consistent, maxLM);
		n_inliers[k] = (int)consistent.size();
		for (int j = 0; j < 7; j++) pose_out[7 * k + j] = pose[j];
	}
	return now_s() - t0;
}

/* Whole RANSAC() on one cluster with the seeded RNG; returns found (0/1). */
int ref_ransac(void *h, int model, const int *members, int n, int maxRansac, int maxLM, int nPtsAlign, int minNPts, float errThr, uint64_t seed, float *pose_out) {
	FtzGuard ftz_guard;
	RefCtx *c = (RefCtx *)h;
	POSE_T alg(maxRansac, maxLM, 1, nPtsAlign, minNPts, errThr);
	HypCtx hc; build_cluster(c, alg, hc, model, members, n);
	ref_srand(seed);
	Pose pose;
	bool found = alg.RANSAC(pose, hc.cl);
	for (int j = 0; j < 7; j++) pose_out[j] = pose[j];
	return found ? 1 : 0;
}

/* project() of moped.hpp:330-354 for an array of points through image 0..: test helper. */
void ref_project(void *h, const float *pose7, const float *xyz, const int *image, int n, float *uv) {
	FtzGuard ftz_guard;
	RefCtx *c = (RefCtx *)h;
	Pose pose; for (int j = 0; j < 7; j++) pose[j] = pose7[j];
	for (int i = 0; i < n; i++) {
		Pt<3> p; p.init(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
		Pt<2> r = project(pose, p, *c->images[image[i]]);
		uv[2 * i] = r[0]; uv[2 * i + 1] = r[1];
	}
}

/* Full reference stage sequence MATCH..FILTER2 (config.hpp:83-120 parameters passed in `p`),
 * per-stage wall clock exactly like moped.cpp:183-191. times[6] = seconds per stage.
 * p = { quality, ratio,  radius, merge, minpts, maxiter,
 *       ransac1, lm1, maxobj1, npts1, minn1, err1,  minPoints1, featDist1, minScore1,
 *       ransac2, lm2, maxobj2, npts2, minn2, err2,  minPoints2, featDist2, minScore2 } */
int ref_run_pipeline(void *h, const float *p, uint64_t seed, double *times) {
	RefCtx *c = (RefCtx *)h;
	ref_clear_frame(h);
	times[0] = ref_run_match(h, p[0], p[1]);
	times[1] = ref_run_cluster(h, p[2], p[3], (int)p[4], (int)p[5]);
	times[2] = ref_run_pose(h, "POSE", (int)p[6], (int)p[7], (int)p[8], (int)p[9], (int)p[10], p[11], seed);
	times[3] = ref_run_filter(h, (int)p[12], p[13], p[14]);
	times[4] = ref_run_pose(h, "POSE2", (int)p[15], (int)p[16], (int)p[17], (int)p[18], (int)p[19], p[20], seed + 1);
	times[5] = ref_run_filter(h, (int)p[21], p[22], p[23]);
	return (int)c->objects.size();
}

} /* extern "C" */
