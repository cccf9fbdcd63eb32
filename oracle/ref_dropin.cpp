/*
 * ref_dropin.cpp — TEST INFRASTRUCTURE: the drop-in proof. Compiled against the reference's OWN headers
 * (moped.hpp, util.hpp and its four CPU stage classes, -std=gnu++98 like the reference) together with the CUDA
 * stage headers of moped_b200/stages/ — i.e. exactly what a maintainer gets after adding
 *     #include <MATCH_CUDA.hpp> ...  pipeline.addAlg( "MATCH_SIFT", new MATCH_CUDA( 128, "SIFT", 5., 0.8) );
 * to moped2/libmoped/src/config.hpp. Runs one frame through the reference MopedPipeline twice,
 * once with the CPU stages (Quality=0: exact matcher) and once with the CUDA stages, and reports stage by stage
 * whether FrameData is the same. Built by oracle/Makefile into oracle/_ref/moped_dropin; run on the GPU box by
 * tests/test_gpu_cpp_stages.py.
 */
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <stdint.h>

#include <moped.hpp>
#include <util.hpp>
#include <ANN.h>
#include <lm.h>

#ifndef MAX_THREADS
#define MAX_THREADS 64
#endif

#include <match/MATCH_ANN_CPU.hpp>
#include <cluster/CLUSTER_MEAN_SHIFT_CPU.hpp>
#include <pose/POSE_RANSAC_LM_DIFF_REPROJECTION_CPU.hpp>
#include <filter/FILTER_PROJECTION_CPU.hpp>
#include <feat/FEAT_SIFT_CPU.hpp>

#include <pipeline_cuda.hpp>
#include <FEAT_SIFT_CUDA.hpp>

using namespace MopedNS;

static vector<float> readf(FILE *f, size_t n) { vector<float> v(n); if (n && fread(&v[0], 4, n, f) != n) exit(3); return v; }
static vector<int> readi(FILE *f, size_t n) { vector<int> v(n); if (n && fread(&v[0], 4, n, f) != n) exit(3); return v; }

static void build_models(vector<SP_Model> &models, const vector<int> &nPts, const vector<float> &xyz, const vector<float> &desc, int D) {
	size_t row = 0;
	for (size_t m = 0; m < nPts.size(); m++) {
		SP_Model mod(new Model);
		mod->name = "obj" + toString(m);
		vector<Model::IP> &ips = mod->IPs["SIFT"];
		ips.resize(nPts[m]);
		for (int i = 0; i < nPts[m]; i++, row++) {
			ips[i].coord3D.init(xyz[3 * row], xyz[3 * row + 1], xyz[3 * row + 2]);
			ips[i].descriptor.assign(desc.begin() + row * D, desc.begin() + (row + 1) * D);
		}
		models.push_back(mod);
	}
}

static void fill_frame(FrameData &fd, list<SP_Object> &objects, SP_Image &im, const vector<float> &qd, const vector<float> &qxy, int Q, int D) {
	fd.objects = &objects;
	fd.images.push_back(im);
	vector<FrameData::DetectedFeature> &feats = fd.detectedFeatures["SIFT"];
	feats.resize(Q);
	for (int i = 0; i < Q; i++) {
		feats[i].imageIdx = 0;
		feats[i].coord2D.init(qxy[2 * i], qxy[2 * i + 1]);
		feats[i].descriptor.assign(qd.begin() + (size_t)i * D, qd.begin() + (size_t)(i + 1) * D);
	}
}

static bool same_matches(const FrameData &a, const FrameData &b) {
	if (a.matches.size() != b.matches.size()) return false;
	for (size_t m = 0; m < a.matches.size(); m++) {
		if (a.matches[m].size() != b.matches[m].size()) return false;
		for (size_t i = 0; i < a.matches[m].size(); i++)
			if (memcmp(&a.matches[m][i], &b.matches[m][i], sizeof(FrameData::Match)) != 0) return false;
	}
	return true;
}
static bool same_clusters(const FrameData &a, const FrameData &b) { return a.clusters == b.clusters; }

/* `moped_dropin --sift case.bin`: step 1 side by side. case.bin = int32 {height, width, n_images, double_size} + pixels.
 * FEAT_SIFT_CPU and FEAT_SIFT_CUDA are registered under the step name "SIFT" in two reference MopedPipelines and run
 * on the same FrameData::images; the detectedFeatures lists are compared entry by entry (same imageIdx, coord2D and
 * descriptor within the printed tolerances, same order). */
static int sift_side_by_side(const char *path) {
	FILE *f = fopen(path, "rb");
	if (!f) return 2;
	vector<int> hdr = readi(f, 4);
	const int H = hdr[0], W = hdr[1], nImages = hdr[2], dbl = hdr[3];
	vector<SP_Image> images;
	for (int i = 0; i < nImages; i++) {
		SP_Image im(new Image);
		im->width = W; im->height = H;
		im->data.resize((size_t)W * H);
		if (fread(&im->data[0], 1, (size_t)W * H, f) != (size_t)W * H) return 3;
		images.push_back(im);
	}
	fclose(f);
	omp_set_num_threads(1);                      /* the reference's list order and duplicate suppression race otherwise */
	MopedPipeline cpu, gpu;
	cpu.addAlg( "SIFT", new FEAT_SIFT_CPU( dbl ? "-1" : "0" ) );
	gpu.addAlg( "SIFT", new FEAT_SIFT_CUDA( dbl ? "-1" : "0" ) );
	map<string,string> cfg;
	list<MopedAlg *> ca = cpu.getAlgs(true), ga = gpu.getAlgs(true);
	foreach( alg, ga ) alg->getConfig(cfg);
	foreach( kv, cfg ) printf("CONFIG %s=%s\n", kv.first.c_str(), kv.second.c_str());
	foreach( alg, ca ) { alg->getConfig(cfg); alg->setConfig(cfg); }     /* FEAT_SIFT_CPU sets libsiftfast's DoubleImSize here (:69-76) */
	FrameData fdCpu, fdGpu;
	fdCpu.images = images; fdGpu.images = images;
	try {
		foreach( alg, ca ) alg->process(fdCpu);
		foreach( alg, ga ) alg->process(fdGpu);
	} catch (string &e) { fprintf(stderr, "ERROR %s\n", e.c_str()); return 1; }
	vector<FrameData::DetectedFeature> &a = fdCpu.detectedFeatures["SIFT"], &b = fdGpu.detectedFeatures["SIFT"];
	size_t n = a.size() < b.size() ? a.size() : b.size(), same = 0;
	float maxdxy = 0, maxdd = 0;
	for (size_t i = 0; i < n; i++) {
		float dxy = fmaxf(fabsf(a[i].coord2D[0] - b[i].coord2D[0]), fabsf(a[i].coord2D[1] - b[i].coord2D[1])), dd = 0;
		for (int k = 0; k < 128; k++) dd = fmaxf(dd, fabsf(a[i].descriptor[k] - b[i].descriptor[k]));
		if (a[i].imageIdx == b[i].imageIdx && dxy < 0.02f && dd < 5e-3f) { same++; maxdxy = fmaxf(maxdxy, dxy); maxdd = fmaxf(maxdd, dd); }
	}
	printf("SIFT cpu=%d gpu=%d same_in_order=%d max_dxy=%g max_ddesc=%g\n", (int)a.size(), (int)b.size(), (int)same, maxdxy, maxdd);
	return 0;
}

int main(int argc, char **argv) {
	if (argc < 2) return 2;
	if (argc >= 3 && !strcmp(argv[1], "--sift")) return sift_side_by_side(argv[2]);
	FILE *f = fopen(argv[1], "rb");
	if (!f) return 2;
	vector<int> hdr = readi(f, 4);
	const int nModels = hdr[0], N = hdr[1], Q = hdr[2], D = hdr[3];
	vector<int> nPts = readi(f, nModels);
	vector<float> xyz = readf(f, 3 * (size_t)N), desc = readf(f, (size_t)N * D), qd = readf(f, (size_t)Q * D), qxy = readf(f, 2 * (size_t)Q);
	fclose(f);
	omp_set_num_threads(1);

	vector<SP_Model> modelsCpu, modelsGpu;
	build_models(modelsCpu, nPts, xyz, desc, D);
	build_models(modelsGpu, nPts, xyz, desc, D);
	SP_Image im(new Image);
	im->width = 640; im->height = 480;
	im->intrinsicLinearCalibration.init(800.f, 800.f, 320.f, 240.f);
	im->cameraPose.rotation.init(0.f, 0.f, 0.f, 1.f); im->cameraPose.translation.init(0.f, 0.f, 0.f);
	im->TM.init(im->cameraPose);

	MopedPipeline cpu, gpu;
	cpu.addAlg( "MATCH_SIFT", new MATCH_ANN_CPU( 128, "SIFT", 0., 0.8) );      /* exact mode: the parity target */
	cpu.addAlg( "CLUSTER", new CLUSTER_MEAN_SHIFT_CPU( 200, 20, 7, 100) );
	cpu.addAlg( "POSE", new POSE_RANSAC_LM_DIFF_REPROJECTION_CPU( 600, 200, 4, 5, 6, 10) );
	cpu.addAlg( "FILTER", new FILTER_PROJECTION_CPU( 5, 4096., 2) );
	cpu.addAlg( "POSE2", new POSE_RANSAC_LM_DIFF_REPROJECTION_CPU( 100, 500, 4, 6, 8, 5) );
	cpu.addAlg( "FILTER2", new FILTER_PROJECTION_CPU( 7, 4096., 3) );
	createCudaRecognitionPipeline(gpu);

	map<string,string> cfg;
	list<MopedAlg *> ga = gpu.getAlgs();
	foreach( alg, ga ) alg->getConfig(cfg);
	foreach( kv, cfg ) printf("CONFIG %s=%s\n", kv.first.c_str(), kv.second.c_str());

	list<SP_Object> objCpu, objGpu;
	FrameData fdCpu, fdGpu;
	fill_frame(fdCpu, objCpu, im, qd, qxy, Q, D);
	fill_frame(fdGpu, objGpu, im, qd, qxy, Q, D);
	list<MopedAlg *> ca = cpu.getAlgs(true); ga = gpu.getAlgs(true);
	foreach( alg, ca ) alg->modelsUpdated(modelsCpu);
	foreach( alg, ga ) alg->modelsUpdated(modelsGpu);

	try {
		list<MopedAlg *>::iterator c = ca.begin(), g = ga.begin();
		for (int step = 0; c != ca.end(); ++c, ++g, ++step) {
			(*c)->process(fdCpu);
			(*g)->process(fdGpu);
			if (step == 0) printf("STEP MATCH same=%d\n", (int)same_matches(fdCpu, fdGpu));
			if (step == 1) printf("STEP CLUSTER same=%d\n", (int)same_clusters(fdCpu, fdGpu));
			if (step >= 2) printf("STEP %s cpu_objects=%d gpu_objects=%d\n", (*c)->_stepName.c_str(), (int)objCpu.size(), (int)objGpu.size());
		}
	} catch (string &e) { fprintf(stderr, "ERROR %s\n", e.c_str()); return 1; }
	foreach( o, objCpu ) printf("CPU %s %.6f %.6f %.6f %.6f %.6f %.6f %.6f score %.4f\n", o->model->name.c_str(), o->pose[0], o->pose[1], o->pose[2], o->pose[3], o->pose[4], o->pose[5], o->pose[6], o->score);
	foreach( o, objGpu ) printf("GPU %s %.6f %.6f %.6f %.6f %.6f %.6f %.6f score %.4f\n", o->model->name.c_str(), o->pose[0], o->pose[1], o->pose[2], o->pose[3], o->pose[4], o->pose[5], o->pose[6], o->score);
	return 0;
}
