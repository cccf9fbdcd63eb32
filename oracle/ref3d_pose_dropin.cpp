/*
 * ref3d_pose_dropin.cpp — TEST INFRASTRUCTURE: the drop-in proof for moped3d's POSE step. Compiled against moped3d's OWN headers
 * (moped.hpp, util.hpp, POSE_RANSAC_LM_DIFF_BACKPROJECTION_DEPTH_CPU.hpp / ..._REPROJECTION_DEPTH_CPU.hpp; -std=gnu++98 like the
 * reference) and its vendored levmar, together with moped_b200/stages/POSE_RANSAC_LM_DIFF_{BACKPROJECTION,REPROJECTION}_DEPTH_CUDA.hpp — what a maintainer gets after
 * replacing
 *     pipeline.addAlg( "POSE", new POSE_RANSAC_LM_DIFF_BACKPROJECTION_DEPTH_CPU( 192, 100, 4, 5, 6, 8, 0.5));   (moped3d/libmoped/src/config.hpp:46)
 * by the same line with ..._CUDA. Runs both stages in two reference MopedPipelines on identical FrameData (matches with depthData,
 * clusters, one camera) and prints the objects each produced; the two draw different random samples (libc rand() vs the seedable
 * stream), so the caller compares object counts per model and poses within the stage's own spread.
 * argv: case file, variant (0 back-projection / 1 reprojection + depth / 2 = the steps after CLUSTER of moped3d's shipped pipeline,
 * POSE -> FILTER -> POSE2 -> FILTER2 with the parameters of moped3d/libmoped/src/config.hpp:46-49, FILTER_PROJECTION_CUDA being the
 * moped2 class unchanged: moped3d's FILTER_PROJECTION_CPU is the same code).
 * Case file: int32 n_models; float K[4]; per model int32 {n_matches, n_clusters}, float {x, y, X, Y, Z, wx, wy, wz, fill} per match,
 * then per cluster int32 size followed by that many match indices.
 */
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <cfloat>
#include <stdint.h>

#include <moped.hpp>
#include <util.hpp>

#ifndef MAX_THREADS
#define MAX_THREADS 64
#endif

#include <lm.h>
#include <pose/POSE_RANSAC_LM_DIFF_BACKPROJECTION_DEPTH_CPU.hpp>
#include <pose/POSE_RANSAC_LM_DIFF_REPROJECTION_DEPTH_CPU.hpp>
#include <filter/FILTER_PROJECTION_CPU.hpp>
#include <POSE_RANSAC_LM_DIFF_BACKPROJECTION_DEPTH_CUDA.hpp>
#include <FILTER_PROJECTION_CUDA.hpp>
#include <pipeline3d_cuda.hpp>
#include <POSE_RANSAC_LM_DIFF_REPROJECTION_DEPTH_CUDA.hpp>

using namespace MopedNS;

static vector<float> readf(FILE *f, size_t n) { vector<float> v(n); if (n && fread(&v[0], 4, n, f) != n) exit(3); return v; }
static vector<int> readi(FILE *f, size_t n) { vector<int> v(n); if (n && fread(&v[0], 4, n, f) != n) exit(3); return v; }

static void dump(const char *tag, const list<SP_Object> &objects) {
	foreach( o, objects ) {
		printf("OBJECT %s %s", tag, o->model->name.c_str());
		for (int j = 0; j < 4; j++) printf(" %.9g", o->pose.rotation[j]);
		for (int j = 0; j < 3; j++) printf(" %.9g", o->pose.translation[j]);
		printf("\n");
	}
}

int main(int argc, char **argv) {
	if (argc < 3) return 2;
	FILE *f = fopen(argv[1], "rb");
	if (!f) return 2;
	const int variant = atoi(argv[2]);
	const int nModels = readi(f, 1)[0];
	vector<float> K = readf(f, 4);
	omp_set_num_threads(1);
	srand(12345);

	vector<SP_Model> models;
	FrameData fdCpu, fdGpu;
	SP_Image im(new Image);
	im->name = "cam"; im->width = 640; im->height = 480;
	im->intrinsicLinearCalibration.init(K[0], K[1], K[2], K[3]);
	im->intrinsicNonlinearCalibration.init(0, 0, 0, 0);
	im->cameraPose.translation.init(0, 0, 0);
	im->cameraPose.rotation.init(0, 0, 0, 1);
	im->TM.init(im->cameraPose);
	fdCpu.images.push_back(im);
	fdCpu.matches.resize(nModels); fdCpu.clusters.resize(nModels);
	for (int m = 0; m < nModels; m++) {
		SP_Model mod(new Model); mod->name = "obj" + toString(m); models.push_back(mod);
		vector<int> hdr = readi(f, 2);
		vector<float> rec = readf(f, 9 * (size_t)hdr[0]);
		fdCpu.matches[m].resize(hdr[0]);
		for (int i = 0; i < hdr[0]; i++) {
			FrameData::Match &ma = fdCpu.matches[m][i];
			ma.imageIdx = 0;
			ma.coord2D.init(rec[9 * i], rec[9 * i + 1]);
			ma.coord3D.init(rec[9 * i + 2], rec[9 * i + 3], rec[9 * i + 4]);
			ma.depthData.depthValid = true;
			ma.depthData.coord3D.init(rec[9 * i + 5], rec[9 * i + 6], rec[9 * i + 7]);
			ma.depthData.depth = rec[9 * i + 7];
			ma.depthData.fillDistance = rec[9 * i + 8];
		}
		fdCpu.clusters[m].resize(hdr[1]);
		for (int c = 0; c < hdr[1]; c++) {
			const int sz = readi(f, 1)[0];
			vector<int> idx = readi(f, sz);
			for (int k = 0; k < sz; k++) fdCpu.clusters[m][c].push_back(idx[k]);
		}
	}
	fclose(f);
	fdGpu.images = fdCpu.images; fdGpu.matches = fdCpu.matches; fdGpu.clusters = fdCpu.clusters;
	list<SP_Object> objCpu, objGpu;
	fdCpu.objects = &objCpu; fdGpu.objects = &objGpu;

	MopedPipeline cpu, gpu;
	if (variant == 3) {                 // registration check of pipeline3d_cuda.hpp: every step of config.hpp:41-49 has its CUDA class
		MopedPipeline all;
		addCudaMatch3d(all); addCudaRecognition3d(all);
		list<MopedAlg *> algs = all.getAlgs(true);
		map<string,string> c;
		foreach( alg, algs ) alg->getConfig(c);
		foreach( kv, c ) printf("CONFIG %s=%s\n", kv.first.c_str(), kv.second.c_str());
		printf("STEP REGISTER algs=%d\n", (int)algs.size());
		return 0;
	}
	if (variant == 0) {
		cpu.addAlg( "POSE", new POSE_RANSAC_LM_DIFF_BACKPROJECTION_DEPTH_CPU( 192, 100, 4, 5, 6, 8, 0.5) );
		gpu.addAlg( "POSE", new POSE_RANSAC_LM_DIFF_BACKPROJECTION_DEPTH_CUDA( 192, 100, 4, 5, 6, 8, 0.5) );
	} else if (variant == 1) {
		cpu.addAlg( "POSE", new POSE_RANSAC_LM_DIFF_REPROJECTION_DEPTH_CPU( 192, 100, 4, 5, 6, 8, 0.5) );
		gpu.addAlg( "POSE", new POSE_RANSAC_LM_DIFF_REPROJECTION_DEPTH_CUDA( 192, 100, 4, 5, 6, 8, 0.5) );
	} else {
		cpu.addAlg( "POSE", new POSE_RANSAC_LM_DIFF_BACKPROJECTION_DEPTH_CPU( 192, 100, 4, 5, 6, 8, 0.5) );
		cpu.addAlg( "FILTER", new FILTER_PROJECTION_CPU( 6, 4096., 2) );
		cpu.addAlg( "POSE2", new POSE_RANSAC_LM_DIFF_BACKPROJECTION_DEPTH_CPU( 64, 250, 4, 6, 8, 5, 0.5) );
		cpu.addAlg( "FILTER2", new FILTER_PROJECTION_CPU( 8, 8192., 1e-4) );
		gpu.addAlg( "POSE", new POSE_RANSAC_LM_DIFF_BACKPROJECTION_DEPTH_CUDA( 192, 100, 4, 5, 6, 8, 0.5) );
		gpu.addAlg( "FILTER", new FILTER_PROJECTION_CUDA( 6, 4096., 2) );
		gpu.addAlg( "POSE2", new POSE_RANSAC_LM_DIFF_BACKPROJECTION_DEPTH_CUDA( 64, 250, 4, 6, 8, 5, 0.5) );
		gpu.addAlg( "FILTER2", new FILTER_PROJECTION_CUDA( 8, 8192., 1e-4) );
	}
	map<string,string> cfg;
	list<MopedAlg *> ca = cpu.getAlgs(true), ga = gpu.getAlgs(true);
	foreach( alg, ga ) { alg->getConfig(cfg); alg->modelsUpdated(models); }
	foreach( alg, ca ) alg->modelsUpdated(models);
	foreach( kv, cfg ) printf("CONFIG %s=%s\n", kv.first.c_str(), kv.second.c_str());
	try {
		foreach( alg, ca ) alg->process(fdCpu);
		dump("cpu", objCpu);
		fflush(stdout);
		foreach( alg, ga ) alg->process(fdGpu);
	} catch (string &e) { fprintf(stderr, "ERROR %s\n", e.c_str()); return 1; }
	dump("cuda", objGpu);
	printf("STEP POSE cpu_objects=%d cuda_objects=%d old_cpu=%d old_cuda=%d\n", (int)objCpu.size(), (int)objGpu.size(),
	       (int)fdCpu.oldObjects.size(), (int)fdGpu.oldObjects.size());
	return 0;
}
