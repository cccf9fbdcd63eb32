/*
 * ref3d_dropin.cpp — TEST INFRASTRUCTURE: the drop-in proof for moped3d's CLUSTER step. Compiled against moped3d's OWN headers
 * (moped.hpp, util.hpp, CLUSTER_LINKAGE_CPU.hpp; -std=gnu++98 like the reference) together with
 * moped_b200/stages/CLUSTER_LINKAGE_CUDA.hpp — what a maintainer gets after replacing
 *     pipeline.addAlg( "CLUSTER", new CLUSTER_LINKAGE_CPU( 0.1, 7, 2, 1, 0.0, 1, -1, -1) );       (moped3d/libmoped/src/config.hpp:45)
 * by the same line with CLUSTER_LINKAGE_CUDA. Runs both stages in two reference MopedPipelines on identical FrameData (depth map,
 * fill-distance map, matches of several models) and reports whether FrameData::clusters is the same.
 * Case file: int32 {W, H, n_models}, int32 n_matches[n_models], float depth[W*H], float distance[W*H], then per match
 * float {x, y, X, Y, Z, wx, wy, wz}.
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

#include <cluster/CLUSTER_LINKAGE_CPU.hpp>
#include <CLUSTER_LINKAGE_CUDA.hpp>

using namespace MopedNS;

static vector<float> readf(FILE *f, size_t n) { vector<float> v(n); if (n && fread(&v[0], 4, n, f) != n) exit(3); return v; }
static vector<int> readi(FILE *f, size_t n) { vector<int> v(n); if (n && fread(&v[0], 4, n, f) != n) exit(3); return v; }

int main(int argc, char **argv) {
	if (argc < 2) return 2;
	FILE *f = fopen(argv[1], "rb");
	if (!f) return 2;
	vector<int> hdr = readi(f, 3);
	const int W = hdr[0], H = hdr[1], nModels = hdr[2];
	vector<int> nm = readi(f, nModels);
	vector<float> depth = readf(f, (size_t)W * H), distance = readf(f, (size_t)W * H);
	omp_set_num_threads(1);

	vector<SP_Model> models;
	FrameData fdCpu, fdGpu;
	SP_Image dm(new Image(IMAGE_TYPE_DEPTH_MAP));
	dm->name = "cam/depth"; dm->width = W; dm->height = H;
	dm->data.assign((size_t)W * H * 4 * sizeof(Float), 0);
	SP_Image pm(new Image(IMAGE_TYPE_PROB_MAP));
	pm->name = dm->name + ".distance"; pm->width = W; pm->height = H;
	pm->data.assign((size_t)W * H * sizeof(Float), 0);
	for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) { dm->setDepth(x, y, depth[(size_t)y * W + x]); pm->setProb(x, y, distance[(size_t)y * W + x]); }
	fdCpu.images.push_back(dm); fdCpu.images.push_back(pm);
	fdCpu.matches.resize(nModels);
	for (int m = 0; m < nModels; m++) {
		SP_Model mod(new Model); mod->name = "obj" + toString(m); models.push_back(mod);
		vector<float> rec = readf(f, 8 * (size_t)nm[m]);
		fdCpu.matches[m].resize(nm[m]);
		for (int i = 0; i < nm[m]; i++) {
			FrameData::Match &ma = fdCpu.matches[m][i];
			ma.imageIdx = 0;
			ma.coord2D.init(rec[8 * i], rec[8 * i + 1]);
			ma.coord3D.init(rec[8 * i + 2], rec[8 * i + 3], rec[8 * i + 4]);
			ma.depthData.depthValid = true;
			ma.depthData.coord3D.init(rec[8 * i + 5], rec[8 * i + 6], rec[8 * i + 7]);
			ma.depthData.depth = rec[8 * i + 7];
			ma.depthData.fillDistance = 0;
		}
	}
	fclose(f);
	fdGpu.images = fdCpu.images; fdGpu.matches = fdCpu.matches;

	MopedPipeline cpu, gpu;
	cpu.addAlg( "CLUSTER", new CLUSTER_LINKAGE_CPU( 0.1, 7, 2, 1, 0.0, 1, -1, -1) );
	gpu.addAlg( "CLUSTER", new CLUSTER_LINKAGE_CUDA( 0.1, 7, 2, 1, 0.0, 1, -1, -1) );
	map<string,string> cfg;
	list<MopedAlg *> ca = cpu.getAlgs(true), ga = gpu.getAlgs(true);
	foreach( alg, ga ) { alg->getConfig(cfg); alg->modelsUpdated(models); }
	foreach( alg, ca ) alg->modelsUpdated(models);
	foreach( kv, cfg ) printf("CONFIG %s=%s\n", kv.first.c_str(), kv.second.c_str());
	try {
		foreach( alg, ca ) alg->process(fdCpu);
		foreach( alg, ga ) alg->process(fdGpu);
	} catch (string &e) { fprintf(stderr, "ERROR %s\n", e.c_str()); return 1; }
	int nc = 0, ng = 0;
	for (size_t m = 0; m < fdCpu.clusters.size(); m++) nc += (int)fdCpu.clusters[m].size();
	for (size_t m = 0; m < fdGpu.clusters.size(); m++) ng += (int)fdGpu.clusters[m].size();
	printf("STEP CLUSTER same=%d cpu_clusters=%d gpu_clusters=%d old_same=%d\n", (int)(fdCpu.clusters == fdGpu.clusters), nc, ng,
	       (int)(fdCpu.oldClusters == fdGpu.oldClusters));
	return 0;
}
