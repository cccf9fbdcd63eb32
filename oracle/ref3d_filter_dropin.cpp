/*
 * ref3d_filter_dropin.cpp — TEST INFRASTRUCTURE: the drop-in proof for moped3d's FILTER_PROJECTION_DEPTH step. Compiled against moped3d's
 * OWN headers (moped.hpp, util.hpp, filter/FILTER_PROJECTION_DEPTH_CPU.hpp; -std=gnu++98, strict IEEE flags so that both sides' float
 * expressions round alike) together with moped_b200/stages/FILTER_PROJECTION_DEPTH_CUDA.hpp — what a maintainer gets after registering
 *     pipeline.addAlg( "FILTER", new FILTER_PROJECTION_DEPTH_CUDA( 5, 4096., 16384., 2., 0.05, 40, 0.2 ) );
 * in place of the CPU class. Runs both stages in two reference MopedPipelines on identical FrameData (camera image, depth map,
 * fill-distance map, matches of several models, an object list) with the same srand() before each, and reports whether the surviving
 * objects (identity, order, score bits) and FrameData::clusters are the same.
 * Case file: int32 {W, H, n_models, n_objects, test_sample_size}, float K[4], int32 n_pts[n_models], int32 n_matches[n_models],
 * float depth[W*H], float distance[W*H], per model float xyz[3 n_pts], per match float {x, y, X, Y, Z}, per object int32 model,
 * float pose[7].
 */
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <cfloat>
#include <stdint.h>
#include <iostream>

#include <moped.hpp>
#include <util.hpp>

#ifndef MAX_THREADS
#define MAX_THREADS 64
#endif

#include <filter/FILTER_PROJECTION_DEPTH_CPU.hpp>
#include <FILTER_PROJECTION_DEPTH_CUDA.hpp>

using namespace MopedNS;

static vector<float> readf(FILE *f, size_t n) { vector<float> v(n); if (n && fread(&v[0], 4, n, f) != n) exit(3); return v; }
static vector<int> readi(FILE *f, size_t n) { vector<int> v(n); if (n && fread(&v[0], 4, n, f) != n) exit(3); return v; }

int main(int argc, char **argv) {
	if (argc < 2) return 2;
	FILE *f = fopen(argv[1], "rb");
	if (!f) return 2;
	vector<int> hdr = readi(f, 5);
	const int W = hdr[0], H = hdr[1], nModels = hdr[2], nObj = hdr[3], sample = hdr[4];
	vector<float> K = readf(f, 4);
	vector<int> np = readi(f, nModels), nm = readi(f, nModels);
	vector<float> depth = readf(f, (size_t)W * H), distance = readf(f, (size_t)W * H);
	omp_set_num_threads(1);

	vector<SP_Model> models;
	FrameData fdCpu, fdGpu;
	SP_Image im(new Image(IMAGE_TYPE_GRAY_IMAGE));
	im->name = "cam"; im->width = W; im->height = H;
	im->intrinsicLinearCalibration.init(K[0], K[1], K[2], K[3]);
	im->cameraPose.rotation.init(0, 0, 0, 1); im->cameraPose.translation.init(0, 0, 0);
	im->TM.init(im->cameraPose);
	SP_Image dm(new Image(IMAGE_TYPE_DEPTH_MAP));
	dm->name = "cam/depth"; dm->width = W; dm->height = H;
	dm->intrinsicLinearCalibration.init(K[0], K[1], K[2], K[3]);
	dm->cameraPose.rotation.init(0, 0, 0, 1); dm->cameraPose.translation.init(0, 0, 0);
	dm->TM.init(dm->cameraPose);
	dm->data.assign((size_t)W * H * 4 * sizeof(Float), 0);
	SP_Image pm(new Image(IMAGE_TYPE_PROB_MAP));
	pm->name = dm->name + ".distance"; pm->width = W; pm->height = H;
	pm->data.assign((size_t)W * H * sizeof(Float), 0);
	for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) { dm->setDepth(x, y, depth[(size_t)y * W + x]); pm->setProb(x, y, distance[(size_t)y * W + x]); }
	fdCpu.images.push_back(im); fdCpu.images.push_back(dm); fdCpu.images.push_back(pm);
	for (int m = 0; m < nModels; m++) {
		SP_Model mod(new Model); mod->name = "obj" + toString(m);
		vector<float> xyz = readf(f, 3 * (size_t)np[m]);
		vector<Model::IP> &ips = mod->IPs["SIFT"];
		ips.resize(np[m]);
		for (int i = 0; i < np[m]; i++) ips[i].coord3D.init(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
		models.push_back(mod);
	}
	fdCpu.matches.resize(nModels);
	for (int m = 0; m < nModels; m++) {
		vector<float> rec = readf(f, 5 * (size_t)nm[m]);
		fdCpu.matches[m].resize(nm[m]);
		for (int i = 0; i < nm[m]; i++) {
			FrameData::Match &ma = fdCpu.matches[m][i];
			ma.imageIdx = 0;
			ma.coord2D.init(rec[5 * i], rec[5 * i + 1]);
			ma.coord3D.init(rec[5 * i + 2], rec[5 * i + 3], rec[5 * i + 4]);
		}
	}
	list<SP_Object> objCpu, objGpu;
	vector<Object *> idCpu, idGpu;
	for (int o = 0; o < nObj; o++) {
		vector<int> mi = readi(f, 1);
		vector<float> p = readf(f, 7);
		for (int side = 0; side < 2; side++) {
			SP_Object ob(new Object);
			ob->model = models[mi[0]];
			ob->pose.rotation.init(p[0], p[1], p[2], p[3]);
			ob->pose.translation.init(p[4], p[5], p[6]);
			ob->score = 0;
			(side ? objGpu : objCpu).push_back(ob);
			(side ? idGpu : idCpu).push_back(ob.get());
		}
	}
	fclose(f);
	fdGpu.images = fdCpu.images; fdGpu.matches = fdCpu.matches;
	fdCpu.objects = &objCpu; fdGpu.objects = &objGpu;

	MopedPipeline cpu, gpu;
	cpu.addAlg( "FILTER", new FILTER_PROJECTION_DEPTH_CPU( 5, 4096., 16384., 2., 0.05, sample, 0.2 ) );
	gpu.addAlg( "FILTER", new FILTER_PROJECTION_DEPTH_CUDA( 5, 4096., 16384., 2., 0.05, sample, 0.2 ) );
	map<string,string> cfg;
	list<MopedAlg *> ca = cpu.getAlgs(true), ga = gpu.getAlgs(true);
	foreach( alg, ga ) { alg->getConfig(cfg); alg->modelsUpdated(models); }
	foreach( alg, ca ) alg->modelsUpdated(models);
	foreach( kv, cfg ) printf("CONFIG %s=%s\n", kv.first.c_str(), kv.second.c_str());
	try {
		std::streambuf *saved = std::cerr.rdbuf(NULL);           /* the CPU class narrates every object on stderr */
		srand(12345);
		foreach( alg, ca ) alg->process(fdCpu);
		std::cerr.rdbuf(saved); std::cerr.clear();
		srand(12345);
		foreach( alg, ga ) alg->process(fdGpu);
	} catch (string &e) { fprintf(stderr, "ERROR %s\n", e.c_str()); return 1; }
	/* survivors by their position in the input list + score bits */
	bool same = objCpu.size() == objGpu.size();
	list<SP_Object>::iterator a = objCpu.begin(), b = objGpu.begin();
	for (; same && a != objCpu.end(); ++a, ++b) {
		int ia = -1, ib = -1;
		for (int o = 0; o < nObj; o++) { if (idCpu[o] == a->get()) ia = o; if (idGpu[o] == b->get()) ib = o; }
		Float sa = (*a)->score, sb = (*b)->score;
		same = ia == ib && memcmp(&sa, &sb, sizeof sa) == 0;
	}
	printf("STEP FILTER same_objects=%d cpu_objects=%d gpu_objects=%d same_clusters=%d input_objects=%d\n", (int)same, (int)objCpu.size(),
	       (int)objGpu.size(), (int)(fdCpu.clusters == fdGpu.clusters), nObj);
	return 0;
}
