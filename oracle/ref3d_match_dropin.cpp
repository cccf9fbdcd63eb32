/*
 * ref3d_match_dropin.cpp — TEST INFRASTRUCTURE: the drop-in proof for moped3d's MATCH step. Compiled against moped3d's OWN headers
 * (moped.hpp, util.hpp, MATCH_ADAPTIVE_FLANN_CPU.hpp; -std=gnu++98) together with moped_b200/stages/MATCH_ADAPTIVE_CUDA.hpp.
 * The reference class sits on OpenCV's FLANN (external, unpinned, absent here — SURVEY.md §8c): this file supplies a stand-in
 * cv::flann::Index / cv::Mat whose knnSearch is an EXHAUSTIVE search in the arithmetic of the reference's exact matcher
 * (sequential fp32 sum of squared differences; ties to the lower row), so that everything else the class does — normalisation,
 * the per-model control points from bounding box / intrinsics / feature count, the depth- and fill-distance-dependent ratio
 * threshold, the depth cut, match assembly — runs UNMODIFIED and pins the CUDA class's host logic.
 * Both stages run on identical FrameData generated from argv[1] (seed); with a GPU the CUDA class runs as shipped
 * (mc_match_adaptive: search and threshold on the device), without one (argv[2] = "host") its host side runs as shipped and the
 * device's two jobs are stood in for by the same exhaustive search and by the kernel's decision header compiled for the host
 * (moped_b200/csrc/adaptive_threshold.cuh). Built with strict IEEE flags so that the two sides' float expressions round alike.
 */
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <cfloat>
#include <stdint.h>
#include <vector>

// ---- stand-in for the OpenCV symbols MATCH_ADAPTIVE_FLANN_CPU uses --------------------------------------------------
#define CV_32F 5
namespace cv {
	struct Mat {
		int rows, cols; std::vector<float> d;
		Mat() : rows(0), cols(0) {}
		Mat(int r, int c, int) : rows(r), cols(c), d((size_t)r * c) {}
		template <typename T> T &at(int r, int c) { return d[(size_t)r * cols + c]; }
	};
	namespace flann {
		struct KDTreeIndexParams { KDTreeIndexParams(int) {} };
		struct SearchParams { SearchParams(int) {} };
		static void exhaustive2nn(const std::vector<float> &data, int dim, const float *q, int *nx, float *dx) {
			int b0 = -1, b1 = -1; float d0 = FLT_MAX, d1 = FLT_MAX;
			const int n = (int)(data.size() / dim);
			for (int r = 0; r < n; r++) {
				float s = 0;
				for (int k = 0; k < dim; k++) { float t = q[k] - data[(size_t)r * dim + k]; s = s + t * t; }
				if (s < d0) { d1 = d0; b1 = b0; d0 = s; b0 = r; }
				else if (s < d1) { d1 = s; b1 = r; }
			}
			nx[0] = b0; nx[1] = b1; dx[0] = d0; dx[1] = d1;
		}
		struct Index {
			std::vector<float> data; int dim;
			Index(const Mat &m, const KDTreeIndexParams &) : data(m.d), dim(m.cols) {}
			void knnSearch(const std::vector<float> &q, std::vector<int> &nx, std::vector<float> &dx, int, const SearchParams &) {
				exhaustive2nn(data, dim, &q[0], &nx[0], &dx[0]);
			}
		};
	}
}

#include <moped.hpp>
#include <util.hpp>

#ifndef MAX_THREADS
#define MAX_THREADS 64
#endif

#include <match/MATCH_ADAPTIVE_FLANN_CPU.hpp>
#include <MATCH_ADAPTIVE_CUDA.hpp>
// "host" mode only: the device kernel's per-feature decision, compiled for the host from the kernel's own header
#include "../moped_b200/csrc/adaptive_threshold.cuh"

using namespace MopedNS;

static uint64_t g_state;
static double urand() { g_state = g_state * 6364136223846793005ULL + 1442695040888963407ULL; return (double)((g_state >> 11) & ((1ULL << 53) - 1)) / (double)(1ULL << 53); }

static void dump(const char *tag, const vector< vector< FrameData::Match > > &matches) {
	for (size_t m = 0; m < matches.size(); m++)
		for (size_t i = 0; i < matches[m].size(); i++) {
			const FrameData::Match &ma = matches[m][i];
			printf("MATCH %s %d %d %.9g %.9g %.9g %.9g %.9g\n", tag, (int)m, ma.imageIdx, ma.coord2D[0], ma.coord2D[1], ma.coord3D[0], ma.coord3D[1], ma.coord3D[2]);
		}
}

int main(int argc, char **argv) {
	if (argc < 2) return 2;
	g_state = strtoull(argv[1], NULL, 10) * 2654435761ULL + 12345;
	const bool hostOnly = argc > 2 && !strcmp(argv[2], "host");
	omp_set_num_threads(1);
	const int D = 128, W = 160, H = 120;

	// models: feature counts on both sides of the density sigmoid's centre (1750), bounding boxes from 5 cm to 40 cm
	const int counts[5] = { 300, 1200, 1750, 2400, 40 };
	vector<SP_Model> modelsCpu, modelsGpu;
	for (int m = 0; m < 5; m++) {
		SP_Model a(new Model), b(new Model);
		a->name = b->name = "obj" + toString(m);
		Float ext[3] = { (Float)(0.05 + 0.35 * urand()), (Float)(0.05 + 0.35 * urand()), (Float)(0.05 + 0.35 * urand()) };
		a->boundingBox[0].init(-ext[0] / 2, -ext[1] / 2, -ext[2] / 2); a->boundingBox[1].init(ext[0] / 2, ext[1] / 2, ext[2] / 2);
		b->boundingBox[0] = a->boundingBox[0]; b->boundingBox[1] = a->boundingBox[1];
		vector<Model::IP> &ipa = a->IPs["SIFT"];
		ipa.resize(counts[m]);
		for (int f = 0; f < counts[m]; f++) {
			ipa[f].descriptor.resize(D);
			for (int k = 0; k < D; k++) { double u = urand(); ipa[f].descriptor[k] = (float)(u * u * u); }
			ipa[f].coord3D.init((Float)(ext[0] * (urand() - 0.5)), (Float)(ext[1] * (urand() - 0.5)), (Float)(ext[2] * (urand() - 0.5)));
		}
		b->IPs["SIFT"] = ipa;
		modelsCpu.push_back(a); modelsGpu.push_back(b);
	}

	FrameData fdCpu;
	SP_Image gray(new Image(IMAGE_TYPE_GRAY_IMAGE));
	gray->name = "cam"; gray->width = W; gray->height = H;
	gray->intrinsicLinearCalibration.init(525.0 / 4, 525.0 / 4, 319.5 / 4, 239.5 / 4);
	SP_Image dm(new Image(IMAGE_TYPE_DEPTH_MAP));
	dm->name = "cam/depth"; dm->width = W; dm->height = H;
	dm->data.assign((size_t)(W + 1) * (H + 1) * 4 * sizeof(Float), 0);
	SP_Image pm(new Image(IMAGE_TYPE_PROB_MAP));
	pm->name = dm->name + ".distance"; pm->width = W; pm->height = H;
	pm->data.assign((size_t)(W + 1) * (H + 1) * sizeof(Float), 0);
	for (int y = 0; y < H; y++)
		for (int x = 0; x < W; x++) {
			dm->setDepth(x, y, (Float)(0.3 + 4.5 * urand() * urand()));            // some beyond MaximumDepth = 4
			pm->setProb(x, y, urand() < 0.6 ? (Float)0 : (Float)(0.5 * urand()));   // measured depth / filled in from up to 0.5 away
		}
	fdCpu.images.push_back(gray); fdCpu.images.push_back(dm); fdCpu.images.push_back(pm);

	// features: noisy copies of model features (a range of noise levels => ratios on both sides of the thresholds) and clutter
	vector<FrameData::DetectedFeature> &feats = fdCpu.detectedFeatures["SIFT"];
	const int Q = 900;
	feats.resize(Q);
	for (int i = 0; i < Q; i++) {
		feats[i].imageIdx = 0;
		feats[i].coord2D.init((Float)(urand() * (W - 1)), (Float)(urand() * (H - 1)));
		feats[i].descriptor.resize(D);
		if (i % 3) {
			const int m = (int)(urand() * 5) % 5;
			const vector<Model::IP> &ips = modelsCpu[m]->IPs["SIFT"];
			const vector<float> &src = ips[(int)(urand() * ips.size()) % ips.size()].descriptor;
			const double noise = 0.02 + 0.5 * urand();
			for (int k = 0; k < D; k++) { double v = src[k] + noise * (urand() - 0.5); feats[i].descriptor[k] = (float)(v < 0 ? 0 : v); }
		} else
			for (int k = 0; k < D; k++) { double u = urand(); feats[i].descriptor[k] = (float)(u * u * u); }
	}
	FrameData fdGpu;
	fdGpu.images = fdCpu.images;
	fdGpu.detectedFeatures = fdCpu.detectedFeatures;

	MopedPipeline cpu, gpu;
	cpu.addAlg( "MATCH_SIFT", new MATCH_ADAPTIVE_FLANN_CPU( 128, "SIFT", 8, 0.6, 0.75, 0.65, 0.8, 150, 50) );     // moped3d/libmoped/src/config.hpp:41
	MATCH_ADAPTIVE_CUDA *cu = new MATCH_ADAPTIVE_CUDA( 128, "SIFT", 8, 0.6, 0.75, 0.65, 0.8, 150, 50);
	gpu.addAlg( "MATCH_SIFT", cu );
	map<string,string> cfg;
	list<MopedAlg *> ca = cpu.getAlgs(true), ga = gpu.getAlgs(true);
	foreach( alg, ga ) { alg->getConfig(cfg); alg->modelsUpdated(modelsGpu); }
	foreach( alg, ca ) alg->modelsUpdated(modelsCpu);
	foreach( kv, cfg ) printf("CONFIG %s=%s\n", kv.first.c_str(), kv.second.c_str());
	try {
		foreach( alg, ca ) alg->process(fdCpu);
		if (!hostOnly) { foreach( alg, ga ) alg->process(fdGpu); }
		else if (cu->prepare(fdGpu, false)) {
			// no device: the class's host side (Update -> rows and ratio curves, gather, emit) runs as shipped; the two things the device does
			// are stood in for by the exhaustive search above and by adaptive_threshold.cuh compiled for the host
			MATCH_ADAPTIVE_CUDA::FrameInputs in;
			cu->gather(fdGpu, in);
			vector<float> dataset;
			for (size_t m = 0; m < modelsGpu.size(); m++) {
				vector<Model::IP> &ips = modelsGpu[m]->IPs["SIFT"];
				for (size_t f = 0; f < ips.size(); f++) dataset.insert(dataset.end(), ips[f].descriptor.begin(), ips[f].descriptor.end());
			}
			vector<int32_t> nnRow(2 * (size_t)Q); vector<float> nnDist(2 * (size_t)Q); vector<uint8_t> accepted(Q);
			mc::AdaptiveParams P; P.maximum_depth = 4.0f; P.default_depth = 1.0f; P.cauchy_scale = 0.1f;
			for (int i = 0; i < Q; i++) {
				int nx[2];
				cv::flann::exhaustive2nn(dataset, D, &in.desc[(size_t)i * D], nx, &nnDist[2 * i]);
				nnRow[2 * i] = nx[0]; nnRow[2 * i + 1] = nx[1];
				const int px = mc::adaptive_pixel(in.xy[2 * i], in.xy[2 * i + 1], in.width, in.height);
				accepted[i] = mc::adaptive_accept(cu->modelCurves()[cu->modelOfRow(nx[0])], in.depth[px], in.fill[px], nnDist[2 * i], nnDist[2 * i + 1], P) ? 1 : 0;
			}
			cu->emit(fdGpu, nnRow, accepted);
		}
	} catch (string &e) { fprintf(stderr, "ERROR %s\n", e.c_str()); return 1; }
	dump("cpu", fdCpu.matches);
	dump(hostOnly ? "host" : "cuda", fdGpu.matches);
	bool same = fdCpu.matches.size() == fdGpu.matches.size(), descSame = true, modelSame = true;
	size_t total = 0;
	for (size_t m = 0; same && m < fdCpu.matches.size(); m++) {
		same = fdCpu.matches[m].size() == fdGpu.matches[m].size();
		for (size_t i = 0; same && i < fdCpu.matches[m].size(); i++)
			same = !memcmp(&fdCpu.matches[m][i].coord2D, &fdGpu.matches[m][i].coord2D, sizeof(Pt<2>)) &&
			       !memcmp(&fdCpu.matches[m][i].coord3D, &fdGpu.matches[m][i].coord3D, sizeof(Pt<3>)) && fdCpu.matches[m][i].imageIdx == fdGpu.matches[m][i].imageIdx;
		total += fdCpu.matches[m].size();
	}
	for (int i = 0; i < Q; i++) descSame = descSame && fdCpu.detectedFeatures["SIFT"][i].descriptor == fdGpu.detectedFeatures["SIFT"][i].descriptor;
	for (size_t m = 0; m < modelsCpu.size(); m++)
		for (size_t f = 0; f < modelsCpu[m]->IPs["SIFT"].size(); f++) modelSame = modelSame && modelsCpu[m]->IPs["SIFT"][f].descriptor == modelsGpu[m]->IPs["SIFT"][f].descriptor;
	printf("STEP MATCH same=%d matches=%d features=%d normalised_features_same=%d normalised_models_same=%d\n", (int)same, (int)total, Q, (int)descSame, (int)modelSame);
	return 0;
}
