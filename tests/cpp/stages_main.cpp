// stages_main.cpp — the CUDA stage classes driven through the plugin API exactly like
// MopedPimpl::processImages drives the CPU ones (moped2/libmoped/src/moped.cpp:166-194), built against
// moped_api.hpp (no reference tree needed). Reads a binary case written by tests/test_gpu_cpp_stages.py,
// prints the recognised objects as text.
#include <cstdio>
#include <cstdlib>
#include <moped_api.hpp>
#include <pipeline_cuda.hpp>
#include <FEAT_SIFT_CUDA.hpp>
#include <cstring>

using namespace MopedNS;

static std::vector<float> readf(FILE *f, size_t n) { std::vector<float> v(n); if (n && fread(&v[0], 4, n, f) != n) exit(3); return v; }
static std::vector<int> readi(FILE *f, size_t n) { std::vector<int> v(n); if (n && fread(&v[0], 4, n, f) != n) exit(3); return v; }

// `stages_main --sift case.bin out.bin`: FEAT_SIFT_CUDA through the plugin API. case.bin = int32 {height, width, n_images,
// double_size} + pixels; out.bin = int32 n, then per feature {int32 imageIdx, float x, float y, float desc[128]}.
static int sift_mode(const char *path, const char *out) {
	FILE *f = fopen(path, "rb");
	if (!f) return 2;
	std::vector<int> hdr = readi(f, 4);
	FrameData fd;
	for (int i = 0; i < hdr[2]; i++) {
		SP_Image im(new Image);
		im->width = hdr[1]; im->height = hdr[0];
		im->data.resize((size_t)hdr[0] * hdr[1]);
		if (fread(&im->data[0], 1, im->data.size(), f) != im->data.size()) return 3;
		fd.images.push_back(im);
	}
	fclose(f);
	MopedPipeline pipeline;
	pipeline.addAlg("SIFT", new FEAT_SIFT_CUDA(hdr[3] ? "-1" : "0"));
	std::map<std::string, std::string> config;
	std::list<MopedAlg *> algs = pipeline.getAlgs(true);
	for (std::list<MopedAlg *>::iterator a = algs.begin(); a != algs.end(); ++a) (*a)->getConfig(config);
	for (std::map<std::string, std::string>::iterator c = config.begin(); c != config.end(); ++c) printf("CONFIG %s=%s\n", c->first.c_str(), c->second.c_str());
	try {
		for (std::list<MopedAlg *>::iterator a = algs.begin(); a != algs.end(); ++a) (*a)->process(fd);
	} catch (std::string &e) { fprintf(stderr, "ERROR %s\n", e.c_str()); return 1; }
	std::vector<FrameData::DetectedFeature> &feats = fd.detectedFeatures["SIFT"];
	FILE *o = fopen(out, "wb");
	if (!o) return 2;
	int n = (int)feats.size();
	fwrite(&n, 4, 1, o);
	for (int i = 0; i < n; i++) {
		float xy[2] = { feats[i].coord2D[0], feats[i].coord2D[1] };
		fwrite(&feats[i].imageIdx, 4, 1, o); fwrite(xy, 4, 2, o); fwrite(&feats[i].descriptor[0], 4, 128, o);
	}
	fclose(o);
	printf("SIFT features %d\n", n);
	return 0;
}

int main(int argc, char **argv) {
	if (argc < 2) return 2;
	if (argc >= 4 && !strcmp(argv[1], "--sift")) return sift_mode(argv[2], argv[3]);
	FILE *f = fopen(argv[1], "rb");
	if (!f) return 2;
	std::vector<int> hdr = readi(f, 4);                   // n_models, N, Q, D
	const int nModels = hdr[0], N = hdr[1], Q = hdr[2], D = hdr[3];
	std::vector<int> nPts = readi(f, nModels);
	std::vector<float> xyz = readf(f, 3 * (size_t)N), desc = readf(f, (size_t)N * D), qd = readf(f, (size_t)Q * D), qxy = readf(f, 2 * (size_t)Q);
	fclose(f);

	std::vector<SP_Model> models;
	size_t row = 0;
	for (int m = 0; m < nModels; m++) {
		SP_Model mod(new Model);
		mod->name = "obj" + toString(m);
		std::vector<Model::IP> &ips = mod->IPs["SIFT"];
		ips.resize(nPts[m]);
		for (int i = 0; i < nPts[m]; i++, row++) {
			ips[i].coord3D.init(xyz[3 * row], xyz[3 * row + 1], xyz[3 * row + 2]);
			ips[i].descriptor.assign(desc.begin() + row * D, desc.begin() + (row + 1) * D);
		}
		models.push_back(mod);
	}
	SP_Image im(new Image);
	im->width = 640; im->height = 480;
	im->intrinsicLinearCalibration.init(800.f, 800.f, 320.f, 240.f);
	im->cameraPose.rotation.init(0.f, 0.f, 0.f, 1.f); im->cameraPose.translation.init(0.f, 0.f, 0.f);
	im->TM.init(im->cameraPose);

	MopedPipeline pipeline;
	createCudaRecognitionPipeline(pipeline);
	std::map<std::string, std::string> config;
	std::list<MopedAlg *> all = pipeline.getAlgs();
	for (std::list<MopedAlg *>::iterator a = all.begin(); a != all.end(); ++a) { (*a)->getConfig(config); (*a)->modelsUpdated(models); }
	for (std::map<std::string, std::string>::iterator c = config.begin(); c != config.end(); ++c) printf("CONFIG %s=%s\n", c->first.c_str(), c->second.c_str());

	for (int rep = 0; rep < 2; rep++) {
		std::list<SP_Object> objects;
		FrameData fd;
		fd.objects = &objects;
		fd.images.push_back(im);
		std::vector<FrameData::DetectedFeature> &feats = fd.detectedFeatures["SIFT"];
		feats.resize(Q);
		for (int i = 0; i < Q; i++) {
			feats[i].imageIdx = 0;
			feats[i].coord2D.init(qxy[2 * i], qxy[2 * i + 1]);
			feats[i].descriptor.assign(qd.begin() + (size_t)i * D, qd.begin() + (size_t)(i + 1) * D);
		}
		std::list<MopedAlg *> algs = pipeline.getAlgs(true);
		try {
			for (std::list<MopedAlg *>::iterator a = algs.begin(); a != algs.end(); ++a) (*a)->process(fd);
		} catch (std::string &e) { fprintf(stderr, "ERROR %s\n", e.c_str()); return 1; }
		size_t nm = 0; for (size_t m = 0; m < fd.matches.size(); m++) nm += fd.matches[m].size();
		printf("FRAME %d matches %zu objects %zu\n", rep, nm, objects.size());
		for (std::list<SP_Object>::iterator o = objects.begin(); o != objects.end(); ++o)
			printf("OBJECT %s %.6f %.6f %.6f %.6f %.6f %.6f %.6f score %.4f\n", (*o)->model->name.c_str(), (*o)->pose[0], (*o)->pose[1], (*o)->pose[2],
			       (*o)->pose[3], (*o)->pose[4], (*o)->pose[5], (*o)->pose[6], (*o)->score);
	}
	return 0;
}
