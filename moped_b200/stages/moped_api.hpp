// moped_api.hpp — stand-alone statement of libmoped's plugin API, for building the CUDA stage classes where
// the reference tree is not available (tests on the GPU box, third-party hosts). Same names, members and
// semantics as the reference's moped2/libmoped/include/moped.hpp:84-365 (value types) and
// moped2/libmoped/src/util.hpp:59-201 (FrameData, MopedAlg, MopedStep, MopedPipeline, GET_CONFIG/SET_CONFIG);
// written from that interface, not copied. Inside the reference tree DO NOT include this file: include the
// reference's own moped.hpp/util.hpp first and then the stage headers — the stage
// headers use only names both provide.
#pragma once
#ifndef MOPED_B200_API_HPP
#define MOPED_B200_API_HPP

#include <cfloat>
#include <cmath>
#include <cstring>
#include <list>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

namespace MopedNS {

using std::list;
using std::map;
using std::string;
using std::vector;

typedef float Float;

template <int N> struct Pt {
	Float p[N];
	Pt<N> &init(Float a) { p[0] = a; return *this; }
	Pt<N> &init(Float a, Float b) { p[0] = a; p[1] = b; return *this; }
	Pt<N> &init(Float a, Float b, Float c) { p[0] = a; p[1] = b; p[2] = c; return *this; }
	Pt<N> &init(Float a, Float b, Float c, Float d) { p[0] = a; p[1] = b; p[2] = c; p[3] = d; return *this; }
	Float &operator[](int n) { return p[n]; }
	const Float &operator[](int n) const { return p[n]; }
	bool operator<(const Pt<N> &o) const { for (int x = 0; x < N; x++) if (p[x] != o.p[x]) return p[x] < o.p[x]; return false; }
	bool operator==(const Pt<N> &o) const { for (int x = 0; x < N; x++) if (p[x] != o.p[x]) return false; return true; }
	Pt<N> &norm() { Float d = 0; for (int x = 0; x < N; x++) d += p[x] * p[x]; d = (Float)(1. / std::sqrt(d)); for (int x = 0; x < N; x++) p[x] *= d; return *this; }
};
typedef Pt<4> Quat;

struct Pose {
	Quat rotation;          // (x, y, z, w)
	Pt<3> translation;
	Float &operator[](int n) { return n < 4 ? rotation[n] : translation[n - 4]; }
	const Float &operator[](int n) const { return n < 4 ? rotation[n] : translation[n - 4]; }
};

struct TransformMatrix {
	Pt<4> p[4];
	void init(const Pose &pose) {
		const Float *q = &pose.rotation[0], *t = &pose.translation[0];
		p[0].init(1 - 2 * q[1] * q[1] - 2 * q[2] * q[2], 2 * q[0] * q[1] - 2 * q[3] * q[2], 2 * q[0] * q[2] + 2 * q[3] * q[1], t[0]);
		p[1].init(2 * q[0] * q[1] + 2 * q[3] * q[2], 1 - 2 * q[0] * q[0] - 2 * q[2] * q[2], 2 * q[1] * q[2] - 2 * q[3] * q[0], t[1]);
		p[2].init(2 * q[0] * q[2] - 2 * q[3] * q[1], 2 * q[1] * q[2] + 2 * q[3] * q[0], 1 - 2 * q[0] * q[0] - 2 * q[1] * q[1], t[2]);
		p[3].init(0, 0, 0, 1);
	}
};

struct Model {
	struct IP { Pt<3> coord3D; vector<float> descriptor; };
	string name;
	map<string, vector<IP> > IPs;
	Pt<3> boundingBox[2];
};
typedef std::shared_ptr<Model> SP_Model;

struct Image {
	vector<unsigned char> data;
	string name;
	int width, height;
	Pt<4> intrinsicLinearCalibration;       // fx, fy, cx, cy
	Pt<4> intrinsicNonlinearCalibration;
	Pose cameraPose;
	TransformMatrix TM;
};
typedef std::shared_ptr<Image> SP_Image;

struct Object {
	SP_Model model;
	Pose pose;
	Float score;
};
typedef std::shared_ptr<Object> SP_Object;

template <typename T> static inline string toString(const T &v) { std::stringstream s; s << v; return s.str(); }
template <typename T> static inline bool fromString(T &var, string &s) { T t = var; std::istringstream is(s.c_str()); is >> var; return t == var; }

// config key = "<step>:<alg index>:<header basename>/<member>", e.g. MATCH_SIFT:0:MATCH_CUDA/Ratio
static inline string moped_cfg_key(const string &step, int alg, const char *file, const char *var) {
	string f(file);
	size_t s = f.find_last_of("/\\");
	f = (s == string::npos) ? f : f.substr(s + 1);
	if (f.size() > 4) f = f.substr(0, f.size() - 4);
	return toString(step) + ":" + toString(alg) + ":" + f + "/" + var;
}
#define GET_CONFIG(varName) config[MopedNS::moped_cfg_key(_stepName, _alg, __FILE__, #varName)] = MopedNS::toString(varName)
#define SET_CONFIG(varName) configUpdated = MopedNS::fromString(varName, config[MopedNS::moped_cfg_key(_stepName, _alg, __FILE__, #varName)]) || configUpdated

struct FrameData {
	struct DetectedFeature { int imageIdx; Pt<2> coord2D; vector<float> descriptor; };
	struct Match { int imageIdx; Pt<2> coord2D; Pt<3> coord3D; };
	typedef list<int> Cluster;
	vector<SP_Image> images;
	map<string, vector<DetectedFeature> > detectedFeatures;
	vector<vector<Match> > matches;
	vector<vector<Cluster> > clusters;
	list<SP_Object> *objects;
	int correctMatches, incorrectMatches;
	vector<vector<Cluster> > oldClusters;
	list<SP_Object> oldObjects;
	map<string, Float> times;
};

class MopedAlg {
public:
	vector<SP_Model> *models;
	bool capable;
	bool configUpdated;
	string _stepName;
	int _alg;
	MopedAlg() : models(0), capable(true), configUpdated(true), _alg(0) {}
	virtual ~MopedAlg() {}
	bool isCapable() const { return capable; }
	void setStepNameAndAlg(string &stepName, int alg) { _stepName = stepName; _alg = alg; }
	virtual void modelsUpdated(vector<SP_Model> &_models) { models = &_models; configUpdated = true; }
	virtual void getConfig(map<string, string> &config) const {}
	virtual void setConfig(map<string, string> &config) {}
	virtual void process(FrameData &frameData) = 0;
};

struct MopedStep : public vector<std::shared_ptr<MopedAlg> > {
	MopedAlg *getAlg() {
		for (size_t i = 0; i < size(); i++) if ((*this)[i]->isCapable()) return (*this)[i].get();
		return 0;
	}
};

struct MopedPipeline : public vector<MopedStep> {
	map<string, int> fromStepNameToIndex;
	void addAlg(string stepName, MopedAlg *alg) {
		int step;
		if (fromStepNameToIndex.find(stepName) == fromStepNameToIndex.end()) { step = (int)fromStepNameToIndex.size(); fromStepNameToIndex[stepName] = step; }
		else step = fromStepNameToIndex[stepName];
		if (step >= (int)size()) resize(step + 1);
		alg->setStepNameAndAlg(stepName, (int)(*this)[step].size());
		(*this)[step].push_back(std::shared_ptr<MopedAlg>(alg));
	}
	list<MopedAlg *> getAlgs(bool onlyActive = false) {
		list<MopedAlg *> algs;
		for (size_t s = 0; s < size(); s++) {
			if (!onlyActive) { for (size_t a = 0; a < (*this)[s].size(); a++) algs.push_back((*this)[s][a].get()); }
			else if ((*this)[s].getAlg()) algs.push_back((*this)[s].getAlg());
		}
		return algs;
	}
};

} // namespace MopedNS
#endif
