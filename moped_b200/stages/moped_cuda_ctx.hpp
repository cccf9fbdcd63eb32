// moped_cuda_ctx.hpp — one libmoped_cuda context shared by the CUDA stage classes of a process, plus the
// FrameData <-> flat-array helpers they share. C++98-compatible on purpose: inside the reference tree these
// headers are compiled with -std=gnu++98 next to the CPU stages.
#pragma once
#ifndef MOPED_CUDA_CTX_HPP
#define MOPED_CUDA_CTX_HPP

#include <moped_cuda.h>

#include <cstdlib>
#include <iostream>
#include <stdexcept>

namespace MopedNS {

struct MopedCuda {
	// The shared context. Throws std::string like the reference wrappers do (moped2/moped_test.cpp:98) when no
	// B200 is usable: there is no CPU fallback. MOPED_CUDA_DEVICE selects the device.
	static mc_ctx *ctx() {
		static mc_ctx *c = NULL;
		if (!c) {
			const char *dev = getenv("MOPED_CUDA_DEVICE");
			if (mc_create(&c, dev ? atoi(dev) : 0) != MC_OK) throw std::string("libmoped_cuda: ") + mc_last_error(NULL);
		}
		return c;
	}
	static void check(mc_status s, const char *what) {
		if (s != MC_OK) throw std::string("libmoped_cuda: ") + what + ": " + mc_last_error(ctx());
	}
	// cameras of this frame -> device (FrameData::images[i]->{intrinsicLinearCalibration, cameraPose})
	static void setCameras(const std::vector<SP_Image> &images) {
		std::vector<float> K(4 * images.size()), P(7 * images.size());
		for (size_t i = 0; i < images.size(); i++) {
			for (int j = 0; j < 4; j++) K[4 * i + j] = images[i]->intrinsicLinearCalibration[j];
			for (int j = 0; j < 7; j++) P[7 * i + j] = images[i]->cameraPose[j];
		}
		if (!images.empty()) check(mc_set_cameras(ctx(), &K[0], &P[0], (int)images.size()), "mc_set_cameras");
	}
	// matches[model][i] -> CSR over models
	static void flattenMatches(const std::vector<std::vector<FrameData::Match> > &matches, size_t nModels, std::vector<int32_t> &off,
	                           std::vector<int32_t> &img, std::vector<float> &xy, std::vector<float> &xyz) {
		off.assign(nModels + 1, 0);
		for (size_t m = 0; m < nModels; m++) off[m + 1] = off[m] + (m < matches.size() ? (int32_t)matches[m].size() : 0);
		img.resize(off[nModels] + 1); xy.resize(2 * off[nModels] + 2); xyz.resize(3 * off[nModels] + 3);
		for (size_t m = 0; m < nModels && m < matches.size(); m++)
			for (size_t i = 0; i < matches[m].size(); i++) {
				const FrameData::Match &ma = matches[m][i];
				const int t = off[m] + (int)i;
				img[t] = ma.imageIdx;
				xy[2 * t] = ma.coord2D[0]; xy[2 * t + 1] = ma.coord2D[1];
				xyz[3 * t] = ma.coord3D[0]; xyz[3 * t + 1] = ma.coord3D[1]; xyz[3 * t + 2] = ma.coord3D[2];
			}
	}
};

} // namespace MopedNS
#endif
