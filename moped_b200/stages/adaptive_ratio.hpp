// adaptive_ratio.hpp — the depth- and model-dependent ratio threshold of moped3d's MATCH_ADAPTIVE_FLANN_CPU
// (moped3d/libmoped/src/match/MATCH_ADAPTIVE_FLANN_CPU.hpp:88-372) as host code shared by MATCH_ADAPTIVE_CUDA.hpp and its CPU
// check (oracle/ref3d_match_dropin.cpp). Per model, Update() derives four control points from the model's bounding box, the
// camera intrinsics and the model's feature count (:139-170); per feature, the threshold blends the ratio at the feature's depth
// with the ratio at a default depth by a Cauchy weight of the depth map's fill distance (:360-372). Float = float as in the
// reference; literals are double exactly where the reference's are. C++98-compatible.
#pragma once
#include <cmath>
#include <vector>

namespace MopedNS {

struct AdaptiveRatio {

	// constants Update() assigns (:107-113)
	Float sigmoidTranslate, sigmoidScale, MaximumDepth, DefaultDepth, CauchyScale;
	// per model, in push order. The reference never clears these vectors: a second Update() appends, and lookups by model
	// number keep hitting the FIRST generation (:158-170). Kept: clear() is only called where the reference constructs the object.
	std::vector<float> maxRatioDepths, minRatioDepths, ratioLows, ratioHighs;

	AdaptiveRatio() : sigmoidTranslate(1750), sigmoidScale(250), MaximumDepth(4.0), DefaultDepth(1.0), CauchyScale(0.1) {}

	static Float canonicalSigmoid( Float x ) { return 1.0 / (1.0 + exp(-1.0 * x)); }                            // :178-180

	static Float getProjectedArea( Pt<4> k, Pt<3> points[4] ) {                                           // :235-252
		Pt<2> projectedPoint;
		Pt<2> minVal, maxVal;
		for( int i = 0; i < 4; i++ ) {
			Pt<3> p = points[i];
			Float u = (k[0]*p[0] + k[2]*p[2]) / p[2];
			Float v = (k[1]*p[1] + k[3]*p[2]) / p[2];
			projectedPoint.init(u,v);
			if( i == 0 ) { minVal = projectedPoint; maxVal = projectedPoint; }
			else { minVal = min(minVal, projectedPoint); maxVal = max(maxVal, projectedPoint); }
		}
		return (maxVal[0] - minVal[0])*(maxVal[1] - minVal[1]);
	}

	static Float getAverageProjectedLength( Pt<3> bbox[2], Pt<4> camIntrinsics, Float depth ) {    // :261-312
		Float minX = bbox[0][0], maxX = bbox[1][0], minY = bbox[0][1], maxY = bbox[1][1], minZ = bbox[0][2], maxZ = bbox[1][2];
		Float xRange = (maxX - minX), yRange = (maxY - minY), zRange = (maxZ - minZ);
		minX = -xRange/2; maxX = xRange/2;
		minY = -yRange/2; maxY = yRange/2;
		minZ = -zRange/2; maxZ = zRange/2;
		Pt<3> surface[4];
		Float zConstSurf = (maxX - minX)*(maxY - minY), yConstSurf = (maxX - minX)*(maxZ - minZ), xConstSurf = (maxY - minY)*(maxZ - minZ);
		if( (zConstSurf >= xConstSurf) && (zConstSurf >= yConstSurf) ) {
			surface[0].init(minX, minY, depth); surface[1].init(minX, maxY, depth);
			surface[2].init(maxX, maxY, depth); surface[3].init(maxX, minY, depth);
			return sqrt(getProjectedArea(camIntrinsics, surface));
		} else if( (yConstSurf >= xConstSurf) && (yConstSurf >= zConstSurf) ) {
			surface[0].init(minX, minZ, depth); surface[1].init(minX, maxZ, depth);
			surface[2].init(maxX, maxZ, depth); surface[3].init(maxX, minZ, depth);
			return sqrt(getProjectedArea(camIntrinsics, surface));
		}
		surface[0].init(minY, minZ, depth); surface[1].init(minY, maxZ, depth);
		surface[2].init(maxY, maxZ, depth); surface[3].init(maxY, minZ, depth);
		return sqrt(getProjectedArea(camIntrinsics, surface));
	}

	// depth at which the model's largest face projects to about targetLength pixels: doubling, then bisection (:321-357)
	static Float solveProjectionDepth( Pt<3> bbox[2], Pt<4> camIntrinsics, Float targetLength, int iters, Float tolerance ) {
		Float left = 0.0, right = 2.0;
		int iter = 0;
		while( iter < iters ) {
			iter++;
			Float length = getAverageProjectedLength(bbox, camIntrinsics, right);
			if( length > targetLength ) right *= 2;
			else break;
		}
		Float maxError = targetLength*tolerance;
		while( iter < iters ) {
			iter++;
			Float middle = (left+right) / 2;
			Float length = getAverageProjectedLength(bbox, camIntrinsics, middle);
			if( fabs(length - targetLength) < maxError ) return middle;
			if( length > targetLength ) left = middle;
			else right = middle;
		}
		return (left+right) / 2;
	}

	// one model's control points (:150-169)
	void addModel( Pt<3> bbox[2], Pt<4> camIntrinsics, int nFeatures, Float MinRatioMin, Float MinRatioMax, Float MaxRatioMin,
	               Float MaxRatioMax, Float DimensionPeak, Float DimensionFade ) {
		Float MinRatioRange = MinRatioMax - MinRatioMin, MaxRatioRange = MaxRatioMax - MaxRatioMin;
		Float depth200Pix = solveProjectionDepth(bbox, camIntrinsics, DimensionPeak, 100, 0.01);
		Float depth100Pix = solveProjectionDepth(bbox, camIntrinsics, DimensionFade, 100, 0.01);
		Float featureCount = (int) nFeatures;
		maxRatioDepths.push_back(depth200Pix);
		minRatioDepths.push_back(depth100Pix);
		Float densityAdjust = canonicalSigmoid( (sigmoidTranslate - featureCount) / sigmoidScale );
		Float MinRatio = MinRatioMin+densityAdjust*MinRatioRange, MaxRatio = MaxRatioMin+densityAdjust*MaxRatioRange;
		ratioLows.push_back(MinRatio);
		ratioHighs.push_back(MaxRatio);
	}

	Float getRatio( Float depth, int modelNum ) const {                                                           // :182-205
		if( depth > MaximumDepth ) return 0.0;
		Float maxRatioDepth = maxRatioDepths[modelNum], minRatioDepth = minRatioDepths[modelNum],
		      ratioLow = ratioLows[modelNum], ratioHigh = ratioHighs[modelNum];
		if( depth < maxRatioDepth ) {
			Float progress = depth / maxRatioDepth;
			return ratioLow + progress*(ratioHigh - ratioLow);
		} else if( depth < minRatioDepth ) {
			return ratioHigh;
		} else if( depth < minRatioDepth*2 ) {
			Float progress = (minRatioDepth*2 - depth) / minRatioDepth;
			return progress*ratioHigh;
		}
		return 0.0;
	}

	// threshold of one feature: depth and fill distance at its pixel, model of its nearest row (:360-372)
	Float getAdjustedRatio( Float depthAtPixel, Float fillDistanceAtPixel, int modelNumber ) const {
		Float weightTerm = fillDistanceAtPixel / CauchyScale;
		Float weight = 1.0 / (1.0 + weightTerm*weightTerm);
		Float putativeRatio = getRatio(depthAtPixel, modelNumber);
		Float defaultRatio = getRatio(DefaultDepth, modelNumber);
		return weight*putativeRatio + (1.0 - weight)*defaultRatio;
	}
};

}
