// MATCH_ADAPTIVE_CUDA.hpp — drop-in CUDA replacement of moped3d's MATCH step, MATCH_ADAPTIVE_FLANN_CPU
// (moped3d/libmoped/src/match/MATCH_ADAPTIVE_FLANN_CPU.hpp:50-520; moped3d/libmoped/src/config.hpp:41).
// Same constructor (DescriptorSize, DescriptorType, NumTrees, MinRatioMin, MinRatioMax, MaxRatioMin, MaxRatioMax, DimensionPeak,
// DimensionFade), the same nine config keys, reads detectedFeatures[DescriptorType] and the IMAGE_TYPE_DEPTH_MAP image with its
// "<name>.distance" probability map, writes matches[model] (imageIdx, coord2D, coord3D — COPY_FEATURE_TO_MATCH, util.hpp:69) in
// feature order, normalises model and query descriptors in place.
// What changes: the two nearest rows come from libmoped_cuda's matcher (mc_match: tensor-core coarse pass + exact fp32 re-rank,
// the bits of an exhaustive search) instead of OpenCV's randomised kd-trees — NumTrees is kept as a parameter and ignored, and
// features beyond MaximumDepth are searched with the rest and dropped afterwards. The depth-dependent ratio threshold is the
// reference's own host arithmetic (adaptive_ratio.hpp). C++98-compatible; include after moped3d's moped.hpp / util.hpp.
#pragma once
#include "moped_cuda_ctx.hpp"
#include "adaptive_ratio.hpp"
#include <algorithm>

namespace MopedNS {

	class MATCH_ADAPTIVE_CUDA : public MopedAlg {

		static inline void norm( vector<float> &d ) {             // :52-55
			float norm=0; for (int x=0; x<(int)d.size(); x++) norm += d[x]*d[x]; norm = 1/sqrtf(norm);
			for (int x=0; x<(int)d.size(); x++) d[x] *=norm;
		}

		int DescriptorSize;
		string DescriptorType;
		int NumTrees;
		Float MinRatioMin, MinRatioMax, MaxRatioMin, MaxRatioMax;
		Float DimensionPeak, DimensionFade;

		bool skipCalculation;
		vector< pair<int, Pt<3> *> > modelPointData;
		AdaptiveRatio adaptive;

		void Update( FrameData &frameData, bool upload ) {        // :105-174

			skipCalculation = true;
			if( models==NULL ) return;

			modelPointData.clear();
			vector<float> dataset, xyz;
			vector<int32_t> rowModel;
			for( int nModel = 0; nModel < (int)models->size(); nModel++ ) {
				vector<Model::IP> &IPs = (*models)[nModel]->IPs[DescriptorType];
				for( int nFeat = 0; nFeat < (int)IPs.size(); nFeat++ ) {
					norm( IPs[nFeat].descriptor );
					for( int i = 0; i < (int)IPs[nFeat].descriptor.size(); i++ ) dataset.push_back( IPs[nFeat].descriptor[i] );
					for( int c = 0; c < 3; c++ ) xyz.push_back( IPs[nFeat].coord3D[c] );
					rowModel.push_back( nModel );
					modelPointData.push_back( make_pair( nModel, &IPs[nFeat].coord3D ) );
				}
			}
			if( modelPointData.size() > 1 ) {
				skipCalculation = false;
				if( upload ) MopedCuda::check( mc_db_upload( MopedCuda::ctx(), &dataset[0], &xyz[0], &rowModel[0], (int64_t)modelPointData.size(), DescriptorSize,
				                                (int)models->size(), 0 ), "mc_db_upload" );
			}

			SP_Image grayImage;
			for( int i = 0; i < (int)frameData.images.size(); i++ )
				if( frameData.images[i]->imageType == IMAGE_TYPE_GRAY_IMAGE ) grayImage = frameData.images[i];

			for( int modelNum = 0; modelNum < (int)models->size(); modelNum++ ) {
				SP_Model model = (*models)[modelNum];
				adaptive.addModel( model->boundingBox, grayImage->intrinsicLinearCalibration, (int)model->IPs[DescriptorType].size(),
				                   MinRatioMin, MinRatioMax, MaxRatioMin, MaxRatioMax, DimensionPeak, DimensionFade );
			}
			configUpdated = false;
		}

	public:

		MATCH_ADAPTIVE_CUDA( int DescriptorSize, string DescriptorType, int NumTrees, Float MinRatioMin, Float MinRatioMax,
		                     Float MaxRatioMin, Float MaxRatioMax, Float DimensionPeak, Float DimensionFade )
		: DescriptorSize(DescriptorSize), DescriptorType(DescriptorType), NumTrees(NumTrees), MinRatioMin(MinRatioMin), MinRatioMax(MinRatioMax),
		  MaxRatioMin(MaxRatioMin), MaxRatioMax(MaxRatioMax), DimensionPeak(DimensionPeak), DimensionFade(DimensionFade), skipCalculation(false) {
		}

		void getConfig( map<string,string> &config ) const {
			GET_CONFIG(MinRatioMin);
			GET_CONFIG(MinRatioMax);
			GET_CONFIG(MaxRatioMin);
			GET_CONFIG(MaxRatioMax);
			GET_CONFIG(DimensionPeak);
			GET_CONFIG(DimensionFade);
			GET_CONFIG(NumTrees);
			GET_CONFIG(DescriptorType);
			GET_CONFIG(DescriptorSize);
		};

		void setConfig( map<string,string> &config ) {
			SET_CONFIG(MinRatioMin);
			SET_CONFIG(MinRatioMax);
			SET_CONFIG(MaxRatioMin);
			SET_CONFIG(MaxRatioMax);
			SET_CONFIG(DimensionPeak);
			SET_CONFIG(DimensionFade);
			SET_CONFIG(NumTrees);
			SET_CONFIG(DescriptorType);
			SET_CONFIG(DescriptorSize);
		};

		// The part of process() after the nearest-neighbour search (:436-470), on the search result of every feature:
		// nnRow / nnDist hold the two nearest rows and their squared distances per feature. Public so that the CPU check
		// (oracle/ref3d_match_dropin.cpp) can drive it with an exhaustive host search where no GPU is present.
		void acceptMatches( FrameData &frameData, const vector<int32_t> &nnRow, const vector<float> &nnDist ) {

			vector< FrameData::DetectedFeature > &corresp = frameData.detectedFeatures[DescriptorType];
			SP_Image depthmap, distanceMap;
			for( int i = 0; i < (int)frameData.images.size(); i++ )
				if( frameData.images[i]->imageType == IMAGE_TYPE_DEPTH_MAP ) depthmap = frameData.images[i];
			for( int i = 0; i < (int)frameData.images.size(); i++ )
				if( frameData.images[i]->imageType == IMAGE_TYPE_PROB_MAP && frameData.images[i]->name == depthmap->name+".distance" ) {
					distanceMap = frameData.images[i]; break;
				}

			vector< vector< FrameData::Match > > &matches = frameData.matches;
			matches.clear();                                  // the reference rebuilds every model's vector from its thread buffers (:474-488)
			matches.resize( models->size() );
			for( int i = 0; i < (int)corresp.size(); i++ ) {
				Pt<2> loc2D = corresp[i].coord2D;
				int x = (int) loc2D[0], y = (int) loc2D[1];
				x = min( max(x,0), depthmap->width );         // sic: the reference clamps to width / height, not width-1 / height-1 (:423-424)
				y = min( max(y,0), depthmap->height );
				Float depth = depthmap->getDepth(x, y);
				if( depth > adaptive.MaximumDepth ) continue;
				int nModel = modelPointData[ nnRow[2*i] ].first;
				Float Ratio = adaptive.getAdjustedRatio( depth, distanceMap->getProb(x, y), nModel );
				if( nnDist[2*i]/nnDist[2*i+1] < Ratio ) {
					matches[nModel].resize( matches[nModel].size() +1 );
					FrameData::Match &match = matches[nModel].back();
					COPY_FEATURE_TO_MATCH(corresp[i], match);
					match.coord3D = *modelPointData[ nnRow[2*i] ].second;
				}
			}
		}

		// query descriptors normalised in place (:427) and packed row-major
		void packQueries( FrameData &frameData, vector<float> &queries ) {
			vector< FrameData::DetectedFeature > &corresp = frameData.detectedFeatures[DescriptorType];
			queries.resize( corresp.size() * (size_t)DescriptorSize );
			for( int i = 0; i < (int)corresp.size(); i++ ) {
				norm( corresp[i].descriptor );
				for( int x = 0; x < DescriptorSize; x++ ) queries[(size_t)i*DescriptorSize + x] = corresp[i].descriptor[x];
			}
		}

		// upload = false: everything except the device upload (the CPU check of the host logic)
		bool prepare( FrameData &frameData, bool upload = true ) {      // :391-397
			if( configUpdated ) Update(frameData, upload);
			if( skipCalculation ) return false;
			return !frameData.detectedFeatures[DescriptorType].empty();
		}

		void process( FrameData &frameData ) {

			if( !prepare(frameData) ) return;
			vector<float> queries;
			packQueries( frameData, queries );
			const int nQueries = (int)frameData.detectedFeatures[DescriptorType].size();
			vector<int32_t> nnRow( 2*(size_t)nQueries );
			vector<float> nnDist( 2*(size_t)nQueries );
			vector<uint8_t> accepted( nQueries );
			// ratio 1: the fixed-ratio test of the device is not used, the adaptive one runs below on the returned distances
			MopedCuda::check( mc_match( MopedCuda::ctx(), &queries[0], nQueries, 1.0f, MC_MATCH_TENSOR, &nnRow[0], &nnDist[0], &accepted[0], NULL ), "mc_match" );
			acceptMatches( frameData, nnRow, nnDist );
		}
	};
};
