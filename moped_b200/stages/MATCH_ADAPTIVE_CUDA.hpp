// MATCH_ADAPTIVE_CUDA.hpp — drop-in CUDA replacement of moped3d's MATCH step, MATCH_ADAPTIVE_FLANN_CPU
// (moped3d/libmoped/src/match/MATCH_ADAPTIVE_FLANN_CPU.hpp:50-520; moped3d/libmoped/src/config.hpp:41).
// Plugin contract kept: constructor (DescriptorSize, DescriptorType, NumTrees, MinRatioMin, MinRatioMax, MaxRatioMin, MaxRatioMax,
// DimensionPeak, DimensionFade), the same nine config keys, reads detectedFeatures[DescriptorType], the IMAGE_TYPE_DEPTH_MAP image
// and its "<name>.distance" probability map, writes matches[model] (COPY_FEATURE_TO_MATCH, util.hpp:69) in feature order,
// normalises model and query descriptors in place.
// How it works here: the class keeps no search structure and evaluates no threshold. A model change uploads the descriptor rows
// (mc_db_upload) and derives one ratio curve per model (mc_adaptive_model_init); a frame hands its descriptors, pixel coordinates,
// the depth plane and the fill-distance plane to mc_match_adaptive, which returns the exact nearest rows (NumTrees is kept as a
// parameter and ignored) and one accept flag per feature, decided on the device. C++98-compatible; include after moped3d's
// moped.hpp / util.hpp.
#pragma once
#include "moped_cuda_ctx.hpp"
#include <algorithm>

namespace MopedNS {

	class MATCH_ADAPTIVE_CUDA : public MopedAlg {

		int DescriptorSize;
		string DescriptorType;
		int NumTrees;
		Float MinRatioMin, MinRatioMax, MaxRatioMin, MaxRatioMax;
		Float DimensionPeak, DimensionFade;

		// database rows in the reference's numbering (models in order, features in order): row -> (model, coord3D)
		struct Row { int model; const Pt<3> *point; };
		vector<Row> rows;
		bool searchable;
		// One curve per model and Update(). The reference appends to its control-point vectors on every Update() and never clears
		// them, while its lookups index by model number (:93-96,158-169): after a second Update() the FIRST generation keeps
		// being used. Same here — the device is given the leading models->size() entries.
		vector<mc_adaptive_model> curves;

		static void unitLength( vector<float> &d ) {             // in-place L2 normalisation, fp32, sequential (:52-55)
			float ss = 0;
			for( size_t k = 0; k < d.size(); k++ ) ss += d[k]*d[k];
			ss = 1/sqrtf(ss);
			for( size_t k = 0; k < d.size(); k++ ) d[k] *= ss;
		}

		static SP_Image firstOfType( FrameData &frameData, int type, bool last ) {
			SP_Image hit;
			for( size_t i = 0; i < frameData.images.size(); i++ )
				if( frameData.images[i]->imageType == type ) { hit = frameData.images[i]; if( !last ) break; }
			return hit;
		}

	public:

		// Rebuild rows and curves after a model or config change. `upload` = false leaves the device alone (host-only checks).
		void Update( FrameData &frameData, bool upload ) {

			searchable = false;
			configUpdated = false;
			if( models == NULL ) return;

			rows.clear();
			vector<float> desc, xyz;
			vector<int32_t> rowModel;
			for( size_t m = 0; m < models->size(); m++ ) {
				vector<Model::IP> &ips = (*models)[m]->IPs[DescriptorType];
				for( size_t f = 0; f < ips.size(); f++ ) {
					unitLength( ips[f].descriptor );
					desc.insert( desc.end(), ips[f].descriptor.begin(), ips[f].descriptor.end() );
					for( int c = 0; c < 3; c++ ) xyz.push_back( ips[f].coord3D[c] );
					rowModel.push_back( (int32_t)m );
					Row r; r.model = (int)m; r.point = &ips[f].coord3D;
					rows.push_back( r );
				}
			}
			if( rows.size() > 1 ) {                       // the reference builds its index only from two rows up (:134)
				searchable = true;
				if( upload ) MopedCuda::check( mc_db_upload( MopedCuda::ctx(), &desc[0], &xyz[0], &rowModel[0], (int64_t)rows.size(), DescriptorSize,
				                                             (int)models->size(), 0 ), "mc_db_upload" );
			}

			// intrinsics of the (last) grey image of the frame that triggers the update (:147-153)
			SP_Image gray = firstOfType( frameData, IMAGE_TYPE_GRAY_IMAGE, true );
			float K[4];
			for( int c = 0; c < 4; c++ ) K[c] = gray->intrinsicLinearCalibration[c];
			for( size_t m = 0; m < models->size(); m++ ) {
				Model &model = *(*models)[m];
				float lo[3], hi[3];
				for( int c = 0; c < 3; c++ ) { lo[c] = model.boundingBox[0][c]; hi[c] = model.boundingBox[1][c]; }
				mc_adaptive_model curve;
				mc_adaptive_model_init( &curve, lo, hi, K, (int)model.IPs[DescriptorType].size(), MinRatioMin, MinRatioMax, MaxRatioMin, MaxRatioMax,
				                        DimensionPeak, DimensionFade );
				curves.push_back( curve );
			}
		}

		MATCH_ADAPTIVE_CUDA( int DescriptorSize, string DescriptorType, int NumTrees, Float MinRatioMin, Float MinRatioMax,
		                     Float MaxRatioMin, Float MaxRatioMax, Float DimensionPeak, Float DimensionFade )
		: DescriptorSize(DescriptorSize), DescriptorType(DescriptorType), NumTrees(NumTrees), MinRatioMin(MinRatioMin), MinRatioMax(MinRatioMax),
		  MaxRatioMin(MaxRatioMin), MaxRatioMax(MaxRatioMax), DimensionPeak(DimensionPeak), DimensionFade(DimensionFade), searchable(false) {
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

		// ---- the pieces of process(), public so that a host-only check can drive them without a device ----

		// true when there is something to match; runs the pending Update()
		bool prepare( FrameData &frameData, bool upload = true ) {
			if( configUpdated ) Update( frameData, upload );
			return searchable && !frameData.detectedFeatures[DescriptorType].empty();
		}

		// what mc_match_adaptive needs from the frame: normalised descriptors (in place, like the reference), pixel coordinates,
		// and the depth / fill-distance planes of the frame's depth map as dense height x width arrays
		struct FrameInputs {
			vector<float> desc, xy, depth, fill;
			int width, height;
		};
		void gather( FrameData &frameData, FrameInputs &in ) {
			vector< FrameData::DetectedFeature > &feats = frameData.detectedFeatures[DescriptorType];
			in.desc.resize( feats.size() * (size_t)DescriptorSize );
			in.xy.resize( feats.size() * 2 );
			for( size_t i = 0; i < feats.size(); i++ ) {
				unitLength( feats[i].descriptor );
				std::copy( feats[i].descriptor.begin(), feats[i].descriptor.begin() + DescriptorSize, in.desc.begin() + i * DescriptorSize );
				in.xy[2*i] = feats[i].coord2D[0]; in.xy[2*i+1] = feats[i].coord2D[1];
			}
			SP_Image depthmap = firstOfType( frameData, IMAGE_TYPE_DEPTH_MAP, true );
			SP_Image fillmap;
			for( size_t i = 0; i < frameData.images.size(); i++ )
				if( frameData.images[i]->imageType == IMAGE_TYPE_PROB_MAP && frameData.images[i]->name == depthmap->name + ".distance" ) {
					fillmap = frameData.images[i]; break;
				}
			in.width = depthmap->width; in.height = depthmap->height;
			in.depth.resize( (size_t)in.width * in.height );
			in.fill.resize( (size_t)in.width * in.height );
			for( int y = 0; y < in.height; y++ )
				for( int x = 0; x < in.width; x++ ) {
					in.depth[(size_t)y * in.width + x] = depthmap->getDepth( x, y );
					in.fill[(size_t)y * in.width + x] = fillmap->getProb( x, y );
				}
		}

		const mc_adaptive_model *modelCurves() const { return &curves[0]; }
		int modelOfRow( int row ) const { return rows[row].model; }

		// accepted features become matches of their nearest row's model, in feature order
		void emit( FrameData &frameData, const vector<int32_t> &nnRow, const vector<uint8_t> &accepted ) {
			vector< FrameData::DetectedFeature > &feats = frameData.detectedFeatures[DescriptorType];
			frameData.matches.clear();                    // the reference reassembles every model's vector from scratch (:474-488)
			frameData.matches.resize( models->size() );
			for( size_t i = 0; i < feats.size(); i++ ) {
				if( !accepted[i] ) continue;
				const Row &r = rows[ nnRow[2*i] ];
				FrameData::Match hit;
				COPY_FEATURE_TO_MATCH( feats[i], hit );
				hit.coord3D = *r.point;
				frameData.matches[ r.model ].push_back( hit );
			}
		}

		void process( FrameData &frameData ) {

			if( !prepare( frameData ) ) return;
			FrameInputs in;
			gather( frameData, in );
			const int nQueries = (int)frameData.detectedFeatures[DescriptorType].size();
			vector<int32_t> nnRow( 2*(size_t)nQueries );
			vector<float> nnDist( 2*(size_t)nQueries );
			vector<uint8_t> accepted( nQueries );
			// 4.0 m / 1.0 m / 0.1: MaximumDepth, DefaultDepth, CauchyScale as Update() sets them in the reference (:106-108)
			MopedCuda::check( mc_match_adaptive( MopedCuda::ctx(), &in.desc[0], &in.xy[0], nQueries, &in.depth[0], &in.fill[0], in.width, in.height,
			                                     modelCurves(), (int)models->size(), 4.0f, 1.0f, 0.1f, &nnRow[0], &nnDist[0], &accepted[0] ),
			                  "mc_match_adaptive" );
			emit( frameData, nnRow, accepted );
		}
	};
};
