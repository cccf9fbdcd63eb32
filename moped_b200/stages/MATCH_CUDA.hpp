// MATCH_CUDA.hpp — drop-in CUDA replacement of the MATCH step (B200, libmoped_cuda).
// Same plugin contract as MATCH_ANN_CPU (moped2/libmoped/src/match/MATCH_ANN_CPU.hpp:52-178): constructor
// (DescriptorSize, DescriptorType, Quality, Ratio), the same four config keys, reads
// frameData.detectedFeatures[DescriptorType], writes frameData.matches[model] in query order, normalises
// model and query descriptors in place. The nearest neighbours are EXACT (the reference's Quality=0
// arithmetic, bit for bit) whatever Quality says: a tensor-core coarse pass + exact fp32 re-rank with a
// per-query exactness certificate replaces the kd-tree. Include after moped.hpp/util.hpp (reference tree) or
// after moped_api.hpp (stand-alone). C++98-compatible.
#pragma once
#include "moped_cuda_ctx.hpp"
#include <algorithm>

namespace MopedNS {

	class MATCH_CUDA : public MopedAlg {

		// L2 normalisation in place, fp32: the host does it (model descriptors once per database build, query
		// descriptors every frame) exactly where the reference does (MATCH_ANN_CPU.hpp:54-57,94,157)
		static inline void normalise( vector<float> &v ) {
			float ss = 0;
			for( size_t k = 0; k < v.size(); k++ ) ss += v[k]*v[k];
			const float inv = 1./sqrtf(ss);
			for( size_t k = 0; k < v.size(); k++ ) v[k] *= inv;
		}

		int DescriptorSize;
		string DescriptorType;
		Float Quality;
		Float Ratio;

		bool databaseReady;
		vector<int32_t> rowModel;           // global row -> model index   (the reference's correspModel)
		vector< const Pt<3> * > rowPoint;    // global row -> its coord3D   (the reference's correspFeat)

		// Rebuild the device database after a model or config change: rows = all descriptors of
		// DescriptorType, models in order, features in order (row ids must equal the reference's, :85-100).
		void uploadDatabase() {

			databaseReady = false;
			configUpdated = false;
			size_t nRows = 0;
			for( size_t m = 0; m < models->size(); m++ ) nRows += (*models)[m]->IPs[DescriptorType].size();
			rowModel.assign( nRows, 0 );
			rowPoint.assign( nRows, NULL );
			if( nRows < 2 ) return;                       // the reference skips matching below two rows (:102)

			vector<float> desc( nRows * DescriptorSize ), xyz( nRows * 3 );
			size_t row = 0;
			for( size_t m = 0; m < models->size(); m++ ) {
				vector<Model::IP> &ips = (*models)[m]->IPs[DescriptorType];
				for( size_t f = 0; f < ips.size(); f++, row++ ) {
					normalise( ips[f].descriptor );
					std::copy( ips[f].descriptor.begin(), ips[f].descriptor.begin() + DescriptorSize, desc.begin() + row * DescriptorSize );
					for( int c = 0; c < 3; c++ ) xyz[row*3+c] = ips[f].coord3D[c];
					rowModel[row] = (int32_t)m;
					rowPoint[row] = &ips[f].coord3D;
				}
			}
			MopedCuda::check( mc_db_upload( MopedCuda::ctx(), &desc[0], &xyz[0], &rowModel[0], (int64_t)nRows, DescriptorSize,
			                                (int)models->size(), 0 ), "mc_db_upload" );
			databaseReady = true;
		}

	public:

		MATCH_CUDA( int DescriptorSize, string DescriptorType, Float Quality, Float Ratio )
		: DescriptorSize(DescriptorSize), DescriptorType(DescriptorType), Quality(Quality), Ratio(Ratio), databaseReady(false) {
		}

		void getConfig( map<string,string> &config ) const {
			GET_CONFIG(DescriptorType);
			GET_CONFIG(DescriptorSize);
			GET_CONFIG(Quality);
			GET_CONFIG(Ratio);
		};

		void setConfig( map<string,string> &config ) {
			SET_CONFIG(DescriptorType);
			SET_CONFIG(DescriptorSize);
			SET_CONFIG(Quality);
			SET_CONFIG(Ratio);
		};

		void process( FrameData &frameData ) {

			if( configUpdated ) uploadDatabase();
			if( !databaseReady ) return;

			vector< FrameData::DetectedFeature > &feats = frameData.detectedFeatures[DescriptorType];
			if( feats.empty() ) return;
			frameData.matches.resize( models->size() );

			const int nQueries = (int)feats.size();
			vector<float> queries( (size_t)nQueries * DescriptorSize );
			for( int i = 0; i < nQueries; i++ ) {
				normalise( feats[i].descriptor );
				std::copy( feats[i].descriptor.begin(), feats[i].descriptor.begin() + DescriptorSize, queries.begin() + (size_t)i * DescriptorSize );
			}
			vector<int32_t> nnRow( 2*(size_t)nQueries );
			vector<float> nnDist( 2*(size_t)nQueries );
			vector<uint8_t> accepted( nQueries );
			MopedCuda::check( mc_match( MopedCuda::ctx(), &queries[0], nQueries, Ratio, MC_MATCH_TENSOR, &nnRow[0], &nnDist[0], &accepted[0], NULL ), "mc_match" );

			// accepted queries (ratio test done on the device) become matches of the nearest row's model, in query order
			for( int i = 0; i < nQueries; i++ ) {
				if( !accepted[i] ) continue;
				const int32_t row = nnRow[2*i];
				FrameData::Match hit;
				hit.imageIdx = feats[i].imageIdx;
				hit.coord2D = feats[i].coord2D;
				hit.coord3D = *rowPoint[row];
				frameData.matches[ rowModel[row] ].push_back( hit );
			}
		}
	};
};
