// FILTER_PROJECTION_CUDA.hpp — drop-in CUDA replacement of the FILTER / FILTER2 steps.
// Same contract as FILTER_PROJECTION_CPU (moped2/libmoped/src/filter/FILTER_PROJECTION_CPU.hpp:50-162):
// constructors (MinPoints, FeatureDistance[, MinScore]), same config keys, scores every object, prunes
// frameData.objects in place (list order kept), rebuilds frameData.clusters[model] for the survivors.
// C++98-compatible.
#pragma once
#include "moped_cuda_ctx.hpp"

namespace MopedNS {

	class FILTER_PROJECTION_CUDA : public MopedAlg {

		int MinPoints;
		Float FeatureDistance;
		Float MinScore;

	public:
		// MinScore is optional
		FILTER_PROJECTION_CUDA( int MinPoints, Float FeatureDistance )
		: MinPoints(MinPoints), FeatureDistance(FeatureDistance), MinScore(0) {
		}

		FILTER_PROJECTION_CUDA( int MinPoints, Float FeatureDistance, Float MinScore )
		: MinPoints(MinPoints), FeatureDistance(FeatureDistance), MinScore(MinScore) {
		}

		void getConfig( map<string,string> &config ) const {
			GET_CONFIG( MinPoints );
			GET_CONFIG( FeatureDistance );
			GET_CONFIG( MinScore );
		}

		void setConfig( map<string,string> &config ) {
			SET_CONFIG( MinPoints );
			SET_CONFIG( FeatureDistance );
			SET_CONFIG( MinScore );
		}

		void process( FrameData &frameData ) {

			vector< vector< FrameData::Match > > &matches = frameData.matches;
			// same sanity check as the reference (:85-87)
			if( matches.size() < models->size() ) return;

			vector<int32_t> off, img; vector<float> xy, xyz;
			MopedCuda::flattenMatches( matches, models->size(), off, img, xy, xyz );
			const int M = off[models->size()];

			// objects in list order; the model index is found by name like the reference does (:96,139,147)
			vector<int32_t> om; vector<float> op;
			for( list<SP_Object>::iterator it = frameData.objects->begin(); it != frameData.objects->end(); ++it ) {
				int mi = -1;
				for( int m=0; m<(int)models->size(); m++ ) if( (*it)->model->name == (*models)[m]->name ) { mi = m; break; }
				om.push_back( mi < 0 ? 0 : mi );
				for( int j=0; j<7; j++ ) op.push_back( (*it)->pose[j] );
			}
			const int nObj = (int)om.size();
			vector<uint8_t> keep( nObj+1 ); vector<float> score( nObj+1 );
			vector<int32_t> co( nObj+2 ), mem( M+2 );
			int32_t nSurv = 0;
			if( om.empty() ) { om.push_back(0); op.resize(7); }
			MopedCuda::setCameras( frameData.images );
			MopedCuda::check( mc_filter_projection( MopedCuda::ctx(), &off[0], &img[0], &xy[0], &xyz[0], (int)models->size(), &om[0], &op[0], nObj,
			                                        MinPoints, FeatureDistance, MinScore, &keep[0], &score[0], &nSurv, &co[0], &mem[0] ), "mc_filter_projection" );

			// prune in place, keep list order (:145-160)
			vector< vector<int> > survivorsOfModel( models->size() );     // object list index per model, list order
			int o = 0;
			for( list<SP_Object>::iterator it = frameData.objects->begin(); it != frameData.objects->end(); o++ ) {
				(*it)->score = score[o];
				if( !keep[o] ) it = frameData.objects->erase( it );
				else { survivorsOfModel[om[o]].push_back( o ); ++it; }
			}
			// clusters of the survivors: the C ABI returns them model-major then list order, like the reference builds them
			frameData.clusters.clear();
			frameData.clusters.resize( models->size() );
			int s = 0;
			for( int m=0; m<(int)models->size(); m++ )
				for( size_t k=0; k<survivorsOfModel[m].size(); k++, s++ ) {
					frameData.clusters[m].resize( frameData.clusters[m].size() + 1 );
					FrameData::Cluster &cl = frameData.clusters[m].back();
					for( int t=co[s]; t<co[s+1]; t++ ) cl.push_back( mem[t] );
				}
		}
	};
};
