// CLUSTER_MEAN_SHIFT_CUDA.hpp — drop-in CUDA replacement of the CLUSTER step.
// Same contract as CLUSTER_MEAN_SHIFT_CPU (moped2/libmoped/src/cluster/CLUSTER_MEAN_SHIFT_CPU.hpp:50-199):
// constructor (Radius, Merge, MinPts, MaxIterations), same config keys, reads frameData.matches and
// frameData.images.size(), writes frameData.clusters[model] (lists of indices into matches[model]) in the
// same order, saves oldClusters when the step is called "CLUSTER". C++98-compatible.
#pragma once
#include "moped_cuda_ctx.hpp"

namespace MopedNS {

	class CLUSTER_MEAN_SHIFT_CUDA : public MopedAlg {

		float Radius;
		float Merge;
		int MinPts;
		int MaxIterations;

	public:

		CLUSTER_MEAN_SHIFT_CUDA( float Radius, float Merge, unsigned int MinPts, unsigned int MaxIterations )
		: Radius(Radius), Merge(Merge), MinPts(MinPts), MaxIterations(MaxIterations) {
		}

		void getConfig( map<string,string> &config ) const {
			GET_CONFIG( Radius );
			GET_CONFIG( Merge );
			GET_CONFIG( MinPts );
			GET_CONFIG( MaxIterations );
		}

		void setConfig( map<string,string> &config ) {
			SET_CONFIG( Radius );
			SET_CONFIG( Merge );
			SET_CONFIG( MinPts );
			SET_CONFIG( MaxIterations );
		}

		void process( FrameData &frameData ) {

			frameData.clusters.resize( models->size() );
			if( !frameData.matches.empty() && !frameData.images.empty() ) {
				vector<int32_t> off, img; vector<float> xy, xyz;
				MopedCuda::flattenMatches( frameData.matches, models->size(), off, img, xy, xyz );
				const int M = off[models->size()];
				if( M > 0 ) {
					vector<int32_t> cm( M+2 ), co( M+2 ), mem( M+2 );
					int32_t nc = 0;
					MopedCuda::check( mc_cluster_meanshift( MopedCuda::ctx(), &off[0], &img[0], &xy[0], (int)models->size(), (int)frameData.images.size(),
					                                        Radius, Merge, MinPts, MaxIterations, &nc, &cm[0], &co[0], &mem[0] ), "mc_cluster_meanshift" );
					for( int c=0; c<nc; c++ ) {
						frameData.clusters[cm[c]].resize( frameData.clusters[cm[c]].size() + 1 );
						FrameData::Cluster &cl = frameData.clusters[cm[c]].back();
						for( int t=co[c]; t<co[c+1]; t++ ) cl.push_back( mem[t] );
					}
				}
			}
			if( _stepName == "CLUSTER" ) frameData.oldClusters = frameData.clusters;
		}
	};
};
