// CLUSTER_LINKAGE_CUDA.hpp — drop-in CUDA replacement of moped3d's CLUSTER step (B200, libmoped_cuda).
// Same plugin contract as CLUSTER_LINKAGE_CPU (moped3d/libmoped/src/cluster/CLUSTER_LINKAGE_CPU.hpp:49-706): the eight
// constructor parameters, the same six config keys, reads frameData.matches[model] (coord2D, coord3D, depthData.coord3D), the
// depth map among frameData.images (IMAGE_TYPE_DEPTH_MAP) and its "<name>.distance" probability map, writes
// frameData.clusters[model] (+ oldClusters when the step is called "CLUSTER"). Include after moped3d's moped.hpp/util.hpp.
// C++98-compatible. Like the reference it expects a depth map to be present.
#pragma once
#include "moped_cuda_ctx.hpp"

namespace MopedNS {

	class CLUSTER_LINKAGE_CUDA : public MopedAlg {

		Float Cutoff;
		int MinPts;
		int Use3DFilter;
		Float WeightGamma;
		Float Alpha;
		int LinkageType;
		Float Sigma2D;
		Float Sigma3D;

	public:

		CLUSTER_LINKAGE_CUDA( Float Cutoff, int MinPts, int Use3DFilter, Float WeightGamma, Float Alpha, int LinkageType, Float Sigma2D, Float Sigma3D )
		: Cutoff(Cutoff), MinPts(MinPts), Use3DFilter(Use3DFilter), WeightGamma(WeightGamma), Alpha(Alpha), LinkageType(LinkageType),
		  Sigma2D(Sigma2D), Sigma3D(Sigma3D) {
		}

		void getConfig( map<string,string> &config ) const {

			GET_CONFIG( Cutoff );
			GET_CONFIG( MinPts );
			GET_CONFIG( Use3DFilter );
			GET_CONFIG( WeightGamma );
			GET_CONFIG( Alpha );
			GET_CONFIG( LinkageType );
		}

		void setConfig( map<string,string> &config ) {

			SET_CONFIG( Cutoff );
			SET_CONFIG( MinPts );
			SET_CONFIG( Use3DFilter );
			SET_CONFIG( WeightGamma );
			SET_CONFIG( Alpha );
			SET_CONFIG( LinkageType );
		}

		void process( FrameData &frameData ) {

			frameData.clusters.resize( models->size() );

			// the depth map and its fill-distance map, found like the reference finds them (:582-597)
			SP_Image depthmap, distanceMap;
			for( int i = 0; i < (int)frameData.images.size(); i++ )
				if( frameData.images[i]->imageType == IMAGE_TYPE_DEPTH_MAP ) { depthmap = frameData.images[i]; break; }
			for( int i = 0; i < (int)frameData.images.size(); i++ )
				if( frameData.images[i]->imageType == IMAGE_TYPE_PROB_MAP && frameData.images[i]->name == depthmap->name + ".distance" ) {
					distanceMap = frameData.images[i]; break;
				}

			const int W = depthmap->width, H = depthmap->height;
			vector<float> depth( (size_t)W*H ), distance( (size_t)W*H );
			for( int y = 0; y < H; y++ )
				for( int x = 0; x < W; x++ ) {
					depth[(size_t)y*W+x] = depthmap->getDepth( x, y );
					distance[(size_t)y*W+x] = distanceMap->getProb( x, y );
				}

			for( int model = 0; model < (int)frameData.matches.size(); model++ ) {

				vector<FrameData::Match> &matches = frameData.matches[model];
				const int n = (int)matches.size();
				vector<FrameData::Cluster> clusters;
				if( n > 0 ) {
					vector<float> xy( 2*(size_t)n ), xyz( 3*(size_t)n ), world( 3*(size_t)n );
					for( int i = 0; i < n; i++ ) {
						xy[2*i] = matches[i].coord2D[0]; xy[2*i+1] = matches[i].coord2D[1];
						for( int c = 0; c < 3; c++ ) { xyz[3*i+c] = matches[i].coord3D[c]; world[3*i+c] = matches[i].depthData.coord3D[c]; }
					}
					int32_t nClusters = 0;
					vector<int32_t> offsets( n + 2 ), members( n + 1 );
					MopedCuda::check( mc_cluster_linkage( MopedCuda::ctx(), &xy[0], &xyz[0], &world[0], n, &depth[0], &distance[0], W, H,
					                                      Cutoff, MinPts, Use3DFilter, LinkageType, Sigma2D, Sigma3D,
					                                      &nClusters, &offsets[0], &members[0], NULL ), "mc_cluster_linkage" );
					clusters.resize( nClusters );
					for( int c = 0; c < nClusters; c++ )
						for( int k = offsets[c]; k < offsets[c+1]; k++ ) clusters[c].push_back( members[k] );
				}
				frameData.clusters[model] = clusters;
			}

			if( _stepName == "CLUSTER" ) frameData.oldClusters = frameData.clusters;
		}
	};
};
