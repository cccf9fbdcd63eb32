// pose_depth_cuda_base.hpp — what the two depth-aware CUDA POSE stage classes share: parameters, getCauchyWeight, process().
// getConfig/setConfig live in the class headers themselves because the reference's GET_CONFIG/SET_CONFIG derive the config key
// from the name of the FILE they are expanded in (moped3d/libmoped/src/util.hpp:62-63). C++98-compatible.
#pragma once
#include "moped_cuda_ctx.hpp"

namespace MopedNS {

	class PoseDepthCudaBase : public MopedAlg {

	protected:

		int MaxRANSACTests;
		int MaxLMTests;
		int MaxObjectsPerCluster;
		int NPtsAlign;
		int MinNPtsObject;
		Float ErrorThreshold;
		Float Alpha;
		int variant;                         // 0 = back-projection, 1 = reprojection + depth
		Float FillInCauchyScale;             // 0.100 / 25.0
		unsigned long long callCounter;      // one RNG stream per process() call

		Float getCauchyWeight( Float distance ) const {
			Float factor = distance / FillInCauchyScale;
			return 1.0 / (1 + factor*factor);
		}

		PoseDepthCudaBase( int variant, Float scale, int MaxRANSACTests, int MaxLMTests, int MaxObjectsPerCluster, int NPtsAlign, int MinNPtsObject,
		                   Float ErrorThreshold, Float Alpha )
		: MaxRANSACTests(MaxRANSACTests), MaxLMTests(MaxLMTests), MaxObjectsPerCluster(MaxObjectsPerCluster), NPtsAlign(NPtsAlign),
		  MinNPtsObject(MinNPtsObject), ErrorThreshold(ErrorThreshold), Alpha(Alpha), variant(variant), FillInCauchyScale(scale), callCounter(0) {
		}

	public:

		void process( FrameData &frameData ) {

			vector< vector< FrameData::Match > > &matches = frameData.matches;
			vector< vector< FrameData::Cluster > > &clusters = frameData.clusters;

			// every (model, cluster) as a contiguous block of correspondences in cluster list order (:404-418)
			vector<int32_t> co( 1, 0 ), img, tie, clModel;
			vector<float> xy, xyz, world, cauchy;
			for( int model=0; model<(int)clusters.size(); model++ )
				for( int cluster=0; cluster<(int)clusters[model].size(); cluster++ ) {
					for( FrameData::Cluster::const_iterator it = clusters[model][cluster].begin(); it != clusters[model][cluster].end(); ++it ) {
						const FrameData::Match &ma = matches[model][*it];
						img.push_back( ma.imageIdx );
						tie.push_back( *it );                    // randSample breaks key ties by LmData address = match index
						xy.push_back( ma.coord2D[0] ); xy.push_back( ma.coord2D[1] );
						for( int c = 0; c < 3; c++ ) { xyz.push_back( ma.coord3D[c] ); world.push_back( ma.depthData.coord3D[c] ); }
						cauchy.push_back( getCauchyWeight( ma.depthData.fillDistance ) );
					}
					co.push_back( (int32_t)img.size() );
					clModel.push_back( model );
				}
			const int nClusters = (int)clModel.size();
			if( nClusters > 0 && !img.empty() ) {
				MopedCuda::setCameras( frameData.images );
				mc_pose_params pp;
				pp.max_ransac_tests = MaxRANSACTests; pp.max_lm_tests = MaxLMTests; pp.max_objects_per_cluster = MaxObjectsPerCluster;
				pp.n_pts_align = NPtsAlign; pp.min_npts_object = MinNPtsObject; pp.error_threshold = ErrorThreshold;
				pp.seed = 0x5DEECE66DULL + (++callCounter) * 0x9E3779B97F4A7C15ULL;
				const int nTasks = nClusters * MaxObjectsPerCluster;
				vector<uint8_t> found( nTasks );
				vector<float> pose( 7*(size_t)nTasks );
				vector<int32_t> nTests( nTasks );
				MopedCuda::check( mc_pose_depth_ransac( MopedCuda::ctx(), variant, &co[0], nClusters, &xy[0], &xyz[0], &world[0], &cauchy[0], &img[0],
				                                        &tie[0], &pp, Alpha, &found[0], &pose[0], &nTests[0] ), "mc_pose_depth_ransac" );
				for( int task=0; task<nTasks; task++ ) {
					if( !found[task] ) continue;
					SP_Object obj(new Object);
					frameData.objects->push_back(obj);
					for( int j=0; j<7; j++ ) obj->pose[j] = pose[7*(size_t)task+j];
					obj->model = (*models)[ clModel[task / MaxObjectsPerCluster] ];
				}
			}
			if( _stepName == "POSE" ) frameData.oldObjects = *frameData.objects;
		}
	};
};
