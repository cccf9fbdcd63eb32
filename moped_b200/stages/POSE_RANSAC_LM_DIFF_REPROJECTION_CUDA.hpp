// POSE_RANSAC_LM_DIFF_REPROJECTION_CUDA.hpp — drop-in CUDA replacement of the POSE / POSE2 steps.
// Same contract as POSE_RANSAC_LM_DIFF_REPROJECTION_CPU
// (moped2/libmoped/src/pose/POSE_RANSAC_LM_DIFF_REPROJECTION_CPU.hpp:57-307): constructor (MaxRANSACTests,
// MaxLMTests, MaxObjectsPerCluster, NPtsAlign, MinNPtsObject, ErrorThreshold), same config keys, reads
// frameData.matches / clusters / images, appends one Object per successful (cluster, try) task, saves
// oldObjects when the step is called "POSE". Tasks run as one CTA each on the device; objects are appended
// in task order (the reference appends in OpenMP completion order). C++98-compatible.
#pragma once
#include "moped_cuda_ctx.hpp"

namespace MopedNS {

	class POSE_RANSAC_LM_DIFF_REPROJECTION_CUDA : public MopedAlg {

		int MaxRANSACTests;
		int MaxLMTests;
		int MaxObjectsPerCluster;
		int NPtsAlign;
		int MinNPtsObject;
		Float ErrorThreshold;
		unsigned long long callCounter;      // one RNG stream per process() call

	public:

		POSE_RANSAC_LM_DIFF_REPROJECTION_CUDA( int MaxRANSACTests, int MaxLMTests, int MaxObjectsPerCluster, int NPtsAlign, int MinNPtsObject, Float ErrorThreshold )
		: MaxRANSACTests(MaxRANSACTests), MaxLMTests(MaxLMTests), MaxObjectsPerCluster(MaxObjectsPerCluster), NPtsAlign(NPtsAlign),
		  MinNPtsObject(MinNPtsObject), ErrorThreshold(ErrorThreshold), callCounter(0) {
		}

		void getConfig( map<string,string> &config ) const {
			GET_CONFIG( MaxRANSACTests );
			GET_CONFIG( MaxLMTests );
			GET_CONFIG( NPtsAlign );
			GET_CONFIG( MinNPtsObject );
			GET_CONFIG( ErrorThreshold );
		}

		void setConfig( map<string,string> &config ) {
			SET_CONFIG( MaxRANSACTests );
			SET_CONFIG( MaxLMTests );
			SET_CONFIG( NPtsAlign );
			SET_CONFIG( MinNPtsObject );
			SET_CONFIG( ErrorThreshold );
		}

		void process( FrameData &frameData ) {

			vector< vector< FrameData::Match > > &matches = frameData.matches;
			vector< vector< FrameData::Cluster > > &clusters = frameData.clusters;

			// every (model, cluster) as a contiguous block of correspondences, in cluster list order (:275-290)
			vector<int32_t> co( 1, 0 ), img, clModel;
			vector<float> xy, xyz;
			for( int model=0; model<(int)clusters.size(); model++ )
				for( int cluster=0; cluster<(int)clusters[model].size(); cluster++ ) {
					for( FrameData::Cluster::const_iterator it = clusters[model][cluster].begin(); it != clusters[model][cluster].end(); ++it ) {
						const FrameData::Match &ma = matches[model][*it];
						img.push_back( ma.imageIdx );
						xy.push_back( ma.coord2D[0] ); xy.push_back( ma.coord2D[1] );
						xyz.push_back( ma.coord3D[0] ); xyz.push_back( ma.coord3D[1] ); xyz.push_back( ma.coord3D[2] );
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
				MopedCuda::check( mc_pose_ransac( MopedCuda::ctx(), &co[0], nClusters, &xy[0], &xyz[0], &img[0], &pp, &found[0], &pose[0], &nTests[0] ), "mc_pose_ransac" );
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
