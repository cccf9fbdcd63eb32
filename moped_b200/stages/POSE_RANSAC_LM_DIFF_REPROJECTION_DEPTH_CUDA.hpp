// POSE_RANSAC_LM_DIFF_REPROJECTION_DEPTH_CUDA.hpp — drop-in CUDA replacement of moped3d's POSE_RANSAC_LM_DIFF_REPROJECTION_DEPTH_CPU
// (moped3d/libmoped/src/pose/POSE_RANSAC_LM_DIFF_REPROJECTION_DEPTH_CPU.hpp:57-435): same constructor (MaxRANSACTests, MaxLMTests,
// MaxObjectsPerCluster, NPtsAlign, MinNPtsObject, ErrorThreshold, Alpha), same five config keys, same FrameData reads (matches incl.
// depthData.coord3D / depthData.fillDistance, clusters, images) and writes (objects appended per successful (cluster, try) task — in
// task order; the reference appends in OpenMP completion order —, oldObjects when the step is called "POSE"). The device side is
// mc_pose_depth_ransac(variant 1): an LM that takes every sum in levmar's order, i.e. the poses a strict-IEEE build of the
// reference computes on the same samples. C++98-compatible: compiles inside moped3d's tree with -std=gnu++98.
#pragma once
#include "pose_depth_cuda_base.hpp"

namespace MopedNS {

	class POSE_RANSAC_LM_DIFF_REPROJECTION_DEPTH_CUDA : public PoseDepthCudaBase {

	public:

		POSE_RANSAC_LM_DIFF_REPROJECTION_DEPTH_CUDA( int MaxRANSACTests, int MaxLMTests, int MaxObjectsPerCluster, int NPtsAlign, int MinNPtsObject, Float ErrorThreshold, Float Alpha )
		: PoseDepthCudaBase( 1, 25.0, MaxRANSACTests, MaxLMTests, MaxObjectsPerCluster, NPtsAlign, MinNPtsObject, ErrorThreshold, Alpha ) {
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
	};
};
