// pipeline3d_cuda.hpp — the recognition steps of moped3d's shipped pipeline (moped3d/libmoped/src/config.hpp:41-49) with every
// stage replaced by its CUDA class, parameters unchanged. The steps around them (UNDISTORTED_IMAGE, DEPTHFILL, SIFT, DEPTHFILTER,
// DEPTHFILTER2, DEPTHPROP — config.hpp:37-43) stay the reference's; register these in their places. Include after moped3d's
// moped.hpp / util.hpp. C++98-compatible.
#pragma once
#include "MATCH_ADAPTIVE_CUDA.hpp"
#include "CLUSTER_LINKAGE_CUDA.hpp"
#include "POSE_RANSAC_LM_DIFF_BACKPROJECTION_DEPTH_CUDA.hpp"
#include "FILTER_PROJECTION_CUDA.hpp"

namespace MopedNS {
	static inline void addCudaMatch3d( MopedPipeline &pipeline ) {
		pipeline.addAlg( "MATCH_SIFT", new MATCH_ADAPTIVE_CUDA( 128, "SIFT", 8, 0.600000, 0.750000, 0.650000, 0.800000, 150, 50) );
	}
	static inline void addCudaRecognition3d( MopedPipeline &pipeline ) {       // after DEPTHPROP
		pipeline.addAlg( "CLUSTER", new CLUSTER_LINKAGE_CUDA( 0.100000, 7, 2, 1, 0.0, 1, -1, -1) );
		pipeline.addAlg( "POSE", new POSE_RANSAC_LM_DIFF_BACKPROJECTION_DEPTH_CUDA( 192, 100, 4, 5, 6, 8, 0.5) );
		pipeline.addAlg( "FILTER", new FILTER_PROJECTION_CUDA( 6, 4096., 2) );
		pipeline.addAlg( "POSE2", new POSE_RANSAC_LM_DIFF_BACKPROJECTION_DEPTH_CUDA( 64, 250, 4, 6, 8, 5, 0.5) );
		pipeline.addAlg( "FILTER2", new FILTER_PROJECTION_CUDA( 8, 8192., 1e-4) );
	}
};
