// pipeline_cuda.hpp — the default pipeline of moped2/libmoped/src/config.hpp:83-120 with every stage of the
// recognition core replaced by its CUDA twin. Parameters unchanged.
#pragma once
#include "MATCH_CUDA.hpp"
#include "CLUSTER_MEAN_SHIFT_CUDA.hpp"
#include "POSE_RANSAC_LM_DIFF_REPROJECTION_CUDA.hpp"
#include "FILTER_PROJECTION_CUDA.hpp"

namespace MopedNS {
	static inline void createCudaRecognitionPipeline( MopedPipeline &pipeline ) {
		pipeline.addAlg( "MATCH_SIFT", new MATCH_CUDA( 128, "SIFT", 5., 0.8) );
		pipeline.addAlg( "CLUSTER", new CLUSTER_MEAN_SHIFT_CUDA( 200, 20, 7, 100) );
		pipeline.addAlg( "POSE", new POSE_RANSAC_LM_DIFF_REPROJECTION_CUDA( 600, 200, 4, 5, 6, 10) );
		pipeline.addAlg( "FILTER", new FILTER_PROJECTION_CUDA( 5, 4096., 2) );
		pipeline.addAlg( "POSE2", new POSE_RANSAC_LM_DIFF_REPROJECTION_CUDA( 100, 500, 4, 6, 8, 5) );
		pipeline.addAlg( "FILTER2", new FILTER_PROJECTION_CUDA( 7, 4096., 3) );
	}
};
