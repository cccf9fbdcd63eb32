// FILTER_PROJECTION_DEPTH_CUDA.hpp — drop-in CUDA replacement of moped3d's FILTER_PROJECTION_DEPTH_CPU
// (moped3d/libmoped/src/filter/FILTER_PROJECTION_DEPTH_CPU.hpp:50-331): same constructor (MinPoints, FeatureDistance, PlausibleSqDistance,
// MinScore, DepthFraction, TestSampleSize, MinKeypointFraction), same config keys, same effect on FrameData: every object is scored by
// reprojection minus a penalty from the depth map, frameData.objects is pruned in place (list order kept) and frameData.clusters[model]
// rebuilt for the survivors. The test points of a model are drawn once, on the first frame, with rand() like the reference does (so a
// process that calls srand() the same way picks the same points). C++98-compatible.
#pragma once
#include "moped_cuda_ctx.hpp"

#include <algorithm>

namespace MopedNS {

	class FILTER_PROJECTION_DEPTH_CUDA : public MopedAlg {

		int MinPoints;
		Float FeatureDistance;
		Float PlausibleSqDistance;
		Float MinScore;
		Float DepthFraction;
		int TestSampleSize;
		Float MinKeypointFraction;

		// test points of all models, model-major (what the reference keeps in `TestPoints`)
		vector<int32_t> testOffsets;
		vector<float> testXYZ;
		bool testPointsSelected;

		void selectTestPoints() {
			testOffsets.assign( 1, 0 );
			testXYZ.clear();
			for( int m=0; m<(int)models->size(); m++ ) {
				// keypoints of every descriptor type, in map order (:102-105)
				vector< Pt<3> > pts;
				const SP_Model &model = (*models)[m];
				for( map< string, vector<Model::IP> >::const_iterator it = model->IPs.begin(); it != model->IPs.end(); ++it )
					for( size_t i=0; i<it->second.size(); i++ ) pts.push_back( it->second[i].coord3D );
				vector<int> chosen;
				if( (int)pts.size() > TestSampleSize ) {
					// randSample (:77-92): one rand() per keypoint in order, the TestSampleSize smallest (key, index) pairs in sorted order
					vector< pair<Float,int> > keyed( pts.size() );
					for( size_t i=0; i<pts.size(); i++ ) keyed[i] = make_pair( (Float)rand(), (int)i );
					std::sort( keyed.begin(), keyed.end() );
					for( int i=0; i<TestSampleSize; i++ ) chosen.push_back( keyed[i].second );
				} else {
					for( size_t i=0; i<pts.size(); i++ ) chosen.push_back( (int)i );
				}
				for( size_t i=0; i<chosen.size(); i++ )
					for( int c=0; c<3; c++ ) testXYZ.push_back( pts[chosen[i]][c] );
				testOffsets.push_back( testOffsets.back() + (int32_t)chosen.size() );
			}
			testPointsSelected = true;
		}

	public:

		FILTER_PROJECTION_DEPTH_CUDA( int MinPoints, Float FeatureDistance, Float PlausibleSqDistance, Float MinScore, Float DepthFraction,
		                              int TestSampleSize, Float MinKeypointFraction )
		: MinPoints(MinPoints), FeatureDistance(FeatureDistance), PlausibleSqDistance(PlausibleSqDistance), MinScore(MinScore),
		  DepthFraction(DepthFraction), TestSampleSize(TestSampleSize), MinKeypointFraction(MinKeypointFraction), testPointsSelected(false) {
		}

		void getConfig( map<string,string> &config ) const {
			GET_CONFIG( MinPoints );
			GET_CONFIG( FeatureDistance );
			GET_CONFIG( MinScore );
			GET_CONFIG( PlausibleSqDistance );
			GET_CONFIG( DepthFraction );
			GET_CONFIG( TestSampleSize );
		}

		void setConfig( map<string,string> &config ) {
			SET_CONFIG( MinPoints );
			SET_CONFIG( FeatureDistance );
			SET_CONFIG( MinScore );
			SET_CONFIG( PlausibleSqDistance );
			SET_CONFIG( DepthFraction );
			SET_CONFIG( TestSampleSize );
		}

		void process( FrameData &frameData ) {

			if( !testPointsSelected ) selectTestPoints();

			vector< vector< FrameData::Match > > &matches = frameData.matches;
			if( matches.size() < models->size() ) return;

			// the depth map and its fill-distance map (:156-176)
			SP_Image depthmap, distanceMap;
			for( size_t i=0; i<frameData.images.size(); i++ )
				if( frameData.images[i]->imageType == IMAGE_TYPE_DEPTH_MAP ) depthmap = frameData.images[i];
			if( !depthmap ) return;
			for( size_t i=0; i<frameData.images.size(); i++ )
				if( frameData.images[i]->imageType == IMAGE_TYPE_PROB_MAP && frameData.images[i]->name == depthmap->name+".distance" ) {
					distanceMap = frameData.images[i];
					break;
				}
			if( !distanceMap ) return;
			const int W = depthmap->width, H = depthmap->height;
			vector<float> depth( (size_t)W*H ), fill( (size_t)W*H );
			for( int y=0; y<H; y++ ) for( int x=0; x<W; x++ ) {
				depth[(size_t)y*W+x] = depthmap->getDepth( x, y );
				fill[(size_t)y*W+x] = distanceMap->getProb( x, y );
			}
			float dK[4], dPose[7];
			for( int j=0; j<4; j++ ) dK[j] = depthmap->intrinsicLinearCalibration[j];
			for( int j=0; j<7; j++ ) dPose[j] = depthmap->cameraPose[j];

			vector<int32_t> off, img; vector<float> xy, xyz;
			MopedCuda::flattenMatches( matches, models->size(), off, img, xy, xyz );
			const int M = off[models->size()];

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
			if( testXYZ.empty() ) testXYZ.resize(3);
			MopedCuda::setCameras( frameData.images );
			MopedCuda::check( mc_filter_projection_depth( MopedCuda::ctx(), &off[0], &img[0], &xy[0], &xyz[0], (int)models->size(), &om[0], &op[0], nObj,
			                                              MinPoints, FeatureDistance, PlausibleSqDistance, MinScore, DepthFraction, MinKeypointFraction,
			                                              &testOffsets[0], &testXYZ[0], dK, dPose, W, H, &depth[0], &fill[0],
			                                              &keep[0], &score[0], &nSurv, &co[0], &mem[0] ), "mc_filter_projection_depth" );

			vector< vector<int> > survivorsOfModel( models->size() );
			int o = 0;
			for( list<SP_Object>::iterator it = frameData.objects->begin(); it != frameData.objects->end(); o++ ) {
				(*it)->score = score[o];
				if( !keep[o] ) it = frameData.objects->erase( it );
				else { survivorsOfModel[om[o]].push_back( o ); ++it; }
			}
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
