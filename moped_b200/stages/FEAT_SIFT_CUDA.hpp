// FEAT_SIFT_CUDA.hpp — drop-in CUDA replacement of the feature-extraction step (B200, libmoped_cuda).
// Same plugin contract as FEAT_SIFT_CPU (moped2/libmoped/src/feat/FEAT_SIFT_CPU.hpp:52-112): constructor (ScaleOrigin),
// the same config key, ScaleOrigin "-1" = libsiftfast's DoubleImSize (:69-76); reads frameData.images[i]->{data, width,
// height} (1 byte per pixel), appends to frameData.detectedFeatures[_stepName] one DetectedFeature per keypoint with
// imageIdx = i, coord2D = (col, row), a 128-float descriptor — in the order FEAT_SIFT_CPU emits them with one OpenMP
// thread (with more, the reference's own order and duplicate suppression race; libsiftfast.cpp:941-951,1188-1196).
// All images of a frame that share a size go to the device as ONE batch. Include after moped.hpp/util.hpp (reference
// tree) or after moped_api.hpp (stand-alone). C++98-compatible.
#pragma once
#include "moped_cuda_ctx.hpp"
#include <cstring>

namespace MopedNS {

	class FEAT_SIFT_CUDA : public MopedAlg {

		string ScaleOrigin;
		int maxKeypoints;                  // per-image slots of the device call; grown when an image has more

	public:

		FEAT_SIFT_CUDA( string ScaleOrigin )
		: ScaleOrigin(ScaleOrigin), maxKeypoints(8192) {
		}

		void getConfig( map<string,string> &config ) const {

			GET_CONFIG( ScaleOrigin );
		}

		void setConfig( map<string,string> &config ) {

			SET_CONFIG( ScaleOrigin );
		}

		void process( FrameData &frameData ) {

			const int doubleSize = ScaleOrigin == "-1" ? 1 : 0;
			vector<FrameData::DetectedFeature> &detectedFeatures = frameData.detectedFeatures[_stepName];

			size_t first = 0;
			while( first < frameData.images.size() ) {

				// a run of images with the same size = one device batch
				const int width = frameData.images[first]->width, height = frameData.images[first]->height;
				size_t last = first + 1;
				while( last < frameData.images.size() && frameData.images[last]->width == width && frameData.images[last]->height == height ) last++;
				const int nImages = (int)(last - first);

				vector<unsigned char> gray( (size_t)nImages * width * height );
				for( int i = 0; i < nImages; i++ )
					memcpy( &gray[(size_t)i * width * height], &frameData.images[first+i]->data[0], (size_t)width * height );

				vector<int32_t> counts( nImages );
				vector<float> xy, desc;
				for(;;) {
					xy.resize( (size_t)nImages * maxKeypoints * 2 );
					desc.resize( (size_t)nImages * maxKeypoints * 128 );
					mc_status st = mc_sift_extract( MopedCuda::ctx(), &gray[0], nImages, height, width, doubleSize, maxKeypoints,
					                                &counts[0], &xy[0], NULL, &desc[0] );
					if( st == MC_ERR_CAPACITY ) {      // counts[] holds what was found: make room and run again
						for( int i = 0; i < nImages; i++ ) if( counts[i] > maxKeypoints ) maxKeypoints = counts[i];
						continue;
					}
					MopedCuda::check( st, "mc_sift_extract" );
					break;
				}

				for( int i = 0; i < nImages; i++ ) {
					for( int k = 0; k < counts[i]; k++ ) {
						const size_t slot = (size_t)i * maxKeypoints + k;
						detectedFeatures.resize( detectedFeatures.size() + 1 );
						FrameData::DetectedFeature &f = detectedFeatures.back();
						f.imageIdx = (int)first + i;
						f.descriptor.assign( &desc[slot*128], &desc[slot*128] + 128 );
						f.coord2D[0] = xy[slot*2];
						f.coord2D[1] = xy[slot*2+1];
					}
				}
				first = last;
			}
		}
	};
};
