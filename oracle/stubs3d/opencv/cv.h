/* stub: the only OpenCV symbols moped3d's util.hpp touches (MopedAlg::gsToIplImage, util.hpp:162-168); never called by the stages on the path */
#pragma once
typedef struct IplImage { char *imageData; int widthStep; } IplImage;
typedef struct CvSize { int width, height; } CvSize;
static inline CvSize cvSize(int w, int h) { CvSize s = { w, h }; return s; }
#define IPL_DEPTH_8U 8
static inline IplImage *cvCreateImage(CvSize s, int, int) { IplImage *i = new IplImage; i->widthStep = s.width; i->imageData = new char[(size_t)s.width * s.height]; return i; }
/* CLUSTER_LINKAGE_CPU::renderMatrix (debug rendering, never called on the path) */
#define cvZero(img) ((void)0)
#define CV_IMAGE_ELEM(img, T, r, c) (((T *)((img)->imageData + (img)->widthStep * (r)))[(c)])
