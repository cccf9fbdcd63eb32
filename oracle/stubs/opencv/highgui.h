/* Empty stand-in: the reference's util.hpp includes <opencv/cv.h>, but none of the four
 * hot-path stage headers uses an OpenCV symbol. OpenCV headers are absent here. */
