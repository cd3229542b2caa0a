// Stand-in for <opencv/cv.h> (see ../opencv2/opencv.hpp).  TEST INFRASTRUCTURE.
#include <opencv2/opencv.hpp>
