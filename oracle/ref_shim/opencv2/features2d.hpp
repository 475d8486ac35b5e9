// Stand-in (NOT OpenCV): include/utils.hpp includes this header and uses nothing from it on the hot path.
#pragma once
#include "opencv.hpp"
