// Stand-in (NOT PCL): include/utils.hpp includes this header and uses nothing from it on the hot path.
#pragma once
#include "../point_cloud.h"
#include "../point_types.h"
