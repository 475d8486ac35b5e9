// Stand-in (NOT PCL; test infrastructure): included by the reference, nothing of it is used on the path.
#pragma once
#include "../point_cloud.h"
#include "../point_types.h"
