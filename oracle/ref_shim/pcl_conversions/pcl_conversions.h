// Stand-in (NOT PCL; test infrastructure): toROSMsg drops the cloud (visualisation only).
#pragma once
#include "../pcl/point_cloud.h"
#include "../sensor_msgs/PointCloud2.h"
namespace pcl { template <typename C> void toROSMsg(const C&, sensor_msgs::PointCloud2&) {} }
