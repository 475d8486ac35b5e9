// Stand-in (NOT ROS; test infrastructure): the message the reference publishes for visualisation only.
#pragma once
#include <string>
#include "../ros/ros.h"
namespace sensor_msgs {
struct PointCloud2 { struct { std::string frame_id; ros::Time stamp; } header; };
}
