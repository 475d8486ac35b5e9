// Stand-in (NOT PCL; test infrastructure): file I/O of the reference's visualisation / export code; aborts if reached.
#pragma once
#include <cstdlib>
#include <string>
#include "../point_cloud.h"
#include "../point_types.h"
namespace pcl { namespace io {
template <typename P> int loadPCDFile(const std::string&, pcl::PointCloud<P>&) { std::abort(); }
template <typename P> int savePCDFileBinary(const std::string&, const pcl::PointCloud<P>&) { return 0; }
template <typename P> int savePCDFile(const std::string&, const pcl::PointCloud<P>&) { return 0; }
template <typename P> int savePLYFileBinary(const std::string&, const pcl::PointCloud<P>&) { return 0; }
} }
