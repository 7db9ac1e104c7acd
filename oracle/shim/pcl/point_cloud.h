#pragma once
#include <vector>
#include <std_msgs/Header.h>
namespace pcl { template <class T> struct PointCloud { std_msgs::Header header; std::vector<T> points; unsigned width = 0, height = 0; }; }
