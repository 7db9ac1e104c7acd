#pragma once
#include <vector>
#include <std_msgs/Header.h>
#include <geometry_msgs/Point32.h>
namespace sensor_msgs { struct PointCloud { std_msgs::Header header; std::vector<geometry_msgs::Point32> points; }; }
