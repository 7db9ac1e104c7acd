#pragma once
#include <vector>
#include <geometry_msgs/Point32.h>
namespace geometry_msgs { struct Polygon { std::vector<Point32> points; }; }
