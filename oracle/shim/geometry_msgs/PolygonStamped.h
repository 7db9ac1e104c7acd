#pragma once
#include <std_msgs/Header.h>
#include <geometry_msgs/Polygon.h>
namespace geometry_msgs { struct PolygonStamped { std_msgs::Header header; Polygon polygon; }; }
