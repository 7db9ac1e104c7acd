#pragma once
#include <std_msgs/Header.h>
#include <geometry_msgs/Point.h>
namespace geometry_msgs {
struct Quaternion { double x = 0, y = 0, z = 0, w = 1; };
struct Pose { Point position; Quaternion orientation; };
struct PoseStamped { std_msgs::Header header; Pose pose; };
}
