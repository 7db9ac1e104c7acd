#pragma once
#include <geometry_msgs/Twist.h>
#include <std_msgs/Header.h>
namespace nav_msgs { struct Odometry { std_msgs::Header header; geometry_msgs::TwistWithCovariance twist; }; }
