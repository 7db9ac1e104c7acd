#pragma once
#include <string>
#include <cstdint>
namespace ros { struct Time { double t = 0; Time() {} explicit Time(double v) : t(v) {} static Time now() { return Time(); } double toSec() const { return t; } }; }
namespace std_msgs { struct Header { uint32_t seq = 0; ros::Time stamp; std::string frame_id; }; }
