#pragma once
namespace geometry_msgs { struct Point32 { float x = 0, y = 0, z = 0; }; }
