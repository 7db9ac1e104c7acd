#pragma once
namespace pcl { struct PointXYZ { float x = 0, y = 0, z = 0; float pad_ = 0; }; }
