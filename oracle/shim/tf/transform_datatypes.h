// Stub of tf/transform_datatypes.h: the planar subset of tf::Stamped<tf::Pose> that TrajectoryPlanner::findBestPath
// touches (trajectory_planner.cpp:908-980).  A Pose is (x, y, z, yaw); getYaw(getRotation()) returns the yaw it was
// built from, bit for bit, so the Eigen::Vector3f rounding of pos / vel in findBestPath sees the caller's values.
#pragma once
#include <string>
namespace tf {
struct Vector3 {
  double v[3];
  Vector3(double x = 0, double y = 0, double z = 0) : v{x, y, z} {}
  double getX() const { return v[0]; }
  double getY() const { return v[1]; }
  double getZ() const { return v[2]; }
  double x() const { return v[0]; }
  double y() const { return v[1]; }
};
struct Quaternion {
  double yaw;
  explicit Quaternion(double y = 0) : yaw(y) {}
};
inline Quaternion createQuaternionFromYaw(double yaw) { return Quaternion(yaw); }
inline double getYaw(const Quaternion& q) { return q.yaw; }
struct Matrix3x3 {
  double yaw = 0;
  void setRotation(const Quaternion& q) { yaw = q.yaw; }
};
struct Pose {
  Vector3 origin;
  double yaw = 0;
  const Vector3& getOrigin() const { return origin; }
  Quaternion getRotation() const { return Quaternion(yaw); }
  void setIdentity() { origin = Vector3(); yaw = 0; }
  void setOrigin(const Vector3& o) { origin = o; }
  void setBasis(const Matrix3x3& m) { yaw = m.yaw; }
};
typedef Pose Transform;
template <class T>
struct Stamped : public T {
  std::string frame_id_;
};
}  // namespace tf
