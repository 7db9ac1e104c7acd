// Stub of tf/transform_datatypes.h for compiling the reference's base_local_planner/src/goal_functions.cpp UNMODIFIED
// (TEST INFRASTRUCTURE; a separate include root from oracle/shim because the planar stub there serves
// trajectory_planner.cpp).  tf is not part of the reference tree: what follows restates, from tf's published source
// (geometry 1.11.x, tf/LinearMath/{Vector3,Matrix3x3,Quaternion,Transform}.h), exactly the members goal_functions.cpp
// touches.  Transform::operator()(v) = (basis row . v) + origin with Vector3::dot = x*x' + y*y' + z*z' left to right;
// Matrix3x3::setRotation / getRotation as published.  PARITY UNPINNED for this arithmetic (see plan_restated.h).
#pragma once
#include <cmath>
#include <stdexcept>
#include <string>
#include <geometry_msgs/PoseStamped.h>
namespace tf {
struct Vector3 {
  double m[3];
  Vector3(double x = 0, double y = 0, double z = 0) : m{x, y, z} {}
  double x() const { return m[0]; }
  double y() const { return m[1]; }
  double z() const { return m[2]; }
  double getX() const { return m[0]; }
  double getY() const { return m[1]; }
  double getZ() const { return m[2]; }
  double dot(const Vector3& v) const { return m[0] * v.m[0] + m[1] * v.m[1] + m[2] * v.m[2]; }
};
struct Quaternion {
  double q[4];  // x, y, z, w
  Quaternion(double x = 0, double y = 0, double z = 0, double w = 1) : q{x, y, z, w} {}
  double x() const { return q[0]; }
  double y() const { return q[1]; }
  double z() const { return q[2]; }
  double w() const { return q[3]; }
  double length2() const { return q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]; }
};
struct Matrix3x3 {
  Vector3 r[3];
  Matrix3x3() { r[0] = Vector3(1, 0, 0); r[1] = Vector3(0, 1, 0); r[2] = Vector3(0, 0, 1); }
  explicit Matrix3x3(const Quaternion& q) { setRotation(q); }
  void setRotation(const Quaternion& q) {  // Matrix3x3.h: setRotation
    const double d = q.length2(), s = 2.0 / d;
    const double xs = q.x() * s, ys = q.y() * s, zs = q.z() * s;
    const double wx = q.w() * xs, wy = q.w() * ys, wz = q.w() * zs;
    const double xx = q.x() * xs, xy = q.x() * ys, xz = q.x() * zs;
    const double yy = q.y() * ys, yz = q.y() * zs, zz = q.z() * zs;
    r[0] = Vector3(1.0 - (yy + zz), xy - wz, xz + wy);
    r[1] = Vector3(xy + wz, 1.0 - (xx + zz), yz - wx);
    r[2] = Vector3(xz - wy, yz + wx, 1.0 - (xx + yy));
  }
  const Vector3& operator[](int i) const { return r[i]; }
  double tdotx(const Vector3& v) const { return r[0].x() * v.x() + r[1].x() * v.y() + r[2].x() * v.z(); }
  double tdoty(const Vector3& v) const { return r[0].y() * v.x() + r[1].y() * v.y() + r[2].y() * v.z(); }
  double tdotz(const Vector3& v) const { return r[0].z() * v.x() + r[1].z() * v.y() + r[2].z() * v.z(); }
  Matrix3x3 operator*(const Matrix3x3& o) const {
    Matrix3x3 out;
    for (int i = 0; i < 3; ++i) out.r[i] = Vector3(o.tdotx(r[i]), o.tdoty(r[i]), o.tdotz(r[i]));
    return out;
  }
  void getRotation(Quaternion& q) const {  // Matrix3x3.h: getRotation
    const double trace = r[0].x() + r[1].y() + r[2].z();
    double t[4];
    if (trace > 0.0) {
      double s = sqrt(trace + 1.0);
      t[3] = s * 0.5;
      s = 0.5 / s;
      t[0] = (r[2].y() - r[1].z()) * s;
      t[1] = (r[0].z() - r[2].x()) * s;
      t[2] = (r[1].x() - r[0].y()) * s;
    } else {
      const int i = r[0].x() < r[1].y() ? (r[1].y() < r[2].z() ? 2 : 1) : (r[0].x() < r[2].z() ? 2 : 0);
      const int j = (i + 1) % 3, k = (i + 2) % 3;
      double s = sqrt(r[i].m[i] - r[j].m[j] - r[k].m[k] + 1.0);
      t[i] = s * 0.5;
      s = 0.5 / s;
      t[3] = (r[k].m[j] - r[j].m[k]) * s;
      t[j] = (r[j].m[i] + r[i].m[j]) * s;
      t[k] = (r[k].m[i] + r[i].m[k]) * s;
    }
    q = Quaternion(t[0], t[1], t[2], t[3]);
  }
};
struct Transform {
  Matrix3x3 basis;
  Vector3 origin;
  Transform() {}
  Transform(const Matrix3x3& b, const Vector3& o) : basis(b), origin(o) {}
  Transform(const Quaternion& q, const Vector3& o) : basis(q), origin(o) {}
  const Vector3& getOrigin() const { return origin; }
  Quaternion getRotation() const { Quaternion q; basis.getRotation(q); return q; }
  Vector3 operator()(const Vector3& x) const {  // Transform.h: operator()
    return Vector3(basis[0].dot(x) + origin.x(), basis[1].dot(x) + origin.y(), basis[2].dot(x) + origin.z());
  }
  Transform operator*(const Transform& t) const { return Transform(basis * t.basis, (*this)(t.origin)); }
};
typedef Transform Pose;
template <class T>
struct Stamped : public T {
  ros::Time stamp_;
  std::string frame_id_;
  Stamped() {}
  void setData(const T& input) { *static_cast<T*>(this) = input; }
};
struct StampedTransform : public Transform {
  ros::Time stamp_;
  std::string frame_id_, child_frame_id_;
};
inline double getYaw(const Quaternion& q) {  // tf::getYaw via Matrix3x3::getRPY's yaw for the planar case
  return atan2(2.0 * (q.w() * q.z() + q.x() * q.y()), 1.0 - 2.0 * (q.y() * q.y() + q.z() * q.z()));
}
inline void poseStampedMsgToTF(const geometry_msgs::PoseStamped& msg, Stamped<Pose>& bt) {
  bt.setData(Transform(Quaternion(msg.pose.orientation.x, msg.pose.orientation.y, msg.pose.orientation.z, msg.pose.orientation.w),
                       Vector3(msg.pose.position.x, msg.pose.position.y, msg.pose.position.z)));
  bt.stamp_ = msg.header.stamp;
  bt.frame_id_ = msg.header.frame_id;
}
inline void poseStampedTFToMsg(const Stamped<Pose>& bt, geometry_msgs::PoseStamped& msg) {
  msg.pose.position.x = bt.getOrigin().x();
  msg.pose.position.y = bt.getOrigin().y();
  msg.pose.position.z = bt.getOrigin().z();
  const Quaternion q = bt.getRotation();
  msg.pose.orientation.x = q.x(); msg.pose.orientation.y = q.y(); msg.pose.orientation.z = q.z(); msg.pose.orientation.w = q.w();
  msg.header.stamp = bt.stamp_;
  msg.header.frame_id = bt.frame_id_;
}
struct TransformException : public std::runtime_error { TransformException(const std::string& s) : std::runtime_error(s) {} };
struct LookupException : public TransformException { LookupException(const std::string& s) : TransformException(s) {} };
struct ConnectivityException : public TransformException { ConnectivityException(const std::string& s) : TransformException(s) {} };
struct ExtrapolationException : public TransformException { ExtrapolationException(const std::string& s) : TransformException(s) {} };
}  // namespace tf
