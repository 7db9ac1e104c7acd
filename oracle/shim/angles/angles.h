// Stub of the `angles` package (ros/angles 1.9.x, angles/include/angles/angles.h, not part of the reference tree):
// the published definitions of the three functions the reference's local planners call.
#pragma once
#include <cmath>
namespace angles {
static inline double normalize_angle_positive(double angle) {
  return fmod(fmod(angle, 2.0 * M_PI) + 2.0 * M_PI, 2.0 * M_PI);
}
static inline double normalize_angle(double angle) {
  double a = normalize_angle_positive(angle);
  if (a > M_PI) a -= 2.0 * M_PI;
  return a;
}
static inline double shortest_angular_distance(double from, double to) { return normalize_angle(to - from); }
}  // namespace angles
