/*
 * scan_ingest_restated.h -- CPU restatement of the observation ingest path.  TEST INFRASTRUCTURE ONLY.
 *
 * PARITY UNPINNED for the third-party part: the arithmetic of this path lives in dependencies that are NOT in the
 * reference tree (package.xml dependencies of costmap_2d: laser_geometry, pcl_ros, pcl_conversions, tf), so there is
 * nothing to compile and no golden vector in the reference for it.  What is restated, from their published sources:
 *   laser_geometry 1.6.x  LaserProjection::projectLaser_ (src/laser_geometry.cpp): range kept when
 *                         range < range_max && range >= range_min; (x, y) = float(double(range) * {cos, sin}(
 *                         angle_min + double(i) * angle_increment)), z = 0
 *   pcl_ros 1.4.x         transformPointCloud (include/pcl_ros/impl/transforms.hpp): Eigen::Quaternionf from the tf
 *                         rotation, Eigen::Vector3f origin, pcl::transformPointCloud with the Affine3f translation *
 *                         rotation -- per row ((m0 * x + m1 * y) + m2 * z) + t in float (PCL 1.7 / Eigen 3.2)
 * and, from the reference itself (this part IS the reference's algorithm):
 *   ObstacleLayer::laserScanValidInfCallback  plugins/obstacle_layer.cpp:277-292  (+inf -> range_max - 0.0001f)
 *   ObstacleLayer::laserScanCallback          :252-275  (target frame = the scan's own frame: identity transform)
 *   ObstacleLayer::pointCloud(2)Callback      :313-339  (is_cloud: the sensor-frame points go straight to bufferCloud)
 *   ObservationBuffer::bufferCloud            src/observation_buffer.cpp:129-195  (origin = transform of (0,0,0);
 *                                             points with z outside [min, max]_obstacle_height dropped, order kept)
 * Included by both checker libraries so that they export the same symbol.
 */
#ifndef NAV_ORACLE_SCAN_INGEST_RESTATED_H_
#define NAV_ORACLE_SCAN_INGEST_RESTATED_H_
#include <cmath>
#include "oracle_api.h"

extern "C" int navo_project_scan(const navo_laser_scan* s, float* xyz_out, int capacity, double origin_out[3]) {
  /* Eigen::Quaternionf(w, x, y, z).toRotationMatrix() in float (Eigen/src/Geometry/Quaternion.h) */
  const float qx = (float)s->rotation_xyzw[0], qy = (float)s->rotation_xyzw[1], qz = (float)s->rotation_xyzw[2],
              qw = (float)s->rotation_xyzw[3];
  const float tx = 2.0f * qx, ty = 2.0f * qy, tz = 2.0f * qz;
  const float twx = tx * qw, twy = ty * qw, twz = tz * qw, txx = tx * qx, txy = ty * qx, txz = tz * qx, tyy = ty * qy,
              tyz = tz * qy, tzz = tz * qz;
  volatile float m[9];
  m[0] = 1.0f - (tyy + tzz); m[1] = txy - twz; m[2] = txz + twy;
  m[3] = txy + twz; m[4] = 1.0f - (txx + tzz); m[5] = tyz - twx;
  m[6] = txz - twy; m[7] = tyz + twx; m[8] = 1.0f - (txx + tyy);
  const float t[3] = {(float)s->translation[0], (float)s->translation[1], (float)s->translation[2]};
  for (int k = 0; k < 3; ++k) origin_out[k] = s->translation[k];
  int count = 0;
  for (int i = 0; i < s->n_ranges; ++i) {
    float x, y, z;
    if (s->is_cloud) {  /* pointCloudCallback / pointCloud2Callback (obstacle_layer.cpp:313-339): no projection */
      x = s->ranges[3 * i]; y = s->ranges[3 * i + 1]; z = s->ranges[3 * i + 2];
    } else {
      float range = s->ranges[i];
      if (s->inf_is_valid && !std::isfinite(range) && range > 0) range = s->range_max - 0.0001f;
      if (!(range < s->range_max && range >= s->range_min)) continue;
      const double ang = (double)s->angle_min + (double)i * (double)s->angle_increment;
      x = (float)((double)range * cos(ang)); y = (float)((double)range * sin(ang)); z = 0.0f;
    }
    volatile float p0, p1, p2, acc;  /* volatile: one float rounding per operation, no contraction */
    float g[3];
    for (int r = 0; r < 3; ++r) {
      p0 = m[3 * r] * x; p1 = m[3 * r + 1] * y; p2 = m[3 * r + 2] * z;
      acc = p0 + p1;
      acc = acc + p2;
      acc = acc + t[r];
      g[r] = acc;
    }
    if (!((double)g[2] <= s->max_obstacle_height && (double)g[2] >= s->min_obstacle_height)) continue;
    if (count < capacity) { xyz_out[3 * count] = g[0]; xyz_out[3 * count + 1] = g[1]; xyz_out[3 * count + 2] = g[2]; }
    ++count;
  }
  return count;
}
#endif
