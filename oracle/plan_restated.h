/*
 * plan_restated.h -- CPU restatement of the local planners' plan preprocessing.  TEST INFRASTRUCTURE ONLY.
 *
 * The loops are the reference's own (base_local_planner/src/goal_functions.cpp): prunePlan :68-84 and the two loops of
 * transformGlobalPlan :118-149.  PINNED against the reference itself: goal_functions.cpp compiles unmodified against
 * the tf stand-in of oracle/shim_tf into oracle/_ref/libgoalref.so (oracle/goal_harness.cpp, `make goalref`), and
 * tests/test_plan_preprocessing.py checks this restatement bit for bit against those functions and against golden
 * vectors generated from them (tests/golden/plan_*.npz).  What stays PARITY UNPINNED is tf's own arithmetic -- tf is NOT
 * part of the reference tree (package.xml dependency `tf`): the stand-in restates, from tf's published source
 * (tf/LinearMath/Transform.h, Matrix3x3.h, Vector3.h, geometry 1.11.x), Transform::operator()(Vector3) =
 * (basis row . v) + origin with Vector3::dot evaluated x*x' + y*y' + z*z' left to right in double.
 * Included by both checker libraries so that they export the same symbols.
 */
#ifndef NAV_ORACLE_PLAN_RESTATED_H_
#define NAV_ORACLE_PLAN_RESTATED_H_
#include "oracle_api.h"

/* transformGlobalPlan for one plan of n poses (x, y, z): robot_xy = robot_pose in the plan's frame (:110-111),
 * m / t = plan_to_global_transform's basis (row-major) and origin (:103-107).  Returns the number of poses pushed to
 * transformed_plan; *first_out = index of the first of them (n when none). */
extern "C" int navo_plan_transform(const double* plan_xyz, int n, const double robot_xy[2], const double m[9],
                                   const double t[3], double dist_threshold, int* first_out, double* out_xyz) {
  unsigned int i = 0;
  const double sq_dist_threshold = dist_threshold * dist_threshold;
  double sq_dist = 0;
  /* :122-130 we need to loop to a point on the plan that is within a certain distance of the robot */
  while (i < (unsigned int)n) {
    const double x_diff = robot_xy[0] - plan_xyz[3 * i], y_diff = robot_xy[1] - plan_xyz[3 * i + 1];
    sq_dist = x_diff * x_diff + y_diff * y_diff;
    if (sq_dist <= sq_dist_threshold) break;
    ++i;
  }
  *first_out = (int)i;
  int pushed = 0;
  /* :135-149 now we'll transform until points are outside of our distance threshold */
  while (i < (unsigned int)n && sq_dist <= sq_dist_threshold) {
    const double* p = plan_xyz + 3 * i;
    double* o = out_xyz + 3 * pushed;
    o[0] = m[0] * p[0] + m[1] * p[1] + m[2] * p[2] + t[0];
    o[1] = m[3] * p[0] + m[4] * p[1] + m[5] * p[2] + t[1];
    o[2] = m[6] * p[0] + m[7] * p[1] + m[8] * p[2] + t[2];
    ++pushed;
    const double x_diff = robot_xy[0] - p[0], y_diff = robot_xy[1] - p[1];
    sq_dist = x_diff * x_diff + y_diff * y_diff;
    ++i;
  }
  return pushed;
}

/* prunePlan :68-84: number of way-points erased from the front of `plan` (and of global_plan) */
extern "C" int navo_plan_prune(const double* plan_xyz, int n, const double robot_xy[2]) {
  int erased = 0;
  while (erased < n) {
    const double x_diff = robot_xy[0] - plan_xyz[3 * erased], y_diff = robot_xy[1] - plan_xyz[3 * erased + 1];
    const double distance_sq = x_diff * x_diff + y_diff * y_diff;
    if (distance_sq < 1) break;
    ++erased;
  }
  return erased;
}
#endif
