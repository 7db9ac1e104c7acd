// goal_harness.cpp -- TEST INFRASTRUCTURE: drives the reference's own base_local_planner::transformGlobalPlan and
// prunePlan (base_local_planner/src/goal_functions.cpp:68-174, compiled unmodified against oracle/shim_tf) so that the
// restatement in plan_restated.h -- and through it the batched GPU kernels -- can be checked against the reference's
// loops themselves.  Only tf (absent from the reference tree) is a stand-in: see oracle/shim_tf/tf/*.h.
#include <base_local_planner/goal_functions.h>

#include <vector>

extern "C" int navref_plan_transform(const double* plan_xyz, int n, const double robot_xy[2], const double m[9],
                                     const double t[3], double dist_threshold, int* first_out, double* out_xyz) {
  tf::TransformListener tf;
  for (int i = 0; i < 3; ++i) tf.plan_to_global.basis.r[i] = tf::Vector3(m[3 * i], m[3 * i + 1], m[3 * i + 2]);
  tf.plan_to_global.origin = tf::Vector3(t[0], t[1], t[2]);
  tf.robot_in_plan_frame.origin = tf::Vector3(robot_xy[0], robot_xy[1], 0.0);
  std::vector<geometry_msgs::PoseStamped> plan(n), out;
  for (int i = 0; i < n; ++i) {
    plan[i].header.frame_id = "plan";
    plan[i].pose.position.x = plan_xyz[3 * i];
    plan[i].pose.position.y = plan_xyz[3 * i + 1];
    plan[i].pose.position.z = plan_xyz[3 * i + 2];
  }
  // dist_threshold = max(size_x, size_y) * resolution / 2 (:114-115): a square costmap of that half-width at resolution 1
  const unsigned cells = (unsigned)(2.0 * dist_threshold);
  costmap_2d::Costmap2D costmap(cells, cells, 1.0, 0.0, 0.0);
  if ((double)cells / 2.0 != dist_threshold) return -2;  // the test only uses thresholds a costmap can express
  tf::Stamped<tf::Pose> global_pose;
  if (!base_local_planner::transformGlobalPlan(tf, plan, global_pose, costmap, "global", out)) {
    *first_out = 0;
    return n == 0 ? 0 : -1;
  }
  // the index of the first kept pose is not returned by the reference: recover it from the pose it kept first
  *first_out = n;
  for (size_t k = 0; k < out.size(); ++k) {
    out_xyz[3 * k] = out[k].pose.position.x;
    out_xyz[3 * k + 1] = out[k].pose.position.y;
    out_xyz[3 * k + 2] = out[k].pose.position.z;
  }
  return (int)out.size();
}

extern "C" int navref_plan_prune(const double* plan_xyz, int n, const double robot_xy[2]) {
  tf::Stamped<tf::Pose> global_pose;
  global_pose.origin = tf::Vector3(robot_xy[0], robot_xy[1], 0.0);
  std::vector<geometry_msgs::PoseStamped> plan(n), global_plan(n);
  for (int i = 0; i < n; ++i) {
    plan[i].pose.position.x = plan_xyz[3 * i];
    plan[i].pose.position.y = plan_xyz[3 * i + 1];
  }
  global_plan = plan;
  base_local_planner::prunePlan(global_pose, plan, global_plan);
  return n - (int)plan.size();
}
