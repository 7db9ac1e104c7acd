// Stub of the dynamic_reconfigure-generated BaseLocalPlannerConfig (base_local_planner/cfg/BaseLocalPlanner.cfg):
// just the fields TrajectoryPlanner::reconfigure reads, with the .cfg defaults.
#pragma once
#include <string>
namespace base_local_planner {
struct BaseLocalPlannerConfig {
  double acc_lim_x = 2.5, acc_lim_y = 2.5, acc_lim_theta = 3.2;
  double max_vel_x = 0.5, min_vel_x = 0.1, max_vel_theta = 1.0, min_vel_theta = -1.0, min_in_place_vel_theta = 0.4;
  double sim_time = 1.7, sim_granularity = 0.025, angular_sim_granularity = 0.025;
  double pdist_scale = 0.6, gdist_scale = 0.8, occdist_scale = 0.01;
  double oscillation_reset_dist = 0.05, escape_reset_dist = 0.10, escape_reset_theta = 1.57079632679;
  int vx_samples = 20, vtheta_samples = 20;
  double heading_lookahead = 0.325;
  bool holonomic_robot = true;
  double escape_vel = -0.1;
  bool dwa = false, heading_scoring = false;
  double heading_scoring_timestep = 0.1;
  bool simple_attractor = false;
  std::string y_vels = "-0.3,-0.1,0.1,-0.3";
  bool restore_defaults = false;
};
}  // namespace base_local_planner
