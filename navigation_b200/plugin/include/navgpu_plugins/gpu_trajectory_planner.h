// GpuTrajectoryPlanner -- the legacy base_local_planner::TrajectoryPlanner with its rollouts on the GPU
// (reference: base_local_planner/include/base_local_planner/trajectory_planner.h:71-212,
//  base_local_planner/src/trajectory_planner.cpp).
//
// Same constructor arguments and the same public members TrajectoryPlannerROS uses (trajectory_planner_ros.cpp:
// tc_->updatePlan, tc_->findBestPath, tc_->checkTrajectory, tc_->scoreTrajectory, tc_->reconfigure semantics through
// setConfig), so TrajectoryPlannerROS switches by changing the type of its `tc_` member.  Per control cycle it replaces
//     path_map_ / goal_map_ wavefronts (MapGrid::setTargetCells / setLocalGoal)               :933-934
//     generateTrajectory for every sample of createTrajectories                                :575-862
// with navgpu_tp_find_best_path (libnavgpu); the selection rules and oscillation / escape flags are applied to the
// device's scores exactly as createTrajectories does.  No CPU fallback: when a device call fails, findBestPath returns
// a trajectory with cost_ = -1 and identity drive velocities (what TrajectoryPlannerROS treats as "no legal command").
#ifndef NAVGPU_PLUGINS_GPU_TRAJECTORY_PLANNER_H_
#define NAVGPU_PLUGINS_GPU_TRAJECTORY_PLANNER_H_

#include <base_local_planner/trajectory.h>
#include <base_local_planner/world_model.h>
#include <costmap_2d/costmap_2d.h>
#include <geometry_msgs/Point.h>
#include <geometry_msgs/PoseStamped.h>
#include <tf/transform_datatypes.h>

#include <vector>

#include "navgpu.h"

namespace navgpu_plugins {

class GpuTrajectoryPlanner {
 public:
  // argument for argument base_local_planner::TrajectoryPlanner's constructor (trajectory_planner.h:106-130);
  // world_model is accepted for source compatibility (the footprint checks run on the costmap, as CostmapModel does)
  GpuTrajectoryPlanner(base_local_planner::WorldModel& world_model, const costmap_2d::Costmap2D& costmap,
                       std::vector<geometry_msgs::Point> footprint_spec, double acc_lim_x = 1.0, double acc_lim_y = 1.0,
                       double acc_lim_theta = 1.0, double sim_time = 1.0, double sim_granularity = 0.025,
                       int vx_samples = 20, int vtheta_samples = 20, double pdist_scale = 0.6, double gdist_scale = 0.8,
                       double occdist_scale = 0.2, double heading_lookahead = 0.325, double oscillation_reset_dist = 0.05,
                       double escape_reset_dist = 0.10, double escape_reset_theta = M_PI_2, bool holonomic_robot = true,
                       double max_vel_x = 0.5, double min_vel_x = 0.1, double max_vel_th = 1.0, double min_vel_th = -1.0,
                       double min_in_place_vel_th = 0.4, double backup_vel = -0.1, bool dwa = false,
                       bool heading_scoring = false, double heading_scoring_timestep = 0.1, bool meter_scoring = true,
                       bool simple_attractor = false, std::vector<double> y_vels = std::vector<double>(0),
                       double stop_time_buffer = 0.2, double sim_period = 0.1, double angular_sim_granularity = 0.025,
                       int device = 0);
  ~GpuTrajectoryPlanner();
  GpuTrajectoryPlanner(const GpuTrajectoryPlanner&) = delete;
  GpuTrajectoryPlanner& operator=(const GpuTrajectoryPlanner&) = delete;

  // TrajectoryPlanner::reconfigure takes the generated BaseLocalPlannerConfig; its fields map 1:1 onto navgpu_tp_config
  const navgpu_tp_config& config() const { return config_; }
  bool setConfig(const navgpu_tp_config& config);

  base_local_planner::Trajectory findBestPath(tf::Stamped<tf::Pose> global_pose, tf::Stamped<tf::Pose> global_vel,
                                              tf::Stamped<tf::Pose>& drive_velocities);
  void updatePlan(const std::vector<geometry_msgs::PoseStamped>& new_plan, bool compute_dists = false);
  bool checkTrajectory(double x, double y, double theta, double vx, double vy, double vtheta, double vx_samp,
                       double vy_samp, double vtheta_samp);
  double scoreTrajectory(double x, double y, double theta, double vx, double vy, double vtheta, double vx_samp,
                         double vy_samp, double vtheta_samp);
  void setFootprint(std::vector<geometry_msgs::Point> footprint);
  std::vector<geometry_msgs::Point> getFootprint() const { return footprint_spec_; }

  int lastStatus() const { return last_status_; }
  int flags() const { return result_.flags; }  // bit layout of navgpu_tp_result::flags

 private:
  bool ensureHandle();
  navgpu_tp* handle_;
  navgpu_tp_config config_;
  const costmap_2d::Costmap2D& costmap_;
  std::vector<geometry_msgs::Point> footprint_spec_;
  int device_;
  unsigned int size_x_, size_y_;
  double resolution_;
  std::vector<double> plan_xy_, points_;
  navgpu_tp_result result_;
  int last_status_;
};

}  // namespace navgpu_plugins
#endif
