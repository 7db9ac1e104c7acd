// GpuScoredSamplingPlanner -- batched replacement for the generator + critics + search that
// dwa_local_planner::DWAPlanner wires together (reference: dwa_local_planner/src/dwa_planner.cpp:116-182, 240-371;
// base_local_planner/include/base_local_planner/trajectory_search.h:49-66).
//
// It implements base_local_planner::TrajectorySearch, so DWAPlanner keeps calling findBestTrajectory(result_traj_,
// &all_explored); what it replaces per control cycle is
//     generator_.initialise(pos, vel, goal, &limits, vsamples_)                       dwa_planner.cpp:309-313
//     scored_sampling_planner_.findBestTrajectory(result_traj_, &all_explored)        dwa_planner.cpp:316-319
//         = 4x MapGridCostFunction::prepare + every sample through SimpleTrajectoryGenerator::generateTrajectory,
//           OscillationCostFunction, ObstacleCostFunction, 4x MapGridCostFunction, first-strictly-smaller argmin
//     oscillation_costs_.updateOscillationFlags(pos, &result_traj_, min_trans_vel)    dwa_planner.cpp:357
// and, per plan update, DWAPlanner::updatePlanAndLocalCosts (:240-286) through setPlan().
// All of it runs in libnavgpu (navgpu_dwa_*); there is no CPU fallback: a failing device call makes
// findBestTrajectory return false with traj.cost_ = -7 (what DWAPlanner treats as "no legal trajectory").
#ifndef NAVGPU_PLUGINS_GPU_SCORED_SAMPLING_PLANNER_H_
#define NAVGPU_PLUGINS_GPU_SCORED_SAMPLING_PLANNER_H_

#include <base_local_planner/local_planner_limits.h>
#include <base_local_planner/trajectory.h>
#include <base_local_planner/trajectory_search.h>
#include <costmap_2d/costmap_2d.h>
#include <geometry_msgs/Point.h>
#include <geometry_msgs/PoseStamped.h>

#include <vector>

#include "navgpu.h"

namespace navgpu_plugins {

class GpuScoredSamplingPlanner : public base_local_planner::TrajectorySearch {
 public:
  // costmap: the local Costmap2D the reference's critics hold a pointer to (dwa_planner.cpp:118-122); it is read
  // (uploaded) at every findBestTrajectory, under the same lock the caller already holds (move_base.cpp:946)
  GpuScoredSamplingPlanner(const navgpu_dwa_config& config, costmap_2d::Costmap2D* costmap, int device = 0);
  ~GpuScoredSamplingPlanner() override;
  GpuScoredSamplingPlanner(const GpuScoredSamplingPlanner&) = delete;
  GpuScoredSamplingPlanner& operator=(const GpuScoredSamplingPlanner&) = delete;

  static navgpu_dwa_config defaultConfig();
  // DWAPlanner::reconfigure (:52-112) with the limits that planner_util_ would hand out
  void reconfigure(const navgpu_dwa_config& config);
  static void applyLimits(navgpu_dwa_config& config, const base_local_planner::LocalPlannerLimits& limits);
  // DWAPlanner::updatePlanAndLocalCosts (:240-286); DWAPlanner::setPlan additionally resets the oscillation flags
  bool setPlan(double pose_x, double pose_y, double pose_yaw, const std::vector<geometry_msgs::PoseStamped>& plan);
  void resetOscillationFlags();
  // the per-cycle inputs of DWAPlanner::findBestPath (:292-313)
  void setState(double pose_x, double pose_y, double pose_yaw, double vel_x, double vel_y, double vel_yaw,
                const std::vector<geometry_msgs::Point>& footprint_spec);

  // base_local_planner::TrajectorySearch
  bool findBestTrajectory(base_local_planner::Trajectory& traj,
                          std::vector<base_local_planner::Trajectory>* all_explored = 0) override;

  // DWAPlanner::checkTrajectory (:213-237): true when the single trajectory for vel_samples is legal
  bool checkTrajectory(double pose_x, double pose_y, double pose_yaw, double vel_x, double vel_y, double vel_yaw,
                       double sample_x, double sample_y, double sample_yaw, double* cost_out = 0);

  int lastStatus() const { return last_status_; }
  int bestIndex() const { return result_.best_index; }
  int samplesScored() const { return result_.n_scored; }
  int oscillationMask() const;

 private:
  bool ensureHandle();
  void firstLoopVelocity(float v[3]) const;
  navgpu_dwa* handle_;
  navgpu_dwa_config config_;
  costmap_2d::Costmap2D* costmap_;
  int device_;
  unsigned int size_x_, size_y_;
  double resolution_;
  double pose_[3], vel_[3];
  std::vector<double> footprint_xy_, plan_xy_, plan_pose_;
  std::vector<double> all_costs_, points_;
  std::vector<float> samples_;
  navgpu_dwa_result result_;
  int last_status_;
};

}  // namespace navgpu_plugins
#endif
