// GpuTrajectoryCostFunction -- the six critics dwa_local_planner::DWAPlanner wires together (dwa_planner.cpp:116-182:
// OscillationCostFunction, ObstacleCostFunction, and the goal_front / alignment / path / goal MapGridCostFunctions) as
// ONE base_local_planner::TrajectoryCostFunction (trajectory_cost_function.h:52-82) whose work runs in libnavgpu.
//
// It is the drop-in for callers that keep the reference's own search and generators:
//     std::vector<TrajectoryCostFunction*> critics(1, &gpu_costs);            // instead of the six CPU critics
//     SimpleScoredSamplingPlanner planner(generator_list, critics);           // dwa_planner.cpp:175-181
// prepare() = the four MapGrid wavefronts of the critics' prepare() (map_grid_cost_function.cpp:59-68) on the costmap as
// it is now; scoreTrajectory(traj) = the sum SimpleScoredSamplingPlanner::scoreTrajectory (:50-79) would have built from
// the six critics in DWAPlanner's order with their scales -- or the first negative code -- so with scale 1 (the
// default) the search selects the very trajectory, with the very cost, it selects with the CPU critics.
// scoreTrajectories() is the batched form the device wants: all trajectories of a sampling run in one launch (one
// warp per trajectory); GpuScoredSamplingPlanner goes one step further and also generates them on the device.
// No CPU fallback: a failing device call makes prepare() return false / scoreTrajectory() return -1.
#ifndef NAVGPU_PLUGINS_GPU_TRAJECTORY_COST_FUNCTION_H_
#define NAVGPU_PLUGINS_GPU_TRAJECTORY_COST_FUNCTION_H_

#include <base_local_planner/trajectory.h>
#include <base_local_planner/trajectory_cost_function.h>
#include <costmap_2d/costmap_2d.h>
#include <geometry_msgs/Point.h>
#include <geometry_msgs/PoseStamped.h>

#include <vector>

#include "navgpu.h"

namespace navgpu_plugins {

class GpuTrajectoryCostFunction : public base_local_planner::TrajectoryCostFunction {
 public:
  // costmap: the local Costmap2D the reference's critics hold a pointer to (dwa_planner.cpp:118-122); it is read
  // (uploaded) at every prepare(), under the lock the caller already holds (move_base.cpp:946)
  GpuTrajectoryCostFunction(const navgpu_dwa_config& config, costmap_2d::Costmap2D* costmap, int device = 0);
  ~GpuTrajectoryCostFunction() override;
  GpuTrajectoryCostFunction(const GpuTrajectoryCostFunction&) = delete;
  GpuTrajectoryCostFunction& operator=(const GpuTrajectoryCostFunction&) = delete;

  // DWAPlanner::reconfigure (:52-112): scales, forward_point_distance, oscillation reset distances
  void reconfigure(const navgpu_dwa_config& config);
  // DWAPlanner::updatePlanAndLocalCosts (:240-286): target poses of the four grid critics
  bool setPlan(double pose_x, double pose_y, double pose_yaw, const std::vector<geometry_msgs::PoseStamped>& plan);
  // ObstacleCostFunction::setFootprint (obstacle_cost_function.cpp:66-68)
  void setFootprint(const std::vector<geometry_msgs::Point>& footprint_spec);
  // OscillationCostFunction's state machine (oscillation_cost_function.cpp:56-164)
  void resetOscillationFlags();
  void updateOscillationFlags(double pose_x, double pose_y, double pose_yaw, base_local_planner::Trajectory* traj);
  int oscillationMask() const;

  // base_local_planner::TrajectoryCostFunction
  bool prepare() override;
  double scoreTrajectory(base_local_planner::Trajectory& traj) override;
  // the same for a whole sampling run in one launch; costs[i] belongs to trajectories[i]
  bool scoreTrajectories(const std::vector<base_local_planner::Trajectory>& trajectories, std::vector<double>* costs);

  int lastStatus() const { return last_status_; }

 private:
  bool ensureHandle();
  navgpu_dwa* handle_;
  navgpu_dwa_config config_;
  costmap_2d::Costmap2D* costmap_;
  int device_;
  unsigned int size_x_, size_y_;
  double resolution_;
  std::vector<double> footprint_xy_, plan_xy_, plan_pose_;
  std::vector<double> points_, vels_, costs_;
  std::vector<int32_t> offsets_;
  int last_status_;
};

}  // namespace navgpu_plugins
#endif
