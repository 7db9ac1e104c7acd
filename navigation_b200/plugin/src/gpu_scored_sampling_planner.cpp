// GpuScoredSamplingPlanner: see include/navgpu_plugins/gpu_scored_sampling_planner.h
#include <navgpu_plugins/gpu_scored_sampling_planner.h>

#include <cmath>
#include <cstring>

#include <ros/console.h>

namespace navgpu_plugins {

GpuScoredSamplingPlanner::GpuScoredSamplingPlanner(const navgpu_dwa_config& config, costmap_2d::Costmap2D* costmap,
                                                   int device)
    : handle_(NULL), config_(config), costmap_(costmap), device_(device), size_x_(0), size_y_(0), resolution_(0),
      last_status_(NAVGPU_OK) {
  memset(&result_, 0, sizeof(result_));
  result_.cost = -7.0;
  result_.best_index = -1;
  for (int k = 0; k < 3; ++k) pose_[k] = vel_[k] = 0.0;
}

GpuScoredSamplingPlanner::~GpuScoredSamplingPlanner() {
  if (handle_) navgpu_dwa_destroy(handle_);
}

navgpu_dwa_config GpuScoredSamplingPlanner::defaultConfig() {
  navgpu_dwa_config c;
  navgpu_dwa_default_config(&c);
  return c;
}

void GpuScoredSamplingPlanner::applyLimits(navgpu_dwa_config& c, const base_local_planner::LocalPlannerLimits& l) {
  c.max_trans_vel = l.max_trans_vel; c.min_trans_vel = l.min_trans_vel;
  c.max_vel_x = l.max_vel_x; c.min_vel_x = l.min_vel_x;
  c.max_vel_y = l.max_vel_y; c.min_vel_y = l.min_vel_y;
  c.max_rot_vel = l.max_rot_vel; c.min_rot_vel = l.min_rot_vel;
  c.acc_lim_x = l.acc_lim_x; c.acc_lim_y = l.acc_lim_y; c.acc_lim_theta = l.acc_lim_theta;
}

// The device handle is tied to the local costmap's geometry; Costmap2DROS may resize it (costmap_2d.cpp:72-85)
bool GpuScoredSamplingPlanner::ensureHandle() {
  const unsigned sx = costmap_->getSizeInCellsX(), sy = costmap_->getSizeInCellsY();
  const double res = costmap_->getResolution();
  if (handle_ && sx == size_x_ && sy == size_y_ && res == resolution_) return true;
  int mask_plan = 0;
  if (handle_) {
    navgpu_dwa_destroy(handle_);
    handle_ = NULL;
    mask_plan = 1;
  }
  last_status_ = navgpu_dwa_create(&handle_, &config_, sx, sy, res, device_);
  if (last_status_ != NAVGPU_OK) {
    ROS_ERROR("GpuScoredSamplingPlanner: navgpu_dwa_create failed (%d): %s", last_status_, navgpu_last_error());
    handle_ = NULL;
    return false;
  }
  size_x_ = sx; size_y_ = sy; resolution_ = res;
  if (mask_plan && !plan_xy_.empty())
    navgpu_dwa_set_plan(handle_, plan_pose_.data(), plan_xy_.data(), (int)(plan_xy_.size() / 2));
  return true;
}

void GpuScoredSamplingPlanner::reconfigure(const navgpu_dwa_config& config) {
  config_ = config;
  if (handle_) last_status_ = navgpu_dwa_reconfigure(handle_, &config_);
}

bool GpuScoredSamplingPlanner::setPlan(double x, double y, double yaw, const std::vector<geometry_msgs::PoseStamped>& plan) {
  plan_xy_.resize(2 * plan.size());
  for (size_t i = 0; i < plan.size(); ++i) {
    plan_xy_[2 * i] = plan[i].pose.position.x;
    plan_xy_[2 * i + 1] = plan[i].pose.position.y;
  }
  plan_pose_.assign(3, 0.0);
  plan_pose_[0] = x; plan_pose_[1] = y; plan_pose_[2] = yaw;
  if (plan.empty() || !ensureHandle()) return false;
  last_status_ = navgpu_dwa_set_plan(handle_, plan_pose_.data(), plan_xy_.data(), (int)plan.size());
  return last_status_ == NAVGPU_OK;
}

void GpuScoredSamplingPlanner::resetOscillationFlags() {
  if (handle_) navgpu_dwa_reset_oscillation(handle_);
}

int GpuScoredSamplingPlanner::oscillationMask() const {
  int m = 0;
  if (handle_) navgpu_dwa_get_oscillation_mask(handle_, &m);
  return m;
}

void GpuScoredSamplingPlanner::setState(double x, double y, double yaw, double vx, double vy, double vyaw,
                                        const std::vector<geometry_msgs::Point>& footprint_spec) {
  pose_[0] = x; pose_[1] = y; pose_[2] = yaw;
  vel_[0] = vx; vel_[1] = vy; vel_[2] = vyaw;
  footprint_xy_.resize(2 * footprint_spec.size());
  for (size_t i = 0; i < footprint_spec.size(); ++i) {
    footprint_xy_[2 * i] = footprint_spec[i].x;
    footprint_xy_[2 * i + 1] = footprint_spec[i].y;
  }
}

bool GpuScoredSamplingPlanner::checkTrajectory(double x, double y, double yaw, double vx, double vy, double vyaw,
                                               double sx, double sy, double syaw, double* cost_out) {
  if (!ensureHandle()) return false;
  // DWAPlanner::checkTrajectory is LatchedStopRotateController's collision check (dwa_planner_ros.cpp:271-288) and runs
  // on cycles WITHOUT a findBestPath: the critics must see the costmap as it is now (dwa_planner.cpp:118-122), not the
  // snapshot of the last search.  The MapGrid critics keep their last prepare(), like the reference's.
  last_status_ = navgpu_dwa_set_costmap(handle_, costmap_->getCharMap(), costmap_->getOriginX(), costmap_->getOriginY());
  if (last_status_ != NAVGPU_OK) {
    ROS_ERROR("GpuScoredSamplingPlanner: navgpu_dwa_set_costmap failed (%d): %s", last_status_, navgpu_last_error());
    return false;
  }
  const double pose[3] = {x, y, yaw}, vel[3] = {vx, vy, vyaw}, samp[3] = {sx, sy, syaw};
  double cost = -1.0;
  last_status_ = navgpu_dwa_check_trajectory(handle_, pose, vel, samp, footprint_xy_.empty() ? NULL : footprint_xy_.data(),
                                             (int)(footprint_xy_.size() / 2), &cost);
  if (last_status_ != NAVGPU_OK) {
    ROS_ERROR("GpuScoredSamplingPlanner: navgpu_dwa_check_trajectory failed (%d): %s", last_status_, navgpu_last_error());
    return false;
  }
  if (cost_out) *cost_out = cost;
  if (cost < 0) ROS_WARN("Invalid Trajectory %f, %f, %f, cost: %f", sx, sy, syaw, cost);
  return cost >= 0;
}

// SimpleTrajectoryGenerator::generateTrajectory's first computeNewVelocities step without use_dwa
// (simple_trajectory_generator.cpp:203-227, 265-276): float state, double time step
void GpuScoredSamplingPlanner::firstLoopVelocity(float v[3]) const {
  const double vmag = hypot((double)v[0], (double)v[1]);
  const int num_steps = (int)ceil(std::max(vmag * config_.sim_time / config_.sim_granularity,
                                           fabs((double)v[2]) * config_.sim_time / config_.angular_sim_granularity));
  if (num_steps <= 0) return;
  const double dt = config_.sim_time / num_steps;
  const float cur[3] = {(float)vel_[0], (float)vel_[1], (float)vel_[2]};
  const float acc[3] = {(float)config_.acc_lim_x, (float)config_.acc_lim_y, (float)config_.acc_lim_theta};
  for (int k = 0; k < 3; ++k)
    v[k] = cur[k] < v[k] ? (float)std::min((double)v[k], cur[k] + acc[k] * dt) : (float)std::max((double)v[k], cur[k] - acc[k] * dt);
}

bool GpuScoredSamplingPlanner::findBestTrajectory(base_local_planner::Trajectory& traj,
                                                  std::vector<base_local_planner::Trajectory>* all_explored) {
  traj.cost_ = -7.0;  // dwa_planner.cpp:316
  if (!ensureHandle()) return false;
  // the critics read whatever the costmap holds now (dwa_planner.cpp:118-122): upload it for this cycle
  last_status_ = navgpu_dwa_set_costmap(handle_, costmap_->getCharMap(), costmap_->getOriginX(), costmap_->getOriginY());
  if (last_status_ != NAVGPU_OK) {
    ROS_ERROR("GpuScoredSamplingPlanner: navgpu_dwa_set_costmap failed (%d): %s", last_status_, navgpu_last_error());
    return false;
  }
  const int nx = config_.vx_samples > 0 ? config_.vx_samples : 1, ny = config_.vy_samples > 0 ? config_.vy_samples : 1,
            nth = config_.vth_samples > 0 ? config_.vth_samples : 1;
  const size_t max_samples = size_t(nx + 1) * (ny + 1) * (nth + 1);  // VelocityIterator may insert a zero per axis
  const int points_capacity = 4096;
  points_.resize(size_t(3) * points_capacity);
  double* costs = NULL;
  if (all_explored) {
    all_costs_.assign(max_samples, 0.0);
    costs = all_costs_.data();
  }
  last_status_ = navgpu_dwa_find_best_path(handle_, pose_, vel_, footprint_xy_.empty() ? NULL : footprint_xy_.data(),
                                           (int)(footprint_xy_.size() / 2), &result_, costs, (int)max_samples,
                                           points_.data(), points_capacity);
  if (last_status_ != NAVGPU_OK) {
    ROS_ERROR("GpuScoredSamplingPlanner: navgpu_dwa_find_best_path failed (%d): %s", last_status_, navgpu_last_error());
    return false;
  }
  if (all_explored) {
    // the reference copies every generated trajectory with its reported cost (simple_scored_sampling_planner.cpp
    // :106-109); samples its generator rejects do not appear (NaN here).  Points of the losers are not materialised.
    all_explored->clear();
    int32_t counts[3] = {0, 0, 0};
    samples_.resize(size_t(nx + 1) + (ny + 1) + (nth + 1));
    navgpu_dwa_get_samples(handle_, counts, samples_.data(), (int)samples_.size());
    const float* xs = samples_.data();
    const float* ys = xs + counts[0];
    const float* ths = ys + counts[1];
    for (int i = 0; i < result_.n_samples; ++i) {
      if (std::isnan(all_costs_[i])) continue;
      base_local_planner::Trajectory t;
      t.cost_ = all_costs_[i];
      // traj.xv_, yv_, thetav_ as generateTrajectory sets them (simple_trajectory_generator.cpp:218-227): the sample
      // itself with use_dwa, else the velocity after the first acceleration step
      const int ith = i % counts[2], iy = (i / counts[2]) % counts[1], ix = i / (counts[2] * counts[1]);
      float v[3] = {xs[ix], ys[iy], ths[ith]};
      if (!config_.use_dwa) firstLoopVelocity(v);
      t.xv_ = v[0]; t.yv_ = v[1]; t.thetav_ = v[2];
      all_explored->push_back(t);
    }
  }
  traj.cost_ = result_.cost;
  if (result_.cost >= 0) {  // simple_scored_sampling_planner.cpp:123-134
    traj.xv_ = result_.xv;
    traj.yv_ = result_.yv;
    traj.thetav_ = result_.thetav;
    traj.resetPoints();
    for (int i = 0; i < result_.n_points && i < points_capacity; ++i)
      traj.addPoint(points_[3 * i], points_[3 * i + 1], points_[3 * i + 2]);
  }
  return result_.cost >= 0;
}

}  // namespace navgpu_plugins
