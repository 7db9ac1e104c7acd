// GpuTrajectoryCostFunction: see include/navgpu_plugins/gpu_trajectory_cost_function.h
#include <navgpu_plugins/gpu_trajectory_cost_function.h>

#include <ros/console.h>

namespace navgpu_plugins {

GpuTrajectoryCostFunction::GpuTrajectoryCostFunction(const navgpu_dwa_config& config, costmap_2d::Costmap2D* costmap,
                                                     int device)
    : handle_(NULL), config_(config), costmap_(costmap), device_(device), size_x_(0), size_y_(0), resolution_(0),
      last_status_(NAVGPU_OK) {}

GpuTrajectoryCostFunction::~GpuTrajectoryCostFunction() {
  if (handle_) navgpu_dwa_destroy(handle_);
}

// The device handle is tied to the local costmap's geometry; Costmap2DROS may resize it (costmap_2d.cpp:72-85)
bool GpuTrajectoryCostFunction::ensureHandle() {
  const unsigned sx = costmap_->getSizeInCellsX(), sy = costmap_->getSizeInCellsY();
  const double res = costmap_->getResolution();
  if (handle_ && sx == size_x_ && sy == size_y_ && res == resolution_) return true;
  int mask = 0;
  const bool had = handle_ != NULL;
  if (handle_) {
    navgpu_dwa_get_oscillation_mask(handle_, &mask);
    navgpu_dwa_destroy(handle_);
    handle_ = NULL;
  }
  last_status_ = navgpu_dwa_create(&handle_, &config_, sx, sy, res, device_);
  if (last_status_ != NAVGPU_OK) {
    ROS_ERROR("GpuTrajectoryCostFunction: navgpu_dwa_create failed (%d): %s", last_status_, navgpu_last_error());
    handle_ = NULL;
    return false;
  }
  size_x_ = sx; size_y_ = sy; resolution_ = res;
  if (had && !plan_xy_.empty())
    navgpu_dwa_set_plan(handle_, plan_pose_.data(), plan_xy_.data(), (int)(plan_xy_.size() / 2));
  return true;
}

void GpuTrajectoryCostFunction::reconfigure(const navgpu_dwa_config& config) {
  config_ = config;
  if (handle_) last_status_ = navgpu_dwa_reconfigure(handle_, &config_);
}

bool GpuTrajectoryCostFunction::setPlan(double x, double y, double yaw, const std::vector<geometry_msgs::PoseStamped>& plan) {
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

void GpuTrajectoryCostFunction::setFootprint(const std::vector<geometry_msgs::Point>& footprint_spec) {
  footprint_xy_.resize(2 * footprint_spec.size());
  for (size_t i = 0; i < footprint_spec.size(); ++i) {
    footprint_xy_[2 * i] = footprint_spec[i].x;
    footprint_xy_[2 * i + 1] = footprint_spec[i].y;
  }
}

void GpuTrajectoryCostFunction::resetOscillationFlags() {
  if (handle_) navgpu_dwa_reset_oscillation(handle_);
}

void GpuTrajectoryCostFunction::updateOscillationFlags(double x, double y, double yaw, base_local_planner::Trajectory* traj) {
  if (!handle_ || !traj) return;
  const double pose[3] = {x, y, yaw};
  navgpu_dwa_update_oscillation(handle_, pose, traj->cost_, traj->xv_, traj->yv_, traj->thetav_);
}

int GpuTrajectoryCostFunction::oscillationMask() const {
  int m = 0;
  if (handle_) navgpu_dwa_get_oscillation_mask(handle_, &m);
  return m;
}

bool GpuTrajectoryCostFunction::prepare() {
  if (!ensureHandle()) return false;
  // the critics read whatever the costmap holds now (dwa_planner.cpp:118-122)
  last_status_ = navgpu_dwa_set_costmap(handle_, costmap_->getCharMap(), costmap_->getOriginX(), costmap_->getOriginY());
  if (last_status_ == NAVGPU_OK) last_status_ = navgpu_dwa_prepare(handle_);
  if (last_status_ != NAVGPU_OK) ROS_ERROR("GpuTrajectoryCostFunction::prepare failed (%d): %s", last_status_, navgpu_last_error());
  return last_status_ == NAVGPU_OK;
}

bool GpuTrajectoryCostFunction::scoreTrajectories(const std::vector<base_local_planner::Trajectory>& trajectories,
                                                  std::vector<double>* costs) {
  if (!costs || !ensureHandle()) return false;
  const size_t n = trajectories.size();
  offsets_.assign(n + 1, 0);
  vels_.resize(3 * n);
  points_.clear();
  for (size_t t = 0; t < n; ++t) {
    const base_local_planner::Trajectory& tr = trajectories[t];
    const unsigned np = tr.getPointsSize();
    for (unsigned i = 0; i < np; ++i) {
      double px, py, pth;
      tr.getPoint(i, px, py, pth);
      points_.push_back(px);
      points_.push_back(py);
      points_.push_back(pth);
    }
    offsets_[t + 1] = offsets_[t] + (int32_t)np;
    vels_[3 * t] = tr.xv_; vels_[3 * t + 1] = tr.yv_; vels_[3 * t + 2] = tr.thetav_;
  }
  costs->assign(n, -1.0);
  last_status_ = navgpu_dwa_score_trajectories(handle_, (int)n, offsets_.data(), points_.data(), vels_.data(),
                                               footprint_xy_.empty() ? NULL : footprint_xy_.data(),
                                               (int)(footprint_xy_.size() / 2), costs->data(), NULL);
  if (last_status_ != NAVGPU_OK) {
    ROS_ERROR("GpuTrajectoryCostFunction: navgpu_dwa_score_trajectories failed (%d): %s", last_status_, navgpu_last_error());
    return false;
  }
  return true;
}

double GpuTrajectoryCostFunction::scoreTrajectory(base_local_planner::Trajectory& traj) {
  std::vector<base_local_planner::Trajectory> one(1, traj);
  if (!scoreTrajectories(one, &costs_)) return -1.0;
  return costs_[0];
}

}  // namespace navgpu_plugins
