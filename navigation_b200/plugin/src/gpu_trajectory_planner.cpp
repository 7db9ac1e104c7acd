// GpuTrajectoryPlanner: see include/navgpu_plugins/gpu_trajectory_planner.h
#include <navgpu_plugins/gpu_trajectory_planner.h>

#include <cmath>
#include <cstring>

#include <ros/console.h>

namespace navgpu_plugins {

namespace {
std::vector<double> flatten(const std::vector<geometry_msgs::Point>& pts) {
  std::vector<double> xy;
  for (size_t i = 0; i < pts.size(); ++i) {
    xy.push_back(pts[i].x);
    xy.push_back(pts[i].y);
  }
  return xy;
}
}  // namespace

GpuTrajectoryPlanner::GpuTrajectoryPlanner(base_local_planner::WorldModel&, const costmap_2d::Costmap2D& costmap,
                                           std::vector<geometry_msgs::Point> footprint_spec, double acc_lim_x,
                                           double acc_lim_y, double acc_lim_theta, double sim_time, double sim_granularity,
                                           int vx_samples, int vtheta_samples, double pdist_scale, double gdist_scale,
                                           double occdist_scale, double heading_lookahead, double oscillation_reset_dist,
                                           double escape_reset_dist, double escape_reset_theta, bool holonomic_robot,
                                           double max_vel_x, double min_vel_x, double max_vel_th, double min_vel_th,
                                           double min_in_place_vel_th, double backup_vel, bool dwa, bool heading_scoring,
                                           double heading_scoring_timestep, bool /*meter_scoring: unused by the
                                           reference's constructor as well*/, bool simple_attractor,
                                           std::vector<double> y_vels, double stop_time_buffer, double sim_period,
                                           double angular_sim_granularity, int device)
    : handle_(NULL), costmap_(costmap), footprint_spec_(footprint_spec), device_(device), size_x_(0), size_y_(0),
      resolution_(0), last_status_(NAVGPU_OK) {
  navgpu_tp_default_config(&config_);
  config_.acc_lim_x = acc_lim_x; config_.acc_lim_y = acc_lim_y; config_.acc_lim_theta = acc_lim_theta;
  config_.sim_time = sim_time; config_.sim_granularity = sim_granularity;
  config_.angular_sim_granularity = angular_sim_granularity; config_.sim_period = sim_period;
  config_.vx_samples = vx_samples; config_.vtheta_samples = vtheta_samples;
  config_.pdist_scale = pdist_scale; config_.gdist_scale = gdist_scale; config_.occdist_scale = occdist_scale;
  config_.heading_lookahead = heading_lookahead; config_.oscillation_reset_dist = oscillation_reset_dist;
  config_.escape_reset_dist = escape_reset_dist; config_.escape_reset_theta = escape_reset_theta;
  config_.holonomic_robot = holonomic_robot; config_.max_vel_x = max_vel_x; config_.min_vel_x = min_vel_x;
  config_.max_vel_th = max_vel_th; config_.min_vel_th = min_vel_th; config_.min_in_place_vel_th = min_in_place_vel_th;
  config_.backup_vel = backup_vel; config_.dwa = dwa; config_.heading_scoring = heading_scoring;
  config_.heading_scoring_timestep = heading_scoring_timestep; config_.simple_attractor = simple_attractor;
  config_.stop_time_buffer = stop_time_buffer;
  config_.n_y_vels = (int)std::min<size_t>(8, y_vels.size());
  for (int i = 0; i < config_.n_y_vels; ++i) config_.y_vels[i] = y_vels[i];
  memset(&result_, 0, sizeof(result_));
  result_.cost = -1.0;
}

GpuTrajectoryPlanner::~GpuTrajectoryPlanner() {
  if (handle_) navgpu_tp_destroy(handle_);
}

// the device handle is tied to the local costmap's geometry, which Costmap2DROS may change (costmap_2d.cpp:72-85)
bool GpuTrajectoryPlanner::ensureHandle() {
  const unsigned sx = costmap_.getSizeInCellsX(), sy = costmap_.getSizeInCellsY();
  const double res = costmap_.getResolution();
  if (handle_ && sx == size_x_ && sy == size_y_ && res == resolution_) return true;
  if (handle_) {
    navgpu_tp_destroy(handle_);  // note: drops the oscillation / escape state, like re-creating the planner does
    handle_ = NULL;
  }
  const std::vector<double> fp = flatten(footprint_spec_);
  last_status_ = navgpu_tp_create(&handle_, &config_, sx, sy, res, fp.data(), (int)footprint_spec_.size(), device_);
  if (last_status_ != NAVGPU_OK) {
    ROS_ERROR("GpuTrajectoryPlanner: navgpu_tp_create failed (%d): %s", last_status_, navgpu_last_error());
    handle_ = NULL;
    return false;
  }
  size_x_ = sx; size_y_ = sy; resolution_ = res;
  if (!plan_xy_.empty()) navgpu_tp_update_plan(handle_, plan_xy_.data(), (int)(plan_xy_.size() / 2));
  return true;
}

bool GpuTrajectoryPlanner::setConfig(const navgpu_tp_config& config) {
  config_ = config;
  if (!handle_) return true;
  last_status_ = navgpu_tp_reconfigure(handle_, &config_);
  return last_status_ == NAVGPU_OK;
}

void GpuTrajectoryPlanner::setFootprint(std::vector<geometry_msgs::Point> footprint) {
  footprint_spec_ = footprint;
  if (handle_) {
    const std::vector<double> fp = flatten(footprint_spec_);
    last_status_ = navgpu_tp_set_footprint(handle_, fp.data(), (int)footprint_spec_.size());
  }
}

void GpuTrajectoryPlanner::updatePlan(const std::vector<geometry_msgs::PoseStamped>& new_plan, bool /*compute_dists*/) {
  // compute_dists only pre-computes the two distance maps, which findBestPath recomputes anyway (:933-934)
  plan_xy_.clear();
  for (size_t i = 0; i < new_plan.size(); ++i) {
    plan_xy_.push_back(new_plan[i].pose.position.x);
    plan_xy_.push_back(new_plan[i].pose.position.y);
  }
  if (handle_) last_status_ = navgpu_tp_update_plan(handle_, plan_xy_.data(), (int)new_plan.size());
}

base_local_planner::Trajectory GpuTrajectoryPlanner::findBestPath(tf::Stamped<tf::Pose> global_pose,
                                                                  tf::Stamped<tf::Pose> global_vel,
                                                                  tf::Stamped<tf::Pose>& drive_velocities) {
  base_local_planner::Trajectory best;
  best.cost_ = -1.0;
  const double pose[3] = {global_pose.getOrigin().getX(), global_pose.getOrigin().getY(), tf::getYaw(global_pose.getRotation())};
  const double vel[3] = {global_vel.getOrigin().getX(), global_vel.getOrigin().getY(), tf::getYaw(global_vel.getRotation())};
  bool ok = ensureHandle();
  if (ok) {
    // the critics read whatever the costmap holds now (the caller holds its lock, move_base.cpp:946)
    last_status_ = navgpu_tp_set_costmap(handle_, costmap_.getCharMap(), costmap_.getOriginX(), costmap_.getOriginY());
    ok = last_status_ == NAVGPU_OK;
  }
  if (ok) {
    points_.resize(3 * 4096);
    last_status_ = navgpu_tp_find_best_path(handle_, pose, vel, &result_, points_.data(), 4096);
    ok = last_status_ == NAVGPU_OK;
    if (!ok) ROS_ERROR("GpuTrajectoryPlanner: navgpu_tp_find_best_path failed (%d): %s", last_status_, navgpu_last_error());
  }
  if (ok) {
    best.xv_ = result_.xv; best.yv_ = result_.yv; best.thetav_ = result_.thetav; best.cost_ = result_.cost;
    for (int i = 0; i < result_.n_points && i < 4096; ++i) best.addPoint(points_[3 * i], points_[3 * i + 1], points_[3 * i + 2]);
  }
  if (best.cost_ < 0) {  // :968-978
    drive_velocities.setIdentity();
  } else {
    tf::Vector3 start(best.xv_, best.yv_, 0);
    drive_velocities.setOrigin(start);
    tf::Matrix3x3 matrix;
    matrix.setRotation(tf::createQuaternionFromYaw(best.thetav_));
    drive_velocities.setBasis(matrix);
  }
  return best;
}

double GpuTrajectoryPlanner::scoreTrajectory(double x, double y, double theta, double vx, double vy, double vtheta,
                                             double vx_samp, double vy_samp, double vtheta_samp) {
  if (!ensureHandle()) return -1.0;
  const double pose[3] = {x, y, theta}, vel[3] = {vx, vy, vtheta}, samp[3] = {vx_samp, vy_samp, vtheta_samp};
  double cost = -1.0;
  last_status_ = navgpu_tp_set_costmap(handle_, costmap_.getCharMap(), costmap_.getOriginX(), costmap_.getOriginY());
  if (last_status_ == NAVGPU_OK) last_status_ = navgpu_tp_score_trajectory(handle_, pose, vel, samp, &cost);
  return last_status_ == NAVGPU_OK ? cost : -1.0;
}

bool GpuTrajectoryPlanner::checkTrajectory(double x, double y, double theta, double vx, double vy, double vtheta,
                                           double vx_samp, double vy_samp, double vtheta_samp) {
  const double cost = scoreTrajectory(x, y, theta, vx, vy, vtheta, vx_samp, vy_samp, vtheta_samp);
  if (cost >= 0) return true;
  ROS_WARN("Invalid Trajectory %f, %f, %f, cost: %f", vx_samp, vy_samp, vtheta_samp, cost);
  return false;
}

}  // namespace navgpu_plugins
