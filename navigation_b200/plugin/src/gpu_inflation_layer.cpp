// GpuInflationLayer: see include/navgpu_plugins/gpu_inflation_layer.h
#include <navgpu_plugins/gpu_inflation_layer.h>

#include <algorithm>
#include <cmath>

#include <costmap_2d/cost_values.h>
#include <pluginlib/class_list_macros.h>

#include "navgpu.h"

PLUGINLIB_EXPORT_CLASS(navgpu_plugins::GpuInflationLayer, costmap_2d::Layer)

namespace navgpu_plugins {

GpuInflationLayer::GpuInflationLayer()
    : inflation_radius_(0),
      inscribed_radius_(0),
      weight_(0),
      resolution_(0),
      cell_inflation_radius_(0),
      last_min_x_(-std::numeric_limits<float>::max()),  // inflation_layer.cpp:63-66
      last_min_y_(-std::numeric_limits<float>::max()),
      last_max_x_(std::numeric_limits<float>::max()),
      last_max_y_(std::numeric_limits<float>::max()),
      need_reinflation_(false),
      device_(0),
      last_status_(NAVGPU_OK) {}

void GpuInflationLayer::onInitialize() {
  {
    std::lock_guard<std::recursive_mutex> lock(inflation_access_);
    ros::NodeHandle nh("~/" + name_);
    current_ = true;
    need_reinflation_ = false;
    // The reference receives these through dynamic_reconfigure, whose server delivers the cfg defaults on start-up
    // (cfg/InflationPlugin.cfg:7-9); plain parameters with the same names and defaults give the same first state.
    bool enabled = true;
    double inflation_radius = 0.55, cost_scaling_factor = 10.0;
    nh.param("enabled", enabled, true);
    nh.param("inflation_radius", inflation_radius, 0.55);
    nh.param("cost_scaling_factor", cost_scaling_factor, 10.0);
    nh.param("device", device_, device_);
    enabled_ = false;  // so that the first setEnabled(true) flags a re-inflation like reconfigureCB does (:104-107)
    setInflationParameters(inflation_radius, cost_scaling_factor);
    setEnabled(enabled);
  }
  matchSize();
}

void GpuInflationLayer::setEnabled(bool enabled) {
  std::lock_guard<std::recursive_mutex> lock(inflation_access_);
  if (enabled_ != enabled) {
    enabled_ = enabled;
    need_reinflation_ = true;
  }
}

void GpuInflationLayer::matchSize() {  // inflation_layer.cpp:110-123 (no seen_ array to size here)
  std::lock_guard<std::recursive_mutex> lock(inflation_access_);
  costmap_2d::Costmap2D* costmap = layered_costmap_->getCostmap();
  resolution_ = costmap->getResolution();
  cell_inflation_radius_ = cellDistance(inflation_radius_);
  computeCaches();
}

void GpuInflationLayer::updateBounds(double, double, double, double* min_x, double* min_y, double* max_x,
                                     double* max_y) {
  if (need_reinflation_) {
    last_min_x_ = *min_x;
    last_min_y_ = *min_y;
    last_max_x_ = *max_x;
    last_max_y_ = *max_y;
    *min_x = -std::numeric_limits<float>::max();
    *min_y = -std::numeric_limits<float>::max();
    *max_x = std::numeric_limits<float>::max();
    *max_y = std::numeric_limits<float>::max();
    need_reinflation_ = false;
  } else {
    const double tmp_min_x = last_min_x_, tmp_min_y = last_min_y_, tmp_max_x = last_max_x_, tmp_max_y = last_max_y_;
    last_min_x_ = *min_x;
    last_min_y_ = *min_y;
    last_max_x_ = *max_x;
    last_max_y_ = *max_y;
    *min_x = std::min(tmp_min_x, *min_x) - inflation_radius_;
    *min_y = std::min(tmp_min_y, *min_y) - inflation_radius_;
    *max_x = std::max(tmp_max_x, *max_x) + inflation_radius_;
    *max_y = std::max(tmp_max_y, *max_y) + inflation_radius_;
  }
}

void GpuInflationLayer::onFootprintChanged() {  // inflation_layer.cpp:160-170
  std::lock_guard<std::recursive_mutex> lock(inflation_access_);
  inscribed_radius_ = layered_costmap_->getInscribedRadius();
  cell_inflation_radius_ = cellDistance(inflation_radius_);
  computeCaches();
  need_reinflation_ = true;
}

void GpuInflationLayer::updateCosts(costmap_2d::Costmap2D& master_grid, int min_i, int min_j, int max_i, int max_j) {
  std::lock_guard<std::recursive_mutex> lock(inflation_access_);
  if (!enabled_ || cell_inflation_radius_ == 0) return;
  // InflationLayer::updateCosts (:172-266): seeds are the LETHAL cells of the window grown by the cell radius,
  // writes may reach one more radius; navgpu_inflate_host stages exactly the rows that can be read or written
  last_status_ = navgpu_inflate_host(master_grid.getCharMap(), master_grid.getSizeInCellsX(),
                                     master_grid.getSizeInCellsY(), min_i, min_j, max_i, max_j, cached_costs_.data(),
                                     cell_inflation_radius_, device_);
  if (last_status_ != NAVGPU_OK) {
    ROS_ERROR("GpuInflationLayer: navgpu_inflate_host failed (%d): %s", last_status_, navgpu_last_error());
    current_ = false;
  } else {
    current_ = true;
  }
}

void GpuInflationLayer::computeCaches() {
  if (resolution_ <= 0) return;
  const unsigned n = cell_inflation_radius_ + 2;
  cached_costs_.assign(size_t(n) * n, 0);
  cached_distances_.assign(size_t(n) * n, 0.0);
  int radius = 0;
  last_status_ = navgpu_build_cost_table(resolution_, inscribed_radius_, inflation_radius_, weight_,
                                         cached_costs_.data(), cached_distances_.data(), (int)(n * n), &radius);
  if (last_status_ != NAVGPU_OK)
    ROS_ERROR("GpuInflationLayer: navgpu_build_cost_table failed (%d): %s", last_status_, navgpu_last_error());
}

unsigned char GpuInflationLayer::computeCost(double distance) const {  // inflation_layer.h:114-129
  unsigned char cost = 0;
  if (distance == 0)
    cost = costmap_2d::LETHAL_OBSTACLE;
  else if (distance * resolution_ <= inscribed_radius_)
    cost = costmap_2d::INSCRIBED_INFLATED_OBSTACLE;
  else {
    const double euclidean_distance = distance * resolution_;
    const double factor = exp(-1.0 * weight_ * (euclidean_distance - inscribed_radius_));
    cost = (unsigned char)((costmap_2d::INSCRIBED_INFLATED_OBSTACLE - 1) * factor);
  }
  return cost;
}

void GpuInflationLayer::setInflationParameters(double inflation_radius, double cost_scaling_factor) {
  if (weight_ != cost_scaling_factor || inflation_radius_ != inflation_radius) {
    std::lock_guard<std::recursive_mutex> lock(inflation_access_);
    inflation_radius_ = inflation_radius;
    if (layered_costmap_ && layered_costmap_->getCostmap()) cell_inflation_radius_ = cellDistance(inflation_radius_);
    weight_ = cost_scaling_factor;
    need_reinflation_ = true;
    computeCaches();
  }
}

}  // namespace navgpu_plugins
