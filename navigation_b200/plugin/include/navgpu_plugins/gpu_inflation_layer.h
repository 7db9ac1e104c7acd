// GpuInflationLayer -- drop-in replacement for costmap_2d::InflationLayer behind the costmap_2d::Layer plugin
// interface (reference: costmap_2d/include/costmap_2d/layer.h:50-130, costmap_2d/include/costmap_2d/inflation_layer.h,
// costmap_2d/plugins/inflation_layer.cpp).  Registered in costmap_plugins.xml as navgpu_plugins::GpuInflationLayer;
// a costmap's `plugins:` list names it instead of costmap_2d::InflationLayer, nothing else changes.
//
// updateBounds keeps the reference's host-side bookkeeping verbatim in behaviour (:125-158); updateCosts hands the
// master grid's affected rows to libnavgpu (navgpu_inflate_host: upload -> k_merge_seed/k_inflate -> download).
// The cached cost table comes from navgpu_build_cost_table, i.e. the reference's computeCost formula evaluated on
// the host (:295-328, inflation_layer.h:114-129).  There is no CPU fallback: a failing device call is reported
// through ROS_ERROR and leaves the master grid untouched, and current_ drops to false so move_base notices.
#ifndef NAVGPU_PLUGINS_GPU_INFLATION_LAYER_H_
#define NAVGPU_PLUGINS_GPU_INFLATION_LAYER_H_

#include <costmap_2d/layer.h>
#include <costmap_2d/layered_costmap.h>
#include <ros/ros.h>

#include <limits>
#include <mutex>
#include <vector>

namespace navgpu_plugins {

class GpuInflationLayer : public costmap_2d::Layer {
 public:
  GpuInflationLayer();
  ~GpuInflationLayer() override {}

  void onInitialize() override;
  void updateBounds(double robot_x, double robot_y, double robot_yaw, double* min_x, double* min_y, double* max_x,
                    double* max_y) override;
  void updateCosts(costmap_2d::Costmap2D& master_grid, int min_i, int min_j, int max_i, int max_j) override;
  void matchSize() override;
  void reset() override { onInitialize(); }
  bool isDiscretized() { return true; }

  // same contract as InflationLayer::computeCost (inflation_layer.h:114-129), answered from the cached table's formula
  unsigned char computeCost(double distance) const;
  // InflationLayer::setInflationParameters (inflation_layer.cpp:356-370)
  void setInflationParameters(double inflation_radius, double cost_scaling_factor);
  void setEnabled(bool enabled);
  void setDevice(int device) { device_ = device; }
  int lastStatus() const { return last_status_; }

 protected:
  void onFootprintChanged() override;

 private:
  void computeCaches();  // InflationLayer::computeCaches (:295-328) through navgpu_build_cost_table
  unsigned int cellDistance(double world_dist) const {
    return layered_costmap_->getCostmap()->cellDistance(world_dist);
  }

  std::recursive_mutex inflation_access_;
  double inflation_radius_, inscribed_radius_, weight_, resolution_;
  unsigned int cell_inflation_radius_;
  std::vector<unsigned char> cached_costs_;  // (R + 2)^2, row-major [dx][dy]
  std::vector<double> cached_distances_;
  double last_min_x_, last_min_y_, last_max_x_, last_max_y_;
  bool need_reinflation_;
  int device_;
  int last_status_;
};

}  // namespace navgpu_plugins
#endif
