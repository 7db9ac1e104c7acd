// GpuLayeredCostmap -- the device-resident counterpart of costmap_2d::LayeredCostmap for stacks that are made of the
// layers libnavgpu implements (static-style grid layers, obstacle layers, inflation): one updateMap() runs the whole
// cycle of costmap_2d/src/layered_costmap.cpp:79-150 on the GPU (bounds of every layer, resetMap, merges in plugin
// order, ray-trace clearing + marking + footprint clearing, inflation) without the master grid leaving HBM, and
// getCostmap() brings an ordinary costmap_2d::Costmap2D for host-side consumers up to date: only the tiles whose cells
// changed since the last call cross PCIe (navgpu_costmap_get_changed), and the host grid is byte-identical to the
// device grid afterwards whatever happened in between (several updateMap calls, rolled origins).
// Header-only RAII over the navgpu_costmap_* C ABI (include/navgpu.h); method names follow LayeredCostmap /
// Costmap2DROS.  Errors: methods return false and keep navgpu_last_error(); there is no CPU fallback.
#ifndef NAVGPU_PLUGINS_GPU_LAYERED_COSTMAP_H_
#define NAVGPU_PLUGINS_GPU_LAYERED_COSTMAP_H_

#include <costmap_2d/costmap_2d.h>
#include <costmap_2d/observation.h>
#include <geometry_msgs/Point.h>

#include <cstring>
#include <vector>

#include "navgpu.h"

namespace navgpu_plugins {

class GpuLayeredCostmap {
 public:
  // LayeredCostmap(global_frame, rolling_window, track_unknown) + resizeMap (layered_costmap.cpp:50-77)
  GpuLayeredCostmap(unsigned int size_x, unsigned int size_y, double resolution, double origin_x, double origin_y,
                    bool rolling_window, bool track_unknown, int device = 0)
      : handle_(NULL), host_(size_x, size_y, resolution, origin_x, origin_y, track_unknown ? 255 : 0) {
    status_ = navgpu_costmap_create(&handle_, size_x, size_y, resolution, origin_x, origin_y, rolling_window,
                                    track_unknown, device);
    bounds_[0] = bounds_[1] = bounds_[2] = bounds_[3] = 0;
    // the host copy is page-locked so that window downloads run at PCIe speed (an optimisation: failure is harmless)
    host_pinned_ = handle_ != NULL && navgpu_host_register(host_.getCharMap(), size_t(size_x) * size_y) == NAVGPU_OK;
  }
  ~GpuLayeredCostmap() {
    if (host_pinned_) navgpu_host_unregister(host_.getCharMap());
    if (handle_) navgpu_costmap_destroy(handle_);
  }
  GpuLayeredCostmap(const GpuLayeredCostmap&) = delete;
  GpuLayeredCostmap& operator=(const GpuLayeredCostmap&) = delete;

  bool ok() const { return handle_ != NULL && status_ == NAVGPU_OK; }
  int lastStatus() const { return status_; }

  // plugins are appended in costmap `plugins:` order (costmap_2d_ros.cpp:115-128); the returned id names the layer
  int addStaticLayer(bool use_maximum = false) {  // StaticLayer::updateCosts policy (static_layer.cpp:287-299)
    int id = -1;
    status_ = navgpu_costmap_add_grid_layer(handle_, use_maximum ? NAVGPU_MAX : NAVGPU_TRUE_OVERWRITE, &id);
    return id;
  }
  int addObstacleLayer(int combination_method = 1, bool footprint_clearing_enabled = true,
                       double max_obstacle_height = 2.0) {  // cfg/ObstaclePlugin.cfg:7-10
    int id = -1;
    status_ = navgpu_costmap_add_obstacle_layer(handle_, combination_method, footprint_clearing_enabled,
                                                max_obstacle_height, &id);
    return id;
  }
  // VoxelLayer with cfg/VoxelPlugin.cfg's parameters (what Costmap2DROS creates for `map_type: voxel`,
  // costmap_2d_ros.cpp:215-228); observations go through setObservations like for an obstacle layer
  int addVoxelLayer(int combination_method = 1, bool footprint_clearing_enabled = true, double max_obstacle_height = 2.0,
                    double origin_z = 0.0, double z_resolution = 0.2, int z_voxels = 10, int unknown_threshold = 15,
                    int mark_threshold = 0) {
    int id = -1;
    status_ = navgpu_costmap_add_voxel_layer(handle_, combination_method, footprint_clearing_enabled, max_obstacle_height,
                                             origin_z, z_resolution, z_voxels, unknown_threshold, mark_threshold, &id);
    return id;
  }
  int addInflationLayer(double inflation_radius = 0.55, double cost_scaling_factor = 10.0) {  // cfg/InflationPlugin.cfg
    int id = -1;
    status_ = navgpu_costmap_add_inflation_layer(handle_, inflation_radius, cost_scaling_factor, &id);
    return id;
  }

  // LayeredCostmap::setFootprint (layered_costmap.cpp:163-173)
  bool setFootprint(const std::vector<geometry_msgs::Point>& footprint_spec) {
    std::vector<double> xy(2 * footprint_spec.size());
    for (size_t i = 0; i < footprint_spec.size(); ++i) {
      xy[2 * i] = footprint_spec[i].x;
      xy[2 * i + 1] = footprint_spec[i].y;
    }
    return check(navgpu_costmap_set_footprint(handle_, xy.data(), (int)footprint_spec.size()));
  }
  // StaticLayer::incomingMap (static_layer.cpp:165-223): nav_msgs/OccupancyGrid data, interpreted on the device
  bool setStaticMap(int layer, const signed char* occupancy, bool track_unknown_space = true,
                    unsigned char unknown_cost_value = 255, unsigned char lethal_threshold = 100, bool trinary = true) {
    return check(navgpu_grid_layer_set_occupancy(handle_, layer, occupancy, track_unknown_space, unknown_cost_value,
                                                 lethal_threshold, trinary));
  }
  bool setLayerCosts(int layer, const unsigned char* costs) { return check(navgpu_grid_layer_set(handle_, layer, costs)); }
  // StaticLayer's has_updated_data_ without new data (static_layer.cpp:263-285): the next updateMap covers the layer's
  // whole extent again
  bool markLayerUpdated(int layer) {
    return check(navgpu_grid_layer_touch(handle_, layer, 0, 0, host_.getSizeInCellsX(), host_.getSizeInCellsY()));
  }
  // ObstacleLayer's marking / clearing observations for the coming cycles (obstacle_layer.cpp:340-365, 450-464);
  // cloud points are float32 x, y, z exactly as pcl::PointXYZ stores them
  bool setObservations(int layer, const std::vector<costmap_2d::Observation>& marking_and_clearing) {
    std::vector<navgpu_observation> obs(marking_and_clearing.size());
    std::vector<std::vector<float> > xyz(marking_and_clearing.size());
    for (size_t i = 0; i < marking_and_clearing.size(); ++i) {
      const costmap_2d::Observation& o = marking_and_clearing[i];
      xyz[i].resize(3 * o.cloud_->points.size());
      for (size_t p = 0; p < o.cloud_->points.size(); ++p) {
        xyz[i][3 * p] = o.cloud_->points[p].x;
        xyz[i][3 * p + 1] = o.cloud_->points[p].y;
        xyz[i][3 * p + 2] = o.cloud_->points[p].z;
      }
      obs[i].origin_x = o.origin_.x; obs[i].origin_y = o.origin_.y; obs[i].origin_z = o.origin_.z;
      obs[i].obstacle_range = o.obstacle_range_;
      obs[i].raytrace_range = o.raytrace_range_;
      obs[i].xyz = xyz[i].data();
      obs[i].n_points = (int)o.cloud_->points.size();
      obs[i].marking = 1;
      obs[i].clearing = 1;
      obs[i].pad_ = 0;
    }
    return check(navgpu_obstacle_set_observations(handle_, layer, obs.data(), (int)obs.size()));
  }
  // On-device observation ingest: hand over the LaserScans themselves instead of projected clouds.  One entry per
  // buffered scan; the caller fills navgpu_laser_scan from sensor_msgs::LaserScan (ranges.data(), angle_min, ...),
  // the tf transform global_frame <- scan frame it would have given bufferCloud, and the ObservationBuffer's
  // min / max_obstacle_height, obstacle_range, raytrace_range.  Replaces ObstacleLayer::laserScanCallback /
  // laserScanValidInfCallback (obstacle_layer.cpp:252-311) + ObservationBuffer::bufferCloud (observation_buffer.cpp:129-195).
  bool setLaserScans(int layer, const std::vector<navgpu_laser_scan>& scans) {
    return check(navgpu_obstacle_set_scans(handle_, layer, scans.data(), (int)scans.size()));
  }
  bool setInflationParameters(int layer, double inflation_radius, double cost_scaling_factor) {
    return check(navgpu_inflation_set_params(handle_, layer, inflation_radius, cost_scaling_factor));
  }
  bool setEnabled(int layer, bool enabled) { return check(navgpu_layer_set_enabled(handle_, layer, enabled)); }

  // LayeredCostmap::updateMap (layered_costmap.cpp:79-150)
  bool updateMap(double robot_x, double robot_y, double robot_yaw) {
    return check(navgpu_costmap_update_map(handle_, robot_x, robot_y, robot_yaw, bounds_));
  }
  // the same without waiting for the device: the cycle is complete (and getBounds valid) after the next getCostmap()
  bool updateMapAsync(double robot_x, double robot_y, double robot_yaw) {
    return check(navgpu_costmap_update_map_async(handle_, robot_x, robot_y, robot_yaw));
  }
  // LayeredCostmap::getBounds (layered_costmap.h:129-135)
  void getBounds(unsigned int* x0, unsigned int* xn, unsigned int* y0, unsigned int* yn) const {
    *x0 = bounds_[0]; *xn = bounds_[1]; *y0 = bounds_[2]; *yn = bounds_[3];
  }
  // LayeredCostmap::getCostmap: brings the host copy up to date with the device grid and returns it.  changed_rects
  // (optional) receives x0, y0, xn, yn of every 128 x 16 tile that was rewritten (empty + *whole_grid = true when the
  // whole grid was copied), e.g. for Costmap2DPublisher's update messages.
  costmap_2d::Costmap2D* getCostmap(std::vector<int>* changed_rects = NULL, bool* whole_grid = NULL) {
    double origin[2];
    navgpu_costmap_get_origin(handle_, origin);
    // the device grid rolled (Costmap2D::updateOrigin, costmap_2d.cpp:264-313): the host copy only takes the new
    // origin; the cells that moved are among the tiles the download below rewrites
    if (origin[0] != host_.getOriginX() || origin[1] != host_.getOriginY()) host_.setOrigin(origin[0], origin[1]);
    int n = 0;
    int* rects = NULL;
    int capacity = 0;
    if (changed_rects) {
      changed_rects->resize(4 * kMaxRects);
      rects = changed_rects->data();
      capacity = kMaxRects;
    }
    if (!check(navgpu_costmap_get_changed(handle_, host_.getCharMap(), host_.getSizeInCellsX(), rects, capacity, &n, NULL)))
      return NULL;
    const bool whole = n > capacity || (n == 1 && rects && rects[2] - rects[0] == (int)host_.getSizeInCellsX() &&
                                        rects[3] - rects[1] == (int)host_.getSizeInCellsY());
    if (changed_rects) changed_rects->resize(whole ? 0 : 4 * size_t(n));
    if (whole_grid) *whole_grid = whole;
    navgpu_costmap_last_window(handle_, bounds_);
    return &host_;
  }
  navgpu_costmap* handle() { return handle_; }

 private:
  bool check(int rc) {
    status_ = rc;
    return rc == NAVGPU_OK;
  }
  // Costmap2D whose origin can follow the device grid without its own copy of updateOrigin's cell shuffle
  struct HostCostmap : public costmap_2d::Costmap2D {
    HostCostmap(unsigned int sx, unsigned int sy, double res, double ox, double oy, unsigned char def)
        : costmap_2d::Costmap2D(sx, sy, res, ox, oy, def) {}
    void setOrigin(double ox, double oy) { origin_x_ = ox; origin_y_ = oy; }
  };
  enum { kMaxRects = 2048 };
  navgpu_costmap* handle_;
  bool host_pinned_;
  HostCostmap host_;
  int bounds_[4];
  int status_;
};

}  // namespace navgpu_plugins
#endif
