// Drop-in check of the C++ adapters in navigation_b200/plugin against the REFERENCE'S OWN classes.
//
// Built only where /root/reference exists (oracle/Makefile target `dropin`): this file and the adapters are compiled
// against the reference's headers (plus the stub headers in oracle/shim for ROS/Boost/PCL) and linked with
// oracle/_ref/libnavref.so (the reference's unmodified hot-path sources) and navigation_b200/libnavgpu.so.
//
//   A1  costmap_2d::LayeredCostmap (reference) + test grid layer + reference costmap_2d::InflationLayer
//       versus the same LayeredCostmap + the same grid layer + navgpu_plugins::GpuInflationLayer as the plugin:
//       master grids must be identical over several cycles (full window, sub-window, parameter change).
//   A2  navgpu_plugins::GpuLayeredCostmap (whole stack on the device) versus the reference stack over several cycles,
//       the host Costmap2D kept in sync by changed tiles only (updateMapAsync + getCostmap, skipped downloads).
//   B   navgpu_plugins::GpuScoredSamplingPlanner as a base_local_planner::TrajectorySearch versus the reference's
//       generator + critics + SimpleScoredSamplingPlanner wired like DWAPlanner (through libnavref's C API).
//   C   navgpu_plugins::GpuTrajectoryPlanner versus the reference's base_local_planner::TrajectoryPlanner, both
//       constructed with the same arguments on the same Costmap2D and driven through the same public calls
//       (updatePlan, findBestPath with tf::Stamped<tf::Pose>, scoreTrajectory, checkTrajectory) for 8 cycles,
//       including cycles where the robot is boxed in and the planner rotates in place / backs up.
//   D   navgpu_plugins::GpuLayeredCostmap fed with LaserScans (setLaserScans: projection, transform and height filter
//       on the device) versus the checker's stack (the reference's LayeredCostmap + its own raytraceLine / MarkCell
//       behind libnavref's C API) fed with the clouds of the restated ingest path: master grids over 3 cycles.
//   E   navgpu_plugins::GpuTrajectoryCostFunction as THE critic of the reference's own SimpleScoredSamplingPlanner fed
//       by the reference's own SimpleTrajectoryGenerator, versus the same planner with the reference's six critics
//       wired like DWAPlanner: same winner, same cost; and the batched scoreTrajectories on every explored trajectory
//       versus the six reference critics' scaled sum.
// Prints one line per check and exits non-zero on any mismatch.  Test infrastructure, not product code.
#include <chrono>
#include <cmath>
#include <limits>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include <costmap_2d/cost_values.h>
#include <costmap_2d/costmap_layer.h>
#include <costmap_2d/inflation_layer.h>
#include <costmap_2d/layered_costmap.h>

#include <navgpu_plugins/gpu_inflation_layer.h>
#include <navgpu_plugins/gpu_layered_costmap.h>
#include <navgpu_plugins/gpu_scored_sampling_planner.h>
#include <navgpu_plugins/gpu_trajectory_cost_function.h>
#include <navgpu_plugins/gpu_trajectory_planner.h>

#include <base_local_planner/costmap_model.h>
#include <base_local_planner/local_planner_limits.h>
#include <base_local_planner/map_grid_cost_function.h>
#include <base_local_planner/obstacle_cost_function.h>
#include <base_local_planner/oscillation_cost_function.h>
#include <base_local_planner/simple_scored_sampling_planner.h>
#include <base_local_planner/simple_trajectory_generator.h>
#include <base_local_planner/trajectory_planner.h>

#include "oracle_api.h"

namespace {

using costmap_2d::Costmap2D;
using costmap_2d::LayeredCostmap;

// A CostmapLayer fed by the test: StaticLayer's non-rolling behaviour (static_layer.cpp:263-299) with TrueOverwrite
class TestGridLayer : public costmap_2d::CostmapLayer {
 public:
  TestGridLayer() : x_(0), y_(0), w_(0), h_(0), updated_(false) {}
  void onInitialize() override {
    current_ = true;
    enabled_ = true;
    default_value_ = costmap_2d::FREE_SPACE;
    matchSize();
  }
  void setData(const std::vector<unsigned char>& d) {
    memcpy(costmap_, d.data(), d.size());
    touchRegion(0, 0, size_x_, size_y_);
  }
  void setCells(unsigned x, unsigned y, unsigned w, unsigned h, unsigned char v) {
    for (unsigned j = y; j < y + h; ++j) memset(costmap_ + j * size_x_ + x, v, w);
    touchRegion(x, y, w, h);
  }
  void touchRegion(unsigned x, unsigned y, unsigned w, unsigned h) { x_ = x; y_ = y; w_ = w; h_ = h; updated_ = true; }
  void updateBounds(double, double, double, double* min_x, double* min_y, double* max_x, double* max_y) override {
    if (!updated_) return;
    double wx, wy;
    mapToWorld(x_, y_, wx, wy);
    *min_x = std::min(wx, *min_x);
    *min_y = std::min(wy, *min_y);
    mapToWorld(x_ + w_, y_ + h_, wx, wy);
    *max_x = std::max(wx, *max_x);
    *max_y = std::max(wy, *max_y);
    updated_ = false;
  }
  void updateCosts(Costmap2D& master, int min_i, int min_j, int max_i, int max_j) override {
    updateWithTrueOverwrite(master, min_i, min_j, max_i, max_j);
  }
  unsigned x_, y_, w_, h_;
  bool updated_;
};

std::vector<geometry_msgs::Point> squareFootprint(double half) {
  std::vector<geometry_msgs::Point> fp(4);
  fp[0].x = half; fp[0].y = half;
  fp[1].x = half; fp[1].y = -half;
  fp[2].x = -half; fp[2].y = -half;
  fp[3].x = -half; fp[3].y = half;
  return fp;
}

// thick axis-aligned blocks: the tie-free map class on which the reference's result does not depend on heap order
std::vector<unsigned char> blockMap(unsigned sx, unsigned sy, unsigned seed) {
  std::mt19937 rng(seed);
  std::vector<unsigned char> g(size_t(sx) * sy, 0);
  for (int k = 0; k < 40; ++k) {
    unsigned w = 3 + rng() % 40, h = 3 + rng() % 40, x = rng() % (sx - w), y = rng() % (sy - h);
    for (unsigned j = y; j < y + h; ++j) memset(&g[size_t(j) * sx + x], 254, w);
  }
  return g;
}

long countDiff(const unsigned char* a, const unsigned char* b, size_t n) {
  long d = 0;
  for (size_t i = 0; i < n; ++i) d += a[i] != b[i];
  return d;
}

int failures = 0;
void report(const char* name, long diff, long total) {
  printf("%-64s %s (%ld of %ld cells differ)\n", name, diff == 0 ? "OK" : "MISMATCH", diff, total);
  if (diff != 0) ++failures;
}

struct Stack {
  LayeredCostmap lc;
  TestGridLayer* grid;
  costmap_2d::Layer* inflation;
  Stack(unsigned sx, unsigned sy, double res, costmap_2d::Layer* infl) : lc("map", false, false), grid(new TestGridLayer), inflation(infl) {
    lc.resizeMap(sx, sy, res, 0.0, 0.0);
    boost::shared_ptr<costmap_2d::Layer> g(grid), i(infl);
    lc.addPlugin(g);
    grid->initialize(&lc, "static", NULL);
    lc.addPlugin(i);
    infl->initialize(&lc, "inflation", NULL);
    lc.setFootprint(squareFootprint(0.325));
  }
};

void testInflationPlugin() {
  const unsigned sx = 400, sy = 300;
  const double res = 0.05;
  costmap_2d::InflationLayer* ref_layer = new costmap_2d::InflationLayer;
  navgpu_plugins::GpuInflationLayer* gpu_layer = new navgpu_plugins::GpuInflationLayer;
  Stack ref(sx, sy, res, ref_layer), gpu(sx, sy, res, gpu_layer);
  const std::vector<unsigned char> map = blockMap(sx, sy, 7);
  ref.grid->setData(map);
  gpu.grid->setData(map);
  const size_t n = size_t(sx) * sy;
  ref.lc.updateMap(5.0, 5.0, 0.0);
  gpu.lc.updateMap(5.0, 5.0, 0.0);
  report("A1 cycle 1: full window, radius 0.55", countDiff(ref.lc.getCostmap()->getCharMap(), gpu.lc.getCostmap()->getCharMap(), n), n);
  // cycle 2: a new block appears in a sub-window
  ref.grid->setCells(120, 80, 30, 12, 254);
  gpu.grid->setCells(120, 80, 30, 12, 254);
  ref.lc.updateMap(5.0, 5.0, 0.0);
  gpu.lc.updateMap(5.0, 5.0, 0.0);
  unsigned a0, a1, a2, a3, b0, b1, b2, b3;
  ref.lc.getBounds(&a0, &a1, &a2, &a3);
  gpu.lc.getBounds(&b0, &b1, &b2, &b3);
  report("A1 cycle 2: sub-window bounds", (a0 != b0) + (a1 != b1) + (a2 != b2) + (a3 != b3), 4);
  report("A1 cycle 2: sub-window update", countDiff(ref.lc.getCostmap()->getCharMap(), gpu.lc.getCostmap()->getCharMap(), n), n);
  // cycle 3: the block disappears again (cleared cells must lose their inflation inside the window)
  ref.grid->setCells(120, 80, 30, 12, 0);
  gpu.grid->setCells(120, 80, 30, 12, 0);
  ref.lc.updateMap(5.0, 5.0, 0.0);
  gpu.lc.updateMap(5.0, 5.0, 0.0);
  report("A1 cycle 3: block removed", countDiff(ref.lc.getCostmap()->getCharMap(), gpu.lc.getCostmap()->getCharMap(), n), n);
  // cycle 4: reconfigure to the C3 inflation (1.0 m => R = 20 cells), which forces a whole-map re-inflation
  ref_layer->setInflationParameters(1.0, 10.0);
  gpu_layer->setInflationParameters(1.0, 10.0);
  ref.lc.updateMap(5.0, 5.0, 0.0);
  gpu.lc.updateMap(5.0, 5.0, 0.0);
  report("A1 cycle 4: setInflationParameters(1.0, 10)", countDiff(ref.lc.getCostmap()->getCharMap(), gpu.lc.getCostmap()->getCharMap(), n), n);
  report("A1 computeCost agrees on 0..25 cells", [&] {
    long d = 0;
    for (int i = 0; i <= 25; ++i) d += ref_layer->computeCost(i) != gpu_layer->computeCost(i);
    return d;
  }(), 26);
}

void testFusedStack() {
  const unsigned sx = 400, sy = 300;
  const double res = 0.05;
  Stack ref(sx, sy, res, new costmap_2d::InflationLayer);
  navgpu_plugins::GpuLayeredCostmap gpu(sx, sy, res, 0.0, 0.0, false, false);
  const int s = gpu.addStaticLayer(false);
  gpu.addInflationLayer(0.55, 10.0);
  gpu.setFootprint(squareFootprint(0.325));
  const std::vector<unsigned char> map = blockMap(sx, sy, 11);
  ref.grid->setData(map);
  gpu.setLayerCosts(s, map.data());
  ref.lc.updateMap(5.0, 5.0, 0.0);
  const bool ok = gpu.ok() && gpu.updateMap(5.0, 5.0, 0.0);
  Costmap2D* out = ok ? gpu.getCostmap() : NULL;
  const size_t n = size_t(sx) * sy;
  if (!out) {
    printf("A2 GpuLayeredCostmap failed: %s\n", navgpu_last_error());
    ++failures;
    return;
  }
  report("A2 GpuLayeredCostmap vs reference LayeredCostmap", countDiff(ref.lc.getCostmap()->getCharMap(), out->getCharMap(), n), n);
  // further cycles: sub-window changes, the host copy refreshed by changed tiles only -- after every cycle, and after
  // two cycles without a download in between
  long bad = 0;
  std::vector<unsigned char> map2 = map;
  for (int c = 0; c < 5; ++c) {
    const unsigned x = 30 + 60 * c, y = 40 + 45 * c;
    for (unsigned j = y; j < y + 9; ++j) memset(&map2[size_t(j) * sx + x], c % 2 ? 0 : 254, 14);
    ref.grid->setData(map2);
    gpu.setLayerCosts(s, map2.data());
    ref.lc.updateMap(5.0, 5.0, 0.0);
    if (!gpu.updateMapAsync(5.0, 5.0, 0.0)) ++bad;
    if (c == 2) continue;  // no getCostmap this cycle: the next one must still bring the host copy up to date
    std::vector<int> rects;
    bool whole = false;
    out = gpu.getCostmap(&rects, &whole);
    if (!out) { ++bad; continue; }
    bad += countDiff(ref.lc.getCostmap()->getCharMap(), out->getCharMap(), n);
    unsigned x0, xn, y0, yn;
    gpu.getBounds(&x0, &xn, &y0, &yn);
    if (x0 != 0 || xn != sx || y0 != 0 || yn != sy) ++bad;  // setData touches the whole layer
    if (whole || rects.size() / 4 > 24) ++bad;               // a 14 x 9 block and its inflation: a handful of tiles
  }
  report("A2 five more cycles, host Costmap2D refreshed by changed tiles", bad, 5 * (long)n);
}

void testScoredSamplingPlanner() {
  // local costmap: corridor + box, inflated by the reference stack
  const unsigned n = 120;
  const double res = 0.05;
  Stack ref(n, n, res, new costmap_2d::InflationLayer);
  std::vector<unsigned char> g(size_t(n) * n, 0);
  for (unsigned y = 26; y < 29; ++y) memset(&g[y * n], 254, n);
  for (unsigned y = 91; y < 94; ++y) memset(&g[y * n], 254, n);
  for (unsigned y = 52; y < 58; ++y) memset(&g[y * n + 84], 254, 6);
  ref.grid->setData(g);
  ref.lc.updateMap(3.0, 3.0, 0.0);
  Costmap2D* local = ref.lc.getCostmap();

  navo_dwa_config rc;
  navo_dwa_default_config(&rc);
  rc.vx_samples = 20; rc.vy_samples = 1; rc.vth_samples = 20; rc.max_vel_y = 0.0; rc.min_vel_y = 0.0;
  navgpu_dwa_config gc = navgpu_plugins::GpuScoredSamplingPlanner::defaultConfig();
  gc.vx_samples = 20; gc.vy_samples = 1; gc.vth_samples = 20; gc.max_vel_y = 0.0; gc.min_vel_y = 0.0;

  void* refp = navo_dwa_create(&rc, n, n, res);
  navo_dwa_set_costmap(refp, local->getCharMap(), 0.0, 0.0);
  navgpu_plugins::GpuScoredSamplingPlanner gpu(gc, local);
  base_local_planner::TrajectorySearch* search = &gpu;  // used through the reference's interface

  std::vector<double> plan_xy;
  std::vector<geometry_msgs::PoseStamped> plan;
  for (int i = 0; i < 120; ++i) {
    geometry_msgs::PoseStamped p;
    p.pose.position.x = 1.0 + 0.05 * i;
    p.pose.position.y = 3.0;
    plan.push_back(p);
    plan_xy.push_back(p.pose.position.x);
    plan_xy.push_back(p.pose.position.y);
  }
  const std::vector<geometry_msgs::Point> fp = squareFootprint(0.3);
  std::vector<double> fp_xy;
  for (size_t i = 0; i < fp.size(); ++i) { fp_xy.push_back(fp[i].x); fp_xy.push_back(fp[i].y); }

  double pose[3] = {1.5, 3.0, 0.0}, vel[3] = {0.3, 0.0, 0.0};
  long bad = 0, cycles = 0;
  for (int c = 0; c < 6; ++c, ++cycles) {
    navo_dwa_set_plan(refp, pose, plan_xy.data(), 120);
    gpu.setPlan(pose[0], pose[1], pose[2], plan);
    navo_dwa_result rr;
    std::vector<double> rcosts(1 << 12), rpts(3 * 4096);
    navo_dwa_find_best_path(refp, pose, vel, fp_xy.data(), 4, &rr, rcosts.data(), (int)rcosts.size(), rpts.data(), 4096);
    gpu.setState(pose[0], pose[1], pose[2], vel[0], vel[1], vel[2], fp);
    base_local_planner::Trajectory traj;
    std::vector<base_local_planner::Trajectory> explored;
    const bool found = search->findBestTrajectory(traj, &explored);
    const bool same = found == (rr.cost >= 0) && gpu.bestIndex() == rr.best_index &&
                      std::fabs(traj.cost_ - rr.cost) <= 1e-5 * std::fabs(rr.cost) && traj.xv_ == rr.xv &&
                      traj.thetav_ == rr.thetav && (int)traj.getPointsSize() == rr.n_points &&
                      (int)explored.size() == rr.n_scored && gpu.oscillationMask() == navo_dwa_get_oscillation_mask(refp);
    if (!same) {
      ++bad;
      printf("  cycle %d: reference best %d cost %.9g (%d pts, %d scored) / gpu best %d cost %.9g (%u pts, %zu explored)\n", c,
             rr.best_index, rr.cost, rr.n_points, rr.n_scored, gpu.bestIndex(), traj.cost_, traj.getPointsSize(), explored.size());
    }
    if (rr.cost >= 0) {  // follow the chosen command for one control period
      double x, y, th;
      traj.getPoint(std::min(2u, traj.getPointsSize() - 1), x, y, th);
      pose[0] = x; pose[1] = y; pose[2] = th;
      vel[0] = rr.xv; vel[1] = rr.yv; vel[2] = rr.thetav;
    }
  }
  report("B  GpuScoredSamplingPlanner vs reference DWA search, 6 cycles", bad, cycles);
  // DWAPlanner::checkTrajectory on cycles without a search (LatchedStopRotateController, dwa_planner_ros.cpp:271-288):
  // the costmap changed since the last findBestTrajectory -- a wall appears right in front of the robot
  {
    long bad2 = 0;
    const double samp[3] = {0.5, 0.0, 0.0};
    double rc0 = 0, gc0 = 0, rc1 = 0, gc1 = 0;
    rc0 = navo_dwa_check_trajectory(refp, pose, vel, samp, fp_xy.data(), 4);
    const bool g0 = gpu.checkTrajectory(pose[0], pose[1], pose[2], vel[0], vel[1], vel[2], samp[0], samp[1], samp[2], &gc0);
    const unsigned wx = (unsigned)((pose[0] + 0.45) / res), wy = (unsigned)(pose[1] / res);
    for (unsigned y = wy - 8; y < wy + 8; ++y) memset(local->getCharMap() + y * n + wx, 254, 3);
    navo_dwa_set_costmap(refp, local->getCharMap(), 0.0, 0.0);
    rc1 = navo_dwa_check_trajectory(refp, pose, vel, samp, fp_xy.data(), 4);
    const bool g1 = gpu.checkTrajectory(pose[0], pose[1], pose[2], vel[0], vel[1], vel[2], samp[0], samp[1], samp[2], &gc1);
    if (g0 != (rc0 >= 0) || g1 != (rc1 >= 0) || std::fabs(gc0 - rc0) > 1e-5 * std::fabs(rc0) || gc1 != rc1 || !(rc0 >= 0) || rc1 >= 0) {
      ++bad2;
      printf("  checkTrajectory: reference %.9g -> %.9g, gpu %.9g -> %.9g\n", rc0, rc1, gc0, gc1);
    }
    report("B  checkTrajectory sees the costmap as it is now", bad2, 2);
  }
  navo_dwa_destroy(refp);
}

// The reference's DWAPlanner wiring (dwa_planner.cpp:116-182, 52-112) with its own classes
struct RefDwa {
  base_local_planner::LocalPlannerLimits limits;
  base_local_planner::ObstacleCostFunction obstacle;
  base_local_planner::MapGridCostFunction path, goal, goal_front, alignment;
  base_local_planner::OscillationCostFunction oscillation;
  base_local_planner::SimpleTrajectoryGenerator generator;
  std::vector<base_local_planner::TrajectoryCostFunction*> critics;
  RefDwa(Costmap2D* cm, const navgpu_dwa_config& c)
      : obstacle(cm), path(cm), goal(cm, 0.0, 0.0, true), goal_front(cm, 0.0, 0.0, true), alignment(cm) {
    goal_front.setStopOnFailure(false);
    alignment.setStopOnFailure(false);
    limits.max_trans_vel = c.max_trans_vel; limits.min_trans_vel = c.min_trans_vel;
    limits.max_vel_x = c.max_vel_x; limits.min_vel_x = c.min_vel_x;
    limits.max_vel_y = c.max_vel_y; limits.min_vel_y = c.min_vel_y;
    limits.max_rot_vel = c.max_rot_vel; limits.min_rot_vel = c.min_rot_vel;
    limits.acc_lim_x = c.acc_lim_x; limits.acc_lim_y = c.acc_lim_y; limits.acc_lim_theta = c.acc_lim_theta;
    generator.setParameters(c.sim_time, c.sim_granularity, c.angular_sim_granularity, c.use_dwa != 0, c.sim_period);
    const double res = cm->getResolution();
    path.setScale(res * c.path_distance_bias * 0.5);
    alignment.setScale(res * c.path_distance_bias * 0.5);
    goal.setScale(res * c.goal_distance_bias * 0.5);
    goal_front.setScale(res * c.goal_distance_bias * 0.5);
    obstacle.setScale(res * c.occdist_scale);
    oscillation.setOscillationResetDist(c.oscillation_reset_dist, c.oscillation_reset_angle);
    oscillation.resetOscillationFlags();
    goal_front.setXShift(c.forward_point_distance);
    alignment.setXShift(c.forward_point_distance);
    obstacle.setParams(c.max_trans_vel, c.max_scaling_factor, c.scaling_speed);
    obstacle.setSumScores(c.sum_scores != 0);
    critics.push_back(&oscillation);
    critics.push_back(&obstacle);
    critics.push_back(&goal_front);
    critics.push_back(&alignment);
    critics.push_back(&path);
    critics.push_back(&goal);
  }
  // DWAPlanner::updatePlanAndLocalCosts (:240-286)
  void setPlan(const double pose[3], const std::vector<geometry_msgs::PoseStamped>& plan, double fpd, double cheat, double res,
               double pdist) {
    path.setTargetPoses(plan);
    goal.setTargetPoses(plan);
    const geometry_msgs::PoseStamped g = plan.back();
    Eigen::Vector3f pos(pose[0], pose[1], pose[2]);
    const double sq_dist = (pos[0] - g.pose.position.x) * (pos[0] - g.pose.position.x) +
                           (pos[1] - g.pose.position.y) * (pos[1] - g.pose.position.y);
    std::vector<geometry_msgs::PoseStamped> front = plan;
    const double a = atan2(g.pose.position.y - pos[1], g.pose.position.x - pos[0]);
    front.back().pose.position.x = front.back().pose.position.x + fpd * cos(a);
    front.back().pose.position.y = front.back().pose.position.y + fpd * sin(a);
    goal_front.setTargetPoses(front);
    if (sq_dist > fpd * fpd * cheat) {
      alignment.setScale(res * pdist * 0.5);
      alignment.setTargetPoses(plan);
    } else {
      alignment.setScale(0.0);
    }
  }
  // SimpleScoredSamplingPlanner::scoreTrajectory (:50-79) without the early exit
  double fullScore(base_local_planner::Trajectory& t) {
    double total = 0.0;
    for (size_t k = 0; k < critics.size(); ++k) {
      if (critics[k]->getScale() == 0) continue;
      double c = critics[k]->scoreTrajectory(t);
      if (c < 0) return c;
      if (c != 0) c *= critics[k]->getScale();
      total += c;
    }
    return total;
  }
};

void testTrajectoryCostFunction() {
  const unsigned n = 120;
  const double res = 0.05;
  Stack ref(n, n, res, new costmap_2d::InflationLayer);
  std::vector<unsigned char> g(size_t(n) * n, 0);
  for (unsigned y = 26; y < 29; ++y) memset(&g[y * n], 254, n);
  for (unsigned y = 91; y < 94; ++y) memset(&g[y * n], 254, n);
  for (unsigned y = 52; y < 58; ++y) memset(&g[y * n + 84], 254, 6);
  ref.grid->setData(g);
  ref.lc.updateMap(3.0, 3.0, 0.0);
  Costmap2D* local = ref.lc.getCostmap();

  navgpu_dwa_config gc = navgpu_plugins::GpuScoredSamplingPlanner::defaultConfig();
  gc.vx_samples = 12; gc.vy_samples = 3; gc.vth_samples = 15;
  RefDwa rd(local, gc);
  navgpu_plugins::GpuTrajectoryCostFunction gpu_costs(gc, local);
  base_local_planner::TrajectoryCostFunction* as_critic = &gpu_costs;  // used through the reference's interface

  std::vector<base_local_planner::TrajectorySampleGenerator*> gens(1, &rd.generator);
  base_local_planner::SimpleScoredSamplingPlanner ref_planner(gens, rd.critics);
  std::vector<base_local_planner::TrajectoryCostFunction*> one(1, as_critic);
  base_local_planner::SimpleScoredSamplingPlanner gpu_planner(gens, one);

  std::vector<geometry_msgs::PoseStamped> plan;
  for (int i = 0; i < 120; ++i) {
    geometry_msgs::PoseStamped p;
    p.pose.position.x = 1.0 + 0.05 * i;
    p.pose.position.y = 3.0;
    plan.push_back(p);
  }
  const std::vector<geometry_msgs::Point> fp = squareFootprint(0.3);
  rd.obstacle.setFootprint(fp);
  gpu_costs.setFootprint(fp);
  Eigen::Vector3f vsamples(gc.vx_samples, gc.vy_samples, gc.vth_samples);

  double pose[3] = {1.5, 3.0, 0.0}, vel[3] = {0.3, 0.0, 0.0};
  long bad = 0, cycles = 0, bad_batch = 0, scored = 0;
  for (int c = 0; c < 6; ++c, ++cycles) {
    rd.setPlan(pose, plan, gc.forward_point_distance, gc.cheat_factor, res, gc.path_distance_bias);
    gpu_costs.setPlan(pose[0], pose[1], pose[2], plan);
    Eigen::Vector3f pos(pose[0], pose[1], pose[2]), v(vel[0], vel[1], vel[2]);
    Eigen::Vector3f goal(plan.back().pose.position.x, plan.back().pose.position.y, 0.0f);
    base_local_planner::Trajectory rt, gt;
    std::vector<base_local_planner::Trajectory> explored;
    rd.generator.initialise(pos, v, goal, &rd.limits, vsamples);
    rt.cost_ = -7;
    ref_planner.findBestTrajectory(rt, &explored);
    rd.generator.initialise(pos, v, goal, &rd.limits, vsamples);
    gt.cost_ = -7;
    gpu_planner.findBestTrajectory(gt, NULL);
    bool same = rt.cost_ == gt.cost_ && rt.xv_ == gt.xv_ && rt.yv_ == gt.yv_ && rt.thetav_ == gt.thetav_ &&
                rt.getPointsSize() == gt.getPointsSize();
    // the batched call on everything the reference explored, against the six critics' full scaled sum
    std::vector<double> costs;
    if (!gpu_costs.scoreTrajectories(explored, &costs)) same = false;
    for (size_t i = 0; i < explored.size() && i < costs.size(); ++i, ++scored) {
      const double want = rd.fullScore(explored[i]);
      if (!(costs[i] == want || std::fabs(costs[i] - want) <= 1e-5 * std::fabs(want))) ++bad_batch;
    }
    rd.oscillation.updateOscillationFlags(pos, &rt, gc.min_trans_vel);
    gpu_costs.updateOscillationFlags(pose[0], pose[1], pose[2], &gt);
    if (!same) {
      ++bad;
      printf("  cycle %d: reference cost %.9g v (%.6g %.6g %.6g) %u pts / gpu critic cost %.9g v (%.6g %.6g %.6g) %u pts\n", c,
             rt.cost_, rt.xv_, rt.yv_, rt.thetav_, rt.getPointsSize(), gt.cost_, gt.xv_, gt.yv_, gt.thetav_, gt.getPointsSize());
    }
    if (rt.cost_ >= 0) {  // alternate the direction to exercise the oscillation flags
      double x, y, th;
      rt.getPoint(std::min(2u, rt.getPointsSize() - 1), x, y, th);
      pose[0] = x; pose[1] = y; pose[2] = th;
      vel[0] = (c % 2 ? 1 : -1) * rt.xv_ * 0.2; vel[1] = rt.yv_; vel[2] = -rt.thetav_;
    }
  }
  report("E  SimpleScoredSamplingPlanner + GpuTrajectoryCostFunction vs + 6 reference critics", bad, cycles);
  report("E  batched scoreTrajectories vs the reference critics' scaled sum", bad_batch, scored);
}

void testTrajectoryPlanner() {
  const unsigned n = 120;
  const double res = 0.05;
  Stack ref(n, n, res, new costmap_2d::InflationLayer);
  std::vector<unsigned char> g(size_t(n) * n, 0);
  for (unsigned y = 26; y < 29; ++y) memset(&g[y * n], 254, n);
  for (unsigned y = 91; y < 94; ++y) memset(&g[y * n], 254, n);
  for (unsigned y = 44; y < 76; ++y) memset(&g[y * n + 70], 254, 4);  // a barrier across most of the corridor
  ref.grid->setData(g);
  ref.lc.updateMap(3.0, 3.0, 0.0);
  Costmap2D* local = ref.lc.getCostmap();
  const std::vector<geometry_msgs::Point> fp = squareFootprint(0.3);
  const std::vector<double> y_vels = {-0.3, -0.1, 0.1, 0.3};
  base_local_planner::CostmapModel world_model(*local);
  // TrajectoryPlannerROS's defaults (trajectory_planner_ros.cpp:116-213)
  base_local_planner::TrajectoryPlanner rtp(world_model, *local, fp, 2.5, 2.5, 3.2, 1.0, 0.025, 6, 20, 0.6, 0.8, 0.01, 0.325,
                                            0.05, 0.10, M_PI_4, true, 0.5, 0.1, 1.0, -1.0, 0.4, -0.1, true, false, 0.8,
                                            false, false, y_vels, 0.2, 0.05, 0.025);
  navgpu_plugins::GpuTrajectoryPlanner gtp(world_model, *local, fp, 2.5, 2.5, 3.2, 1.0, 0.025, 6, 20, 0.6, 0.8, 0.01, 0.325,
                                           0.05, 0.10, M_PI_4, true, 0.5, 0.1, 1.0, -1.0, 0.4, -0.1, true, false, 0.8,
                                           false, false, y_vels, 0.2, 0.05, 0.025);
  std::vector<geometry_msgs::PoseStamped> plan;
  for (int i = 0; i < 98; ++i) {
    geometry_msgs::PoseStamped p;
    p.pose.position.x = 1.0 + 0.05 * i;
    p.pose.position.y = 3.0;
    plan.push_back(p);
  }
  rtp.updatePlan(plan, false);
  gtp.updatePlan(plan, false);
  double pose[3] = {1.5, 3.0, 0.0}, vel[3] = {0.3, 0.0, 0.0};
  long bad = 0, cycles = 0;
  for (int c = 0; c < 8; ++c, ++cycles) {
    if (c == 4) { pose[0] = 3.15; pose[1] = 3.0; pose[2] = 0.1; vel[0] = 0.0; }  // nose against the barrier
    tf::Stamped<tf::Pose> gp, gv, rd, gd;
    gp.setOrigin(tf::Vector3(pose[0], pose[1], 0.0));
    gv.setOrigin(tf::Vector3(vel[0], vel[1], 0.0));
    tf::Matrix3x3 m;
    m.setRotation(tf::createQuaternionFromYaw(pose[2]));
    gp.setBasis(m);
    m.setRotation(tf::createQuaternionFromYaw(vel[2]));
    gv.setBasis(m);
    base_local_planner::Trajectory rt = rtp.findBestPath(gp, gv, rd);
    base_local_planner::Trajectory gt = gtp.findBestPath(gp, gv, gd);
    const double rs = rtp.scoreTrajectory(pose[0], pose[1], pose[2], vel[0], vel[1], vel[2], 0.3, 0.0, 0.2);
    const double gs = gtp.scoreTrajectory(pose[0], pose[1], pose[2], vel[0], vel[1], vel[2], 0.3, 0.0, 0.2);
    const bool rc = rtp.checkTrajectory(pose[0], pose[1], pose[2], vel[0], vel[1], vel[2], 0.5, 0.0, -1.0);
    const bool gc = gtp.checkTrajectory(pose[0], pose[1], pose[2], vel[0], vel[1], vel[2], 0.5, 0.0, -1.0);
    const bool same = std::fabs(rt.cost_ - gt.cost_) <= 1e-5 * std::fabs(rt.cost_) && rt.xv_ == gt.xv_ && rt.yv_ == gt.yv_ &&
                      rt.thetav_ == gt.thetav_ && rt.getPointsSize() == gt.getPointsSize() &&
                      std::fabs(rs - gs) <= 1e-5 * std::fabs(rs) && rc == gc &&
                      rd.getOrigin().getX() == gd.getOrigin().getX() && tf::getYaw(rd.getRotation()) == tf::getYaw(gd.getRotation());
    if (!same) {
      ++bad;
      printf("  cycle %d: reference cost %.9g v (%.6g %.6g %.6g) %u pts score %.9g / gpu cost %.9g v (%.6g %.6g %.6g) %u pts score %.9g\n",
             c, rt.cost_, rt.xv_, rt.yv_, rt.thetav_, rt.getPointsSize(), rs, gt.cost_, gt.xv_, gt.yv_, gt.thetav_,
             gt.getPointsSize(), gs);
    }
    if (rt.cost_ >= 0 && rt.getPointsSize() > 0) {  // follow the chosen command for a few steps
      double x, y, th;
      rt.getPoint(std::min(4u, rt.getPointsSize() - 1), x, y, th);
      pose[0] = x; pose[1] = y; pose[2] = th;
      vel[0] = rt.xv_; vel[1] = rt.yv_; vel[2] = rt.thetav_;
    }
  }
  report("C  GpuTrajectoryPlanner vs reference TrajectoryPlanner, 8 cycles", bad, cycles);
}

void testLaserScanIngest() {
  const unsigned sx = 240, sy = 200;
  const double res = 0.05;
  const std::vector<unsigned char> map = blockMap(sx, sy, 3);
  navgpu_plugins::GpuLayeredCostmap gpu(sx, sy, res, 0.0, 0.0, false, false);
  const int gs = gpu.addStaticLayer(false);
  const int go = gpu.addObstacleLayer(1, true, 2.0);
  gpu.addInflationLayer(0.55, 10.0);
  gpu.setFootprint(squareFootprint(0.325));
  gpu.setLayerCosts(gs, map.data());
  void* ref = navo_costmap_create(sx, sy, res, 0.0, 0.0, 0, 0);
  const int rs = navo_costmap_add_grid_layer(ref, NAVO_TRUE_OVERWRITE);
  const int ro = navo_costmap_add_obstacle_layer(ref, 1, 1, 2.0);
  navo_costmap_add_inflation_layer(ref, 0.55, 10.0);
  const double fp[8] = {0.325, 0.325, 0.325, -0.325, -0.325, -0.325, -0.325, 0.325};
  navo_costmap_set_footprint(ref, fp, 4);
  navo_grid_layer_set(ref, rs, map.data());
  std::mt19937 rng(5);
  std::uniform_real_distribution<float> range(0.05f, 6.0f);
  long bad = 0, cells = 0;
  for (int c = 0; c < 3; ++c) {
    const int n = 360;
    std::vector<float> ranges(n);
    for (int i = 0; i < n; ++i) ranges[i] = range(rng);
    ranges[7] = std::numeric_limits<float>::infinity();
    ranges[11] = std::numeric_limits<float>::quiet_NaN();
    const double yaw = 0.4 * c, px = 5.0 + 0.5 * c, py = 4.0 + 0.3 * c;
    navgpu_laser_scan gsn;
    memset(&gsn, 0, sizeof(gsn));
    gsn.ranges = ranges.data(); gsn.n_ranges = n; gsn.inf_is_valid = c & 1;
    gsn.angle_min = -3.14159f; gsn.angle_increment = 6.28318f / n; gsn.range_min = 0.1f; gsn.range_max = 5.5f;
    gsn.sensor_to_global_translation[0] = px; gsn.sensor_to_global_translation[1] = py; gsn.sensor_to_global_translation[2] = 0.3;
    gsn.sensor_to_global_rotation_xyzw[2] = sin(yaw / 2); gsn.sensor_to_global_rotation_xyzw[3] = cos(yaw / 2);
    gsn.min_obstacle_height = 0.0; gsn.max_obstacle_height = 2.0;
    gsn.obstacle_range = 2.5; gsn.raytrace_range = 3.0; gsn.marking = 1; gsn.clearing = 1;
    gpu.setLaserScans(go, std::vector<navgpu_laser_scan>(1, gsn));
    navo_laser_scan rsn;
    memset(&rsn, 0, sizeof(rsn));
    rsn.ranges = ranges.data(); rsn.n_ranges = n; rsn.inf_is_valid = gsn.inf_is_valid;
    rsn.angle_min = gsn.angle_min; rsn.angle_increment = gsn.angle_increment; rsn.range_min = gsn.range_min; rsn.range_max = gsn.range_max;
    for (int k = 0; k < 3; ++k) rsn.translation[k] = gsn.sensor_to_global_translation[k];
    for (int k = 0; k < 4; ++k) rsn.rotation_xyzw[k] = gsn.sensor_to_global_rotation_xyzw[k];
    rsn.min_obstacle_height = 0.0; rsn.max_obstacle_height = 2.0;
    std::vector<float> cloud(3 * n);
    double origin[3];
    const int np = navo_project_scan(&rsn, cloud.data(), n, origin);
    navo_observation ob;
    memset(&ob, 0, sizeof(ob));
    ob.origin_x = origin[0]; ob.origin_y = origin[1]; ob.origin_z = origin[2];
    ob.obstacle_range = 2.5; ob.raytrace_range = 3.0; ob.xyz = cloud.data(); ob.n_points = np; ob.marking = 1; ob.clearing = 1;
    navo_obstacle_set_observations(ref, ro, &ob, 1);
    int32_t w[4];
    navo_costmap_update_map(ref, px, py, yaw, w);
    std::vector<unsigned char> want(size_t(sx) * sy);
    navo_costmap_get(ref, want.data());
    const bool ok = gpu.ok() && gpu.updateMap(px, py, yaw);
    Costmap2D* out = ok ? gpu.getCostmap() : NULL;
    if (!out) {
      printf("D GpuLayeredCostmap::setLaserScans failed: %s\n", navgpu_last_error());
      ++failures;
      navo_costmap_destroy(ref);
      return;
    }
    // getCostmap() downloads the updated window; outside it both grids still hold the previous cycle's values
    bad += countDiff(want.data(), out->getCharMap(), want.size());
    cells += (long)want.size();
  }
  navo_costmap_destroy(ref);
  report("D  GpuLayeredCostmap fed with LaserScans vs the checker's stack, 3 cycles", bad, cells);
}

// informational: what a host-side C++ consumer of the adapter sees for a full-window update of a 4000 x 4000 map with
// an obstacle layer whose scans move every cycle (ray-cast against the static blocks): setObservations + markLayerUpdated
// + updateMapAsync + getCostmap (changed tiles only) per cycle
void timeFusedStackEndToEnd() {
  const unsigned n = 4000;
  const double res = 0.05;
  navgpu_plugins::GpuLayeredCostmap gpu(n, n, res, 0.0, 0.0, false, false);
  const int s = gpu.addStaticLayer(false);
  const int o = gpu.addObstacleLayer(1, true, 2.0);
  gpu.addInflationLayer(1.0, 10.0);
  gpu.setFootprint(squareFootprint(0.325));
  const std::vector<unsigned char> map = blockMap(n, n, 21);
  if (!gpu.ok()) return;
  gpu.setLayerCosts(s, map.data());
  // four scan sets of 8 x 360 beams each, cast from positions a few cells apart
  std::vector<std::vector<costmap_2d::Observation> > sets(4);
  std::vector<std::vector<pcl::PointCloud<pcl::PointXYZ> > > clouds(4, std::vector<pcl::PointCloud<pcl::PointXYZ> >(8));
  for (int k = 0; k < 4; ++k)
    for (int b = 0; b < 8; ++b) {
      const double ox = 100.0 + 0.35 * k + 0.6 * b, oy = 100.0;
      pcl::PointCloud<pcl::PointXYZ>& c = clouds[k][b];
      for (int i = 0; i < 360; ++i) {
        const double a = i * (2 * M_PI / 360), dx = cos(a), dy = sin(a);
        double t = 0.05;
        for (; t < 10.0; t += 0.025) {
          const int cx = (int)((ox + dx * t) / res), cy = (int)((oy + dy * t) / res);
          if (cx < 0 || cy < 0 || cx >= (int)n || cy >= (int)n || map[size_t(cy) * n + cx] == 254) break;
        }
        pcl::PointXYZ p;
        p.x = (float)(ox + dx * t); p.y = (float)(oy + dy * t); p.z = 0.3f;
        c.points.push_back(p);
      }
      geometry_msgs::Point origin;
      origin.x = ox; origin.y = oy; origin.z = 0.3;
      sets[k].push_back(costmap_2d::Observation(origin, c, 10.0, 10.0));
    }
  double best = 1e30, sum = 0;
  int counted = 0;
  size_t tiles = 0;
  for (int c = 0; c < 24; ++c) {
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<int> rects;
    bool whole = false;
    const bool ok = gpu.setObservations(o, sets[c % 4]) && gpu.markLayerUpdated(s) && gpu.updateMapAsync(100.0, 100.0, 0.0) &&
                    gpu.getCostmap(&rects, &whole) != NULL;
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (!ok) return;
    if (c >= 4) { best = std::min(best, ms); sum += ms; ++counted; tiles += rects.size() / 4; }
  }
  double idle = 1e30;
  for (int c = 0; c < 50; ++c) {  // nothing changed in between: launch + host wait of the diff alone
    const auto t0 = std::chrono::steady_clock::now();
    gpu.getCostmap();
    idle = std::min(idle, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
  }
  printf("   (getCostmap() with nothing to fetch: %.3f ms)\n", idle);
  printf("   (C++ end to end, 4000 x 4000 full window, 8 x 360 beams moving every cycle: setObservations + updateMapAsync + "
         "getCostmap: mean %.3f ms, best %.3f ms, %.0f changed tiles per cycle)\n", sum / counted, best, double(tiles) / counted);
}

}  // namespace

int main() {
  if (navgpu_device_count() == 0) {
    printf("no CUDA device: the adapters have no CPU fallback\n");
    return 2;
  }
  testInflationPlugin();
  testFusedStack();
  testScoredSamplingPlanner();
  testTrajectoryPlanner();
  testLaserScanIngest();
  testTrajectoryCostFunction();
  timeFusedStackEndToEnd();
  printf("%s\n", failures ? "DROP-IN FAILED" : "DROP-IN OK");
  return failures ? 1 : 0;
}
