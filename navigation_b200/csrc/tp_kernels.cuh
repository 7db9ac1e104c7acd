// tp_kernels.cuh -- the legacy base_local_planner::TrajectoryPlanner's rollout scoring on the device
// (base_local_planner/src/trajectory_planner.cpp:214-370 generateTrajectory), the second consumer of the footprint /
// MapGrid machinery of dwa_kernels.cuh.
//
// One warp per velocity sample.  The trajectory is a sequential fp64 recurrence (acceleration-limited velocities,
// then x / y / theta), replayed by every lane in rounds of 32 steps; lane l keeps step l of the round, evaluates its
// trig, and the position prefix is accumulated in the reference's order through shuffles.  The per-step checks
// (worldToMap, footprint, path / goal distance) then run one step per lane, the footprint's (step, edge) walks
// spread over all 32 lanes.  The reference returns at the FIRST failing step; here every step of a round is
// evaluated and the lowest failing step decides, which is the same value.
#pragma once

#include "dwa_kernels.cuh"

namespace navgpu {

struct TpSampleResult {
  double cost;         // traj.cost_: >= 0 legal, -1 footprint / off the map, -2 no path to the goal
  double ahead_gdist;  // goal_map_ at the end point pushed heading_lookahead ahead (:672-681), -1 when off the map
  int n_points;        // points the trajectory holds (steps before the failing one)
  int pad_;
};

struct TpScoreArgs {
  DwaGeom g;
  const uint32_t* path_dist;  // path_map_ target_dist, sx*sy
  const uint32_t* goal_dist;  // goal_map_
  const double* samples;      // n x (vx_samp, vy_samp, vtheta_samp)
  int n;
  double x, y, theta, vx, vy, vtheta;  // robot state (already rounded through Eigen::Vector3f, :911-912)
  double acc_x, acc_y, acc_theta;
  double sim_time, sim_granularity, angular_sim_granularity;
  int simple_attractor;
  double goal_x, goal_y;  // last pose of the global plan (simple_attractor_)
  int heading_scoring;    // heading_scoring_: path / goal distance and the heading difference of ONE step only
  double heading_scoring_timestep;
  const double* plan_xy;  // the global plan as received (not resolution-adjusted), n_plan (x, y) pairs
  int n_plan;
  double pdist_scale, gdist_scale, occdist_scale;
  double heading_lookahead;
  int allow_unknown;
  int nfp;
  double fpx[kMaxFootprint], fpy[kMaxFootprint];
  TpSampleResult* out;
  double* points;     // nullable: n x points_stride x 3
  int points_stride;  // points kept per sample
};

// TrajectoryPlanner::computeNewVelocity, trajectory_planner.h:369-375
__device__ __forceinline__ double tp_new_velocity(double vg, double vi, double a_max, double dt) {
  if ((vg - vi) >= 0) return fmin(vg, vi + a_max * dt);
  return fmax(vg, vi - a_max * dt);
}

// TrajectoryPlanner::lineCost + pointCost (trajectory_planner.cpp:389-475): Bresenham from (x0, y0) to (x1, y1), both
// ends included; false as soon as a cell is LETHAL, INSCRIBED or (NO_INFORMATION and unknown is not allowed)
__device__ bool tp_line_is_clear(const TpScoreArgs& a, int x0, int x1, int y0, int y1) {
  const int dx = abs(x1 - x0), dy = abs(y1 - y0);
  const int xinc = x1 >= x0 ? 1 : -1, yinc = y1 >= y0 ? 1 : -1;
  const bool xmajor = dx >= dy;
  const int den = xmajor ? dx : dy, numadd = xmajor ? dy : dx;
  int num = den / 2, x = x0, y = y0;
  for (int k = 0; k <= den; ++k) {
    const int c = a.g.cost[y * (int)a.g.pitch + x];
    if (c == kLethal || c == kInscribed || (c == kNoInfo && !a.allow_unknown)) return false;
    num += numadd;
    if (num >= den) {
      num -= den;
      if (xmajor) y += yinc; else x += xinc;
    }
    if (xmajor) x += xinc; else y += yinc;
  }
  return true;
}

// TrajectoryPlanner::headingDiff (:372-387) with the whole warp: the farthest pose of the global plan that is on the
// map and in clear line of sight of the robot's cell; lanes test 32 candidates at a time, from the end of the plan.
// All lanes pass the same arguments and get the same result.
__device__ double tp_heading_diff(const TpScoreArgs& a, int cell_x, int cell_y, double x, double y, double heading, int lane) {
  for (int top = a.n_plan - 1; top >= 0; top -= 32) {
    const int i = top - lane;
    bool clear = false;
    int gx_c = 0, gy_c = 0;
    if (i >= 0 && dwa_world_to_map(a.g, a.plan_xy[2 * i], a.plan_xy[2 * i + 1], gx_c, gy_c))
      clear = tp_line_is_clear(a, cell_x, gx_c, cell_y, gy_c);
    const unsigned found = __ballot_sync(0xffffffffu, clear);
    if (found) {
      const int src = __ffs(found) - 1;  // the lowest lane holds the highest plan index
      gx_c = __shfl_sync(0xffffffffu, gx_c, src);
      gy_c = __shfl_sync(0xffffffffu, gy_c, src);
      const double gx = a.g.ox + (gx_c + 0.5) * a.g.res, gy = a.g.oy + (gy_c + 0.5) * a.g.res;  // Costmap2D::mapToWorld
      // angles::shortest_angular_distance(heading, atan2(...)) = normalize_angle(to - from)
      double d = fmod(fmod(atan2(gy - y, gx - x) - heading, 2.0 * M_PI) + 2.0 * M_PI, 2.0 * M_PI);
      if (d > M_PI) d -= 2.0 * M_PI;
      return fabs(d);
    }
  }
  return 1.7976931348623157e308;  // DBL_MAX
}

__device__ void tp_score_sample(const TpScoreArgs& a, int sample, int lane, double* warp_scratch) {
  const double vx_samp = a.samples[3 * sample], vy_samp = a.samples[3 * sample + 1], vth_samp = a.samples[3 * sample + 2];
  const double vmag = hypot(vx_samp, vy_samp);
  int num_steps = a.heading_scoring ? (int)(a.sim_time / a.sim_granularity + 0.5)
                                    : (int)(fmax((vmag * a.sim_time) / a.sim_granularity, fabs(vth_samp) / a.angular_sim_granularity) + 0.5);
  if (num_steps == 0) num_steps = 1;
  const double dt = a.sim_time / num_steps;
  const uint32_t n_cells = a.g.sx * a.g.sy;  // path_map_.obstacleCosts(): the "impossible" cost (:535, :591)

  double sx = a.x, sy = a.y, sth = a.theta;  // state at the first step of the current round
  double vxi = a.vx, vyi = a.vy, vthi = a.vtheta;
  double occ_cost = 0.0, path_dist = 0.0, goal_dist = 0.0, ahead = -1.0, heading_diff = 0.0, time = 0.0;
  double fail_code = 0.0;
  int n_points = num_steps;
  double* out_pts = a.points ? a.points + (size_t)sample * a.points_stride * 3 : nullptr;

  for (int base = 0; base < num_steps; base += 32) {
    const int cnt = min(32, num_steps - base);
    // velocity / heading recurrence (:352-361): the point of step l uses theta_l; the move to step l + 1 uses the
    // velocities already advanced to l + 1 and, for x / y, still theta_l
    double th = sth, my_th = sth, my_vx = 0.0, my_vy = 0.0, my_time = 0.0;
    for (int l = 0; l < cnt; ++l) {
      if (lane == l) { my_th = th; my_time = time; }
      time += dt;  // `time += dt` once per step, as the reference accumulates it (:369)
      vxi = tp_new_velocity(vx_samp, vxi, a.acc_x, dt);
      vyi = tp_new_velocity(vy_samp, vyi, a.acc_y, dt);
      vthi = tp_new_velocity(vth_samp, vthi, a.acc_theta, dt);
      if (lane == l) { my_vx = vxi; my_vy = vyi; }
      th = th + vthi * dt;
    }
    double c, s;
    sincos(my_th, &s, &c);
    // computeNewXPosition / computeNewYPosition, trajectory_planner.h:332-347
    double ddx, ddy;
    if (my_vy != 0.0) {
      double c2, s2;
      sincos(M_PI_2 + my_th, &s2, &c2);
      ddx = (my_vx * c + my_vy * c2) * dt;
      ddy = (my_vx * s + my_vy * s2) * dt;
    } else {
      ddx = (my_vx * c + 0.0) * dt;
      ddy = (my_vx * s + 0.0) * dt;
    }
    double x = sx, y = sy, px = sx, py = sy;
    for (int l = 0; l < cnt; ++l) {
      if (lane == l) { px = x; py = y; }
      x = x + __shfl_sync(0xffffffffu, ddx, l);
      y = y + __shfl_sync(0xffffffffu, ddy, l);
    }
    sx = x;
    sy = y;
    sth = th;
    const bool active = lane < cnt;

    // ---- footprint of every step of the round: (step, vertex) cells, then (step, edge) walks, over all lanes
    double* pose_s = warp_scratch;
    unsigned* edge_max = reinterpret_cast<unsigned*>(warp_scratch + 128);
    pose_s[lane] = px;
    pose_s[32 + lane] = py;
    pose_s[64 + lane] = c;
    pose_s[96 + lane] = s;
    edge_max[lane] = 0;
    __syncwarp();
    if (a.nfp >= 3) {
      int* vcell = reinterpret_cast<int*>(warp_scratch + 128 + 16);
      const int items = cnt * a.nfp;
      const unsigned inv_nfp = (65536u + (unsigned)a.nfp - 1u) / (unsigned)a.nfp;  // exact for it < 512, nfp <= 16
      for (int it = lane; it < items; it += 32) {
        const int p = (int)(((unsigned)it * inv_nfp) >> 16), v = it - p * a.nfp;
        vcell[it] = footprint_vertex_cell(a, pose_s[p], pose_s[32 + p], pose_s[64 + p], pose_s[96 + p], v);
      }
      __syncwarp();
      for (int it = lane; it < items; it += 32) {
        const int p = (int)(((unsigned)it * inv_nfp) >> 16), e = it - p * a.nfp;
        const int next = e + 1 < a.nfp ? it + 1 : it - e;  // packed [point][vertex] rows of nfp words
        const int ec = footprint_edge_cost(a, vcell[it], vcell[next]);
        atomicMax(&edge_max[p], ec < 0 ? 0x80000000u : (unsigned)ec);
      }
      __syncwarp();
    }
    // ---- my step
    double code = 0.0, occ = 0.0, pd = 0.0, gd = 0.0;
    bool heading_step = false;
    int my_cx = 0, my_cy = 0;
    if (active) {
      int cx, cy;
      if (!dwa_world_to_map(a.g, px, py, cx, cy)) code = -1.0;  // off the known map (:274-277)
      else {
        const int centre = a.g.cost[cy * (int)a.g.pitch + cx];
        int f = (int)edge_max[lane];
        if (a.nfp < 3)  // CostmapModel::footprintCost without a polygon: the centre cell alone (costmap_model.cpp:61-67)
          f = (centre == kLethal || centre == kInscribed || (centre == kNoInfo && !a.allow_unknown)) ? -1 : centre;
        if (f < 0) code = -1.0;  // the footprint hits an obstacle (:283-285)
        else {
          occ = fmax((double)f, (double)centre);
          if (a.simple_attractor) {
            gd = (px - a.goal_x) * (px - a.goal_x) + (py - a.goal_y) * (py - a.goal_y);
          } else {
            // with heading scoring only the step inside [timestep, timestep + dt) looks at the distance maps (:318-331)
            heading_step = a.heading_scoring && my_time >= a.heading_scoring_timestep && my_time < a.heading_scoring_timestep + dt;
            if (!a.heading_scoring || heading_step) {
              const uint32_t pdi = a.path_dist[cy * (int)a.g.sx + cx], gdi = a.goal_dist[cy * (int)a.g.sx + cx];
              pd = (double)pdi;
              gd = (double)gdi;
              if (n_cells <= gdi || n_cells <= pdi) code = -2.0;  // no clear path to the goal (:337-342)
            }
            my_cx = cx;
            my_cy = cy;
          }
        }
      }
    }
    __syncwarp();
    const unsigned failing = __ballot_sync(0xffffffffu, code != 0.0);
    if (out_pts && active && (!failing || lane < __ffs(failing) - 1) && base + lane < a.points_stride) {
      out_pts[3 * (base + lane)] = px;
      out_pts[3 * (base + lane) + 1] = py;
      out_pts[3 * (base + lane) + 2] = my_th;
    }
    if (failing) {
      const int first = __ffs(failing) - 1;
      fail_code = __shfl_sync(0xffffffffu, code, first);
      n_points = base + first;
      break;
    }
    double m = active ? occ : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    occ_cost = fmax(occ_cost, m);
    if (!a.heading_scoring || a.simple_attractor) {
      path_dist = __shfl_sync(0xffffffffu, pd, cnt - 1);
      goal_dist = __shfl_sync(0xffffffffu, gd, cnt - 1);
    } else {
      const unsigned hs = __ballot_sync(0xffffffffu, heading_step);
      if (hs) {  // at most one step of a trajectory qualifies
        const int src = __ffs(hs) - 1;
        path_dist = __shfl_sync(0xffffffffu, pd, src);
        goal_dist = __shfl_sync(0xffffffffu, gd, src);
        heading_diff = tp_heading_diff(a, __shfl_sync(0xffffffffu, my_cx, src), __shfl_sync(0xffffffffu, my_cy, src),
                                       __shfl_sync(0xffffffffu, px, src), __shfl_sync(0xffffffffu, py, src),
                                       __shfl_sync(0xffffffffu, my_th, src), lane);
      }
    }
    if (base + cnt >= num_steps) {  // the end point, pushed heading_lookahead ahead (:672-681 / :760-769)
      double ag = -1.0;
      if (lane == cnt - 1) {
        int cx, cy;
        if (dwa_world_to_map(a.g, px + a.heading_lookahead * c, py + a.heading_lookahead * s, cx, cy))
          ag = (double)a.goal_dist[cy * (int)a.g.sx + cx];
      }
      ahead = __shfl_sync(0xffffffffu, ag, cnt - 1);
    }
  }
  if (lane == 0) {
    TpSampleResult r;
    const double legal = a.heading_scoring
                             ? a.occdist_scale * occ_cost + a.pdist_scale * path_dist + 0.3 * heading_diff + goal_dist * a.gdist_scale
                             : a.pdist_scale * path_dist + goal_dist * a.gdist_scale + a.occdist_scale * occ_cost;  // :362-367
    r.cost = fail_code != 0.0 ? fail_code : legal;
    r.ahead_gdist = fail_code != 0.0 ? -1.0 : ahead;
    r.n_points = n_points;
    r.pad_ = 0;
    a.out[sample] = r;
  }
}

__global__ void __launch_bounds__(kDwaWarpsPerBlock * 32) k_tp_score(TpScoreArgs a) {
  __shared__ double scratch[kDwaWarpsPerBlock][kWarpScratchDoubles];
  cudaGridDependencySynchronize();  // launched behind the two MapGrid wavefronts
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sample = blockIdx.x * kDwaWarpsPerBlock + warp;
  if (sample >= a.n) return;
  tp_score_sample(a, sample, lane, scratch[warp]);
}

}  // namespace navgpu
