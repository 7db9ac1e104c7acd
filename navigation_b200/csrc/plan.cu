// plan.cu -- plan preprocessing of the local planners, batched:
//   base_local_planner::transformGlobalPlan   base_local_planner/src/goal_functions.cpp:86-174
//   base_local_planner::prunePlan             base_local_planner/src/goal_functions.cpp:68-84
// Both are sequential loops over the poses of ONE plan in the reference (a few hundred poses, once per control cycle:
// dwa_planner_ros.cpp:184-199).  A fleet runs them for thousands of robots per cycle; here one warp takes one plan:
// the lanes stride over its poses, the two loop exits of transformGlobalPlan ("first pose within the threshold", "first
// pose beyond it after that") are warp-wide minimum reductions over pose indices, and the kept poses are transformed
// in parallel.  tf's part -- looking the transform up and expressing the robot pose in the plan's frame -- stays with
// the caller (tf is not part of the reference tree); the call takes the resulting rigid transform and robot position.
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace navgpu {

__device__ __forceinline__ int warp_min(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// One warp per plan.  transformGlobalPlan :118-149: skip poses until one lies within dist_threshold of the robot (both in
// the plan's frame), then keep poses while the PREVIOUS kept pose was within the threshold -- so the first pose beyond
// it is still kept -- and express every kept pose in the global frame: plan_to_global_transform * pose, i.e.
// basis row . position + origin, evaluated left to right like tf::Transform::operator() (no FMA: -fmad=false).
__global__ void k_plans_transform(int n_plans, const int* __restrict__ offsets, const double* __restrict__ xyz,
                                  const double* __restrict__ robot_xy, const navgpu_rigid_transform* __restrict__ tfm,
                                  const double* __restrict__ threshold, int* __restrict__ first_out,
                                  int* __restrict__ count_out, double* __restrict__ out_xyz) {
  const int plan = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (plan >= n_plans) return;
  const int base = offsets[plan], n = offsets[plan + 1] - base;
  const double rx = robot_xy[2 * plan], ry = robot_xy[2 * plan + 1];
  const double thr = threshold[plan], sq_thr = thr * thr;
  auto sq_dist = [&](int i) {
    const double x_diff = rx - xyz[3 * (size_t)(base + i)], y_diff = ry - xyz[3 * (size_t)(base + i) + 1];
    return x_diff * x_diff + y_diff * y_diff;
  };
  int i0 = n;
  for (int i = lane; i < n && i0 == n; i += 32)
    if (sq_dist(i) <= sq_thr) i0 = i;
  i0 = warp_min(i0);
  int i1 = n - 1;  // last kept pose: the first one beyond the threshold behind i0, else the end of the plan
  for (int i = i0 + 1 + lane; i < n && i1 == n - 1; i += 32)
    if (!(sq_dist(i) <= sq_thr)) i1 = i;
  i1 = warp_min(i1);
  const int count = i0 < n ? i1 - i0 + 1 : 0;
  if (lane == 0) {
    first_out[plan] = i0;
    count_out[plan] = count;
  }
  const navgpu_rigid_transform t = tfm[plan];
  for (int k = lane; k < count; k += 32) {
    const double* p = xyz + 3 * (size_t)(base + i0 + k);
    double* o = out_xyz + 3 * (size_t)(base + k);
    o[0] = t.m[0] * p[0] + t.m[1] * p[1] + t.m[2] * p[2] + t.t[0];
    o[1] = t.m[3] * p[0] + t.m[4] * p[1] + t.m[5] * p[2] + t.t[1];
    o[2] = t.m[6] * p[0] + t.m[7] * p[1] + t.m[8] * p[2] + t.t[2];
  }
}

// prunePlan :68-84: way-points are erased from the front until one lies less than 1 m from the robot
__global__ void k_plans_prune(int n_plans, const int* __restrict__ offsets, const double* __restrict__ xyz,
                              const double* __restrict__ robot_xy, int* __restrict__ erase_out) {
  const int plan = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (plan >= n_plans) return;
  const int base = offsets[plan], n = offsets[plan + 1] - base;
  const double rx = robot_xy[2 * plan], ry = robot_xy[2 * plan + 1];
  int first = n;
  for (int i = lane; i < n && first == n; i += 32) {
    const double x_diff = rx - xyz[3 * (size_t)(base + i)], y_diff = ry - xyz[3 * (size_t)(base + i) + 1];
    if (x_diff * x_diff + y_diff * y_diff < 1) first = i;
  }
  first = warp_min(first);
  if (lane == 0) erase_out[plan] = first;
}

namespace {

struct PlanContext {  // per thread and device, like the plugin-seam calls of costmap.cu
  int device = -1;
  cudaStream_t stream = nullptr;
  char* d_buf = nullptr;
  size_t capacity = 0;
};

int plan_context(int device, PlanContext** out, size_t need) {
  static thread_local PlanContext ctx[16];
  if (device < 0 || device >= 16) return fail(NAVGPU_ERR_INVALID, "bad device %d", device);
  if (navgpu_device_count() <= device) return fail(NAVGPU_ERR_CUDA, "no CUDA device %d (libnavgpu has no CPU fallback)", device);
  PlanContext& c = ctx[device];
  NAVGPU_CUDA(cudaSetDevice(device));
  if (c.device < 0) {
    NAVGPU_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    c.device = device;
  }
  if (need > c.capacity) {
    if (c.d_buf) cudaFree(c.d_buf);
    c.d_buf = nullptr;
    NAVGPU_CUDA(cudaMalloc(&c.d_buf, 2 * need));
    c.capacity = 2 * need;
  }
  *out = &c;
  return NAVGPU_OK;
}

int check_offsets(int n_plans, const int32_t* offsets) {
  if (offsets[0] != 0) return fail(NAVGPU_ERR_INVALID, "plan offsets must start at 0");
  for (int p = 0; p < n_plans; ++p)
    if (offsets[p + 1] < offsets[p]) return fail(NAVGPU_ERR_INVALID, "plan offsets must not decrease");
  return NAVGPU_OK;
}

size_t align16(size_t b) { return (b + 15) & ~size_t(15); }

}  // namespace
}  // namespace navgpu

using namespace navgpu;

extern "C" {

int navgpu_plans_transform(int n_plans, const int32_t* offsets, const double* plan_xyz, const double* robot_xy,
                           const navgpu_rigid_transform* plan_to_global, const double* dist_threshold,
                           int32_t* first_out, int32_t* count_out, double* transformed_xyz, int device) {
  if (n_plans < 0 || (n_plans > 0 && (!offsets || !robot_xy || !plan_to_global || !dist_threshold || !first_out || !count_out)))
    return fail(NAVGPU_ERR_INVALID, "bad arguments");
  if (n_plans == 0) return NAVGPU_OK;
  NAVGPU_TRY(check_offsets(n_plans, offsets));
  const size_t total = (size_t)offsets[n_plans];
  if (total > 0 && (!plan_xyz || !transformed_xyz)) return fail(NAVGPU_ERR_INVALID, "bad arguments");
  const size_t b_xyz = align16(total * 3 * sizeof(double)), b_off = align16((size_t)(n_plans + 1) * sizeof(int32_t)),
               b_rob = align16((size_t)n_plans * 2 * sizeof(double)), b_tf = align16((size_t)n_plans * sizeof(navgpu_rigid_transform)),
               b_thr = align16((size_t)n_plans * sizeof(double)), b_int = align16((size_t)n_plans * sizeof(int32_t));
  PlanContext* c;
  NAVGPU_TRY(plan_context(device, &c, 2 * b_xyz + b_off + b_rob + b_tf + b_thr + 2 * b_int));
  char* p = c->d_buf;
  double* d_xyz = reinterpret_cast<double*>(p); p += b_xyz;
  double* d_out = reinterpret_cast<double*>(p); p += b_xyz;
  navgpu_rigid_transform* d_tf = reinterpret_cast<navgpu_rigid_transform*>(p); p += b_tf;
  double* d_rob = reinterpret_cast<double*>(p); p += b_rob;
  double* d_thr = reinterpret_cast<double*>(p); p += b_thr;
  int* d_off = reinterpret_cast<int*>(p); p += b_off;
  int* d_first = reinterpret_cast<int*>(p); p += b_int;
  int* d_count = reinterpret_cast<int*>(p);
  if (total) NAVGPU_CUDA(cudaMemcpyAsync(d_xyz, plan_xyz, total * 3 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  NAVGPU_CUDA(cudaMemcpyAsync(d_off, offsets, (size_t)(n_plans + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  NAVGPU_CUDA(cudaMemcpyAsync(d_rob, robot_xy, (size_t)n_plans * 2 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  NAVGPU_CUDA(cudaMemcpyAsync(d_tf, plan_to_global, (size_t)n_plans * sizeof(navgpu_rigid_transform), cudaMemcpyHostToDevice, c->stream));
  NAVGPU_CUDA(cudaMemcpyAsync(d_thr, dist_threshold, (size_t)n_plans * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  const int warps_per_block = 8;
  k_plans_transform<<<(n_plans + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, c->stream>>>(
      n_plans, d_off, d_xyz, d_rob, d_tf, d_thr, d_first, d_count, d_out);
  NAVGPU_LAUNCHED(1);
  NAVGPU_CUDA(cudaGetLastError());
  NAVGPU_CUDA(cudaMemcpyAsync(first_out, d_first, (size_t)n_plans * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  NAVGPU_CUDA(cudaMemcpyAsync(count_out, d_count, (size_t)n_plans * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  if (total) NAVGPU_CUDA(cudaMemcpyAsync(transformed_xyz, d_out, total * 3 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  NAVGPU_CUDA(cudaStreamSynchronize(c->stream));
  return NAVGPU_OK;
}

int navgpu_plans_prune(int n_plans, const int32_t* offsets, const double* plan_xyz, const double* robot_xy,
                       int32_t* erase_count_out, int device) {
  if (n_plans < 0 || (n_plans > 0 && (!offsets || !robot_xy || !erase_count_out))) return fail(NAVGPU_ERR_INVALID, "bad arguments");
  if (n_plans == 0) return NAVGPU_OK;
  NAVGPU_TRY(check_offsets(n_plans, offsets));
  const size_t total = (size_t)offsets[n_plans];
  if (total > 0 && !plan_xyz) return fail(NAVGPU_ERR_INVALID, "bad arguments");
  const size_t b_xyz = align16(total * 3 * sizeof(double)), b_off = align16((size_t)(n_plans + 1) * sizeof(int32_t)),
               b_rob = align16((size_t)n_plans * 2 * sizeof(double)), b_int = align16((size_t)n_plans * sizeof(int32_t));
  PlanContext* c;
  NAVGPU_TRY(plan_context(device, &c, b_xyz + b_off + b_rob + b_int));
  char* p = c->d_buf;
  double* d_xyz = reinterpret_cast<double*>(p); p += b_xyz;
  double* d_rob = reinterpret_cast<double*>(p); p += b_rob;
  int* d_off = reinterpret_cast<int*>(p); p += b_off;
  int* d_erase = reinterpret_cast<int*>(p);
  if (total) NAVGPU_CUDA(cudaMemcpyAsync(d_xyz, plan_xyz, total * 3 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  NAVGPU_CUDA(cudaMemcpyAsync(d_off, offsets, (size_t)(n_plans + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  NAVGPU_CUDA(cudaMemcpyAsync(d_rob, robot_xy, (size_t)n_plans * 2 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  const int warps_per_block = 8;
  k_plans_prune<<<(n_plans + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, c->stream>>>(n_plans, d_off, d_xyz,
                                                                                                    d_rob, d_erase);
  NAVGPU_LAUNCHED(1);
  NAVGPU_CUDA(cudaGetLastError());
  NAVGPU_CUDA(cudaMemcpyAsync(erase_count_out, d_erase, (size_t)n_plans * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  NAVGPU_CUDA(cudaStreamSynchronize(c->stream));
  return NAVGPU_OK;
}

}  // extern "C"
