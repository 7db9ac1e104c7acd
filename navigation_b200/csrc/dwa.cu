// dwa.cu -- Path B (placeholder while the kernels are being written): every entry point reports UNSUPPORTED.
#include "common.cuh"
using namespace navgpu;
struct navgpu_dwa { int unused; };
extern "C" {
void navgpu_dwa_default_config(navgpu_dwa_config* c) { memset(c, 0, sizeof(*c)); }
#define NB return fail(NAVGPU_ERR_UNSUPPORTED, "Path B not built yet")
int navgpu_dwa_create(navgpu_dwa**, const navgpu_dwa_config*, uint32_t, uint32_t, double, int) { NB; }
int navgpu_dwa_destroy(navgpu_dwa*) { return NAVGPU_OK; }
int navgpu_dwa_reconfigure(navgpu_dwa*, const navgpu_dwa_config*) { NB; }
int navgpu_dwa_set_costmap(navgpu_dwa*, const uint8_t*, double, double) { NB; }
int navgpu_dwa_set_costmap_device(navgpu_dwa*, const uint8_t*, uint32_t, double, double) { NB; }
int navgpu_dwa_set_plan(navgpu_dwa*, const double*, const double*, int) { NB; }
int navgpu_dwa_reset_oscillation(navgpu_dwa*) { NB; }
int navgpu_dwa_get_oscillation_mask(navgpu_dwa*, int*) { NB; }
int navgpu_dwa_find_best_path(navgpu_dwa*, const double*, const double*, const double*, int, navgpu_dwa_result*, double*, int, double*, int) { NB; }
int navgpu_dwa_score_range(navgpu_dwa*, const double*, const double*, const double*, int, int64_t, int64_t, double*, int64_t*, int64_t*) { NB; }
int navgpu_dwa_finish_sharded(navgpu_dwa*, const double*, const double*, const double*, const int64_t*, int, navgpu_dwa_result*, double*, int) { NB; }
int navgpu_dwa_get_grid(navgpu_dwa*, int, double*) { NB; }
int navgpu_dwa_find_best_path_async(navgpu_dwa*, const double*, const double*, const double*, int) { NB; }
int navgpu_dwa_synchronize(navgpu_dwa*) { NB; }
void* navgpu_dwa_stream(navgpu_dwa*) { return nullptr; }
}
