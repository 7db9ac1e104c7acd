// Stub of the generated message base_local_planner/Position2DInt (msg/Position2DInt.msg: int64 x, int64 y).
#pragma once
#include <stdint.h>
namespace base_local_planner {
struct Position2DInt {
  int64_t x, y;
  Position2DInt() : x(0), y(0) {}
};
}  // namespace base_local_planner
