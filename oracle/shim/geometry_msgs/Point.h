#pragma once
// The real message headers pull in <cmath>/<cstddef> transitively; sources on the path rely on that.
#include <cmath>
#include <cstddef>
namespace geometry_msgs { struct Point { double x = 0, y = 0, z = 0; }; }
