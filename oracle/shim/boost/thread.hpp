// Stub of <boost/thread.hpp> for compiling the reference's hot-path sources without Boost.
// TEST INFRASTRUCTURE ONLY (oracle build). Maps the handful of Boost names the sources use onto std::.
#pragma once
#include <mutex>
#include <memory>
#include <functional>
#include <climits>
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <cstdint>
#include <cassert>
#include <limits>
#include <string>
#include <vector>
#include <algorithm>
namespace boost {
using std::recursive_mutex;
struct mutex : std::mutex {
  struct scoped_lock : std::unique_lock<std::mutex> {
    explicit scoped_lock(::boost::mutex& m) : std::unique_lock<std::mutex>(m) {}
  };
};
using std::unique_lock;
using std::shared_ptr;
using std::bind;
}
using namespace std::placeholders;
