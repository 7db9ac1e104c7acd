#pragma once
#include <cassert>
#include <cstddef>
#define ROS_DEBUG(...) ((void)0)
#define ROS_INFO(...) ((void)0)
#define ROS_WARN(...) ((void)0)
#define ROS_ERROR(...) ((void)0)
#define ROS_FATAL(...) ((void)0)
#define ROS_DEBUG_NAMED(...) ((void)0)
#define ROS_WARN_NAMED(...) ((void)0)
#define ROS_WARN_THROTTLE(...) ((void)0)
#define ROS_ASSERT(c) assert(c)
#define ROS_ASSERT_MSG(c, ...) assert(c)
