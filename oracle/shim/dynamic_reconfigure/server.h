#pragma once
#include <functional>
#include <cstdint>
#include <ros/ros.h>
namespace dynamic_reconfigure {
// setCallback immediately delivers the default configuration, like the real server does on start-up.
template <class C> struct Server {
  typedef std::function<void(C&, uint32_t)> CallbackType;
  Server() {}
  explicit Server(const ros::NodeHandle&) {}
  void clearCallback() { cb_ = CallbackType(); }
  void setCallback(const CallbackType& cb) { cb_ = cb; C c; cb_(c, ~0u); }
  CallbackType cb_;
};
}
