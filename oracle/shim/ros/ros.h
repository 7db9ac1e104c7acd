#pragma once
#include <string>
#include <ros/console.h>
#include <std_msgs/Header.h>
namespace XmlRpc { struct XmlRpcValue {}; }
namespace ros {
// parameter server stand-in: every lookup answers with the caller's default
struct NodeHandle {
  NodeHandle() {}
  NodeHandle(const std::string&) {}
  template <class T> bool param(const std::string&, T& value, const T& default_value) const { value = default_value; return false; }
};
}
