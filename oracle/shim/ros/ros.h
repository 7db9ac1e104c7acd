#pragma once
#include <string>
#include <ros/console.h>
#include <std_msgs/Header.h>
namespace XmlRpc { struct XmlRpcValue {}; }
namespace ros { struct NodeHandle { NodeHandle() {} NodeHandle(const std::string&) {} }; }
