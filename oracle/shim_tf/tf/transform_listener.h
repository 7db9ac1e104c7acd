// Stub of tf/transform_listener.h (see transform_datatypes.h next to it): a listener that answers every look-up with
// the ONE transform the test installed, and expresses poses in the plan's frame with its inverse's stand-in --
// the test hands over robot_in_plan_frame directly, which is what tf.transformPose would have produced.
#pragma once
#include <tf/transform_datatypes.h>
namespace ros {
struct Duration { double d; explicit Duration(double v = 0) : d(v) {} };
struct Publisher { template <class M> void publish(const M&) const {} };
}
namespace tf {
struct TransformListener {
  Transform plan_to_global;
  Stamped<Pose> robot_in_plan_frame;
  bool waitForTransform(const std::string&, const ros::Time&, const std::string&, const ros::Time&, const std::string&,
                        const ros::Duration&) const { return true; }
  void lookupTransform(const std::string&, const ros::Time&, const std::string&, const ros::Time&, const std::string&,
                       StampedTransform& t) const {
    static_cast<Transform&>(t) = plan_to_global;
  }
  void transformPose(const std::string&, const Stamped<Pose>&, Stamped<Pose>& out) const { out = robot_in_plan_frame; }
};
}  // namespace tf
