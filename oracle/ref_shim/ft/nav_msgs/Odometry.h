// Stand-in for nav_msgs/Odometry (header stamp + twist.twist.linear, node.cpp:209-214).
#pragma once
#include <std_msgs/Header.h>
namespace nav_msgs {
struct Odometry {
  typedef std::shared_ptr<const Odometry> ConstPtr;
  std_msgs::Header header;
  struct {
    struct {
      struct { double x = 0, y = 0, z = 0; } linear, angular;
    } twist;
  } twist;
};
}  // namespace nav_msgs
