// Stand-in for feature_tracker/src/dvs_msgs/Event.h (a ROS-generated message header that needs
// roscpp's serialization): the fields of dvs_msgs::Event (Event.h:42-52).
// TEST INFRASTRUCTURE, see oracle/Makefile (_ref/libesvio_ref_ft.so, libesvio_ref_node.so).
#pragma once
#include <cstdint>
#include <vector>
#include <ros/ros.h>
namespace dvs_msgs {
struct Event {
  uint16_t x = 0, y = 0;
  ros::Time ts;
  uint8_t polarity = 0;
};
}  // namespace dvs_msgs
