// Stand-in for feature_tracker/src/dvs_msgs/Event.h (a ROS-generated message header that needs
// roscpp's serialization): the fields of dvs_msgs::Event (Event.h:42-52) and ros::Time::toSec().
// TEST INFRASTRUCTURE, see oracle/Makefile (_ref/libesvio_ref_ft.so).
#pragma once
#include <cstdint>
#include <vector>
namespace ros_shim {
struct Time {
  uint32_t sec = 0, nsec = 0;
  // ros::TimeBase::toSec(): (double)sec + 1e-9 * (double)nsec
  double toSec() const { return (double)sec + 1e-9 * (double)nsec; }
};
struct Header {
  uint32_t seq = 0;
  Time stamp;
};
}  // namespace ros_shim
namespace dvs_msgs {
struct Event {
  uint16_t x = 0, y = 0;
  ros_shim::Time ts;
  uint8_t polarity = 0;
};
}  // namespace dvs_msgs
