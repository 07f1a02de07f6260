// Stand-in for sensor_msgs/Imu (the fields stereo_event_tracker_node.cpp:113-126,241-248 reads).
#pragma once
#include <std_msgs/Header.h>
namespace sensor_msgs {
struct Imu {
  std_msgs::Header header;
  struct { double x = 0, y = 0, z = 0; } angular_velocity, linear_acceleration;
};
typedef std::shared_ptr<const Imu> ImuConstPtr;
}  // namespace sensor_msgs
