// Stand-in for feature_tracker/src/utility/visualization.h (ROS publishers / markers; nothing of
// it is on the tracking path).  feature_tracker.h:238-244 names sensor_msgs::Imu in a template
// that the tracker never instantiates.
#pragma once
namespace sensor_msgs {
struct Imu {
  struct { double x, y, z; } angular_velocity, linear_acceleration;
};
}  // namespace sensor_msgs
