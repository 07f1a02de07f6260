// Stand-in for sensor_msgs/PointCloud, ChannelFloat32 and geometry_msgs/Point32 (fields only):
// the message the node publishes to esvio_estimator (stereo_event_tracker_node.cpp:272-339).
#pragma once
#include <std_msgs/Header.h>
namespace geometry_msgs {
struct Point32 {
  float x = 0, y = 0, z = 0;
};
}  // namespace geometry_msgs
namespace sensor_msgs {
struct ChannelFloat32 {
  std::string name;
  std::vector<float> values;
};
struct PointCloud {
  std_msgs::Header header;
  std::vector<geometry_msgs::Point32> points;
  std::vector<ChannelFloat32> channels;
};
typedef std::shared_ptr<PointCloud> PointCloudPtr;
}  // namespace sensor_msgs
