// Stand-in for sensor_msgs/Image (visualisation topics only).
#pragma once
#include <std_msgs/Header.h>
namespace sensor_msgs {
struct Image {
  std_msgs::Header header;
};
typedef std::shared_ptr<Image> ImagePtr;
typedef std::shared_ptr<const Image> ImageConstPtr;
}  // namespace sensor_msgs
