// Stand-in for sensor_msgs/image_encodings.h: the three names the image node passes to cv_bridge.
#pragma once
#include <string>
namespace sensor_msgs {
namespace image_encodings {
const std::string MONO8 = "mono8";
const std::string BGR8 = "bgr8";
const std::string RGB8 = "rgb8";
}  // namespace image_encodings
}  // namespace sensor_msgs
