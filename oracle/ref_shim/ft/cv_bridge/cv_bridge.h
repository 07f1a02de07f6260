// Stand-in for cv_bridge (pubTrackImage, SHOW_TRACK only).
#pragma once
#include <opencv2/core/core.hpp>
#include <sensor_msgs/Image.h>
namespace cv_bridge {
struct CvImage {
  CvImage(const std_msgs::Header&, const char*, const cv::Mat&) {}
  sensor_msgs::ImagePtr toImageMsg() const { return std::make_shared<sensor_msgs::Image>(); }
};
}  // namespace cv_bridge
