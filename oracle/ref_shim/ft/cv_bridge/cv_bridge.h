// Stand-in for cv_bridge: toCvCopy of a one-channel 8-bit image is a copy of its rows into a
// cv::Mat (what cv_bridge does for mono8 -> MONO8); colour inputs are not exercised here (the
// synthetic frames are grey) and come back as their first channel.  CvImage::toImageMsg serves
// the SHOW_TRACK visualisation only.
#pragma once
#include <cstring>
#include <opencv2/core/core.hpp>
#include <sensor_msgs/Image.h>
namespace cv_bridge {
struct CvImage {
  std_msgs::Header header;
  std::string encoding;
  cv::Mat image;
  CvImage() {}
  CvImage(const std_msgs::Header& h, const char* enc, const cv::Mat& img) : header(h), encoding(enc), image(img) {}
  sensor_msgs::ImagePtr toImageMsg() const { return std::make_shared<sensor_msgs::Image>(); }
};
typedef std::shared_ptr<CvImage> CvImagePtr;
typedef std::shared_ptr<const CvImage> CvImageConstPtr;
inline CvImagePtr toCvCopy(const sensor_msgs::Image& m, const std::string& enc) {
  CvImagePtr p = std::make_shared<CvImage>();
  p->header = m.header;
  p->encoding = enc;
  p->image = cv::Mat::zeros(cv::Size((int)m.width, (int)m.height), CV_8UC1);
  const size_t ch = m.width ? m.step / m.width : 1;
  for (uint32_t y = 0; y < m.height; ++y)
    for (uint32_t x = 0; x < m.width; ++x) p->image.at<unsigned char>((int)y, (int)x) = m.data[(size_t)y * m.step + x * ch];
  return p;
}
inline CvImagePtr toCvCopy(const sensor_msgs::ImageConstPtr& m, const std::string& enc) { return toCvCopy(*m, enc); }
}  // namespace cv_bridge
