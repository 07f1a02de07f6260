// Stand-in for camodocal's Camera interface (camera_model/include/camodocal/camera_models/
// Camera.h): the three virtuals feature_tracker.cpp calls.  TEST INFRASTRUCTURE.
#pragma once
#include <memory>
#include <string>
#include <eigen3/Eigen/Dense>
#include <opencv2/core/core.hpp>
namespace camodocal {
class Camera {
 public:
  virtual ~Camera() {}
  virtual void liftProjective(const Eigen::Vector2d& p, Eigen::Vector3d& P) const = 0;
  virtual void spaceToPlane(const Eigen::Vector3d&, Eigen::Vector2d&) const {}  // visualisation only
  virtual cv::Mat initUndistortRectifyMap(cv::Mat&, cv::Mat&) const { return cv::Mat(); }
};
typedef std::shared_ptr<Camera> CameraPtr;  // boost::shared_ptr in camodocal
}  // namespace camodocal
