// Stand-in for feature_tracker/src/feature_tracker.h while compiling the reference's
// event_detector.cc on its own (TEST INFRASTRUCTURE, see oracle/Makefile).  The real header
// drags in the whole FeatureTracker class (camodocal, ROS message headers, OpenCV video /
// calib3d); event_detector.{h,cc} only needs from it: the std / Eigen using-directives, the
// dvs_msgs::EventArray type of one unused member, and the two template helpers
// vectorToSkewMat / ConvertToHomogeneous -- those are NOT restated here: the Makefile cuts them
// out of the real header (feature_tracker.h:205-235) into ft_helpers.inc at build time.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <iostream>
#include <map>
#include <queue>
#include <string>
#include <vector>

#include <eigen3/Eigen/Core>
#include <eigen3/Eigen/Dense>
#include <eigen3/Eigen/Geometry>
#include <eigen3/unsupported/Eigen/MatrixFunctions>
#include <opencv2/core/core.hpp>

#include "parameters.h"  // the reference's own (symlinked next to this file)

namespace dvs_msgs {  // feature_tracker/src/dvs_msgs/Event.h:42-52, EventArray.h:44-54 (fields only)
struct Event {
  uint16_t x, y;
  struct { uint32_t sec, nsec; } ts;
  uint8_t polarity;
};
struct EventArray {
  uint32_t height, width;
  std::vector<Event> events;
};
}  // namespace dvs_msgs

using namespace std;
using namespace Eigen;

using Motion_correction_value = std::pair<bool, std::pair<std::pair<Eigen::Vector4d, Eigen::Vector3f>, std::pair<Eigen::Vector2d, std::pair<Eigen::Vector3f, Eigen::Vector3f>>>>;

#include "ft_helpers.inc"
