// Stand-in for camodocal::CameraFactory: FeatureTracker::stereo_readIntrinsicParameter
// (feature_tracker.cpp:966-978) is compiled but not called by the harness, which installs the
// cameras directly (no cv::FileStorage here).
#pragma once
#include "Camera.h"
namespace camodocal {
class CameraFactory {
 public:
  static CameraFactory* instance() {
    static CameraFactory f;
    return &f;
  }
  CameraPtr generateCameraFromYamlFile(const std::string&) { return CameraPtr(); }
};
}  // namespace camodocal
