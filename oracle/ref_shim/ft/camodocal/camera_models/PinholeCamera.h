// Stand-in for camodocal::PinholeCamera: parameter storage + the derived members its constructor
// fills (PinholeCamera.cc:272-296).  liftProjective / distortion are NOT restated: the Makefile
// cuts their definitions out of the reference's camera_model/src/camera_models/PinholeCamera.cc
// (:450-510, :646-662) into pinhole_lift.inc, which ref_ft_api.cc compiles.  TEST INFRASTRUCTURE.
#pragma once
#include "Camera.h"
namespace camodocal {
class PinholeCamera : public Camera {
 public:
  class Parameters {
   public:
    double m_k1 = 0, m_k2 = 0, m_p1 = 0, m_p2 = 0, m_fx = 1, m_fy = 1, m_cx = 0, m_cy = 0;
    double k1() const { return m_k1; }
    double k2() const { return m_k2; }
    double p1() const { return m_p1; }
    double p2() const { return m_p2; }
    double fx() const { return m_fx; }
    double fy() const { return m_fy; }
    double cx() const { return m_cx; }
    double cy() const { return m_cy; }
  };
  PinholeCamera(double fx, double fy, double cx, double cy, double k1, double k2, double p1, double p2) {
    mParameters.m_fx = fx, mParameters.m_fy = fy, mParameters.m_cx = cx, mParameters.m_cy = cy;
    mParameters.m_k1 = k1, mParameters.m_k2 = k2, mParameters.m_p1 = p1, mParameters.m_p2 = p2;
    // PinholeCamera.cc:280-295
    m_noDistortion = (mParameters.k1() == 0.0) && (mParameters.k2() == 0.0) && (mParameters.p1() == 0.0) &&
                     (mParameters.p2() == 0.0);
    m_inv_K11 = 1.0 / mParameters.fx();
    m_inv_K13 = -mParameters.cx() / mParameters.fx();
    m_inv_K22 = 1.0 / mParameters.fy();
    m_inv_K23 = -mParameters.cy() / mParameters.fy();
  }
  void liftProjective(const Eigen::Vector2d& p, Eigen::Vector3d& P) const override;
  void distortion(const Eigen::Vector2d& p_u, Eigen::Vector2d& d_u) const;

 private:
  Parameters mParameters;
  double m_inv_K11, m_inv_K13, m_inv_K22, m_inv_K23;
  bool m_noDistortion;
};
}  // namespace camodocal
