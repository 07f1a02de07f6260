// ref_ft_api.cc -- TEST INFRASTRUCTURE: C entry points around the reference's own FeatureTracker
// (feature_tracker/src/feature_tracker.cpp + event_detector/event_detector.cc, both compiled
// UNMODIFIED from /root/reference by oracle/Makefile into oracle/_ref/libesvio_ref_ft.so).
//
// What runs as the reference wrote it: FeatureTracker::trackEvent (both overloads) and
// trackImage with everything they call inside feature_tracker.cpp -- Event_FeaturesToTrack,
// Event_setMask / Image_setMask, inBorder(_event), reduceVector, rejectWithF_event's lifting,
// undistortedPts, ptsVelocity, the id counter, the state roll -- plus EventDetector, plus
// camodocal's PinholeCamera::liftProjective / distortion (cut out of PinholeCamera.cc at build
// time, see pinhole_lift.inc in the Makefile).  What is supplied from outside: the third-party
// OpenCV algorithms (mini_cv.h forwards them to the oracle's cv2-pinned restatements) and the
// Eigen stand-in (mini_eigen.h).  Used only by tests/test_oracle_ref_tracker.py to pin the
// oracle's restatement of feature_tracker.cpp's bookkeeping on the reference's code.
#include "feature_tracker.h"

#include <cstring>
#include <new>

#include "camodocal/camera_models/PinholeCamera.h"

// ---- globals of feature_tracker/src/parameters.cpp (parameters.h:6-65) that the two files read
int ROW = 480, COL = 640, ROW_event = 480, COL_event = 640, FOCAL_LENGTH = 460;
int STEREO = 1, system_mode = 0;
int MAX_CNT = 150, MAX_CNT_IMG = 150, MIN_DIST = 10, MIN_DIST_IMG = 30, WINDOW_SIZE = 20, FREQ = 10, FREQ_IMG = 10;
double F_THRESHOLD = 1.0, TS_LK_THRESHOLD = 128.0;
int para_ignore_polarity = 0;
double para_decay_ms = 20.0, para_decay_loop_ms = 20.0;
int para_median_blur_kernel_size = 0;
double para_feature_filter_threshold = 0.01;
int Do_motion_correction = 0;
double fx = 0, fy = 0, cx = 0, cy = 0, fx_event = 0, fy_event = 0, cx_event = 0, cy_event = 0;
int SHOW_TRACK = 0, FLOW_BACK = 1, STEREO_TRACK = 1, EQUALIZE = 0, FISHEYE = 0;
bool PUB_THIS_FRAME = false;
Eigen::Matrix3d Eeesntial_matrix, Eeesntial_matrix_event;
int Num_of_thread = 1;

extern esvio::EventDetector detector;  // feature_tracker.cpp:7

// Matrix3f::exp() of the Eigen stand-in: the oracle's restatement of Eigen 3.3's Pade kernels
extern "C" void ora_mat3_exp_f(const float* A, float* R);
extern "C" void esvio_ref_shim_mat3_exp_f(const float* a, float* o) { ora_mat3_exp_f(a, o); }

// ---- camodocal::PinholeCamera::liftProjective / distortion, as the reference wrote them
namespace camodocal {
#include "pinhole_lift.inc"
}

#define REF_API extern "C" __attribute__((visibility("default")))

// ---- real OpenCV behind the stand-ins (mini_cv.h): installed from Python, null = the oracle's restatements
extern "C" {
esvio_ref_lk_hook esvio_ref_hook_lk = nullptr;
esvio_ref_fm_hook esvio_ref_hook_fm = nullptr;
esvio_ref_img_hook esvio_ref_hook_clahe = nullptr, esvio_ref_hook_normalize = nullptr;
esvio_ref_gftt_hook esvio_ref_hook_gftt = nullptr;
}
REF_API void ref_ft_set_cv_hooks(esvio_ref_lk_hook lk, esvio_ref_fm_hook fm, esvio_ref_img_hook clahe,
                                 esvio_ref_img_hook normalize, esvio_ref_gftt_hook gftt) {
  esvio_ref_hook_lk = lk;
  esvio_ref_hook_fm = fm;
  esvio_ref_hook_clahe = clahe;
  esvio_ref_hook_normalize = normalize;
  esvio_ref_hook_gftt = gftt;
}

struct RefTracker {
  FeatureTracker ft;
};

// the parameter globals, the process-wide detector and id counters, as a fresh process has them
void esvio_ref_apply_config(const int* cfg, const double* dcfg) {
  ROW = ROW_event = cfg[1];
  COL = COL_event = cfg[0];
  MAX_CNT = MAX_CNT_IMG = cfg[2];
  MIN_DIST = MIN_DIST_IMG = cfg[3];
  FLOW_BACK = cfg[4];
  EQUALIZE = cfg[5];
  para_ignore_polarity = cfg[6];
  para_median_blur_kernel_size = cfg[7];
  FOCAL_LENGTH = cfg[8];
  F_THRESHOLD = dcfg[0];
  TS_LK_THRESHOLD = dcfg[1];
  para_decay_ms = dcfg[2];
  para_feature_filter_threshold = dcfg[3];
  SHOW_TRACK = 0;
  FISHEYE = 0;
  detector.~EventDetector();  // not assignable (const members): rebuilt in place
  new (&detector) esvio::EventDetector();
  FeatureTracker::n_id = 0;
  FeatureTracker::n_id_right = 0;
  // fx, fy, cx, cy: stereo_readIntrinsicParameter leaves the LAST camera's values there
  // (feature_tracker.cpp:972-976); the harness passes the values the detector should get
  fx = dcfg[4], fy = dcfg[5], cx = dcfg[6], cy = dcfg[7];
}

// what stereo_readIntrinsicParameter (feature_tracker.cpp:966-978) leaves in stereo_m_camera
void esvio_ref_install_cameras(FeatureTracker& ft, const double* dcfg) {
  ft.stereo_m_camera.clear();
  for (int c = 0; c < 2; ++c) {
    const double* k = dcfg + 4 + 8 * c;
    ft.stereo_m_camera.push_back(
        camodocal::CameraPtr(new camodocal::PinholeCamera(k[0], k[1], k[2], k[3], k[4], k[5], k[6], k[7])));
  }
}

// cfg: W, H, max_cnt, min_dist, flow_back, equalize, ignore_polarity, median_k, focal_length (ints);
// dcfg: f_threshold, ts_lk_threshold, decay_ms, filter_threshold, then 2 x (fx fy cx cy k1 k2 p1 p2).
// One tracker at a time: `detector`, n_id and the parameter globals are process-wide in the
// reference (feature_tracker.cpp:7-9, parameters.cpp).
REF_API void* ref_ft_create(const int* cfg, const double* dcfg, int image_mode) {
  (void)image_mode;
  esvio_ref_apply_config(cfg, dcfg);
  RefTracker* t = new RefTracker();
  esvio_ref_install_cameras(t->ft, dcfg);
  return t;
}

REF_API void ref_ft_set_intrinsics(double fx_, double fy_, double cx_, double cy_) {
  fx = fx_, fy = fy_, cx = cx_, cy = cy_;
}

REF_API void ref_ft_destroy(void* p) { delete static_cast<RefTracker*>(p); }

static void fill(dvs_msgs::EventArray& a, const uint16_t* x, const uint16_t* y, const uint32_t* sec,
                 const uint32_t* nsec, const uint8_t* pol, size_t n, uint32_t st_sec, uint32_t st_nsec) {
  a.events.resize(n);
  for (size_t i = 0; i < n; ++i) {
    dvs_msgs::Event& e = a.events[i];
    e.x = x[i], e.y = y[i], e.ts.sec = sec[i], e.ts.nsec = nsec[i], e.polarity = pol[i];
  }
  a.header.stamp.sec = st_sec;
  a.header.stamp.nsec = st_nsec;
}

// FeatureTracker::trackEvent(cur_time, event_left, event_right[, measurements]) with PUB_THIS_FRAME
// set as the node does (stereo_event_tracker_node.cpp:179,188).  motion == NULL: the plain overload;
// else state4[4], v_pre[3], accel[3], omega[3] of Motion_correction_value (feature_tracker.h:35).
REF_API void ref_ft_track(void* p, double cur_time, const uint16_t* lx, const uint16_t* ly, const uint32_t* lsec,
                          const uint32_t* lnsec, const uint8_t* lp, size_t nl, const uint16_t* rx,
                          const uint16_t* ry, const uint32_t* rsec, const uint32_t* rnsec, const uint8_t* rp,
                          size_t nr, int pub, uint32_t stamp_sec, uint32_t stamp_nsec, const double* state4,
                          const float* v_pre, const float* accel, const float* omega) {
  RefTracker* t = static_cast<RefTracker*>(p);
  dvs_msgs::EventArray L, R;
  fill(L, lx, ly, lsec, lnsec, lp, nl, stamp_sec, stamp_nsec);
  fill(R, rx, ry, rsec, rnsec, rp, nr, stamp_sec, stamp_nsec);
  PUB_THIS_FRAME = pub != 0;
  if (!state4) {
    t->ft.trackEvent(cur_time, L, R);
    return;
  }
  Eigen::Vector4d State;
  Eigen::Vector3f vp, a, w;
  for (int i = 0; i < 4; ++i) State[i] = state4[i];
  for (int i = 0; i < 3; ++i) vp[i] = v_pre[i], a[i] = accel[i], w[i] = omega[i];
  const Motion_correction_value m = std::make_pair(
      true, std::make_pair(std::make_pair(State, vp), std::make_pair(Eigen::Vector2d(0.0, 0.0), std::make_pair(a, w))));
  std::streambuf* old = std::cout.rdbuf(nullptr);  // detector.init(..., fx, ...) prints the matrix
  t->ft.trackEvent(cur_time, L, R, m);
  std::cout.rdbuf(old);
}

// FeatureTracker::trackImage(cur_time, img_left, img_right); right == NULL: an empty cv::Mat
REF_API void ref_ft_track_image(void* p, double cur_time, const uint8_t* left, const uint8_t* right, int pub) {
  RefTracker* t = static_cast<RefTracker*>(p);
  cv::Mat l = cv::Mat::zeros(cv::Size(COL, ROW), CV_8UC1), r;
  memcpy(l.ptr(), left, (size_t)COL * ROW);
  if (right) {
    r = cv::Mat::zeros(cv::Size(COL, ROW), CV_8UC1);
    memcpy(r.ptr(), right, (size_t)COL * ROW);
  }
  PUB_THIS_FRAME = pub != 0;
  t->ft.trackImage(cur_time, l, r);
}

// the public result vectors (feature_tracker.h:126-135); counts[0] = left, counts[1] = right
REF_API void ref_ft_counts(void* p, int* counts) {
  RefTracker* t = static_cast<RefTracker*>(p);
  counts[0] = (int)t->ft.ids.size();
  counts[1] = (int)t->ft.ids_right.size();
  counts[2] = FeatureTracker::n_id;
  counts[3] = (int)t->ft.cur_pts.size();
  counts[4] = (int)t->ft.cur_un_pts.size();
  counts[5] = (int)t->ft.pts_velocity.size();
  counts[6] = (int)t->ft.track_cnt.size();
  counts[7] = (int)t->ft.cur_right_pts.size();
  counts[8] = (int)t->ft.cur_un_right_pts.size();
  counts[9] = (int)t->ft.right_pts_velocity.size();
}

static void put(float* dst, const std::vector<cv::Point2f>& v) {
  for (size_t i = 0; i < v.size(); ++i) dst[2 * i] = v[i].x, dst[2 * i + 1] = v[i].y;
}

REF_API void ref_ft_get(void* p, int* ids, int* track_cnt, float* pts, float* un_pts, float* vel, int* ids_right,
                        float* rpts, float* run_pts, float* rvel) {
  RefTracker* t = static_cast<RefTracker*>(p);
  const FeatureTracker& f = t->ft;
  std::copy(f.ids.begin(), f.ids.end(), ids);
  std::copy(f.track_cnt.begin(), f.track_cnt.end(), track_cnt);
  put(pts, f.cur_pts);
  put(un_pts, f.cur_un_pts);
  put(vel, f.pts_velocity);
  std::copy(f.ids_right.begin(), f.ids_right.end(), ids_right);
  put(rpts, f.cur_right_pts);
  put(run_pts, f.cur_un_right_pts);
  put(rvel, f.right_pts_velocity);
}

// the image the tracker handed to LK (cur_img_left / cur_img_right after the optional CLAHE)
REF_API void ref_ft_lk_image(void* p, int cam, uint8_t* out) {
  RefTracker* t = static_cast<RefTracker*>(p);
  const cv::Mat& m = cam == 0 ? t->ft.prev_img_left : t->ft.cur_img_right;
  if (!m.empty()) memcpy(out, m.ptr(), (size_t)m.rows * m.cols);
}

// std::sort itself (this toolchain's libstdc++, the one the reference is built with on its ROS
// platform) on the element type and comparator of Event_setMask (feature_tracker.cpp:127-135):
// order[k] = original index of the element at position k.  depth_limit >= 0 runs the library's
// __introsort_loop with that budget instead of 2*floor(log2 n) (to reach its heap-sort branch).
REF_API void ref_std_sort_order(const int* key, int n, int depth_limit, int* order) {
  vector<pair<int, pair<cv::Point2f, int>>> v;
  for (int i = 0; i < n; ++i) v.push_back(make_pair(key[i], make_pair(cv::Point2f((float)i, 0.f), i)));
  auto cmp = [](const pair<int, pair<cv::Point2f, int>>& a, const pair<int, pair<cv::Point2f, int>>& b) {
    return a.first > b.first;
  };
  if (depth_limit < 0) {
    sort(v.begin(), v.end(), cmp);
  } else if (n > 1) {
    std::__introsort_loop(v.begin(), v.end(), (long)depth_limit, __gnu_cxx::__ops::__iter_comp_iter(cmp));
    std::__final_insertion_sort(v.begin(), v.end(), __gnu_cxx::__ops::__iter_comp_iter(cmp));
  }
  for (int i = 0; i < n; ++i) order[i] = v[i].second.second;
}
