// ref_api.cc -- TEST INFRASTRUCTURE: C entry points around the reference's own EventDetector
// (feature_tracker/src/event_detector/event_detector.cc, compiled UNMODIFIED from
// /root/reference by oracle/Makefile into oracle/_ref/libesvio_ref.so).  Used only by
// tests/test_oracle_ref.py to pin the oracle (and through it the CUDA path) on the reference's
// code for createSAE_*, SAEtoTimeSurface_*, isCorner and the motion-compensated createSAE_*.
#define private public  // the SAE planes are private members (event_detector.h:74-79)
#include "event_detector/event_detector.h"
#undef private

#include <cstring>

// globals of feature_tracker/src/parameters.cpp that event_detector.cc reads
// (parameters.h:29,36-40)
int MIN_DIST = 10;
int para_ignore_polarity = 0;
double para_decay_ms = 20.0;
int para_median_blur_kernel_size = 0;
double para_feature_filter_threshold = 0.01;

static void (*g_exp_hook)(const float*, float*) = nullptr;
extern "C" void esvio_ref_shim_mat3_exp_f(const float* a, float* o) {
  if (!g_exp_hook) {
    fprintf(stderr, "esvio_ref: Matrix3f::exp() called without a hook (ref_set_exp_hook)\n");
    abort();
  }
  g_exp_hook(a, o);
}

#define REF_API extern "C" __attribute__((visibility("default")))

struct RefHandle {
  esvio::EventDetector det;
  int W, H;
};

REF_API void ref_set_exp_hook(void (*fn)(const float*, float*)) { g_exp_hook = fn; }

// detector.init(COL_event, ROW_event[, fx, fy, cx, cy]) as FeatureTracker::trackEvent does on
// its first call (feature_tracker.cpp:347-350, :612-619)
REF_API void* ref_create(int W, int H, double decay_ms, int ignore_polarity, int median_k,
                         double filter_threshold, int min_dist, int with_intrinsics, double fx,
                         double fy, double cx, double cy) {
  para_decay_ms = decay_ms;
  para_ignore_polarity = ignore_polarity;
  para_median_blur_kernel_size = median_k;
  para_feature_filter_threshold = filter_threshold;
  MIN_DIST = min_dist;
  RefHandle* h = new RefHandle();
  h->W = W;
  h->H = H;
  if (with_intrinsics) {
    std::streambuf* old = std::cout.rdbuf(nullptr);  // init() prints the matrix
    h->det.init(W, H, fx, fy, cx, cy);
    std::cout.rdbuf(old);
  } else {
    h->det.init(W, H);
  }
  h->det.cur_event_mat_left = cv::Mat::zeros(cv::Size(W, H), CV_8UC3);   // feature_tracker.cpp:352-353
  h->det.cur_event_mat_right = cv::Mat::zeros(cv::Size(W, H), CV_8UC3);
  return h;
}

REF_API void ref_destroy(void* p) { delete static_cast<RefHandle*>(p); }

// feature_tracker.cpp:356-362: createSAE_left / createSAE_right for every event in order
REF_API void ref_update(void* p, int cam, const uint16_t* x, const uint16_t* y, const double* t,
                        const uint8_t* pol, size_t n) {
  RefHandle* h = static_cast<RefHandle*>(p);
  for (size_t i = 0; i < n; ++i) {
    if (cam == 0) h->det.createSAE_left(t[i], x[i], y[i], pol[i] != 0);
    else h->det.createSAE_right(t[i], x[i], y[i], pol[i] != 0);
  }
}

// the motion-compensated overloads; use_mc[i] selects the overload per event (the caller applies
// the window rule of feature_tracker.cpp:628-642)
REF_API void ref_update_mc(void* p, int cam, const uint16_t* x, const uint16_t* y, const double* t,
                           const uint8_t* pol, const uint8_t* use_mc, size_t n,
                           const double* state4, const float* v_pre, const float* accel,
                           const float* omega, double t0, double t1) {
  RefHandle* h = static_cast<RefHandle*>(p);
  Eigen::Vector4d State;
  Eigen::Vector3f vp, a, w;
  for (int i = 0; i < 4; ++i) State[i] = state4[i];
  for (int i = 0; i < 3; ++i) vp[i] = v_pre[i], a[i] = accel[i], w[i] = omega[i];
  const esvio::Motion_correction_value m = std::make_pair(
      true, std::make_pair(std::make_pair(State, vp),
                           std::make_pair(Eigen::Vector2d(t0, t1), std::make_pair(a, w))));
  for (size_t i = 0; i < n; ++i) {
    const bool ep = pol[i] != 0;
    if (cam == 0) {
      if (use_mc[i]) h->det.createSAE_left(t[i], x[i], y[i], ep, m);
      else h->det.createSAE_left(t[i], x[i], y[i], ep);
    } else {
      if (use_mc[i]) h->det.createSAE_right(t[i], x[i], y[i], ep, m);
      else h->det.createSAE_right(t[i], x[i], y[i], ep);
    }
  }
}

REF_API void ref_time_surface(void* p, int cam, double t_ref, uint8_t* out) {
  RefHandle* h = static_cast<RefHandle*>(p);
  cv::Mat m = cam == 0 ? h->det.SAEtoTimeSurface_left(t_ref) : h->det.SAEtoTimeSurface_right(t_ref);
  memcpy(out, m.buf->data(), (size_t)h->W * h->H);
}

REF_API void ref_corner_flags(void* p, const uint16_t* x, const uint16_t* y, const double* t,
                              const uint8_t* pol, size_t n, uint8_t* flags) {
  RefHandle* h = static_cast<RefHandle*>(p);
  for (size_t i = 0; i < n; ++i) flags[i] = h->det.isCorner(t[i], x[i], y[i], pol[i] != 0) ? 1 : 0;
}

// which: 0 = sae (last accepted), 1 = sae_latest; out[y * W + x] (MatrixXd(W, H)(x, y) is
// column-major, i.e. offset x + y * W)
REF_API void ref_get_plane(void* p, int cam, int which, int pol, double* out) {
  RefHandle* h = static_cast<RefHandle*>(p);
  const Eigen::MatrixXd& m = cam == 0 ? (which ? h->det.sae_latest_[pol] : h->det.sae_[pol])
                                      : (which ? h->det.sae_latest_right[pol] : h->det.sae_right[pol]);
  memcpy(out, m.data(), sizeof(double) * (size_t)h->W * h->H);
}
