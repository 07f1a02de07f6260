/*
 * esvio_oracle.h -- CPU oracle for the ESVIO event front-end.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the shipped
 * product path; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library, and only as the
 * checker or the timed CPU baseline.
 *
 * It is a plain-C, single-thread restatement of the reference algorithm
 * (paths relative to /root/reference):
 *   feature_tracker/src/event_detector/event_detector.cc:149-166,212-228  SAE update
 *   feature_tracker/src/event_detector/event_detector.cc:230-305          time surface
 *   feature_tracker/src/event_detector/event_detector.cc:308-544          Arc* corner test
 *   feature_tracker/src/event_detector/event_detector.cc:102-147,168-210,547-591  motion-compensated SAE
 *   feature_tracker/src/feature_tracker.cpp:13-38                          corner selection
 *   feature_tracker/src/feature_tracker.cpp:48-72                          border test / compaction
 *   feature_tracker/src/feature_tracker.cpp:123-151                        min-distance mask
 *   feature_tracker/src/feature_tracker.cpp:340-603                        per-window orchestration
 *   feature_tracker/src/feature_tracker.cpp:375-382                        CLAHE + normalize (EQUALIZE)
 *   feature_tracker/src/feature_tracker.cpp:910-947                        F-matrix rejection
 *   feature_tracker/src/feature_tracker.cpp:991-1045                       undistort, velocity
 *   camera_model/src/camera_models/PinholeCamera.cc:450-510,646-662        liftProjective
 * and of the third-party arithmetic the path calls but the reference tree does
 * not contain (OpenCV, version unpinned by feature_tracker/CMakeLists.txt:17;
 * restated from the published algorithm of modules/video/src/lkpyramid.cpp,
 * modules/imgproc/src/pyramids.cpp, modules/imgproc/src/drawing.cpp and
 * modules/calib3d/src/{fundam,ptsetreg}.cpp and pinned against cv2 4.13.0 in
 * the build container by tests/golden/make_golden.py).
 *
 * Parity status: the reference ships no tests, fixtures or golden vectors for
 * this path.  What pins this oracle instead:
 *   - SAE update, time surface, Arc* and the motion-compensated update are checked
 *     against the REFERENCE'S OWN CODE: oracle/_ref/libesvio_ref.so is
 *     event_detector.cc compiled unmodified from /root/reference against stand-in
 *     Eigen/OpenCV/ROS headers (oracle/ref_shim/, recipe in oracle/Makefile);
 *     tests/test_oracle_ref.py demands identical planes, time surfaces and corner
 *     decisions on the synthetic streams.  (For motion compensation the stand-in's
 *     Matrix3f arithmetic is this file's statement of Eigen's kernels, so there the
 *     reference's control flow is pinned and Eigen's float kernels are not.)
 *   - the OpenCV-derived stages are pinned against real OpenCV (cv2) outputs
 *     committed under tests/golden/.
 *   - the whole per-window tracker -- trackEvent (both overloads) and trackImage with the
 *     bookkeeping of feature_tracker.cpp (mask, selection order, compaction, velocity,
 *     ids, state roll) -- is checked against the REFERENCE'S OWN FeatureTracker:
 *     oracle/_ref/libesvio_ref_ft.so is feature_tracker.cpp + event_detector.cc compiled
 *     unmodified (liftProjective cut out of PinholeCamera.cc), with the OpenCV algorithms
 *     supplied by this file's cv2-pinned restatements; tests/test_oracle_ref_tracker.py
 *     demands bit-exact equality of every public result vector on every window.
 *   - the order std::sort leaves equal track counts in (Event_setMask / Image_setMask) is
 *     libstdc++'s, transcribed in ora_std_sort_order and checked against the real std::sort.
 *   - what stays unpinned: Eigen's Matrix3f::exp() rounding (motion compensation, see above).
 */
#ifndef ESVIO_ORACLE_H
#define ESVIO_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------- Surface of Active Events (one camera) ---------------- */
typedef struct ora_sae {
  int W, H;
  double *sae[2];    /* last ACCEPTED event time per polarity, index x + y*W */
  double *latest[2]; /* last event time per polarity */
} ora_sae;

ora_sae *ora_sae_create(int W, int H);
void ora_sae_destroy(ora_sae *s);
void ora_sae_reset(ora_sae *s);
/* event_detector.cc:149-166 / 212-228 */
void ora_sae_update(ora_sae *s, const uint16_t *x, const uint16_t *y, const double *t,
                    const uint8_t *p, size_t n, double filter_threshold);
/* event_detector.cc:230-305 (convertTo(CV_8U) rounding = round-half-even) */
void ora_time_surface(const ora_sae *s, double t_ref, double decay_ms, int ignore_polarity,
                      uint8_t *out);
/* event_detector.cc:308-544 */
int ora_is_corner(const ora_sae *s, double t, int x, int y, int p, double filter_threshold,
                  int min_dist);
void ora_corner_flags(const ora_sae *s, const uint16_t *x, const uint16_t *y, const double *t,
                      const uint8_t *p, size_t n, double filter_threshold, int min_dist,
                      uint8_t *flags);

/* ---------------- motion-compensated SAE (event_detector.cc:102-147,168-210,547-591) ------ */
/* The fields of the reference's Motion_correction_value (feature_tracker.h:35) that the path
 * reads, plus the detector's intrinsics_matrix (event_detector.cc:95-97). */
typedef struct ora_motion {
  double state_v[3]; /* State[0..2]: current velocity (double, cast to float, ed.cc:113-116) */
  float v_pre[3];    /* previous velocity */
  float accel[3];    /* "accel_avg_": the velocity-differenced acceleration (node.cpp:230-232) */
  float omega[3];    /* IMU angular velocity */
  double t1;         /* left EventArray header stamp (feature_tracker.cpp:620) */
  float K[4];        /* fx, fy, cx, cy of intrinsics_matrix */
} ora_motion;
void ora_mat3_exp_f(const float *A /*9, row-major*/, float *R);
void ora_motion_correct(const ora_motion *m, int W, int H, double ex, double ey, double dt,
                        int *ox, int *oy);
int ora_motion_active(const ora_motion *m);
void ora_sae_update_mc(ora_sae *s, const uint16_t *x, const uint16_t *y, const double *t,
                       const uint8_t *p, size_t n, double filter_threshold, const ora_motion *m,
                       double t0);

/* ---------------- raster helpers (OpenCV drawing.cpp Circle, filled) ---------------- */
/* half_width[k] for k=0..r : row offset k from the centre is filled on [cx-hw, cx+hw]. */
void ora_disc_half_widths(int r, int *half_width);
void ora_fill_disc_u8(uint8_t *mask, int W, int H, int cx, int cy, int r, uint8_t value);

/* std::sort of libstdc++ (introsort) as a permutation: order[k] = index of the element at
 * position k after sorting by key descending; depth_limit < 0 = the library's 2*floor(log2 n) */
void ora_std_sort_order(const int *key, int n, int depth_limit, int *order);
/* feature_tracker.cpp:123-151 ; ties in track_cnt end up where libstdc++'s std::sort puts them */
int ora_set_mask(int W, int H, int min_dist, int n, float *pts /*2n*/, int *ids, int *track_cnt,
                 uint8_t *mask /*W*H out: 0 / 255*/);
/* feature_tracker.cpp:13-38 */
int ora_features_to_track(const ora_sae *left, const uint16_t *x, const uint16_t *y,
                          const double *t, const uint8_t *p, size_t n, int max_corners,
                          int min_dist, const uint8_t *mask, const uint8_t *ts,
                          double ts_lk_threshold, double filter_threshold,
                          float *out_pts /*2*max_corners*/, uint8_t *mask_out /*nullable*/);

/* ---------------- OpenCV pyramid + pyramidal LK ---------------- */
/* sizes of level l ; returns the number of levels actually built (<= max_level+1) */
int ora_pyramid_sizes(int W, int H, int max_level, int win, int *w_out, int *h_out);
void ora_pyr_down(const uint8_t *src, int sw, int sh, uint8_t *dst, int dw, int dh);
void ora_scharr_deriv(const uint8_t *src, int w, int h, int16_t *dst /* 2ch */);
void ora_calc_optical_flow_pyr_lk(const uint8_t *prev, const uint8_t *next, int W, int H,
                                  const float *prev_pts, float *next_pts, int n, uint8_t *status,
                                  int win, int max_level, int max_count, double epsilon,
                                  int use_initial_flow, double min_eig_threshold);

/* ---------------- optional image conditioning (OpenCV imgproc) ---------------- */
/* cv::medianBlur, CV_8U, BORDER_REPLICATE (event_detector.cc:262-264) */
void ora_median_blur_u8(const uint8_t *src, int W, int H, int ksize, uint8_t *dst);
/* cv::createCLAHE(clip, Size(tiles,tiles))->apply (feature_tracker.cpp:377-379) */
void ora_clahe_u8(const uint8_t *src, int W, int H, double clip_limit, int tiles, uint8_t *dst);
/* cv::normalize(0, 255, NORM_MINMAX), CV_8U (feature_tracker.cpp:380-381) */
void ora_normalize_minmax_u8(const uint8_t *src, size_t n, uint8_t *dst);
/* CLAHE defaults (40, 8x8) + normalize: the EQUALIZE branch of trackEvent */
void ora_equalize_u8(const uint8_t *src, int W, int H, uint8_t *dst);

/* ---------------- cv::goodFeaturesToTrack (frame path, feature_tracker.cpp:228) ---------------- */
/* cv::cornerMinEigenVal(img, blockSize 3, ksize 3), CV_8U -> CV_32F, BORDER_REFLECT_101 */
void ora_corner_min_eigen_val_u8(const uint8_t *img, int W, int H, float *eig);
/* cv::goodFeaturesToTrack(img, maxCorners, quality, minDistance, mask, blockSize 3, Harris off);
 * mask NULL or W*H (non-zero = allowed); returns the number of (x, y) pairs written */
int ora_good_features_to_track(const uint8_t *img, int W, int H, const uint8_t *mask,
                               int max_corners, double quality, double min_distance,
                               float *out_xy);

/* ---------------- camodocal pinhole ---------------- */
typedef struct ora_pinhole {
  double fx, fy, cx, cy, k1, k2, p1, p2;
} ora_pinhole;
void ora_lift_projective(const ora_pinhole *cam, double u, double v, double *x, double *y);

/* ---------------- OpenCV findFundamentalMat(FM_RANSAC) mask ---------------- */
int ora_solve_cubic(const double *c /*4*/, double *roots /*3*/);
int ora_run_7point(const float *m1 /*14*/, const float *m2 /*14*/, double *F /*27*/);
/* returns 1 if a model was found; mask gets 0/1 per point */
int ora_find_fundamental_mask(const float *pts1, const float *pts2, int n, double thresh,
                              double confidence, int max_iters, uint8_t *mask);

/* ---------------- whole-window tracker (FeatureTracker::trackEvent) ---------------- */
typedef void (*ora_lk_fn)(const uint8_t *prev, const uint8_t *next, int W, int H,
                          const float *prev_pts, float *next_pts, int n, uint8_t *status,
                          int max_level, int use_initial_flow);
typedef int (*ora_fmat_fn)(const float *pts1, const float *pts2, int n, double thresh,
                           uint8_t *mask);
typedef void (*ora_equalize_fn)(const uint8_t *src, int W, int H, uint8_t *dst);

typedef struct ora_config {
  int width, height, max_cnt, min_dist, flow_back, equalize;
  double f_threshold, ts_lk_threshold, decay_ms;
  int ignore_polarity, median_blur_kernel_size;
  double feature_filter_threshold;
  double focal_length;
  ora_pinhole cam[2];
} ora_config;

typedef struct ora_tracks {
  int n_left;
  int *id;
  int *track_cnt;
  float *u, *v, *un_x, *un_y, *vx, *vy;
  int n_right;
  int *id_right;
  float *ru, *rv, *run_x, *run_y, *rvx, *rvy;
  /* stage counters */
  int n_prev, n_after_temporal, n_after_ransac, n_after_mask, n_new;
} ora_tracks;

typedef struct ora_tracker ora_tracker;
ora_tracker *ora_tracker_create(const ora_config *cfg);
void ora_tracker_destroy(ora_tracker *t);
void ora_tracker_set_hooks(ora_tracker *t, ora_lk_fn lk, ora_fmat_fn fm, ora_equalize_fn eq);
void ora_tracker_disable_ransac(ora_tracker *t, int disable);
/* out arrays must have capacity >= cfg.max_cnt */
int ora_tracker_track(ora_tracker *t, double cur_time, const uint16_t *lx, const uint16_t *ly,
                      const double *lt, const uint8_t *lp, size_t nl, const uint16_t *rx,
                      const uint16_t *ry, const double *rt, const uint8_t *rp, size_t nr,
                      int pub_this_frame, ora_tracks *out);
int ora_tracker_track_mc(ora_tracker *t, double cur_time, const uint16_t *lx, const uint16_t *ly,
                         const double *lt, const uint8_t *lp, size_t nl, const uint16_t *rx,
                         const uint16_t *ry, const double *rt, const uint8_t *rp, size_t nr,
                         int pub_this_frame, const ora_motion *mc, ora_tracks *out);
/* FeatureTracker::trackImage (feature_tracker.cpp:164-338) on the same tracker state:
 * cfg.max_cnt / min_dist play MAX_CNT_IMG / MIN_DIST_IMG; right == NULL: img_right.empty() */
int ora_tracker_track_image(ora_tracker *t, double cur_time, const uint8_t *left,
                            const uint8_t *right, int pub, ora_tracks *out);
/* views of internal state, for stage-level parity checks */
const ora_sae *ora_tracker_sae(const ora_tracker *t, int cam);
const uint8_t *ora_tracker_time_surface(const ora_tracker *t, int cam);
/* the image handed to LK (the time surface after the optional CLAHE + normalize) */
const uint8_t *ora_tracker_lk_image(const ora_tracker *t, int cam);
/* stage timers (seconds, accumulated): 0 sae,1 ts,2 temporal lk,3 ransac+mask+select,4 stereo lk,5 other */
void ora_tracker_timers(const ora_tracker *t, double *out6);
int ora_tracker_next_id(const ora_tracker *t); /* FeatureTracker::n_id */

#ifdef __cplusplus
}
#endif
#endif
