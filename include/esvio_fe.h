/*
 * esvio_fe.h -- C ABI of the B200-native ESVIO event front-end (libesvio_fe.so).
 *
 * Drop-in boundary for the hot path of arclab-hku/ESVIO's stereo_event_tracker
 * node.  One `esvio_fe` handle replaces the reference's global
 * `FeatureTracker trackerData` + `esvio::EventDetector detector`
 * (feature_tracker/src/stereo_event_tracker_node.cpp:45,
 *  feature_tracker/src/feature_tracker.cpp:7) for ONE stereo event stream.
 *
 * Plain pointers and sizes only; no C++/torch types.  Every entry point
 * returns an esvio_status (0 = ok) and never throws.  A handle is not
 * thread-safe (the reference has exactly one caller thread, sync_process,
 * stereo_event_tracker_node.cpp:366); many handles may live in one process.
 * All device work of a handle runs on its own CUDA stream on `device_id`.
 *
 * There is no CPU fallback: esvio_fe_create fails with ESVIO_FE_ENODEV when no
 * CUDA device is usable.
 */
#ifndef ESVIO_FE_H
#define ESVIO_FE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ESVIO_FE_ABI_VERSION 3

typedef enum esvio_status {
  ESVIO_FE_OK = 0,
  ESVIO_FE_EINVAL = 1,    /* bad argument / config */
  ESVIO_FE_ENODEV = 2,    /* no usable CUDA device */
  ESVIO_FE_ECUDA = 3,     /* CUDA runtime error (see esvio_fe_last_error) */
  ESVIO_FE_ECAPACITY = 4, /* more events than max_events_per_window / tracks capacity too small */
  ESVIO_FE_ESTATE = 5     /* call sequence error (e.g. wait without submit) */
} esvio_status;

/* camodocal PINHOLE calibration (config/<set>/event{0,1}_esvio.yaml;
 * camera_model/src/camera_models/PinholeCamera.cc:292-295) */
typedef struct esvio_pinhole {
  double fx, fy, cx, cy, k1, k2, p1, p2;
} esvio_pinhole;

/* Mirrors the globals filled by readParameters_event
 * (feature_tracker/src/parameters.cpp:201-229,273-275). */
typedef struct esvio_fe_config {
  int32_t width;                   /* COL_event */
  int32_t height;                  /* ROW_event */
  int32_t max_cnt;                 /* MAX_CNT */
  int32_t min_dist;                /* MIN_DIST */
  int32_t flow_back;               /* FLOW_BACK */
  int32_t equalize;                /* EQUALIZE: CLAHE(40, 8x8) + normalize(0,255,MINMAX) before LK */
  double f_threshold;              /* F_THRESHOLD */
  double ts_lk_threshold;          /* TS_LK_THRESHOLD */
  double decay_ms;                 /* para_decay_ms */
  int32_t ignore_polarity;         /* para_ignore_polarity */
  int32_t median_blur_kernel_size; /* para_median_blur_kernel_size k: medianBlur(2k+1), 0..7 */
  double feature_filter_threshold; /* para_feature_filter_threshold */
  int32_t do_motion_correction;    /* Do_motion_correction: 1 allows esvio_fe_track_mc (allocates the warp scratch) */
  double focal_length;             /* FOCAL_LENGTH = 460 (parameters.cpp:274) */
  esvio_pinhole cam[2];            /* left, right */
  int32_t device_id;
  int32_t max_events_per_window;   /* per camera; staging capacity */
  int32_t use_ransac;              /* 1 = rejectWithF_event enabled (reference behaviour) */
  int32_t reserved0;
  /* EventDetector::intrinsics_matrix of the motion compensation (event_detector.cc:95-97):
   * the globals fx, fy, cx, cy left by stereo_readIntrinsicParameter
   * (feature_tracker.cpp:963-976), i.e. the LAST camera's rectified K.  mc_fx <= 0 selects
   * that default: (float)cam[1].fx, (float)cam[1].fy, width / 2, height / 2 (integer halves,
   * PinholeCamera.cc:750-767). */
  double mc_fx, mc_fy, mc_cx, mc_cy;
  int32_t reserved[6];
} esvio_fe_config;

/* The fields of the reference's Motion_correction_value (feature_tracker.h:35, packed at
 * stereo_event_tracker_node.cpp:252) that the SAE update reads. */
typedef struct esvio_motion {
  double state_v[3]; /* State_[0..2]: current linear velocity from odometry (node.cpp:209-216) */
  float v_pre[3];    /* previous velocity (node.cpp:218-220) */
  float accel[3];    /* temp_a: velocity-differenced acceleration (node.cpp:230-232); the warp
                      * is applied only when its norm exceeds 5 m/s^2 (event_detector.cc:125) */
  float omega[3];    /* IMU angular velocity (node.cpp:243-245) */
  double t1;         /* left EventArray header stamp (feature_tracker.cpp:620); the window's
                      * first left event time t0 is read from the events themselves (:619) */
} esvio_motion;

/* One camera's events for one window (dvs_msgs/EventArray.events).
 * Either SoA (x,y,t,p all non-NULL) or AoS (`aos` non-NULL): an array of
 * 16-byte dvs_msgs::Event records {u16 x; u16 y; u32 sec; u32 nsec; u8 polarity; pad[3]}
 * (feature_tracker/src/dvs_msgs/Event.h:42-52), converted on the GPU with
 * ros::Time::toSec() arithmetic.  Caller-owned, read-only, valid for the call.
 * Events with x >= width or y >= height are dropped and counted. */
typedef struct esvio_events {
  const uint16_t *x;
  const uint16_t *y;
  const double *t;  /* seconds, e.ts.toSec() */
  const uint8_t *p; /* polarity 0/1 */
  const void *aos;
  size_t n;
  int32_t on_device; /* 1: pointers are device pointers on the handle's device */
  int32_t flags;     /* ESVIO_EVENTS_* (0 when unused) */
} esvio_events;
/* set on BOTH cameras' structs: the eight arrays lie in ONE host allocation laid out by
 * esvio_fe_soa_layout_stereo (the library then moves the window with a single transfer) */
#define ESVIO_EVENTS_STEREO_BLOCK 1

typedef struct esvio_stats {
  int32_t n_events[2];
  int32_t n_dropped[2];       /* out-of-range events */
  int32_t n_prev;             /* tracks entering the window */
  int32_t n_after_temporal;   /* after LK + backward check + border test */
  int32_t n_after_ransac;
  int32_t n_after_mask;
  int32_t n_new;              /* new Arc* corners appended */
  int32_t n_corner_flags;     /* left events flagged corner && TS != threshold */
  int32_t ransac_iters;
  int32_t reserved[5];
} esvio_stats;

/* Caller-allocated outputs; replaces the public std::vectors the node reads
 * after trackEvent (feature_tracker/src/feature_tracker.h:126-135;
 * stereo_event_tracker_node.cpp:289-323).  Every array holds >= capacity
 * entries, capacity >= config.max_cnt. */
typedef struct esvio_tracks {
  int32_t capacity;
  int32_t n_left;
  int32_t *id;        /* ids */
  int32_t *track_cnt; /* track_cnt */
  float *u, *v;       /* cur_pts */
  float *un_x, *un_y; /* cur_un_pts */
  float *vx, *vy;     /* pts_velocity */
  int32_t n_right;
  int32_t *id_right;    /* ids_right */
  float *ru, *rv;       /* cur_right_pts */
  float *run_x, *run_y; /* cur_un_right_pts */
  float *rvx, *rvy;     /* right_pts_velocity */
  esvio_stats stats;
} esvio_tracks;

/* ---- lifecycle ---- */
typedef struct esvio_fe esvio_fe;

int esvio_fe_abi_version(void);
/* Values common to every shipped config (config/<set>/es*io.yaml front-end block):
 * max_cnt 150, min_dist 10, flow_back 1, F_threshold 1, TS_LK_threshold 128,
 * decay_ms 20, feature_filter_threshold 0.01, focal_length 460, use_ransac 1. */
void esvio_fe_default_config(esvio_fe_config *cfg, int32_t width, int32_t height);
/* replaces detector.init(COL_event, ROW_event) (feature_tracker.cpp:347-350) and
 * stereo_readIntrinsicParameter (feature_tracker.cpp:963-976) */
int esvio_fe_create(const esvio_fe_config *cfg, esvio_fe **out);
void esvio_fe_destroy(esvio_fe *fe);
/* clears SAE, previous image, tracks and the id counter */
int esvio_fe_reset(esvio_fe *fe);
const char *esvio_fe_strerror(int status);
const char *esvio_fe_last_error(const esvio_fe *fe);

/* ---- the hot path ---- */
/* replaces FeatureTracker::trackEvent(cur_time, event_left, event_right)
 * (feature_tracker.h:51; feature_tracker.cpp:340-603; call site
 * stereo_event_tracker_node.cpp:193).  `pub_this_frame` is the reference's
 * global PUB_THIS_FRAME (stereo_event_tracker_node.cpp:179,188).
 * Synchronous: returns after the <= 2*max_cnt track records are on the host. */
int esvio_fe_track(esvio_fe *fe, double cur_time, const esvio_events *left,
                   const esvio_events *right, int32_t pub_this_frame, esvio_tracks *out);
/* The same call split in two so that consecutive windows overlap on the GPU (a window is a
 * graph of short kernels on several streams; only the SAE update, the temporal chain and the
 * packing are serial from window to window): at most esvio_fe_pipeline_depth() windows may be
 * in flight; waits return results in order.  Host event buffers passed to submit must stay
 * valid and unmodified until the matching wait returns (the copy is asynchronous).  A wait
 * whose `out` is too small (ESVIO_FE_ECAPACITY) leaves the window waitable. */
int esvio_fe_pipeline_depth(void);
int esvio_fe_track_submit(esvio_fe *fe, double cur_time, const esvio_events *left,
                          const esvio_events *right, int32_t pub_this_frame);
int esvio_fe_track_wait(esvio_fe *fe, esvio_tracks *out);
/* replaces FeatureTracker::trackEvent(cur_time, event_left, event_right, measurements)
 * (feature_tracker.h:52; feature_tracker.cpp:605-877; call site
 * stereo_event_tracker_node.cpp:254): the SAE update warps every event of the window back to
 * the first event's time (EventDetector::motioncorrection, event_detector.cc:547-591) when
 * |accel| > 5; everything after the SAE update is the plain path.  `mc` == NULL is the plain
 * call.  Needs config.do_motion_correction = 1. */
int esvio_fe_track_mc(esvio_fe *fe, double cur_time, const esvio_events *left,
                      const esvio_events *right, int32_t pub_this_frame, const esvio_motion *mc,
                      esvio_tracks *out);
int esvio_fe_track_submit_mc(esvio_fe *fe, double cur_time, const esvio_events *left,
                             const esvio_events *right, int32_t pub_this_frame,
                             const esvio_motion *mc);

/* ---- the frame front-end ---- */
/* replaces FeatureTracker::trackImage(cur_time, img_left, img_right)
 * (feature_tracker.h:49; feature_tracker.cpp:164-338; call site
 * stereo_image_tracker_node.cpp:99): forward + full backward pyramidal LK on the previous
 * frame, Image_setMask (:91-121), cv::goodFeaturesToTrack(img, MAX_CNT_IMG - n, 0.01,
 * MIN_DIST_IMG, mask) (:228), stereo LK against the right frame, undistortion, velocities.
 * Images are CV_8UC1, width x height of the config, `stride` bytes per row; config.max_cnt /
 * min_dist play MAX_CNT_IMG / MIN_DIST_IMG.  right == NULL is img_right.empty() (mono).
 * config.equalize = 1 applies the image node's EQUALIZE step (cv::createCLAHE()->apply on both
 * frames, stereo_image_tracker_node.cpp:93-97) on the GPU first: a node that keeps its own CLAHE
 * call must create the frame handle with equalize = 0.
 * Use a handle of its own for frames (the reference runs them in a separate node,
 * stereo_image_tracker_node.cpp:45).  Results come back through esvio_fe_track_wait.  The
 * submit form copies the frames asynchronously: they must stay valid and unmodified until the
 * matching esvio_fe_track_wait returns. */
int esvio_fe_track_image(esvio_fe *fe, double cur_time, const uint8_t *left, size_t left_stride,
                         const uint8_t *right, size_t right_stride, int32_t pub_this_frame,
                         esvio_tracks *out);
int esvio_fe_track_image_submit(esvio_fe *fe, double cur_time, const uint8_t *left,
                                size_t left_stride, const uint8_t *right, size_t right_stride,
                                int32_t pub_this_frame);

/* ---- groups: several independent stereo streams on one GPU ---- */
/* S handles of the same configuration whose event stage (binning, SAE update + time surface,
 * pyramids) runs as one batched launch sequence per window -- one k_sae_update_ts launch covers
 * all 2S cameras -- while the per-stream tracking stages run concurrently on the members' own
 * streams (SURVEY.md 8e "independent streams"; BASELINE configs[4]).  Results are identical to
 * S separate handles.  The reference would need one process per stream (its tracker state is
 * global, feature_tracker.cpp:7-9).  1 <= n_streams <= 8; equalize / median blur / motion
 * compensation are not available in a group.  All arrays below have n_streams entries. */
typedef struct esvio_fe_group esvio_fe_group;
int esvio_fe_group_create(const esvio_fe_config *cfg, int32_t n_streams, esvio_fe_group **out);
void esvio_fe_group_destroy(esvio_fe_group *g);
int esvio_fe_group_reset(esvio_fe_group *g); /* esvio_fe_reset for every member */
/* member i, for the read-only queries (esvio_fe_time_surface, esvio_fe_get_sae, ...) */
esvio_fe *esvio_fe_group_member(esvio_fe_group *g, int32_t i);
int esvio_fe_group_track(esvio_fe_group *g, const double *cur_time, const esvio_events *left,
                         const esvio_events *right, const int32_t *pub_this_frame,
                         esvio_tracks *out);
int esvio_fe_group_track_submit(esvio_fe_group *g, const double *cur_time,
                                const esvio_events *left, const esvio_events *right,
                                const int32_t *pub_this_frame);
int esvio_fe_group_track_wait(esvio_fe_group *g, esvio_tracks *out);
int esvio_fe_group_kernel_launches(esvio_fe_group *g, int64_t *count);
/* CUDA-event milliseconds of the batched k_sae_update_ts launch of the last completed window */
int esvio_fe_group_sae_ts_ms(esvio_fe_group *g, float *ms);

/* replaces FeatureTracker::gettimesurface() (feature_tracker.cpp:894-897): the
 * CV_8U time surface of the last window; dst has `stride` bytes per row. */
int esvio_fe_time_surface(esvio_fe *fe, int32_t cam, uint8_t *dst, size_t stride);

/* ---- memory helpers ---- */
/* Byte offsets of x, y, t, p inside ONE block that holds the four SoA arrays of n events (each
 * array 16-byte aligned).  SoA events whose host pointers follow this layout cross PCIe as a
 * single copy; any other placement is copied array by array. */
void esvio_fe_soa_layout(size_t n, size_t *offsets /* 4 */, size_t *total_bytes);
/* The same for BOTH cameras of a window in one block: the left camera's arrays as above, the
 * right camera's behind them (its block starts at the first multiple of 256 bytes).  A window
 * whose eight host pointers follow this layout inside one pinned allocation, declared with
 * ESVIO_EVENTS_STEREO_BLOCK in both structs' flags, is copied with a single transfer (4.3 MB at 50 GB/s instead of two 2.2 MB transfers at 45 GB/s on a B200 host;
 * the end-to-end rate of esvio_fe_track_submit is bound by exactly this). */
void esvio_fe_soa_layout_stereo(size_t n_left, size_t n_right, size_t *offsets_left /* 4 */,
                                size_t *offsets_right /* 4 */, size_t *total_bytes);
void *esvio_fe_host_alloc(size_t bytes); /* pinned host memory for event staging */
void esvio_fe_host_free(void *p);
int esvio_fe_device_alloc(esvio_fe *fe, size_t bytes, void **out);
int esvio_fe_device_free(esvio_fe *fe, void *p);
int esvio_fe_copy_to_device(esvio_fe *fe, void *dst, const void *src, size_t bytes);

/* ---- multi-GPU plumbing ---- */
/* Device-resident packed track records of the last completed window
 * (header + 15 arrays of max_cnt int32/float), for one collective
 * (all-gather) per window; and the handle's cudaStream_t. */
int esvio_fe_result_device_ptr(esvio_fe *fe, void **ptr, size_t *bytes);
int esvio_fe_stream(esvio_fe *fe, void **cuda_stream);
/* The same block for a consumer on a stream of its own (the all-gather of a publish window,
 * kept off the tracking streams): _acquire makes `consumer_stream` (cudaStream_t) wait for the
 * most recently submitted window's packed records and returns their device address (every
 * in-flight window has its own block); after enqueuing its reads the caller calls _release on
 * the same stream, and the block is not rewritten (a pipeline depth later) before those reads
 * have finished.  No host synchronisation on either side. */
int esvio_fe_result_acquire(esvio_fe *fe, void *consumer_stream, void **ptr, size_t *bytes);
int esvio_fe_result_release(esvio_fe *fe, void *consumer_stream);

/* The replica mode's one collective (independent stereo streams, one per
 * GPU): an all-gather of every rank's packed track block of a publish window, so that any rank
 * (or rank 0's adapter) can publish all clouds.  NCCL is resolved at run time (libnccl.so.2; the
 * copy already loaded in the process, if any; a process that brings its own NCCL, e.g. through
 * PyTorch, has to load it before the first call below), the library does not link against it.
 *   rank 0:     esvio_fe_nccl_unique_id(id)          (ncclGetUniqueId; ship the 128 bytes)
 *   every rank: esvio_fe_comm_init(fe, id, rank, world)   or  _comm_attach(fe, ncclComm_t, ...)
 *   per publish window, after esvio_fe_track_submit: esvio_fe_allgather_tracks(fe)
 * The collective runs on a stream of its own behind the window's packing, into one of two
 * alternating receive buffers; the tracking streams never wait for it.  _gathered_tracks
 * returns the latest receive buffer ([world] blocks of bytes_per_rank, device memory) and the
 * cudaStream_t it is ordered on. */
int esvio_fe_nccl_unique_id(void *id128);
int esvio_fe_comm_init(esvio_fe *fe, const void *id128, int32_t rank, int32_t world);
int esvio_fe_comm_attach(esvio_fe *fe, void *nccl_comm, int32_t rank, int32_t world);
int esvio_fe_allgather_tracks(esvio_fe *fe);
int esvio_fe_gathered_tracks(esvio_fe *fe, void **dev_blocks, size_t *bytes_per_rank, void **cuda_stream);

/* Left/right split of ONE stereo stream over two GPUs.  The right camera
 * only feeds its pyramid to the stereo LK (feature_tracker.cpp:475-495), so its createSAE_right /
 * SAEtoTimeSurface_right / pyramid (feature_tracker.cpp:358-368) can run on a second GPU:
 *   right GPU:  esvio_fe_split_image_submit(fe_r, t, &right_events, xs, &img, &n)
 *               then send the n bytes at img (NCCL send / peer copy) on stream xs
 *   left GPU:   esvio_fe_split_right_buffer(fe_l, &buf, &n); receive into buf on stream xs;
 *               esvio_fe_track_submit_split(fe_l, t, &left_events, pub, xs); esvio_fe_track_wait
 * `exchange_stream` is the cudaStream_t (0 = legacy default stream) the caller moves the image
 * on: the library orders its own streams against it with CUDA events, no host synchronisation.
 * The image block holds all pyramid levels (level 0 = the CV_8U time surface, row pitch
 * 32-byte aligned); both handles must have the same width/height.  The right handle keeps the
 * camera in its plane 0 (esvio_fe_get_sae(fe_r, 0, ...)); its cam[0] calibration is unused
 * (undistortion of the right points happens on the left GPU with cam[1]).  Results are
 * identical to esvio_fe_track on one GPU.  Motion compensation is not available in a split. */
int esvio_fe_split_image_submit(esvio_fe *fe, double cur_time, const esvio_events *events,
                                void *exchange_stream, void **image, size_t *bytes);
int esvio_fe_split_right_buffer(esvio_fe *fe, void **image, size_t *bytes);
int esvio_fe_track_submit_split(esvio_fe *fe, double cur_time, const esvio_events *left,
                                int32_t pub_this_frame, void *exchange_stream);

/* Time-window shard: consecutive windows' SAE / time-surface / corner
 * stages on different GPUs, the serial track chain on one.  createSAE_*'s acceptance test reads
 * only sae_latest_ (event_detector.cc:149-166), so a window can be replayed from a carry-in that
 * is the element-wise maximum of the windows before it; esvio_b200/shard.py TimeWindowShard owns
 * the protocol and the collectives, these are its building blocks (all stream-ordered, no host
 * synchronisation; `cuda_stream` is a cudaStream_t):
 *   _state_device_ptrs  the handle's sae / sae_latest planes (double2[2 cams][H][W], .x = polarity 0)
 *   _shard_merge_max    dst[i] = max_k srcs[k][i] over n_src <= 10 planes of n_doubles doubles
 *   _shard_event_stage  binning + SAE update + time surface + pyramids of one window, ordered behind
 *                       and in front of the caller's stream; images: _shard_images
 *   _shard_corner_candidates  the window's Arc* candidates (per 128 events: pixels x | y << 16 in
 *                       stream order, and their count) on the device; sizes: _shard_sizes
 *   _external_buffers + _track_submit_external  track a window whose images and candidate lists
 *                       were produced elsewhere and written into the returned buffers on the
 *                       caller's stream; results through esvio_fe_track_wait. */
int esvio_fe_state_device_ptrs(esvio_fe *fe, void **sae, void **lat, size_t *bytes);
int esvio_fe_shard_merge_max(esvio_fe *fe, void *dst, const void *const *srcs, int32_t n_src,
                             size_t n_doubles, void *cuda_stream);
int esvio_fe_shard_event_stage(esvio_fe *fe, double t_ref, const esvio_events *left,
                               const esvio_events *right, void *cuda_stream);
int esvio_fe_shard_corner_candidates(esvio_fe *fe, const esvio_events *left, void *cuda_stream,
                                     void **cand, void **cand_cnt);
int esvio_fe_shard_sizes(esvio_fe *fe, size_t *image_bytes, size_t *cand_bytes, size_t *cand_cnt_bytes);
int esvio_fe_shard_images(esvio_fe *fe, void **left_img, void **right_img);
int esvio_fe_external_buffers(esvio_fe *fe, void **left_img, void **right_img, void **cand,
                              void **cand_cnt);
int esvio_fe_track_submit_external(esvio_fe *fe, double cur_time, int32_t n_left_events,
                                   int32_t pub_this_frame, void *cuda_stream);

/* ---- profiling ---- */
#define ESVIO_FE_NUM_STAGES 9
/* CUDA-event milliseconds of the last completed window when profiling is on:
 * 0 wait for the h2d copies (own stream), 1 bin_events (3 kernels), 2 sae_update_ts (1 kernel), 3 pyramid, 4 corner flags,
 * 5 temporal LK + filter, 6 select (ransac+mask+corners), 7 stereo LK + pack, 8 d2h */
int esvio_fe_set_profiling(esvio_fe *fe, int32_t on);
int esvio_fe_get_stage_ms(esvio_fe *fe, float *ms /* ESVIO_FE_NUM_STAGES */);
/* The same window as a timeline: milliseconds since esvio_fe_set_profiling(fe, 1) of the
 * markers, each on the stream it reports on:  0 submit, 1 events landed (copy stream) | 12
 * binned | 2 SAE kernel start, 3 time surface done | 4 pyramids done | 5 corner flags done |
 * 10 temporal chain start, 6 temporal LK + filter done, 7 selection done | 11 stereo LK start
 * | 8 packed, 9 result on the host (result stream). */
#define ESVIO_FE_NUM_MARKS 13
int esvio_fe_get_stage_marks(esvio_fe *fe, float *ms /* ESVIO_FE_NUM_MARKS */);
int esvio_fe_kernel_launches(esvio_fe *fe, int64_t *count); /* kernels launched so far */

/* ---- stage-level entry points (used by the parity tests) ---- */
/* plane: 0 sae[0], 1 sae[1], 2 sae_latest[0], 3 sae_latest[1]; dst = H*W doubles,
 * index x + y*W (event_detector.h:74-79) */
int esvio_fe_get_sae(esvio_fe *fe, int32_t cam, int32_t plane, double *dst);
/* Teacher forcing (parity tests): replace the tracker's carried state -- what trackEvent keeps
 * from one call to the next (feature_tracker.cpp:585-590: prev_pts, ids, track_cnt,
 * prev_un_pts_map, prev_un_right_pts_map, prev_time; :9 n_id) -- by the caller's, e.g. the
 * reference's after the same window.  SAE state and images are left alone (they are exact).
 * pts / un are n (x, y) pairs, un_r are n_r pairs keyed by ids_r. */
int esvio_fe_stage_set_tracks(esvio_fe *fe, double prev_time, int32_t next_id, int32_t n,
                              const float *pts, const int32_t *ids, const int32_t *track_cnt,
                              const float *un, int32_t n_r, const int32_t *ids_r, const float *un_r);
/* createSAE_* + SAEtoTimeSurface_* + pyramids only (feature_tracker.cpp:356-368) */
int esvio_fe_stage_update(esvio_fe *fe, double t_ref, const esvio_events *left,
                          const esvio_events *right);
/* createSAE_*(..., measurements) + SAEtoTimeSurface_* + pyramids (feature_tracker.cpp:628-645) */
int esvio_fe_stage_update_mc(esvio_fe *fe, double t_ref, const esvio_events *left,
                             const esvio_events *right, const esvio_motion *mc);
/* EventDetector::motioncorrection (event_detector.cc:547-591) on n caller triples
 * (x, y, dt): out_xy gets n (x, y) int pairs */
int esvio_fe_stage_motion_correct(esvio_fe *fe, const esvio_motion *mc, const float *xy_dt,
                                  int32_t n, int32_t *out_xy);
/* EventDetector::isCorner for every left event against the current SAE
 * (event_detector.cc:308-544); flags[i] in {0,1}.  If and_ts_test != 0 the
 * TS != TS_LK_threshold test of feature_tracker.cpp:26 is ANDed in. */
int esvio_fe_stage_corner_flags(esvio_fe *fe, const esvio_events *left, int32_t and_ts_test,
                                uint8_t *flags);
/* which: 0 = current left, 1 = current right, 2 = previous left */
int esvio_fe_get_pyramid_level(esvio_fe *fe, int32_t which, int32_t level, uint8_t *dst,
                               int32_t *w, int32_t *h);
/* cv::calcOpticalFlowPyrLK(prev, next, winSize 21x21, maxLevel) as the reference
 * calls it (feature_tracker.cpp:410,417-418,490,495) on caller images (W*H u8). */
int esvio_fe_stage_lk(esvio_fe *fe, const uint8_t *prev_img, const uint8_t *next_img,
                      const float *prev_pts, float *next_pts, int32_t n, uint8_t *status,
                      int32_t max_level, int32_t use_initial_flow);
/* The optional conditioning of the time surface on a caller image (W*H u8):
 * cv::medianBlur(src, ksize) when median_ksize > 1 (event_detector.cc:262-264, ksize = 2k+1),
 * then cv::createCLAHE()->apply + cv::normalize(0,255,NORM_MINMAX) when equalize != 0
 * (feature_tracker.cpp:375-382).  The handle must have been created with equalize or
 * median_blur_kernel_size set (that allocates the scratch images). */
int esvio_fe_stage_condition(esvio_fe *fe, const uint8_t *src, int32_t median_ksize,
                             int32_t equalize, uint8_t *dst);
/* cv::goodFeaturesToTrack(img, max_corners, 0.01, min_distance, mask, blockSize 3) on a host
 * image (width x height, contiguous); mask NULL or width x height bytes (non-zero = allowed).
 * *out_n = corners found, the first min(*out_n, capacity) (x, y) pairs go to out_xy; eig
 * (nullable) receives the cv::cornerMinEigenVal plane.  min_distance > 1 needs
 * 1 <= max_corners <= 1024. */
int esvio_fe_stage_good_features(esvio_fe *fe, const uint8_t *img, const uint8_t *mask,
                                 int32_t max_corners, double min_distance, float *out_xy,
                                 int32_t capacity, int32_t *out_n, float *eig);
/* cv::findFundamentalMat(p1, p2, FM_RANSAC, thresh, 0.99, status)
 * (feature_tracker.cpp:935) */
int esvio_fe_stage_fmat_mask(esvio_fe *fe, const float *p1, const float *p2, int32_t n,
                             double thresh, uint8_t *mask, int32_t *iters);
/* Event_setMask + Event_FeaturesToTrack + id assignment
 * (feature_tracker.cpp:446-468) against the current SAE / time surface. In:
 * n tracked points; out: kept + new points (capacity max_cnt). */
int esvio_fe_stage_select(esvio_fe *fe, const esvio_events *left, int32_t n, const float *pts,
                          const int32_t *ids, const int32_t *track_cnt, int32_t *n_out,
                          float *pts_out, int32_t *ids_out, int32_t *track_cnt_out,
                          int32_t *n_kept);
/* The visiting order of Event_setMask / Image_setMask: `sort(..., a.first > b.first)`
 * (feature_tracker.cpp:100-103,132-135) as GCC's libstdc++ runs it -- order[k] = index of the
 * element at position k; ties land where the library's introsort puts them.  depth_limit < 0:
 * the library's own recursion budget (test runs force 0..30 to reach its heap-sort branch);
 * n <= 1024. */
int esvio_fe_stage_sort_order(esvio_fe *fe, const int32_t *key, int32_t n, int32_t depth_limit,
                              int32_t *order);
/* PinholeCamera::liftProjective (PinholeCamera.cc:450-510) -> (x/z, y/z) as f32 */
int esvio_fe_stage_undistort(esvio_fe *fe, int32_t cam, const float *uv, int32_t n, float *out);

#ifdef __cplusplus
}
#endif
#endif
