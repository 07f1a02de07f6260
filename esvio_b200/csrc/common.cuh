// common.cuh -- shared types and device helpers of libesvio_fe (sm_100a only).
//
// HBM layout of one stereo stream (all row-major, index x + y*W as the reference's
// Eigen::MatrixXd(W,H) planes, feature_tracker/src/event_detector/event_detector.h:74-79):
//   sae  : double2[2 cams][H][W]   .x = sae[0] (last ACCEPTED negative), .y = sae[1]
//   lat  : double2[2 cams][H][W]   .x = sae_latest[0],                  .y = sae_latest[1]
//   pyr  : u8 image pyramids (level 0 = the CV_8U time surface) of cur-left (ping-pong
//          with prev-left) and cur-right
// Events of a window sit in HBM either as SoA (x u16, y u16, t f64, p u8) or as the raw
// 16-byte dvs_msgs::Event records, and are re-ordered once per window into per-tile
// runs (stable counting sort by 16x8 fine tile) so that one CTA owns one 32x8 tile in shared
// memory and each of its warps one fine tile's run.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace esvio {

constexpr int kTileW = 32;                    // 32x8-pixel tile: one TMA box of the SAE state,
constexpr int kTileH = 8;                     // one CTA of K1
constexpr int kTilePx = kTileW * kTileH;      // 256 pixels: binned key = local pixel | pol << 8
constexpr int kPolShift = 8;
constexpr int kFineW = 16;                    // events are binned by 16x8 fine tile ...
constexpr int kFine = kTileW / kFineW;        // ... 2 per tile, one warp each in K1
constexpr int kChunkThreads = 256;
constexpr int kChunkSteps = 8;                // events per thread in the binning kernels
constexpr int kChunk = kChunkThreads * kChunkSteps;  // 2048 events per CTA
constexpr int kMaxLevels = 4;                 // maxLevel = 3 (feature_tracker.cpp:410)
constexpr int kWin = 21;                      // cv::Size(21, 21)
constexpr int kHalfWin = 10;
constexpr int kMaxCnt = 1024;                 // hard cap on MAX_CNT
constexpr int kMaxCams = 16;                  // cameras per launch: 8 stereo streams of one group
constexpr int kSlots = 6;                     // windows in flight: a window is a ~0.4 ms chain of short
                                              // kernels, a new one can start every ~0.08 ms
constexpr int kCornerBlock = 128;             // events per CTA of k_corner_flags = per candidate list
constexpr int kResultHdr = 32;                // int32 words in front of the result arrays
constexpr int kResultArrays = 15;

struct DevEvents {
  const uint16_t* x;
  const uint16_t* y;
  const double* t;
  const uint8_t* p;
  const uint4* aos;  // dvs_msgs::Event records (feature_tracker/src/dvs_msgs/Event.h:42-52)
  int n;
  // motion-compensated pixel of every event (k_warp_events); null = use x, y as they are.
  // Only the SAE update sees these; corner detection keeps the raw coordinates like the
  // reference (feature_tracker.cpp:740 passes event_left itself).
  const uint16_t* wx;
  const uint16_t* wy;
};

struct Ev {
  int x, y, p;
  double t;
};

// e.ts.toSec() == (double)sec + 1e-9 * (double)nsec (ros::Time); compiled with -fmad=false
// so the multiply and the add round separately like the reference's x86 build.
__device__ __forceinline__ Ev load_event(const DevEvents& e, int i) {
  Ev r;
  if (e.aos) {
    const uint4 v = __ldg(e.aos + i);
    r.x = (int)(v.x & 0xffffu);
    r.y = (int)(v.x >> 16);
    r.t = (double)v.y + 1e-9 * (double)v.z;
    r.p = (v.w & 0xffu) ? 1 : 0;
  } else {
    r.x = __ldg(e.x + i);
    r.y = __ldg(e.y + i);
    r.t = __ldg(e.t + i);
    r.p = __ldg(e.p + i) ? 1 : 0;
  }
  if (e.wx) {
    r.x = __ldg(e.wx + i);
    r.y = __ldg(e.wy + i);
  }
  return r;
}

struct PyrDesc {
  int levels;
  int w[kMaxLevels], h[kMaxLevels], pitch[kMaxLevels];
  uint32_t off[kMaxLevels];  // byte offset of the level inside one pyramid buffer
  uint32_t bytes;            // size of one pyramid buffer
};

struct Pinhole {
  double fx, fy, cx, cy, k1, k2, p1, p2;
};

// Per-stream tracker bookkeeping that lives on the device between windows.
struct TrackState {
  int n_prev;   // tracks carried in from the previous window (prev_pts.size())
  int n_cur;    // cur_pts.size() as the window progresses
  int n_right;  // matched right points of this window
  int next_id;  // FeatureTracker::n_id (feature_tracker.cpp:9)
  int n_prev_un;    // entries of prev_un_pts_map
  int n_prev_un_r;  // entries of prev_un_right_pts_map
  int stat_after_temporal, stat_after_ransac, stat_after_mask, stat_new;
  int stat_corner_flags, stat_ransac_iters, stat_n_prev;
  int pad[3];
};

// ---------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = (i < 0) ? -i : 2 * n - 2 - i;
  return i;
}

// The same fold without a loop, for |overshoot| < 2n - 2 (two folds): LK only ever reaches at
// most 43 pixels beyond an image whose sides are >= 22.  Being branch-free matters: a batch of
// loads whose addresses go through reflect101()'s loop is issued one by one behind the loop's
// branches, each paying its own trip to L2 (measured: 32 staging loads took 9 800 cycles,
// profiles/r2_lk_phases.txt).
__device__ __forceinline__ int reflect101_nb(int i, int n) {
  int a = abs(i);
  a = min(a, 2 * n - 2 - a);
  a = abs(a);
  return min(a, 2 * n - 2 - a);
}

// cvRound(float): round half to even
__device__ __forceinline__ int cv_round(float v) { return __float2int_rn(v); }

// ---------------------------------------------------------------------------------
// TMA / mbarrier wrappers (PTX; SASS shows UTMALDG / UTMASTG)
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)
               : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  const uint32_t addr = smem_u32(bar);
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* smem_src, int c0,
                                             int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
          reinterpret_cast<uint64_t>(map)),
      "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// L2 prefetch of one box (no shared memory, no barrier): the tile a CTA launched later will load
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read0() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------------------------
// Programmatic dependent launch: a window is a chain of ~17 short kernels on three streams, and
// the hand-over from one kernel to the next (drain, launch latency, first instruction) costs as
// much as many of them run.  Every kernel of the chain releases its successor right away
// (pdl_launch_dependents) and waits for its predecessor's completion and memory flush
// (pdl_wait) before it touches anything, so the successor's launch overlaps the predecessor's
// execution while the data dependencies stay exactly those of plain stream order.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
// CAUTION: a plain load through a `const T* __restrict__` parameter compiles to ld.global.nc,
// which nvcc may hoist ABOVE griddepcontrol.wait (it did in k_gftt_pick): whatever the
// predecessor kernel produces must be read with __ldcg / through a non-const pointer.
// tests/test_sass_pdl.py scans the SASS of every kernel for a global load in front of ACQBULK.
#define PDL_PROLOGUE()       \
  do {                       \
    pdl_launch_dependents(); \
    pdl_wait();              \
  } while (0)

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at;
  at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------------------------
// launch wrappers implemented in the .cu files
// ---------------------------------------------------------------------------------
struct BinLayout {
  int W, H, tiles_x, tiles_y, n_tiles;  // coarse 32x8 tiles
  int n_bins;                           // kFine * n_tiles fine tiles; bin n_bins = out of range
  int max_chunks;                       // chunk capacity of `counts`
};

// Every event-stage kernel covers `n_cams` cameras per launch (blockIdx.y / .z = camera): the
// two cameras of one stereo stream, or the 2S cameras of a group of S streams.
struct EventStageBuffers {
  int n_cams;
  uint32_t* counts;     // [2 passes][n_cams][max_chunks][128]: per-chunk digit histograms
  uint32_t* bin_total;  // [n_cams][n_bins+1] events per fine tile (cleared by the sort itself)
  uint32_t* bin_start;  // [n_cams][n_bins+2]
  double* bt[kMaxCams];    // binned event times
  uint16_t* bk[kMaxCams];  // binned keys: local pixel (8 bits) | polarity << 8
  double* it[kMaxCams];    // after the first pass (sorted by x / 16): time,
  uint16_t* ik[kMaxCams];  // key,
  uint8_t* im[kMaxCams];   // y / 8
};
int event_stage_alloc(const BinLayout& L, int n_cams, int cap, EventStageBuffers* E);
void event_stage_free(EventStageBuffers* E);
EventStageBuffers event_stage_cam_view(const BinLayout& L, const EventStageBuffers& E, int cam);
void event_stage_clear(const BinLayout& L, const EventStageBuffers& E, cudaStream_t s);
int bin_configure(const BinLayout& L);

// by-value kernel parameter (__grid_constant__: indexed by camera without a local copy)
struct CamBatch {
  int n_cams;
  int n_chunks[kMaxCams];
  DevEvents ev[kMaxCams];
  double* bt[kMaxCams];
  uint16_t* bk[kMaxCams];
  double* it[kMaxCams];
  uint16_t* ik[kMaxCams];
  uint8_t* im[kMaxCams];
};

void launch_bin_events(const BinLayout& L, const EventStageBuffers& B, const DevEvents* ev,
                       cudaStream_t s, int64_t* launches);

// Motion_correction_value (feature_tracker.h:35) as the SAE update reads it
struct McParams {
  float v_cur[3], v_pre[3], omega[3];
  float K[4];   // fx, fy, cx, cy of EventDetector::intrinsics_matrix
  double t1;    // left header stamp
  int W, H;
};
// createSAE_*(..., measurements) coordinates (event_detector.cc:102-147,168-210,547-591) for
// both cameras: wx/wy[cam][i]; t0 = time of the first left event
void launch_warp_events(const McParams& P, const DevEvents ev[2], uint16_t* const wx[2],
                        uint16_t* const wy[2], cudaStream_t s, int64_t* launches);
// the warp on caller-supplied (x, y, dt) triples (parity tests)
void launch_warp_points(const McParams& P, const float* xy_dt, int n, int* out_xy, cudaStream_t s,
                        int64_t* launches);

struct SaeTsParams {
  int W, H, tiles_x, n_tiles, n_cams;
  double decay_sec, inv_decay, filter_threshold;  // inv_decay = RN(1 / decay_sec)
  int ignore_polarity;
  const uint32_t* bin_start;  // [n_cams][kFine*n_tiles+2]
  double t_ref[kMaxCams];     // per camera (the two cameras of a stream share theirs)
  const double* bt[kMaxCams];
  const uint16_t* bk[kMaxCams];
  uint8_t* ts[kMaxCams];      // level-0 images
  int ts_pitch;
  int prefetch_dist;          // CTAs ahead whose tile state is prefetched into L2 (0 = off)
  // filled in by launch_sae_update_ts:
  int pf_dx, pf_dy, pf_dz;    // prefetch_dist split into grid coordinates (x + gx * (y + gy * z))
  float ts_karg;              // -log2(e) / decay_sec: exponent of the float time-surface estimate
};
void launch_sae_update_ts(const SaeTsParams& P, const CUtensorMap& map_sae,
                          const CUtensorMap& map_lat, cudaStream_t s, int64_t* launches);

struct CornerParams {
  int W, H, min_dist;
  double filter_threshold, ts_lk_threshold;
  const double2* sae;  // left camera planes
  const double2* lat;
  const uint8_t* ts;   // left level-0 image (may be null when and_ts_test == 0)
  int ts_pitch;
  int and_ts_test;
  // per CTA of kCornerBlock events: the flagged events' pixels (x | y << 16) in stream order
  // and their number; null = flags only
  uint32_t* cand;   // [ceil(n / kCornerBlock) * kCornerBlock]
  int* cand_cnt;    // [ceil(n / kCornerBlock)]
  // per-pixel plane of k_corner_plane ([H][W] bytes: bit p = Arc* verdict of polarity p, bit 2 + p =
  // "evaluated"); null: every event evaluates its own pixel
  uint8_t* plane;
};
void launch_corner_flags(const CornerParams& P, const DevEvents& ev, uint8_t* flags,
                         cudaStream_t s, int64_t* launches);

void launch_pyramids(const PyrDesc& pd, uint8_t* const* pyr, int n_img, cudaStream_t s,
                     int64_t* launches);

// dst[i] = max over the n_src planes of srcs[k][i] (time-window shard: times only grow, 0 =
// never, so "the last event before me" over earlier windows is an element-wise maximum)
constexpr int kMaxMergeSrc = 10;
void launch_merge_max(double* dst, const double* const* srcs, int n_src, size_t n, cudaStream_t s,
                      int64_t* launches);

// optional conditioning of the time surface (imgops.cu)
void launch_median(const uint8_t* const src[2], uint8_t* const dst[2], int n_img, int W, int H,
                   int pitch, int ksize, cudaStream_t s, int64_t* launches);
size_t clahe_lut_bytes();
void launch_equalize(const uint8_t* const src[2], uint8_t* const tmp[2], uint8_t* const dst[2],
                     int n_img, int W, int H, int pitch, uint8_t* lut, int* minmax,
                     cudaStream_t s, int64_t* launches);

void launch_clahe_inplace(uint8_t* const img[2], int n_img, int W, int H, int pitch, uint8_t* lut,
                          int* minmax, cudaStream_t s, int64_t* launches);

// cv::calcOpticalFlowPyrLK(I, J, prev, next, status, err, Size(21,21), max_level[, 30/0.01,
// USE_INITIAL_FLOW]); n is read on the device.
// mode 0: that call alone.  mode 1: followed, per point and in the same launch, by the
// temporal backward check (J->I, maxLevel 1, initial flow = prev).  mode 2: followed by the
// stereo backward check (J->I, same maxLevel, no initial flow).
void launch_lk(const PyrDesc& pd, const uint8_t* I, const uint8_t* J, const float2* prev_pts,
               float2* next_pts, uint8_t* status, float2* rev_pts, uint8_t* rev_status,
               const int* n_ptr, int n_max, int max_level, int use_initial_flow, int mode,
               cudaStream_t s, int64_t* launches);

struct TrackBuffers {
  TrackState* st;
  float2 *prev_pts, *cur_pts, *rev_pts;
  // stereo LK outputs, one set per in-flight slot ([kSlots][max_cnt]): the stereo LK of window
  // k+1 runs while window k is still being packed
  float2 *right_pts, *rev_left_pts;
  int *ids, *cnt;
  uint8_t *st_fwd, *st_bwd, *st_sf, *st_sb;
  int* prev_un_ids;
  float2* prev_un;
  int* prev_un_r_ids;
  float2* prev_un_r;
  int32_t* result;  // [kSlots] blocks of kResultHdr ints + kResultArrays * max_cnt words: window k
                    // packs into block k % kSlots, so a consumer on another stream (the
                    // all-gather) can still read it while the next windows finalize
  int result_words;
  // snapshot of (cur_pts, ids, track_cnt, counters) taken at the end of the temporal/selection
  // stage of a window, one per in-flight slot: the stereo stage of window k reads it while
  // the temporal stage of window k+1 already rewrites cur_pts / ids / cnt
  float2* snap_pts;  // [kSlots][max_cnt]
  int* snap_ids;     // [kSlots][max_cnt]
  int* snap_cnt;     // [kSlots][max_cnt]
  int* snap_hdr;     // [kSlots][16]: n, stat_n_prev, after_temporal, after_ransac, after_mask,
                     //               new, corner_flags, ransac_iters, next_id
  void* rs;                   // RansacScratch (ransac.cu)
  const uint32_t* rng_draws;  // first ransac_num_draws() raw outputs of cv::RNG(-1)
  // the 7 sample indices of every RANSAC attempt for every point count rs_cache_lo..max_cnt
  // (they depend on the count only): [n - rs_cache_lo][1280 attempts][8] u16, and the number of
  // attempts the draw table affords
  const uint16_t* rs_idx_cache;
  const int* rs_natt_cache;
  int rs_cache_lo;
};

struct TrackParams {
  int W, H, max_cnt, min_dist, flow_back;
  double focal_length, f_threshold;
  Pinhole cam[2];
};

// frame path: scratch of cv::goodFeaturesToTrack (frames.cu), allocated on first use
struct GfttBuffers {
  float* cov[3];                 // Dx*Dx, Dx*Dy, Dy*Dy                      [H][W]
  float* eig;                    // cornerMinEigenVal                        [H][W]
  uint32_t* blocked;             // 1 bit per pixel: mask == 0               [H][(W+31)/32]
  float* thr;                    // maxVal * qualityLevel
  unsigned long long* keys;      // float_order(eig) << 32 | pixel of the candidates, compacted [H*W]
  unsigned long long* keys_sorted;  // the same, best first, then a zero key [H*W]
  int* n_cand;                   // number of candidates of the frame
  float2* out_xy;                // stage entry: picked corners              [H*W] (capacity)
  int* out_n;
};
void launch_gftt_eig(const GfttBuffers& G, const uint8_t* img, int pitch, int W, int H,
                     cudaStream_t s, int64_t* launches);
void launch_gftt_thr(const GfttBuffers& G, int W, int H, bool use_mask, cudaStream_t s,
                     int64_t* launches);
int launch_gftt_candidates(const GfttBuffers& G, int W, int H, bool use_mask, cudaStream_t s,
                           int64_t* launches);
// Image_setMask (feature_tracker.cpp:91-121): survivors compacted in place, their discs
// rastered into G.blocked
void launch_image_set_mask(const TrackParams& P, const TrackBuffers& B, const GfttBuffers& G,
                           cudaStream_t s, int64_t* launches);
// greedy minimum-distance pick over G.keys_sorted.  Tracker form: appends up to
// max_cnt - n_cur new points (ids from next_id, track_cnt 1), updates the counters and takes
// the snapshot of slot `snap_slot`.  Stage form: writes up to max_corners (<= 0: all) corners
// to G.out_xy / G.out_n.
void launch_gftt_pick_tracks(const TrackParams& P, const TrackBuffers& B, const GfttBuffers& G,
                             int snap_slot, cudaStream_t s, int64_t* launches);
void launch_gftt_pick_stage(const GfttBuffers& G, int W, int H, int max_corners,
                            double min_distance, cudaStream_t s, int64_t* launches);
// the right-camera velocity map survives a frame without a right image
// (feature_tracker.cpp:245: the whole block is skipped): stash / restore its size
void launch_right_map_keep(const TrackBuffers& B, int restore, cudaStream_t s, int64_t* launches);

// snap_slot >= 0: the kernel is the last one of the window's temporal stage and also takes the
// per-slot snapshot (cur_pts, ids, track_cnt, counters) the stereo stage works from
void launch_post_temporal(const TrackParams& P, const TrackBuffers& B, int snap_slot,
                          cudaStream_t s, int64_t* launches);
// cand / cand_cnt: the candidate lists k_corner_flags left for the window's n_events left events
void launch_select(const TrackParams& P, const TrackBuffers& B, int n_events, const uint32_t* cand,
                   const int* cand_cnt, int snap_slot, cudaStream_t s, int64_t* launches);
// test entry: order[k] = index libstdc++'s std::sort (key descending) leaves at position k
void launch_sort_order(const int* key, int n, int depth_limit, int* order, cudaStream_t s, int64_t* launches);
void launch_finalize(const TrackParams& P, const TrackBuffers& B, int slot, double cur_time,
                     double prev_time, cudaStream_t s, int64_t* launches);
void launch_ransac(const TrackParams& P, const TrackBuffers& B, cudaStream_t s,
                   int64_t* launches);
// stage entry: F-RANSAC mask on caller points (device pointers)
void launch_ransac_stage(const TrackParams& P, const TrackBuffers& B, const float2* p1,
                         const float2* p2, int n, double thresh, uint8_t* mask, int* iters,
                         cudaStream_t s, int64_t* launches);
size_t ransac_scratch_bytes();
size_t ransac_cache_idx_bytes(int n_lo, int n_hi);
int ransac_build_cache(const TrackParams& P, const TrackBuffers& B, const float2* dummy, int n_lo,
                       int n_hi, uint16_t* cache_idx, int* cache_natt, cudaStream_t s);
int ransac_num_draws();
void ransac_fill_draw_table(uint32_t* host_table);
void launch_undistort(const Pinhole& cam, const float2* uv, int n, float2* out, cudaStream_t s,
                      int64_t* launches);

size_t select_smem_bytes(int W, int H);

// Raise a kernel's dynamic shared-memory limit on the CURRENT device to at least `bytes`
// (cudaFuncSetAttribute is per device and a process may hold handles of several sizes on
// several GPUs: the limit only ever grows).  `state` is the caller's static table.
struct SmemLimit {
  size_t bytes[64];  // per device ordinal
};
template <typename K>
inline int raise_dyn_smem(K kernel, size_t bytes, SmemLimit* state) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
  // benign race between handles created from different threads: both would set a sufficient value
  if (bytes <= state->bytes[dev]) return 0;
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  state->bytes[dev] = bytes;
  return 0;
}

void prefer_shared_lk();  // k_lk's shared-memory carve-out preference (lk.cu)
void prefer_shared_events();   // the same for k_sae_update_ts (events.cu)

}  // namespace esvio
