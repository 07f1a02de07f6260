// esvio_fe.cu -- the C ABI of libesvio_fe.so (include/esvio_fe.h): handle lifecycle, HBM
// layout, host<->device staging and the per-window launch sequence that replaces
// FeatureTracker::trackEvent (feature_tracker/src/feature_tracker.cpp:340-603).
// There is no CPU path in this library: every stage is a kernel from events.cu,
// pyramid.cu, lk.cu, select.cu, ransac.cu.
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "../../include/esvio_fe.h"
#include "common.cuh"

using namespace esvio;

namespace esvio {
int select_configure(int W, int H);

}

#define FE_API extern "C" __attribute__((visibility("default")))

constexpr int kLeftBufs = kSlots + 1, kRightBufs = kSlots, kRightBase = kLeftBufs, kScratchBase = kLeftBufs + kRightBufs,
              kNumPyr = kScratchBase + 2;

// ------------------------------------------------------------------------------------------
// Staging of pageable host buffers.  The reference node holds a window's events in
// std::vector<dvs_msgs::Event> (stereo_event_tracker_node.cpp:128-142): ordinary memory, which
// cudaMemcpyAsync moves through the driver's own bounce buffers on one thread at ~10 GB/s --
// 0.5 ms for the 5.3 MB of a 640x480 window at 5 Mev/s per camera, longer than all the kernels of
// the window together.  Instead the calling thread and a few helpers copy the caller's buffer
// into a pinned ring of the handle, chunk by chunk, and the DMA of a chunk starts as soon as it
// and the chunks before it have landed; the call returns when the last chunk is enqueued, so
// the caller's buffer is free again.  One pool per process, one job at a time.
// ESVIO_FE_STAGE_THREADS=<n> sets the number of helpers (0: the driver's path).
// ------------------------------------------------------------------------------------------
class HostStager {
 public:
  static HostStager& get() {
    static HostStager s;
    return s;
  }
  int helpers() const { return (int)workers_.size(); }
  // dst_pinned / dst_dev / src: `bytes` each.  The helpers and the calling thread copy chunks
  // into pinned memory; only the calling thread talks to CUDA (concurrent cudaMemcpyAsync calls
  // from several threads serialise on the driver's lock: measured, slower than no helpers at
  // all): it enqueues the contiguous prefix of finished chunks host -> device on `stream`.
  cudaError_t run(uint8_t* dst_pinned, uint8_t* dst_dev, const uint8_t* src, size_t bytes, cudaStream_t stream) {
    std::lock_guard<std::mutex> one_job(job_mutex_);
    const int n_chunks = (int)((bytes + kChunk - 1) / kChunk);
    if (n_chunks > kMaxChunks) return cudaMemcpyAsync(dst_dev, src, bytes, cudaMemcpyHostToDevice, stream);
    {
      std::unique_lock<std::mutex> lk(m_);
      done_cv_.wait(lk, [&] { return working_ == 0; });  // stragglers of the job before
      pinned_ = dst_pinned, src_ = src, bytes_ = bytes, n_chunks_ = n_chunks;
      for (int c = 0; c < n_chunks; ++c) done_[c].store(0, std::memory_order_relaxed);
      next_.store(0);
      working_ = (int)workers_.size();
      ++generation_;
    }
    cv_.notify_all();
    cudaError_t err = cudaSuccess;
    int enqueued = 0;
    while (enqueued < n_chunks) {
      const int c = next_.fetch_add(1);
      if (c < n_chunks) copy_chunk(c);
      int ready = enqueued;
      while (ready < n_chunks && done_[ready].load(std::memory_order_acquire)) ++ready;
      if (ready > enqueued) {
        const size_t off = (size_t)enqueued * kChunk;
        const size_t len = (ready == n_chunks ? bytes : (size_t)ready * kChunk) - off;
        const cudaError_t e = cudaMemcpyAsync(dst_dev + off, dst_pinned + off, len, cudaMemcpyHostToDevice, stream);
        if (e != cudaSuccess) err = e;
        enqueued = ready;
      } else if (c >= n_chunks) {
        std::this_thread::yield();  // the helpers hold the chunks that are still missing
      }
    }
    return err;
  }

 private:
  static constexpr size_t kChunk = 256 << 10;
  static constexpr int kMaxChunks = 4096;
  HostStager() : done_(kMaxChunks) {
    int n = 1;  // measured on the B200 boxes (640x480 @5 Mev/s, 5.3 MB per window): 0 helpers 0.75 ms per
                // synchronous call, 1: 0.66, 2: 0.74, 4: 0.95 -- more threads lose to their wake-ups
    if (const char* e = getenv("ESVIO_FE_STAGE_THREADS")) n = atoi(e);
    const int hw = (int)std::thread::hardware_concurrency();
    if (hw > 0 && n > hw - 1) n = hw - 1;
    if (n < 0) n = 0;
    for (int i = 0; i < n; ++i) workers_.emplace_back([this] { loop(); });
  }
  ~HostStager() {
    {
      std::lock_guard<std::mutex> lk(m_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  void copy_chunk(int c) {
    const size_t off = (size_t)c * kChunk, len = bytes_ - off < kChunk ? bytes_ - off : kChunk;
    memcpy(pinned_ + off, src_ + off, len);
    done_[c].store(1, std::memory_order_release);
  }
  void loop() {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return stop_ || generation_ != seen; });
        if (stop_) return;
        seen = generation_;
      }
      for (;;) {
        const int c = next_.fetch_add(1);
        if (c >= n_chunks_) break;
        copy_chunk(c);
      }
      {
        std::lock_guard<std::mutex> lk(m_);
        --working_;
      }
      done_cv_.notify_all();
    }
  }
  std::vector<std::thread> workers_;
  std::vector<std::atomic<uint8_t>> done_;
  std::mutex m_, job_mutex_;
  std::condition_variable cv_, done_cv_;
  bool stop_ = false;
  uint64_t generation_ = 0;
  int working_ = 0, n_chunks_ = 0;
  uint8_t* pinned_ = nullptr;
  const uint8_t* src_ = nullptr;
  size_t bytes_ = 0;
  std::atomic<int> next_{0};
};

// ordinary (pageable) host memory?  Pinned and registered buffers go straight to the DMA engine.
static bool is_pageable(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return true;
  }
  return a.type == cudaMemoryTypeUnregistered;
}

struct esvio_fe {
  esvio_fe_config cfg;
  int dev;
  cudaStream_t stream;
  char err[256];
  int W, H;
  size_t npx;
  BinLayout bl;
  double2 *sae, *lat;  // [2][H][W]
  CUtensorMap map_sae, map_lat;
  CUtensorMap map_sae_r, map_lat_r;  // the right camera's planes alone (left-first windows)
  cudaEvent_t cl_done[kSlots], k1l_done[kSlots];  // left-first windows: left copy landed / left K1 done
  int left_first;   // ESVIO_LEFT_FIRST (default 1)
  PyrDesc pd;
  // Pyramid buffers: kLeftBufs left, kRightBufs right, 2 scratch for esvio_fe_stage_lk.  Up to
  // kSlots windows are in flight.  The left image of window k is read by the temporal LK of k
  // (as cur) and of k+1 (as prev) and by the stereo LK of k; when window k+kSlots+1 overwrites
  // it, windows <= k+1 have been waited for -- hence kSlots+1 left buffers.  The right image of
  // window k is only read by its own stereo LK: kSlots buffers.
  uint8_t* pyr[kNumPyr];
  // image conditioning (median blur / CLAHE + normalize): [stage][camera] scratch images with
  // the layout of pyramid level 0; ts_sel[cam] = the time surface the corner selection and
  // gettimesurface() see (after the median blur, before CLAHE)
  uint8_t* aux[3][2];
  uint8_t* clahe_lut;
  int* clahe_minmax;
  const uint8_t* ts_sel[2];
  uint16_t* warp_xy[2][2];  // [x|y][camera]: motion-compensated pixels of the window's events
  float mc_K[4];
  GfttBuffers gftt;  // frame path (esvio_fe_track_image): allocated on first use
  void* gftt_block;
  int cur_left;   // index of the newest left pyramid; prev_left is the one before it
  int prev_left;
  int cur_right;
  int windows;      // windows processed since create/reset
  int cap;          // events per camera per window
  uint8_t* raw[kSlots][2];  // [slot][camera] 16 B * cap: SoA carve-out or dvs_msgs::Event records
  EventStageBuffers esb[2];  // binned events of even / odd windows: binning of window k+1 runs
                             // while the SAE kernel of window k still reads window k's
  uint8_t* flags[kSlots];   // [slot] Arc* corner flags of the left events
  uint8_t* corner_plane;    // [H][W] per-pixel Arc* verdicts of the window being flagged (k_corner_plane)
  uint32_t* cand[kSlots];   // [slot] the flagged events' pixels, one list per kCornerBlock events
  int* cand_cnt[kSlots];
  // A window is a graph of short kernels; every node that has no data dependency on another
  // gets its own stream, and windows overlap wherever the data allow it (dependencies are
  // CUDA events, never host synchronisation):
  //   copy  -> bin (K0) -> SAE + time surface (K1) -+-> pyramids ------> temporal LK -> filter
  //                                                 +-> corner flags -.              [-> F-RANSAC]
  //                                                                    `-----------> [-> selection]
  //   -> stereo LK (two alternating streams) -> pack + D2H (`stream`, the one esvio_fe_stream()
  //   hands out).  Across windows only three chains are serial: K1 (the SAE state; K1 of k+1
  //   also waits for the corner flags of k, which read that state), the temporal chain
  //   (prev_pts = cur_pts) and the packing (velocity maps).
  cudaStream_t stream_c;   // host -> device copies of a window's events
  cudaStream_t stream_b;   // K0: binning (+ the motion-compensation warp)
  cudaStream_t stream_e;   // K1: SAE update + time surface (+ median / CLAHE)
  cudaStream_t stream_p;   // pyramids (frame path)
  cudaStream_t stream_r;   // left-first windows: the right camera's event stage
  cudaStream_t stream_f;   // Arc* corner flags (publish windows)
  cudaStream_t stream_t1;  // temporal chain: temporal LK, filter, F-RANSAC, selection
  cudaStream_t stream_s[2];  // stereo LK of even / odd slots
  cudaEvent_t c_done[kSlots];   // [slot] events landed
  cudaEvent_t b_done[kSlots];   // binned
  cudaEvent_t k1_done[kSlots];  // SAE + time surface (+ conditioning) done
  cudaEvent_t p_done[kSlots];   // pyramids done
  cudaEvent_t f_done[kSlots];   // corner flags done
  cudaEvent_t t1_done[kSlots];  // temporal chain done (snapshot taken)
  cudaEvent_t s_done[kSlots];   // stereo LK done
  cudaEvent_t x_ready[kSlots];  // left/right split: marks on the caller's exchange stream
  int f_pending;                // slot whose corner flags the next K1 has to wait for, or -1
  int shard_seq;                // esvio_fe_shard_event_stage calls so far (slot rotation)
  TrackBuffers tb;
  TrackParams tp;
  int32_t* h_result[kSlots];
  uint8_t* h_stage[kSlots][2];  // pinned ring for pageable input (allocated on first use), 16 B x cap each
  size_t result_words;
  int* d_scratch_n;
  float2 *d_scratch_p0, *d_scratch_p1;
  uint8_t* d_scratch_st;
  double prev_time;
  int64_t launches;
  int q_head, q_count;  // in-flight windows (results land in h_result[(q_head + k) % kSlots])
  cudaEvent_t q_done[kSlots];
  // esvio_fe_result_acquire / _release: a consumer stream (the per-window all-gather) reads
  // the device result block of slot `last_slot`; the slot's next finalize waits for r_free
  int last_slot;
  cudaEvent_t r_free[kSlots];
  int r_held[kSlots];
  int profiling;
  struct esvio_fe_group* group;  // non-null: the event stage is run by the group, batched
  // one set of markers per in-flight slot (ESVIO_FE_NUM_MARKS, include/esvio_fe.h)
  cudaEvent_t pev[kSlots][ESVIO_FE_NUM_MARKS];
  int pev_slot;
  int pev_valid[kSlots];
  int stage_ms_valid;
  float stage_ms[ESVIO_FE_NUM_STAGES];
  cudaEvent_t pev_ref;  // recorded by esvio_fe_set_profiling(on): origin of stage_marks
  float stage_marks[ESVIO_FE_NUM_MARKS];
  // multi-GPU replicas: the all-gather of the packed track records (esvio_fe_comm_*)
  void* nccl_comm;
  int comm_owned, comm_rank, comm_world;
  cudaStream_t stream_g;       // the collective's own stream
  int32_t* gathered[2];        // [world][result_words] each, alternating
  int gather_seq;
};

static int fail(esvio_fe* fe, int code, const char* what, cudaError_t ce) {
  if (fe) snprintf(fe->err, sizeof(fe->err), "%s: %s", what, ce == cudaSuccess ? "-" : cudaGetErrorString(ce));
  return code;
}
#define CU(call)                                                   \
  do {                                                             \
    cudaError_t ce_ = (call);                                      \
    if (ce_ != cudaSuccess) return fail(fe, ESVIO_FE_ECUDA, #call, ce_); \
  } while (0)

FE_API int esvio_fe_abi_version(void) { return ESVIO_FE_ABI_VERSION; }
FE_API int esvio_fe_pipeline_depth(void) { return kSlots; }

FE_API const char* esvio_fe_strerror(int s) {
  switch (s) {
    case ESVIO_FE_OK: return "ok";
    case ESVIO_FE_EINVAL: return "invalid argument or configuration";
    case ESVIO_FE_ENODEV: return "no usable CUDA device";
    case ESVIO_FE_ECUDA: return "CUDA runtime error";
    case ESVIO_FE_ECAPACITY: return "capacity exceeded";
    case ESVIO_FE_ESTATE: return "call sequence error";
    default: return "unknown status";
  }
}

FE_API const char* esvio_fe_last_error(const esvio_fe* fe) { return fe ? fe->err : ""; }

FE_API void esvio_fe_default_config(esvio_fe_config* c, int32_t width, int32_t height) {
  memset(c, 0, sizeof(*c));
  c->width = width;
  c->height = height;
  c->max_cnt = 150;
  c->min_dist = 10;
  c->flow_back = 1;
  c->equalize = 0;
  c->f_threshold = 1.0;
  c->ts_lk_threshold = 128.0;
  c->decay_ms = 20.0;
  c->ignore_polarity = 0;
  c->median_blur_kernel_size = 0;
  c->feature_filter_threshold = 0.01;
  c->do_motion_correction = 0;
  c->focal_length = 460.0;
  for (int i = 0; i < 2; ++i) {
    c->cam[i].fx = c->cam[i].fy = 460.0;
    c->cam[i].cx = width / 2.0;
    c->cam[i].cy = height / 2.0;
  }
  c->device_id = 0;
  c->max_events_per_window = 1 << 20;
  c->use_ransac = 1;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

FE_API void esvio_fe_soa_layout(size_t n, size_t* offsets, size_t* total_bytes) {
  size_t o = 0;
  const size_t w[4] = {2, 2, 8, 1};
  for (int i = 0; i < 4; ++i) {
    if (offsets) offsets[i] = o;
    o = align_up(o + w[i] * n, 16);
  }
  if (total_bytes) *total_bytes = o;
}

FE_API void esvio_fe_soa_layout_stereo(size_t n_left, size_t n_right, size_t* offsets_left,
                                       size_t* offsets_right, size_t* total_bytes) {
  size_t tl = 0, tr = 0, ol[4], orr[4];
  esvio_fe_soa_layout(n_left, ol, &tl);
  esvio_fe_soa_layout(n_right, orr, &tr);
  const size_t base_r = align_up(tl, 256);
  for (int i = 0; i < 4; ++i) {
    if (offsets_left) offsets_left[i] = ol[i];
    if (offsets_right) offsets_right[i] = base_r + orr[i];
  }
  if (total_bytes) *total_bytes = base_r + tr;
}

static void build_pyr_desc(int W, int H, PyrDesc* pd) {
  int w = W, h = H;
  size_t off = 0;
  pd->levels = 0;
  for (int l = 0; l < kMaxLevels; ++l) {
    pd->w[l] = w;
    pd->h[l] = h;
    pd->pitch[l] = (int)align_up(w, 32);
    pd->off[l] = (uint32_t)off;
    off = align_up(off + (size_t)pd->pitch[l] * h, 256);
    pd->levels = l + 1;
    w = (w + 1) / 2;
    h = (h + 1) / 2;
    if (w <= kWin || h <= kWin) break;  // buildOpticalFlowPyramid stops here
  }
  for (int l = pd->levels; l < kMaxLevels; ++l) pd->w[l] = pd->h[l] = pd->pitch[l] = 0, pd->off[l] = 0;
  pd->bytes = (uint32_t)off;
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_state_map(esvio_fe* fe, double2* base, CUtensorMap* map, int n_cams = 2) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CU(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (!fn || q != cudaDriverEntryPointSuccess)
    return fail(fe, ESVIO_FE_ECUDA, "cuTensorMapEncodeTiled unavailable", cudaSuccess);
  const cuuint64_t dims[3] = {(cuuint64_t)fe->W * 2, (cuuint64_t)fe->H, (cuuint64_t)n_cams};
  const cuuint64_t strides[2] = {(cuuint64_t)fe->W * 16, (cuuint64_t)fe->W * fe->H * 16};
  const cuuint32_t box[3] = {2 * kTileW, kTileH, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = ((encode_tiled_fn)fn)(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dims,
                                           strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                           CU_TENSOR_MAP_SWIZZLE_NONE,
                                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(fe->err, sizeof(fe->err), "cuTensorMapEncodeTiled failed: %d", (int)r);
    return ESVIO_FE_ECUDA;
  }
  return ESVIO_FE_OK;
}

// ---------------------------------------------------------------------------------------
// NCCL, resolved at run time: the library has no link-time dependency on it, and a process that
// already carries a copy of libnccl.so.2 (e.g. PyTorch's) gets that copy
// ---------------------------------------------------------------------------------------
struct NcclUniqueIdBytes {
  char b[128];  // ncclUniqueId (nccl.h: NCCL_UNIQUE_ID_BYTES)
};
struct NcclApi {
  int (*GetUniqueId)(NcclUniqueIdBytes*);
  int (*CommInitRank)(void**, int, NcclUniqueIdBytes, int);
  int (*CommDestroy)(void*);
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t);
  const char* (*GetErrorString)(int);
};
static const NcclApi* nccl_api() {
  static NcclApi api;
  static int state = 0;  // 0 untried, 1 ok, -1 unavailable
  if (state == 0) {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (h) {
      api.GetUniqueId = (int (*)(NcclUniqueIdBytes*))dlsym(h, "ncclGetUniqueId");
      api.CommInitRank = (int (*)(void**, int, NcclUniqueIdBytes, int))dlsym(h, "ncclCommInitRank");
      api.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
      api.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(h, "ncclAllGather");
      api.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
    }
    state = (h && api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather) ? 1 : -1;
  }
  return state == 1 ? &api : nullptr;
}

static void comm_release(esvio_fe* fe) {
  if (fe->stream_g) cudaStreamSynchronize(fe->stream_g);
  if (fe->nccl_comm && fe->comm_owned && nccl_api()) nccl_api()->CommDestroy(fe->nccl_comm);
  fe->nccl_comm = nullptr;
  cudaFree(fe->gathered[0]);
  cudaFree(fe->gathered[1]);
  fe->gathered[0] = fe->gathered[1] = nullptr;
  if (fe->stream_g) cudaStreamDestroy(fe->stream_g);
  fe->stream_g = nullptr;
}

static void free_all(esvio_fe* fe) {
  if (!fe) return;
  cudaSetDevice(fe->dev);
  for (cudaStream_t st : {fe->stream_c, fe->stream_b, fe->stream_e, fe->stream_p, fe->stream_r, fe->stream_f,
                          fe->stream_t1, fe->stream_s[0], fe->stream_s[1], fe->stream})
    if (st) cudaStreamSynchronize(st);
  cudaFree(fe->sae);
  cudaFree(fe->lat);
  for (int i = 0; i < kNumPyr; ++i) cudaFree(fe->pyr[i]);
  for (int i = 0; i < 3; ++i) cudaFree(fe->aux[i][0]), cudaFree(fe->aux[i][1]);
  cudaFree(fe->clahe_lut);
  cudaFree(fe->clahe_minmax);
  cudaFree(fe->gftt_block);
  for (int i = 0; i < 2; ++i) cudaFree(fe->warp_xy[i][0]), cudaFree(fe->warp_xy[i][1]);
  for (int i = 0; i < kSlots; ++i) {
    cudaFree(fe->raw[i][0]);  // raw[i][1] is the second half of the same block
    cudaFree(fe->flags[i]);
    if (i == 0) cudaFree(fe->corner_plane);
    cudaFree(fe->cand[i]);
    cudaFree(fe->cand_cnt[i]);
    for (cudaEvent_t ev : {fe->c_done[i], fe->b_done[i], fe->k1_done[i], fe->p_done[i], fe->f_done[i],
                           fe->t1_done[i], fe->s_done[i], fe->x_ready[i], fe->cl_done[i], fe->k1l_done[i]})
      if (ev) cudaEventDestroy(ev);
    if (fe->h_result[i]) cudaFreeHost(fe->h_result[i]);
    for (int c = 0; c < 2; ++c)
      if (fe->h_stage[i][c]) cudaFreeHost(fe->h_stage[i][c]);
    if (fe->q_done[i]) cudaEventDestroy(fe->q_done[i]);
    if (fe->r_free[i]) cudaEventDestroy(fe->r_free[i]);
  }
  for (int b = 0; b < 2; ++b) event_stage_free(&fe->esb[b]);
  cudaFree(fe->tb.snap_pts);
  cudaFree(fe->tb.snap_ids);
  cudaFree(fe->tb.snap_hdr);
  cudaFree(fe->tb.st);
  cudaFree(fe->tb.right_pts);
  cudaFree(fe->tb.prev_pts);
  cudaFree(fe->tb.ids);
  cudaFree(fe->tb.st_fwd);
  cudaFree(fe->tb.result);
  cudaFree(fe->tb.rs);
  cudaFree((void*)fe->tb.rng_draws);
  cudaFree((void*)fe->tb.rs_idx_cache);
  cudaFree((void*)fe->tb.rs_natt_cache);
  cudaFree(fe->d_scratch_n);
  cudaFree(fe->d_scratch_p0);
  cudaFree(fe->d_scratch_st);
  for (int k = 0; k < kSlots; ++k)
    for (int i = 0; i < ESVIO_FE_NUM_MARKS; ++i)
      if (fe->pev[k][i]) cudaEventDestroy(fe->pev[k][i]);
  if (fe->pev_ref) cudaEventDestroy(fe->pev_ref);
  comm_release(fe);
  for (cudaStream_t st : {fe->stream_c, fe->stream_b, fe->stream_e, fe->stream_p, fe->stream_r, fe->stream_f,
                          fe->stream_t1, fe->stream_s[0], fe->stream_s[1], fe->stream})
    if (st) cudaStreamDestroy(st);
  free(fe);
}

static int reset_state(esvio_fe* fe) {
  const size_t plane = fe->npx * 2 * sizeof(double2);
  CU(cudaMemsetAsync(fe->sae, 0, plane, fe->stream));
  CU(cudaMemsetAsync(fe->lat, 0, plane, fe->stream));
  for (int i = 0; i < kNumPyr; ++i) CU(cudaMemsetAsync(fe->pyr[i], 0, fe->pd.bytes, fe->stream));
  CU(cudaMemsetAsync(fe->tb.snap_hdr, 0, sizeof(int) * 16 * kSlots, fe->stream));
  CU(cudaMemsetAsync(fe->tb.st, 0, sizeof(TrackState), fe->stream));
  CU(cudaMemsetAsync(fe->tb.result, 0, fe->result_words * 4 * kSlots, fe->stream));
  for (int b = 0; b < 2; ++b)
    if (fe->esb[b].bin_total) event_stage_clear(fe->bl, fe->esb[b], fe->stream);
  CU(cudaStreamSynchronize(fe->stream));
  fe->cur_left = fe->prev_left = 0;
  fe->cur_right = kRightBase;
  fe->ts_sel[0] = fe->ts_sel[1] = nullptr;
  fe->windows = 0;
  fe->prev_time = 0.0;
  fe->q_head = fe->q_count = 0;
  fe->last_slot = -1;
  fe->f_pending = -1;
  fe->shard_seq = 0;
  for (int k = 0; k < kSlots; ++k) fe->pev_valid[k] = 0, fe->r_held[k] = 0;
  fe->stage_ms_valid = 0;
  return ESVIO_FE_OK;
}

// ------------------------------------------------------------------------------------------
// SM partition (CUDA green contexts).  Inside the pipeline the kernels of the temporal chain --
// the one stage that is serial from window to window and so sets the period -- share every SM's
// issue slots, shared memory and L1 with the event stage and the stereo LKs of five other
// windows and run 1.3-1.5x slower than alone.  With ESVIO_T1_SMS=<n> (n a multiple of 8) the
// device's SMs are split into two green contexts: the temporal chain's stream runs on n SMs of
// its own, every other stream of the handle on the rest.  Same address space, same events; only
// the streams differ.  One pair of contexts per device and per n, shared by all handles.
// ------------------------------------------------------------------------------------------
static thread_local bool g_creating_group_member = false;  // groups keep the whole device
struct SmPartition {
  CUgreenCtx chain = nullptr, rest = nullptr;
  int n = 0;
  bool tried = false;
  CUresult (*stream_create)(CUstream*, CUgreenCtx, unsigned int, int) = nullptr;
};
static SmPartition g_partition[64];
static std::mutex g_partition_mutex;

static SmPartition* sm_partition(int dev, int n_chain) {
  if (dev < 0 || dev >= 64 || n_chain <= 0) return nullptr;
  std::lock_guard<std::mutex> lk(g_partition_mutex);
  SmPartition& P = g_partition[dev];
  if (P.tried) return (P.chain && P.n == n_chain) ? &P : nullptr;
  P.tried = true;
  auto entry = [](const char* name) -> void* {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      return nullptr;
    return fn;
  };
  auto dev_get = (CUresult (*)(CUdevice*, int))entry("cuDeviceGet");
  auto get_res = (CUresult (*)(CUdevice, CUdevResource*, CUdevResourceType))entry("cuDeviceGetDevResource");
  auto split = (CUresult (*)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int,
                             unsigned int))entry("cuDevSmResourceSplitByCount");
  auto gen = (CUresult (*)(CUdevResourceDesc*, CUdevResource*, unsigned int))entry("cuDevResourceGenerateDesc");
  auto create = (CUresult (*)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int))entry("cuGreenCtxCreate");
  P.stream_create = (CUresult (*)(CUstream*, CUgreenCtx, unsigned int, int))entry("cuGreenCtxStreamCreate");
  if (!dev_get || !get_res || !split || !gen || !create || !P.stream_create) return nullptr;
  CUdevice cd;
  CUdevResource all, part, rem;
  unsigned int n_groups = 1;
  CUdevResourceDesc d_chain, d_rest;
  if (dev_get(&cd, dev) != CUDA_SUCCESS || get_res(cd, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) return nullptr;
  if (split(&part, &n_groups, &all, &rem, 0, (unsigned)n_chain) != CUDA_SUCCESS || n_groups != 1) return nullptr;
  if (gen(&d_chain, &part, 1) != CUDA_SUCCESS || gen(&d_rest, &rem, 1) != CUDA_SUCCESS) return nullptr;
  CUgreenCtx a = nullptr, b = nullptr;
  if (create(&a, d_chain, cd, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return nullptr;
  if (create(&b, d_rest, cd, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return nullptr;
  P.chain = a, P.rest = b, P.n = n_chain;
  if (getenv("ESVIO_FE_VERBOSE"))
    fprintf(stderr, "esvio_fe: SM partition on device %d: %u SMs for the temporal chain, %u for the rest\n", dev,
            part.sm.smCount, rem.sm.smCount);
  return &P;
}

// a non-blocking stream of the given priority: in the partition's green context when there is one
static cudaError_t make_stream(cudaStream_t* out, SmPartition* P, bool chain, int priority) {
  if (P) {
    CUstream st = nullptr;
    if (P->stream_create(&st, chain ? P->chain : P->rest, CU_STREAM_NON_BLOCKING, priority) == CUDA_SUCCESS) {
      *out = st;
      return cudaSuccess;
    }
    return cudaErrorUnknown;
  }
  return cudaStreamCreateWithPriority(out, cudaStreamNonBlocking, priority);
}

FE_API int esvio_fe_create(const esvio_fe_config* cfg, esvio_fe** out) {
  if (!cfg || !out) return ESVIO_FE_EINVAL;
  *out = nullptr;
  if (cfg->width < 64 || cfg->height < 64 || cfg->width > 8192 || cfg->height > 8192)
    return ESVIO_FE_EINVAL;
  if (cfg->max_cnt < 1 || cfg->max_cnt > kMaxCnt) return ESVIO_FE_EINVAL;
  if (cfg->min_dist < 1 || cfg->min_dist > 64) return ESVIO_FE_EINVAL;
  if (cfg->median_blur_kernel_size < 0 || cfg->median_blur_kernel_size > 7) return ESVIO_FE_EINVAL;
  if (cfg->max_events_per_window < 1) return ESVIO_FE_EINVAL;
  if (!(cfg->decay_ms > 0.0)) return ESVIO_FE_EINVAL;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || cfg->device_id < 0 ||
      cfg->device_id >= ndev) {
    cudaGetLastError();
    return ESVIO_FE_ENODEV;
  }
  esvio_fe* fe = (esvio_fe*)calloc(1, sizeof(esvio_fe));
  if (!fe) return ESVIO_FE_EINVAL;
  fe->cfg = *cfg;
  fe->dev = cfg->device_id;
  fe->W = cfg->width;
  fe->H = cfg->height;
  fe->npx = (size_t)fe->W * fe->H;
  int rc = ESVIO_FE_OK;
#define CUC(call)                                              \
  do {                                                         \
    cudaError_t ce_ = (call);                                  \
    if (ce_ != cudaSuccess) {                                  \
      fprintf(stderr, "esvio_fe_create: %s: %s\n", #call, cudaGetErrorString(ce_)); \
      free_all(fe);                                            \
      return ce_ == cudaErrorNoDevice ? ESVIO_FE_ENODEV : ESVIO_FE_ECUDA; \
    }                                                          \
  } while (0)
  CUC(cudaSetDevice(fe->dev));
  cudaDeviceProp prop;
  CUC(cudaGetDeviceProperties(&prop, fe->dev));
  if (prop.major < 10) {
    fprintf(stderr, "esvio_fe_create: device sm_%d%d is not Blackwell (sm_100a build)\n",
            prop.major, prop.minor);
    free_all(fe);
    return ESVIO_FE_ENODEV;
  }
  {
    int prio_lo = 0, prio_hi = 0;
    CUC(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    static const int chain_sms = getenv("ESVIO_T1_SMS") ? atoi(getenv("ESVIO_T1_SMS")) : 0;
    SmPartition* part = g_creating_group_member ? nullptr : sm_partition(fe->dev, chain_sms);
    CUC(make_stream(&fe->stream, part, false, prio_lo));
    // the event-stage kernels are short, wide and latency-bound: let their CTAs go first when
    // the long one-CTA-per-point LK kernels of other windows are also pending
    CUC(make_stream(&fe->stream_b, part, false, prio_hi));
    CUC(make_stream(&fe->stream_e, part, false, prio_hi));
    CUC(make_stream(&fe->stream_p, part, false, prio_hi));
    // the right camera's event stage of a left-first window: behind the temporal LK's CTAs
    CUC(make_stream(&fe->stream_r, part, false, prio_lo));
    CUC(make_stream(&fe->stream_f, part, false, prio_hi));
    // The temporal chain is the one stage that is serial from window to window (it sets the
    // period): its CTAs go first whenever an SM has room, ahead of the stereo LKs of the two
    // windows before.  ESVIO_T1_PRIO=0 (experiments): default priority, as before.
    static const int t1_hi = getenv("ESVIO_T1_PRIO") ? atoi(getenv("ESVIO_T1_PRIO")) : 1;
    CUC(make_stream(&fe->stream_t1, part, true, t1_hi ? prio_hi : prio_lo));
    CUC(make_stream(&fe->stream_s[0], part, false, prio_lo));
    CUC(make_stream(&fe->stream_s[1], part, false, prio_lo));
  }
  CUC(cudaStreamCreateWithFlags(&fe->stream_c, cudaStreamNonBlocking));

  fe->esb[0].n_cams = fe->esb[1].n_cams = 2;
  BinLayout& L = fe->bl;
  L.W = fe->W;
  L.H = fe->H;
  L.tiles_x = (fe->W + kTileW - 1) / kTileW;
  L.tiles_y = (fe->H + kTileH - 1) / kTileH;
  L.n_tiles = L.tiles_x * L.tiles_y;
  L.n_bins = L.n_tiles * kFine;
  fe->cap = (int)align_up((size_t)cfg->max_events_per_window, 64);
  L.max_chunks = (fe->cap + kChunk - 1) / kChunk;

  CUC(cudaMalloc(&fe->sae, fe->npx * 2 * sizeof(double2)));
  CUC(cudaMalloc(&fe->lat, fe->npx * 2 * sizeof(double2)));
  build_pyr_desc(fe->W, fe->H, &fe->pd);
  for (int i = 0; i < kNumPyr; ++i) CUC(cudaMalloc(&fe->pyr[i], fe->pd.bytes));
  if (cfg->equalize || cfg->median_blur_kernel_size) {
    for (int i = 0; i < 3; ++i)
      for (int c = 0; c < 2; ++c) {
        CUC(cudaMalloc(&fe->aux[i][c], fe->pd.bytes));
        CUC(cudaMemset(fe->aux[i][c], 0, fe->pd.bytes));
      }
    CUC(cudaMalloc(&fe->clahe_lut, clahe_lut_bytes()));
    CUC(cudaMalloc(&fe->clahe_minmax, 4 * sizeof(int)));
  }
  for (int c = 0; c < kSlots; ++c) {
    // both cameras' raw events of a slot in ONE block: a window staged as one host block
    // (esvio_fe_soa_layout_stereo) crosses PCIe as one copy
    CUC(cudaMalloc(&fe->raw[c][0], (size_t)fe->cap * 32));
    fe->raw[c][1] = fe->raw[c][0] + (size_t)fe->cap * 16;
    CUC(cudaMalloc(&fe->flags[c], (size_t)fe->cap + 16));
    if (c == 0) CUC(cudaMalloc(&fe->corner_plane, fe->npx));
    CUC(cudaMalloc(&fe->cand[c], ((size_t)fe->cap + kCornerBlock) * sizeof(uint32_t)));
    CUC(cudaMalloc(&fe->cand_cnt[c], ((size_t)fe->cap / kCornerBlock + 2) * sizeof(int)));
    for (cudaEvent_t* ev : {&fe->c_done[c], &fe->b_done[c], &fe->k1_done[c], &fe->p_done[c], &fe->f_done[c],
                            &fe->t1_done[c], &fe->s_done[c], &fe->x_ready[c], &fe->cl_done[c], &fe->k1l_done[c]})
      CUC(cudaEventCreateWithFlags(ev, cudaEventDisableTiming));
  }
  if (cfg->do_motion_correction)
    for (int i = 0; i < 2; ++i)
      for (int c = 0; c < 2; ++c) CUC(cudaMalloc(&fe->warp_xy[i][c], (size_t)fe->cap * sizeof(uint16_t)));
  for (int b = 0; b < 2; ++b)
    CUC(event_stage_alloc(L, 2, fe->cap, &fe->esb[b]) == 0 ? cudaSuccess : cudaErrorMemoryAllocation);

  const int M = cfg->max_cnt;
  TrackBuffers& B = fe->tb;
  CUC(cudaMalloc(&B.st, sizeof(TrackState)));
  float2* f2 = nullptr;
  CUC(cudaMalloc(&f2, sizeof(float2) * (size_t)M * 5));
  CUC(cudaMemset(f2, 0, sizeof(float2) * (size_t)M * 5));
  B.prev_pts = f2;
  B.cur_pts = f2 + M;
  B.rev_pts = f2 + 2 * M;
  B.prev_un = f2 + 3 * M;
  B.prev_un_r = f2 + 4 * M;
  CUC(cudaMalloc(&B.right_pts, (sizeof(float2) * 2 + 2) * (size_t)M * kSlots));
  CUC(cudaMemset(B.right_pts, 0, (sizeof(float2) * 2 + 2) * (size_t)M * kSlots));
  B.rev_left_pts = B.right_pts + (size_t)M * kSlots;
  B.st_sf = reinterpret_cast<uint8_t*>(B.rev_left_pts + (size_t)M * kSlots);
  B.st_sb = B.st_sf + (size_t)M * kSlots;
  int* i4 = nullptr;
  CUC(cudaMalloc(&i4, sizeof(int) * (size_t)M * 4));
  CUC(cudaMemset(i4, 0, sizeof(int) * (size_t)M * 4));
  B.ids = i4;
  B.cnt = i4 + M;
  B.prev_un_ids = i4 + 2 * M;
  B.prev_un_r_ids = i4 + 3 * M;
  uint8_t* u4 = nullptr;
  CUC(cudaMalloc(&u4, (size_t)M * 2));
  CUC(cudaMemset(u4, 0, (size_t)M * 2));
  B.st_fwd = u4;
  B.st_bwd = u4 + M;
  CUC(cudaMalloc(&B.snap_pts, sizeof(float2) * (size_t)M * kSlots));
  CUC(cudaMalloc(&B.snap_ids, sizeof(int) * (size_t)M * kSlots * 2));
  B.snap_cnt = B.snap_ids + (size_t)M * kSlots;
  CUC(cudaMalloc(&B.snap_hdr, sizeof(int) * 16 * kSlots));
  fe->result_words = kResultHdr + (size_t)kResultArrays * M;
  B.result_words = (int)fe->result_words;
  CUC(cudaMalloc(&B.result, fe->result_words * 4 * kSlots));
  for (int i = 0; i < kSlots; ++i) {
    CUC(cudaHostAlloc(&fe->h_result[i], fe->result_words * 4, cudaHostAllocDefault));
    CUC(cudaEventCreateWithFlags(&fe->q_done[i], cudaEventDisableTiming));
    CUC(cudaEventCreateWithFlags(&fe->r_free[i], cudaEventDisableTiming));
  }
  CUC(cudaMalloc(&B.rs, ransac_scratch_bytes()));
  CUC(cudaMemset(B.rs, 0, ransac_scratch_bytes()));
  {
    const int nd = ransac_num_draws();
    uint32_t* tab = (uint32_t*)malloc(sizeof(uint32_t) * nd);
    ransac_fill_draw_table(tab);
    uint32_t* dtab = nullptr;
    cudaError_t ce = cudaMalloc(&dtab, sizeof(uint32_t) * nd);
    if (ce == cudaSuccess) ce = cudaMemcpy(dtab, tab, sizeof(uint32_t) * nd, cudaMemcpyHostToDevice);
    free(tab);
    B.rng_draws = dtab;
    CUC(ce);
  }
  CUC(cudaMalloc(&fe->d_scratch_n, 64));
  CUC(cudaMalloc(&fe->d_scratch_p0, sizeof(float2) * 2 * (size_t)kMaxCnt));
  fe->d_scratch_p1 = fe->d_scratch_p0 + kMaxCnt;
  CUC(cudaMalloc(&fe->d_scratch_st, kMaxCnt));
  for (int k = 0; k < kSlots; ++k)
    for (int i = 0; i < ESVIO_FE_NUM_MARKS; ++i) CUC(cudaEventCreate(&fe->pev[k][i]));
  CUC(cudaEventCreate(&fe->pev_ref));
#undef CUC

  if (cfg->mc_fx > 0.0) {
    fe->mc_K[0] = (float)cfg->mc_fx, fe->mc_K[1] = (float)cfg->mc_fy;
    fe->mc_K[2] = (float)cfg->mc_cx, fe->mc_K[3] = (float)cfg->mc_cy;
  } else {
    fe->mc_K[0] = (float)cfg->cam[1].fx, fe->mc_K[1] = (float)cfg->cam[1].fy;
    fe->mc_K[2] = (float)(fe->W / 2), fe->mc_K[3] = (float)(fe->H / 2);
  }
  TrackParams& P = fe->tp;
  P.W = fe->W;
  P.H = fe->H;
  P.max_cnt = M;
  P.min_dist = cfg->min_dist;
  P.flow_back = cfg->flow_back;
  P.focal_length = cfg->focal_length;
  P.f_threshold = cfg->f_threshold;
  for (int c = 0; c < 2; ++c) {
    const esvio_pinhole& s = cfg->cam[c];
    P.cam[c] = Pinhole{s.fx, s.fy, s.cx, s.cy, s.k1, s.k2, s.p1, s.p2};
  }
  {
    static const int lf = getenv("ESVIO_LEFT_FIRST") ? atoi(getenv("ESVIO_LEFT_FIRST")) : 1;
    fe->left_first = lf;
  }
  if ((rc = make_state_map(fe, fe->sae, &fe->map_sae)) != ESVIO_FE_OK ||
      (rc = make_state_map(fe, fe->lat, &fe->map_lat)) != ESVIO_FE_OK ||
      (rc = make_state_map(fe, fe->sae + fe->npx, &fe->map_sae_r, 1)) != ESVIO_FE_OK ||
      (rc = make_state_map(fe, fe->lat + fe->npx, &fe->map_lat_r, 1)) != ESVIO_FE_OK) {
    fprintf(stderr, "esvio_fe_create: %s\n", fe->err);
    free_all(fe);
    return rc;
  }
  prefer_shared_lk();
  prefer_shared_events();
  cudaGetLastError();
  if (select_configure(fe->W, fe->H) != 0 || bin_configure(L) != 0) {
    fprintf(stderr, "esvio_fe_create: sensor too large for the shared-memory mask / histogram\n");
    free_all(fe);
    return ESVIO_FE_EINVAL;
  }
  if ((rc = reset_state(fe)) != ESVIO_FE_OK) {
    fprintf(stderr, "esvio_fe_create: %s\n", fe->err);
    free_all(fe);
    return rc;
  }
  if (cfg->use_ransac && M >= 8) {
    // the sample indices of every RANSAC attempt, for every track count this handle can see
    uint16_t* c_idx = nullptr;
    int* c_natt = nullptr;
    cudaError_t ce = cudaMalloc(&c_idx, ransac_cache_idx_bytes(8, M));
    if (ce == cudaSuccess) ce = cudaMalloc(&c_natt, sizeof(int) * (size_t)(M - 8 + 1));
    if (ce == cudaSuccess) ce = cudaMemsetAsync(fe->d_scratch_p0, 0, sizeof(float2) * 2 * (size_t)kMaxCnt, fe->stream);
    fe->tb.rs_idx_cache = c_idx;
    fe->tb.rs_natt_cache = c_natt;
    fe->tb.rs_cache_lo = 8;
    if (ce == cudaSuccess &&
        ransac_build_cache(fe->tp, fe->tb, fe->d_scratch_p0, 8, M, c_idx, c_natt, fe->stream) != 0)
      ce = cudaErrorUnknown;
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(fe->stream);
    if (ce != cudaSuccess) {
      fprintf(stderr, "esvio_fe_create: RANSAC sample cache: %s\n", cudaGetErrorString(ce));
      free_all(fe);
      return ESVIO_FE_ECUDA;
    }
  }
  *out = fe;
  return ESVIO_FE_OK;
}

FE_API void esvio_fe_destroy(esvio_fe* fe) {
  if (fe && fe->group) return;  // owned by its group: esvio_fe_group_destroy
  free_all(fe);
}

static int sync_all(esvio_fe* fe) {
  for (cudaStream_t st : {fe->stream_c, fe->stream_b, fe->stream_e, fe->stream_p, fe->stream_r, fe->stream_f,
                          fe->stream_t1, fe->stream_s[0], fe->stream_s[1], fe->stream})
    CU(cudaStreamSynchronize(st));
  return ESVIO_FE_OK;
}

FE_API int esvio_fe_reset(esvio_fe* fe) {
  if (!fe) return ESVIO_FE_EINVAL;
  if (fe->group) return fail(fe, ESVIO_FE_ESTATE, "handle belongs to a group: use esvio_fe_group_reset", cudaSuccess);
  CU(cudaSetDevice(fe->dev));
  if (sync_all(fe) != ESVIO_FE_OK) return ESVIO_FE_ECUDA;
  return reset_state(fe);
}

// ---------------------------------------------------------------------------------------
// event staging
// ---------------------------------------------------------------------------------------
static int stage_events(esvio_fe* fe, int slot, int cam, const esvio_events* e, DevEvents* d) {
  memset(d, 0, sizeof(*d));
  if (!e || e->n == 0) return ESVIO_FE_OK;
  if (e->n > (size_t)fe->cap) return fail(fe, ESVIO_FE_ECAPACITY, "events > max_events_per_window", cudaSuccess);
  const bool soa = e->x && e->y && e->t && e->p;
  if (!soa && !e->aos) return fail(fe, ESVIO_FE_EINVAL, "events: need x,y,t,p or aos", cudaSuccess);
  d->n = (int)e->n;
  if (e->on_device) {
    if (e->aos) d->aos = (const uint4*)e->aos;
    else d->x = e->x, d->y = e->y, d->t = e->t, d->p = e->p;
    return ESVIO_FE_OK;
  }
  uint8_t* raw = fe->raw[slot][cam];
  cudaStream_t se = fe->stream_c;
  const size_t n = e->n, cap = (size_t)fe->cap;
  // host -> raw + off: pinned memory goes straight to the DMA engine, large pageable buffers
  // through the handle's pinned ring on the staging threads (HostStager)
  auto h2d = [&](size_t off, const void* src, size_t bytes) -> int {
    if (bytes >= (128u << 10) && HostStager::get().helpers() > 0 && is_pageable(src)) {
      if (!fe->h_stage[slot][cam])
        CU(cudaHostAlloc(&fe->h_stage[slot][cam], cap * 16, cudaHostAllocDefault));
      CU(HostStager::get().run(fe->h_stage[slot][cam] + off, raw + off, (const uint8_t*)src, bytes, se));
    } else {
      CU(cudaMemcpyAsync(raw + off, src, bytes, cudaMemcpyHostToDevice, se));
    }
    return ESVIO_FE_OK;
  };
  int rc;
  if (e->aos) {
    if ((rc = h2d(0, e->aos, n * 16)) != ESVIO_FE_OK) return rc;
    d->aos = (const uint4*)raw;
  } else {
    size_t off[4], total;
    esvio_fe_soa_layout(n, off, &total);
    const uint8_t* hx = (const uint8_t*)e->x;
    if ((const uint8_t*)e->y == hx + off[1] && (const uint8_t*)e->t == hx + off[2] &&
        (const uint8_t*)e->p == hx + off[3]) {
      // the four arrays sit in one block laid out by esvio_fe_soa_layout: one copy
      if ((rc = h2d(0, hx, total)) != ESVIO_FE_OK) return rc;
      d->x = (uint16_t*)raw, d->y = (uint16_t*)(raw + off[1]), d->t = (double*)(raw + off[2]),
      d->p = raw + off[3];
    } else {
      uint16_t* dx = (uint16_t*)raw;
      uint16_t* dy = (uint16_t*)(raw + 2 * cap);
      double* dt = (double*)(raw + 4 * cap);
      uint8_t* dp = raw + 12 * cap;
      if ((rc = h2d(0, e->x, n * 2)) != ESVIO_FE_OK) return rc;
      if ((rc = h2d(2 * cap, e->y, n * 2)) != ESVIO_FE_OK) return rc;
      if ((rc = h2d(4 * cap, e->t, n * 8)) != ESVIO_FE_OK) return rc;
      if ((rc = h2d(12 * cap, e->p, n)) != ESVIO_FE_OK) return rc;
      d->x = dx, d->y = dy, d->t = dt, d->p = dp;
    }
  }
  return ESVIO_FE_OK;
}

// Both cameras of a window in one pinned host block laid out by esvio_fe_soa_layout_stereo: ONE
// copy (two 2.2 MB copies back to back reach 45 GB/s on the B200 boxes, one 4.3 MB copy 50 GB/s:
// 96 vs 86 us per 640x480 window at 5 Mev/s per camera, which is what bounds the end-to-end rate).
// Returns 1 when it took the window, 0 when the caller has to stage the cameras one by one.
static int stage_events_stereo_block(esvio_fe* fe, int slot, const esvio_events* l, const esvio_events* r,
                                     DevEvents* d, int* took, cudaEvent_t after_left = nullptr) {
  *took = 0;
  if (!l || !r || !(l->flags & r->flags & ESVIO_EVENTS_STEREO_BLOCK)) return ESVIO_FE_OK;  // the caller's promise
  if (l->n == 0 || r->n == 0 || l->on_device || r->on_device || l->aos || r->aos) return ESVIO_FE_OK;
  if (!(l->x && l->y && l->t && l->p && r->x && r->y && r->t && r->p)) return ESVIO_FE_OK;
  if (l->n > (size_t)fe->cap || r->n > (size_t)fe->cap) return ESVIO_FE_OK;  // stage_events reports it
  size_t ol[4], orr[4], total;
  esvio_fe_soa_layout_stereo(l->n, r->n, ol, orr, &total);
  const uint8_t* h = (const uint8_t*)l->x;
  if ((const uint8_t*)l->y != h + ol[1] || (const uint8_t*)l->t != h + ol[2] || (const uint8_t*)l->p != h + ol[3] ||
      (const uint8_t*)r->x != h + orr[0] || (const uint8_t*)r->y != h + orr[1] ||
      (const uint8_t*)r->t != h + orr[2] || (const uint8_t*)r->p != h + orr[3])
    return ESVIO_FE_OK;
  if (total > (size_t)fe->cap * 32 || is_pageable(h)) return ESVIO_FE_OK;
  uint8_t* raw = fe->raw[slot][0];
  if (after_left) {  // left-first window: the left camera's arrays cross first and are marked
    CU(cudaMemcpyAsync(raw, h, orr[0], cudaMemcpyHostToDevice, fe->stream_c));
    CU(cudaEventRecord(after_left, fe->stream_c));
    CU(cudaMemcpyAsync(raw + orr[0], h + orr[0], total - orr[0], cudaMemcpyHostToDevice, fe->stream_c));
  } else {
    CU(cudaMemcpyAsync(raw, h, total, cudaMemcpyHostToDevice, fe->stream_c));
  }
  const size_t* off[2] = {ol, orr};
  const size_t n[2] = {l->n, r->n};
  for (int c = 0; c < 2; ++c) {
    memset(&d[c], 0, sizeof(d[c]));
    d[c].n = (int)n[c];
    d[c].x = (uint16_t*)(raw + off[c][0]), d[c].y = (uint16_t*)(raw + off[c][1]);
    d[c].t = (double*)(raw + off[c][2]), d[c].p = raw + off[c][3];
  }
  *took = 1;
  return ESVIO_FE_OK;
}

// the copies of a window's events were enqueued on the copy stream: mark where they end
static int staging_done(esvio_fe* fe, int slot) {
  CU(cudaEventRecord(fe->c_done[slot], fe->stream_c));
  return ESVIO_FE_OK;
}

// profiling markers (include/esvio_fe.h, ESVIO_FE_NUM_MARKS), each recorded on the stream whose
// progress it reports
enum { kMarkSubmit = 0, kMarkLanded = 1, kMarkK1Start = 2, kMarkK1Done = 3, kMarkPyrDone = 4,
       kMarkFlagsDone = 5, kMarkTemporalLk = 6, kMarkSelect = 7, kMarkPacked = 8, kMarkResult = 9,
       kMarkTemporalStart = 10, kMarkStereoStart = 11, kMarkBinned = 12 };
static void prof_mark(esvio_fe* fe, int i, cudaStream_t st) {
  if (!fe->profiling) return;
  cudaEventRecord(fe->pev[fe->pev_slot][i], st);
}

// createSAE_* + SAEtoTimeSurface_* + pyramids (feature_tracker.cpp:356-368) into the
// pyramid buffers `left_idx` / right
static McParams mc_params(const esvio_fe* fe, const esvio_motion* mc) {
  McParams p;
  for (int i = 0; i < 3; ++i) {
    p.v_cur[i] = (float)mc->state_v[i];  // temp_v[i] = State[i] (event_detector.cc:113-116)
    p.v_pre[i] = mc->v_pre[i];
    p.omega[i] = mc->omega[i];
  }
  for (int i = 0; i < 4; ++i) p.K[i] = fe->mc_K[i];
  p.t1 = mc->t1;
  p.W = fe->W;
  p.H = fe->H;
  return p;
}

// sqrt(pow(a0,2) + pow(a1,2) + pow(a2,2)) > a_motion_compensation_threshold
// (event_detector.cc:125, event_detector.h:51), in double like std::pow(float, int)
static bool mc_active(const esvio_motion* mc) {
  const double a0 = mc->accel[0], a1 = mc->accel[1], a2 = mc->accel[2];
  return sqrt(a0 * a0 + a1 * a1 + a2 * a2) > 5.0;
}

// createSAE_* + SAEtoTimeSurface_* + pyramids (feature_tracker.cpp:356-368) of the window in
// `slot` into the pyramid buffers `left_idx` / `right_idx`: binning on stream_b behind the
// window's copies, K1 (+ conditioning) and the pyramids on stream_e; records b_done,
// k1_done and p_done of the slot.  n_cams = 1 (left/right split over two GPUs): only camera 0
// of this handle -- ev_in[0], state plane 0, image `left_idx` -- is processed.
static int run_event_stage(esvio_fe* fe, int slot, double t_ref, const DevEvents ev_in[2], int left_idx,
                           int right_idx, const esvio_motion* mc = nullptr, int n_cams = 2) {
  cudaStream_t sb = fe->stream_b, se = fe->stream_e, s_pyr = fe->stream_e;
  DevEvents ev[2] = {ev_in[0], n_cams == 2 ? ev_in[1] : DevEvents{}};
  // ---- K0 on stream_b.  The binned-event buffers alternate between even and odd slots; the
  // set of this slot was last read by the K1 two windows back.
  EventStageBuffers esb = fe->esb[slot & 1];
  esb.n_cams = n_cams;
  CU(cudaStreamWaitEvent(sb, fe->c_done[slot], 0));
  CU(cudaStreamWaitEvent(sb, fe->k1_done[(slot + kSlots - 2) % kSlots], 0));
  if (mc && n_cams == 2 && mc_active(mc) && ev[0].n > 0) {
    launch_warp_events(mc_params(fe, mc), ev_in, fe->warp_xy[0], fe->warp_xy[1], sb, &fe->launches);
    for (int c = 0; c < 2; ++c) ev[c].wx = fe->warp_xy[0][c], ev[c].wy = fe->warp_xy[1][c];
  }
  launch_bin_events(fe->bl, esb, ev, sb, &fe->launches);
  prof_mark(fe, kMarkBinned, sb);
  CU(cudaEventRecord(fe->b_done[slot], sb));
  // ---- K1 on stream_e: behind the binning, and behind the corner flags of the window before,
  // which read the SAE state this launch rewrites
  CU(cudaStreamWaitEvent(se, fe->b_done[slot], 0));
  if (fe->f_pending >= 0) {
    CU(cudaStreamWaitEvent(se, fe->f_done[fe->f_pending], 0));
    fe->f_pending = -1;
  }
  prof_mark(fe, kMarkK1Start, se);
  SaeTsParams sp;
  sp.W = fe->W;
  sp.H = fe->H;
  sp.tiles_x = fe->bl.tiles_x;
  sp.n_tiles = fe->bl.n_tiles;
  sp.n_cams = n_cams;
  for (int c = 0; c < kMaxCams; ++c) {
    sp.t_ref[c] = t_ref;
    sp.bt[c] = nullptr, sp.bk[c] = nullptr, sp.ts[c] = nullptr;
  }
  sp.decay_sec = fe->cfg.decay_ms / 1000.0;
  sp.inv_decay = 1.0 / sp.decay_sec;
  sp.filter_threshold = fe->cfg.feature_filter_threshold;
  sp.ignore_polarity = fe->cfg.ignore_polarity;
  sp.bin_start = esb.bin_start;
  sp.bt[0] = esb.bt[0];
  sp.bt[1] = esb.bt[1];
  sp.bk[0] = esb.bk[0];
  sp.bk[1] = esb.bk[1];
  // time surface -> [median blur] -> (selection / gettimesurface see this) -> [CLAHE +
  // normalize] -> pyramid level 0 (event_detector.cc:260-264, feature_tracker.cpp:370-388)
  uint8_t* imgs[2] = {fe->pyr[left_idx], fe->pyr[right_idx]};
  const int med = fe->cfg.median_blur_kernel_size, eq = fe->cfg.equalize;
  uint8_t* const* a_out = (med || eq) ? fe->aux[0] : imgs;                 // K1 writes here
  uint8_t* const* b_out = med ? (eq ? fe->aux[1] : imgs) : a_out;          // after the median
  sp.ts[0] = a_out[0];
  sp.ts[1] = a_out[1];
  sp.ts_pitch = fe->pd.pitch[0];
  launch_sae_update_ts(sp, fe->map_sae, fe->map_lat, se, &fe->launches);
  if (med) {
    const uint8_t* src[2] = {a_out[0], a_out[1]};
    launch_median(src, b_out, n_cams, fe->W, fe->H, fe->pd.pitch[0], 2 * med + 1, se, &fe->launches);
  }
  fe->ts_sel[0] = b_out[0];
  if (n_cams == 2) fe->ts_sel[1] = b_out[1];
  if (eq) {
    const uint8_t* src[2] = {b_out[0], b_out[1]};
    launch_equalize(src, fe->aux[2], imgs, n_cams, fe->W, fe->H, fe->pd.pitch[0], fe->clahe_lut,
                    fe->clahe_minmax, se, &fe->launches);
  }
  prof_mark(fe, kMarkK1Done, se);
  CU(cudaEventRecord(fe->k1_done[slot], se));
  // ---- pyramids right behind K1 on the same stream (a programmatic dependent launch instead
  // of an event hop to another stream: the temporal LK of a synchronous call starts ~8 us
  // earlier; the next window's K1 queues 9 us later, which the pipeline does not notice --
  // its period is set by the temporal chain)
  launch_pyramids(fe->pd, imgs, n_cams, s_pyr, &fe->launches);
  prof_mark(fe, kMarkPyrDone, s_pyr);
  CU(cudaEventRecord(fe->p_done[slot], s_pyr));
  CU(cudaGetLastError());
  return ESVIO_FE_OK;
}

// ---- left-first windows -------------------------------------------------------------------------
// A window submitted while nothing else of the handle is in flight -- the synchronous call the
// reference node makes (stereo_event_tracker_node.cpp:193) -- has nothing to overlap with but
// itself.  The temporal chain (temporal LK, filter, F-RANSAC, selection) needs the LEFT camera
// only (feature_tracker.cpp:405-468); the right image is not read before the stereo LK (:490).
// So the left camera's events cross PCIe first and its binning, SAE update, time surface and
// pyramid run as launches of their own, while the right camera's copy and event stage follow
// on another stream next to the temporal LK.  Same kernels, same per-camera arithmetic, same
// results; per-camera views of the two-camera buffers (event_stage_cam_view), a tensor map on
// the right camera's planes.  Not taken with motion compensation, median blur or CLAHE (their
// launches cover both cameras), nor while the per-stage events are on (they time the
// two-camera launches).  ESVIO_LEFT_FIRST=0 switches it off.
static int run_event_stage_cam(esvio_fe* fe, int slot, double t_ref, const DevEvents& ev, int cam, int img_idx,
                               cudaStream_t sb, cudaStream_t se, cudaEvent_t copied, cudaEvent_t binned,
                               cudaEvent_t k1, cudaEvent_t pyr) {
  const EventStageBuffers esb = event_stage_cam_view(fe->bl, fe->esb[slot & 1], cam);
  CU(cudaStreamWaitEvent(sb, copied, 0));
  launch_bin_events(fe->bl, esb, &ev, sb, &fe->launches);
  if (cam == 0) prof_mark(fe, kMarkBinned, sb);
  if (sb != se) {
    CU(cudaEventRecord(binned, sb));
    CU(cudaStreamWaitEvent(se, binned, 0));
  }
  if (cam == 0 && fe->f_pending >= 0) {  // the corner flags of the window before read the left SAE state
    CU(cudaStreamWaitEvent(se, fe->f_done[fe->f_pending], 0));
    fe->f_pending = -1;
  }
  SaeTsParams sp;
  sp.W = fe->W;
  sp.H = fe->H;
  sp.tiles_x = fe->bl.tiles_x;
  sp.n_tiles = fe->bl.n_tiles;
  sp.n_cams = 1;
  for (int c = 0; c < kMaxCams; ++c) {
    sp.t_ref[c] = t_ref;
    sp.bt[c] = nullptr, sp.bk[c] = nullptr, sp.ts[c] = nullptr;
  }
  sp.decay_sec = fe->cfg.decay_ms / 1000.0;
  sp.inv_decay = 1.0 / sp.decay_sec;
  sp.filter_threshold = fe->cfg.feature_filter_threshold;
  sp.ignore_polarity = fe->cfg.ignore_polarity;
  sp.bin_start = esb.bin_start;
  sp.bt[0] = esb.bt[0];
  sp.bk[0] = esb.bk[0];
  sp.ts[0] = fe->pyr[img_idx];
  sp.ts_pitch = fe->pd.pitch[0];
  if (cam == 0) prof_mark(fe, kMarkK1Start, se);
  launch_sae_update_ts(sp, cam ? fe->map_sae_r : fe->map_sae, cam ? fe->map_lat_r : fe->map_lat, se, &fe->launches);
  fe->ts_sel[cam] = fe->pyr[img_idx];
  if (cam == 0) prof_mark(fe, kMarkK1Done, se);
  CU(cudaEventRecord(k1, se));
  uint8_t* imgs[1] = {fe->pyr[img_idx]};
  launch_pyramids(fe->pd, imgs, 1, se, &fe->launches);
  if (cam == 0) prof_mark(fe, kMarkPyrDone, se);
  CU(cudaEventRecord(pyr, se));
  CU(cudaGetLastError());
  return ESVIO_FE_OK;
}

static int run_event_stage_left_first(esvio_fe* fe, int slot, double t_ref, const DevEvents ev[2], int left_idx,
                                      int right_idx) {
  int rc;
  // left: one chain on stream_e (K0 too: nothing of an earlier window is there to overlap with, and a
  // launch behind its predecessor on the same stream starts earlier than one behind an event)
  if ((rc = run_event_stage_cam(fe, slot, t_ref, ev[0], 0, left_idx, fe->stream_e, fe->stream_e, fe->cl_done[slot],
                                nullptr, fe->k1l_done[slot], fe->p_done[slot])) != ESVIO_FE_OK)
    return rc;
  // right: one chain on stream_r behind the whole copy; x_ready[slot] = its pyramid (the stereo LK waits for it)
  if ((rc = run_event_stage_cam(fe, slot, t_ref, ev[1], 1, right_idx, fe->stream_r, fe->stream_r, fe->c_done[slot],
                                nullptr, fe->x_ready[slot], fe->x_ready[slot])) != ESVIO_FE_OK)
    return rc;
  // k1_done[slot] keeps its meaning for the windows behind (both cameras' state and binned-event
  // buffers are free again), and stream_e stays the one stream that orders the SAE updates
  CU(cudaStreamWaitEvent(fe->stream_e, fe->x_ready[slot], 0));
  CU(cudaEventRecord(fe->k1_done[slot], fe->stream_e));
  return ESVIO_FE_OK;
}

static CornerParams corner_params(esvio_fe* fe, int left_idx, int and_ts, int slot) {
  CornerParams cp;
  cp.W = fe->W;
  cp.H = fe->H;
  cp.min_dist = fe->cfg.min_dist;
  cp.filter_threshold = fe->cfg.feature_filter_threshold;
  cp.ts_lk_threshold = fe->cfg.ts_lk_threshold;
  cp.sae = fe->sae;
  cp.lat = fe->lat;
  cp.ts = fe->ts_sel[0] ? fe->ts_sel[0] : fe->pyr[left_idx];
  cp.ts_pitch = fe->pd.pitch[0];
  cp.and_ts_test = and_ts;
  cp.cand = fe->cand[slot];
  cp.cand_cnt = fe->cand_cnt[slot];
  cp.plane = fe->corner_plane;
  return cp;
}

FE_API int esvio_fe_track_submit(esvio_fe* fe, double cur_time, const esvio_events* left,
                                 const esvio_events* right, int32_t pub_this_frame) {
  return esvio_fe_track_submit_mc(fe, cur_time, left, right, pub_this_frame, nullptr);
}

// Which buffers a new window uses: its in-flight slot and the rotating pyramid images.
struct WindowPlan {
  int slot, cur, prev, rcur;
};

static int plan_window(esvio_fe* fe, WindowPlan* w) {
  if (fe->q_count >= kSlots)
    return fail(fe, ESVIO_FE_ESTATE, "too many windows in flight (esvio_fe_pipeline_depth)", cudaSuccess);
  w->slot = (fe->q_head + fe->q_count) % kSlots;
  w->cur = fe->windows == 0 ? 0 : (fe->cur_left + 1) % kLeftBufs;
  w->prev = fe->windows == 0 ? 0 : fe->cur_left;  // first window: prev_img = cur_img
  w->rcur = fe->windows == 0 ? kRightBase : kRightBase + (fe->cur_right - kRightBase + 1) % kRightBufs;
  return ESVIO_FE_OK;
}

// Everything of a window behind the SAE / time surface / pyramids.  `after_k1` / `after_pyr`:
// the events that mark the window's time surface + SAE state and its pyramids as ready (the
// handle's own k1_done / p_done, or its group's).  `wait_exchange`: the stereo LK also waits
// for x_ready[slot] (left/right split: the right image arrives on the caller's stream).
static int submit_tracking(esvio_fe* fe, const WindowPlan& w, const DevEvents& ev_left,
                           double cur_time, int32_t pub_this_frame, cudaEvent_t after_k1,
                           cudaEvent_t after_pyr, bool wait_exchange = false) {
  cudaStream_t sf = fe->stream_f, s1 = fe->stream_t1, ss = fe->stream_s[w.slot & 1], s2 = fe->stream;
  const int slot = w.slot, cur = w.cur, prev = w.prev, rcur = w.rcur;
  const TrackBuffers& B = fe->tb;
  const int M = fe->cfg.max_cnt;
  // ---------------- Arc* corner flags (feature_tracker.cpp:458 -> event_detector.cc:308), off the
  // LK path: selection is their only reader
  if (pub_this_frame) {
    CU(cudaStreamWaitEvent(sf, after_k1, 0));
    launch_corner_flags(corner_params(fe, cur, 1, slot), ev_left, fe->flags[slot], sf, &fe->launches);
    prof_mark(fe, kMarkFlagsDone, sf);
    CU(cudaEventRecord(fe->f_done[slot], sf));
    fe->f_pending = slot;
  } else {
    prof_mark(fe, kMarkFlagsDone, fe->stream_e);
  }

  // ---------------- temporal chain (feature_tracker.cpp:405-468), in order behind window k-1's
  CU(cudaStreamWaitEvent(s1, after_pyr, 0));
  prof_mark(fe, kMarkTemporalStart, s1);
  launch_lk(fe->pd, fe->pyr[prev], fe->pyr[cur], B.prev_pts, B.cur_pts, B.st_fwd, B.rev_pts,
            B.st_bwd, &B.st->n_prev, M, 3, 0, fe->cfg.flow_back ? 1 : 0, s1, &fe->launches);
  launch_post_temporal(fe->tp, B, pub_this_frame ? -1 : slot, s1, &fe->launches);
  prof_mark(fe, kMarkTemporalLk, s1);
  if (pub_this_frame) {
    if (fe->cfg.use_ransac) launch_ransac(fe->tp, B, s1, &fe->launches);
    CU(cudaStreamWaitEvent(s1, fe->f_done[slot], 0));
    launch_select(fe->tp, B, ev_left.n, fe->cand[slot], fe->cand_cnt[slot], slot, s1, &fe->launches);
  }
  prof_mark(fe, kMarkSelect, s1);
  CU(cudaEventRecord(fe->t1_done[slot], s1));

  // ---------------- stereo LK (feature_tracker.cpp:490,495) on the snapshot; consecutive windows
  // alternate between two streams, so their stereo LKs overlap
  CU(cudaStreamWaitEvent(ss, fe->t1_done[slot], 0));
  if (wait_exchange) CU(cudaStreamWaitEvent(ss, fe->x_ready[slot], 0));
  prof_mark(fe, kMarkStereoStart, ss);
  launch_lk(fe->pd, fe->pyr[cur], fe->pyr[rcur], B.snap_pts + (size_t)slot * M,
            B.right_pts + (size_t)slot * M, B.st_sf + (size_t)slot * M,
            B.rev_left_pts + (size_t)slot * M, B.st_sb + (size_t)slot * M, B.snap_hdr + slot * 16, M, 3,
            0, fe->cfg.flow_back ? 2 : 0, ss, &fe->launches);
  CU(cudaEventRecord(fe->s_done[slot], ss));

  // ---------------- stereo check, undistortion, velocities, packing (:470-473, 496-590): in
  // window order on the result stream (the velocity maps roll from window to window)
  CU(cudaStreamWaitEvent(s2, fe->s_done[slot], 0));
  if (fe->r_held[slot]) {  // a consumer stream may still read this slot's previous block
    CU(cudaStreamWaitEvent(s2, fe->r_free[slot], 0));
    fe->r_held[slot] = 0;
  }
  launch_finalize(fe->tp, B, slot, cur_time, fe->prev_time, s2, &fe->launches);
  prof_mark(fe, kMarkPacked, s2);
  CU(cudaMemcpyAsync(fe->h_result[slot], B.result + (size_t)slot * fe->result_words, fe->result_words * 4,
                     cudaMemcpyDeviceToHost, s2));
  prof_mark(fe, kMarkResult, s2);
  CU(cudaEventRecord(fe->q_done[slot], s2));
  CU(cudaGetLastError());
  fe->q_count++;
  fe->last_slot = slot;
  fe->prev_left = prev;
  fe->cur_left = cur;
  fe->cur_right = rcur;
  fe->windows++;
  fe->prev_time = cur_time;
  return ESVIO_FE_OK;
}

FE_API int esvio_fe_track_submit_mc(esvio_fe* fe, double cur_time, const esvio_events* left,
                                    const esvio_events* right, int32_t pub_this_frame,
                                    const esvio_motion* mc) {
  if (!fe) return ESVIO_FE_EINVAL;
  if (fe->group) return fail(fe, ESVIO_FE_ESTATE, "handle belongs to a group: use esvio_fe_group_track_submit", cudaSuccess);
  if (mc && !fe->cfg.do_motion_correction)
    return fail(fe, ESVIO_FE_ESTATE, "motion compensation needs config.do_motion_correction", cudaSuccess);
  CU(cudaSetDevice(fe->dev));
  WindowPlan w;
  int rc;
  if ((rc = plan_window(fe, &w)) != ESVIO_FE_OK) return rc;
  fe->pev_slot = w.slot;
  // Everything a later node of the window's graph reads from an earlier one is either per-slot
  // (raw events, flags, snapshot, stereo outputs, result) or rotates (pyramids, binned events),
  // and a slot is only reused after esvio_fe_track_wait returned its window.
  prof_mark(fe, kMarkSubmit, fe->stream_c);
  DevEvents ev[2];
  int one_block = 0;
  const bool left_first = fe->left_first && fe->q_count == 0 && (!fe->profiling || fe->left_first == 2) && !mc &&
                          !fe->cfg.median_blur_kernel_size && !fe->cfg.equalize && left && right &&
                          left->n > 0 && right->n > 0;
  if ((rc = stage_events_stereo_block(fe, w.slot, left, right, ev, &one_block,
                                      left_first ? fe->cl_done[w.slot] : nullptr)) != ESVIO_FE_OK)
    return rc;
  if (!one_block) {
    if ((rc = stage_events(fe, w.slot, 0, left, &ev[0])) != ESVIO_FE_OK) return rc;
    if (left_first) CU(cudaEventRecord(fe->cl_done[w.slot], fe->stream_c));
    if ((rc = stage_events(fe, w.slot, 1, right, &ev[1])) != ESVIO_FE_OK) return rc;
  }
  prof_mark(fe, kMarkLanded, fe->stream_c);
  if ((rc = staging_done(fe, w.slot)) != ESVIO_FE_OK) return rc;
  if (left_first) {
    if ((rc = run_event_stage_left_first(fe, w.slot, cur_time, ev, w.cur, w.rcur)) != ESVIO_FE_OK) return rc;
    if ((rc = submit_tracking(fe, w, ev[0], cur_time, pub_this_frame, fe->k1l_done[w.slot], fe->p_done[w.slot],
                              true)) != ESVIO_FE_OK)
      return rc;
    fe->pev_valid[w.slot] = fe->profiling;  // ESVIO_LEFT_FIRST=2 (diagnosis): the marks follow the left camera
    return ESVIO_FE_OK;
  }
  if ((rc = run_event_stage(fe, w.slot, cur_time, ev, w.cur, w.rcur, mc)) != ESVIO_FE_OK) return rc;
  if ((rc = submit_tracking(fe, w, ev[0], cur_time, pub_this_frame, fe->k1_done[w.slot],
                            fe->p_done[w.slot])) != ESVIO_FE_OK)
    return rc;
  fe->pev_valid[w.slot] = fe->profiling;
  return ESVIO_FE_OK;
}

FE_API int esvio_fe_track_wait(esvio_fe* fe, esvio_tracks* out) {
  if (!fe || !out) return ESVIO_FE_EINVAL;
  if (fe->q_count <= 0) return fail(fe, ESVIO_FE_ESTATE, "wait without submit", cudaSuccess);
  const int M = fe->cfg.max_cnt;
  // checked before the window is dequeued: a too-small `out` leaves it waitable
  if (out->capacity < M) return fail(fe, ESVIO_FE_ECAPACITY, "esvio_tracks.capacity < max_cnt", cudaSuccess);
  CU(cudaSetDevice(fe->dev));
  const int slot = fe->q_head;
  CU(cudaEventSynchronize(fe->q_done[slot]));
  fe->q_head = (fe->q_head + 1) % kSlots;
  fe->q_count--;
  const int32_t* r = fe->h_result[slot];
  const int nl = r[0], nr = r[1];
  out->n_left = nl;
  out->n_right = nr;
  const int32_t* a = r + kResultHdr;
  void* dst[kResultArrays] = {out->id, out->track_cnt, out->u,     out->v,     out->un_x,
                              out->un_y, out->vx,      out->vy,    out->id_right, out->ru,
                              out->rv, out->run_x,     out->run_y, out->rvx,   out->rvy};
  for (int k = 0; k < kResultArrays; ++k) {
    const int cnt = k < 8 ? nl : nr;
    if (dst[k] && cnt > 0) memcpy(dst[k], a + (size_t)k * M, (size_t)cnt * 4);
  }
  esvio_stats& s = out->stats;
  memset(&s, 0, sizeof(s));
  s.n_prev = r[2];
  s.n_after_temporal = r[3];
  s.n_after_ransac = r[4];
  s.n_after_mask = r[5];
  s.n_new = r[6];
  s.n_corner_flags = r[7];
  s.ransac_iters = r[8];
  if (fe->pev_valid[slot]) {
    // h2d, bin_events, sae_update_ts, pyramid, corner_flags, lk_temporal, select, lk_stereo, d2h
    static const int kFrom[ESVIO_FE_NUM_STAGES] = {kMarkSubmit, kMarkLanded, kMarkK1Start, kMarkK1Done, kMarkK1Done,
                                                   kMarkTemporalStart, kMarkTemporalLk, kMarkStereoStart, kMarkPacked};
    static const int kTo[ESVIO_FE_NUM_STAGES] = {kMarkLanded, kMarkBinned, kMarkK1Done, kMarkPyrDone, kMarkFlagsDone,
                                                 kMarkTemporalLk, kMarkSelect, kMarkPacked, kMarkResult};
    for (int i = 0; i < ESVIO_FE_NUM_STAGES; ++i)
      if (cudaEventElapsedTime(&fe->stage_ms[i], fe->pev[slot][kFrom[i]], fe->pev[slot][kTo[i]]) != cudaSuccess) {
        fe->stage_ms[i] = 0.f;
        cudaGetLastError();
      }
    for (int i = 0; i < ESVIO_FE_NUM_MARKS; ++i)
      if (cudaEventElapsedTime(&fe->stage_marks[i], fe->pev_ref, fe->pev[slot][i]) != cudaSuccess) {
        fe->stage_marks[i] = -1.f;
        cudaGetLastError();
      }
    fe->stage_ms_valid = 1;
    fe->pev_valid[slot] = 0;
  }
  return ESVIO_FE_OK;
}

FE_API int esvio_fe_track(esvio_fe* fe, double cur_time, const esvio_events* left,
                          const esvio_events* right, int32_t pub_this_frame, esvio_tracks* out) {
  return esvio_fe_track_mc(fe, cur_time, left, right, pub_this_frame, nullptr, out);
}

FE_API int esvio_fe_track_mc(esvio_fe* fe, double cur_time, const esvio_events* left,
                             const esvio_events* right, int32_t pub_this_frame,
                             const esvio_motion* mc, esvio_tracks* out) {
  if (!fe || !out) return ESVIO_FE_EINVAL;
  if (fe->q_count != 0) return fail(fe, ESVIO_FE_ESTATE, "track while windows are in flight", cudaSuccess);
  const int rc = esvio_fe_track_submit_mc(fe, cur_time, left, right, pub_this_frame, mc);
  if (rc != ESVIO_FE_OK) return rc;
  return esvio_fe_track_wait(fe, out);
}

// ---------------------------------------------------------------------------------------
// left/right split over two GPUs
// ---------------------------------------------------------------------------------------
// The right camera only feeds its pyramid to the stereo LK (feature_tracker.cpp:475-495), so
// it can live on another GPU: that GPU runs createSAE_right / SAEtoTimeSurface_right and the
// pyramid (esvio_fe_split_image_submit), the caller moves the image block (NCCL send/recv or a
// peer copy on a stream of its own) into the buffer esvio_fe_split_right_buffer names, and the
// left GPU tracks with it (esvio_fe_track_submit_split).  Ordering against the caller's
// exchange stream is by events recorded on / waited for by that stream, never by host syncs
// (exchange_stream is a cudaStream_t; 0 = the legacy default stream, which is what
// torch.cuda.current_stream().cuda_stream is unless the caller switched streams).
static int split_ok(esvio_fe* fe) {
  if (fe->group) return fail(fe, ESVIO_FE_ESTATE, "handle belongs to a group", cudaSuccess);
  return ESVIO_FE_OK;
}

FE_API int esvio_fe_split_image_submit(esvio_fe* fe, double cur_time, const esvio_events* ev,
                                       void* exchange_stream, void** image, size_t* bytes) {
  if (!fe || !image || !bytes) return ESVIO_FE_EINVAL;
  int rc;
  if ((rc = split_ok(fe)) != ESVIO_FE_OK) return rc;
  if (fe->q_count != 0) return fail(fe, ESVIO_FE_ESTATE, "split image while windows are in flight", cudaSuccess);
  CU(cudaSetDevice(fe->dev));
  cudaStream_t xs = (cudaStream_t)exchange_stream;
  const int slot = fe->windows % kSlots;
  const int idx = fe->windows == 0 ? kRightBase : kRightBase + (fe->cur_right - kRightBase + 1) % kRightBufs;
  // nobody waits for results on this GPU, so buffer reuse is ordered by events: the staging
  // buffer of this slot was last read by the binning kSlots windows ago ...
  CU(cudaStreamWaitEvent(fe->stream_c, fe->b_done[slot], 0));
  // ... and the sends of earlier windows, enqueued on the exchange stream, have read the image
  // buffers before K1 writes the next time surface into one of them
  CU(cudaEventRecord(fe->x_ready[slot], xs));
  CU(cudaStreamWaitEvent(fe->stream_e, fe->x_ready[slot], 0));
  DevEvents d[2];
  memset(d, 0, sizeof(d));
  if ((rc = stage_events(fe, slot, 0, ev, &d[0])) != ESVIO_FE_OK) return rc;
  if ((rc = staging_done(fe, slot)) != ESVIO_FE_OK) return rc;
  if ((rc = run_event_stage(fe, slot, cur_time, d, idx, idx, nullptr, 1)) != ESVIO_FE_OK) return rc;
  CU(cudaStreamWaitEvent(xs, fe->p_done[slot], 0));
  fe->cur_right = idx;
  fe->windows++;
  fe->prev_time = cur_time;
  *image = fe->pyr[idx];
  *bytes = fe->pd.bytes;
  return ESVIO_FE_OK;
}

FE_API int esvio_fe_split_right_buffer(esvio_fe* fe, void** image, size_t* bytes) {
  if (!fe || !image || !bytes) return ESVIO_FE_EINVAL;
  int rc;
  if ((rc = split_ok(fe)) != ESVIO_FE_OK) return rc;
  WindowPlan w;
  // the stereo LK that last read this buffer (kSlots windows ago) has been waited for
  if ((rc = plan_window(fe, &w)) != ESVIO_FE_OK) return rc;
  *image = fe->pyr[w.rcur];
  *bytes = fe->pd.bytes;
  return ESVIO_FE_OK;
}

FE_API int esvio_fe_track_submit_split(esvio_fe* fe, double cur_time, const esvio_events* left,
                                       int32_t pub_this_frame, void* exchange_stream) {
  if (!fe) return ESVIO_FE_EINVAL;
  int rc;
  if ((rc = split_ok(fe)) != ESVIO_FE_OK) return rc;
  CU(cudaSetDevice(fe->dev));
  WindowPlan w;
  if ((rc = plan_window(fe, &w)) != ESVIO_FE_OK) return rc;
  fe->pev_slot = w.slot;
  prof_mark(fe, kMarkSubmit, fe->stream_c);
  DevEvents ev[2];
  memset(ev, 0, sizeof(ev));
  if ((rc = stage_events(fe, w.slot, 0, left, &ev[0])) != ESVIO_FE_OK) return rc;
  prof_mark(fe, kMarkLanded, fe->stream_c);
  if ((rc = staging_done(fe, w.slot)) != ESVIO_FE_OK) return rc;
  if ((rc = run_event_stage(fe, w.slot, cur_time, ev, w.cur, w.rcur, nullptr, 1)) != ESVIO_FE_OK) return rc;
  // the stereo LK reads the right image the caller wrote on its exchange stream
  cudaStream_t xs = (cudaStream_t)exchange_stream;
  CU(cudaEventRecord(fe->x_ready[w.slot], xs));
  if ((rc = submit_tracking(fe, w, ev[0], cur_time, pub_this_frame, fe->k1_done[w.slot], fe->p_done[w.slot],
                            true)) != ESVIO_FE_OK)
    return rc;
  fe->pev_valid[w.slot] = fe->profiling;
  return ESVIO_FE_OK;
}

// ---------------------------------------------------------------------------------------
// time-window shard: building blocks
// ---------------------------------------------------------------------------------------
// createSAE_*'s acceptance test reads only sae_latest_ (event_detector.cc:149-166), which after
// a window is "time of the last event per pixel and polarity": consecutive windows can therefore
// be replayed on different GPUs once each knows the element-wise maximum of the windows before
// it (its carry-in), and the accepted times (sae_) merge the same way afterwards.  The caller
// (esvio_b200/shard.py TimeWindowShard) owns the protocol and the collectives; the library
// provides: the state planes as device memory, an element-wise maximum over planes, the event
// stage ordered on the caller's stream, the corner candidates of a window on the device, and
// tracking from event-stage products computed elsewhere.
FE_API int esvio_fe_state_device_ptrs(esvio_fe* fe, void** sae, void** lat, size_t* bytes) {
  if (!fe || !sae || !lat || !bytes) return ESVIO_FE_EINVAL;
  *sae = fe->sae;
  *lat = fe->lat;
  *bytes = fe->npx * 2 * sizeof(double2);
  return ESVIO_FE_OK;
}

FE_API int esvio_fe_shard_merge_max(esvio_fe* fe, void* dst, const void* const* srcs, int32_t n_src,
                                    size_t n_doubles, void* cuda_stream) {
  if (!fe || !dst || !srcs || n_src < 1 || n_src > kMaxMergeSrc) return ESVIO_FE_EINVAL;
  CU(cudaSetDevice(fe->dev));
  launch_merge_max((double*)dst, (const double* const*)srcs, n_src, n_doubles, (cudaStream_t)cuda_stream,
                   &fe->launches);
  CU(cudaGetLastError());
  return ESVIO_FE_OK;
}

// esvio_fe_stage_update without host synchronisation: binning + SAE update + time surface +
// pyramids of one window on the handle's own streams, ordered behind and in front of the
// caller's stream.  Events must be device-resident (or stay valid until the stream passed).
FE_API int esvio_fe_shard_event_stage(esvio_fe* fe, double t_ref, const esvio_events* left,
                                      const esvio_events* right, void* cuda_stream) {
  if (!fe) return ESVIO_FE_EINVAL;
  if (fe->group) return fail(fe, ESVIO_FE_ESTATE, "handle belongs to a group", cudaSuccess);
  if (fe->q_count != 0) return fail(fe, ESVIO_FE_ESTATE, "windows in flight", cudaSuccess);
  CU(cudaSetDevice(fe->dev));
  cudaStream_t us = (cudaStream_t)cuda_stream;
  const int slot = fe->shard_seq % kSlots;
  fe->shard_seq++;
  DevEvents ev[2];
  int rc;
  // everything the caller enqueued so far (state writes, earlier reads of the images) first
  CU(cudaEventRecord(fe->x_ready[slot], us));
  for (cudaStream_t st : {fe->stream_c, fe->stream_b, fe->stream_e, fe->stream_p})
    CU(cudaStreamWaitEvent(st, fe->x_ready[slot], 0));
  if ((rc = stage_events(fe, slot, 0, left, &ev[0])) != ESVIO_FE_OK) return rc;
  if ((rc = stage_events(fe, slot, 1, right, &ev[1])) != ESVIO_FE_OK) return rc;
  if ((rc = staging_done(fe, slot)) != ESVIO_FE_OK) return rc;
  if ((rc = run_event_stage(fe, slot, t_ref, ev, fe->cur_left, fe->cur_right)) != ESVIO_FE_OK) return rc;
  CU(cudaStreamWaitEvent(us, fe->p_done[slot], 0));
  return ESVIO_FE_OK;
}

// Arc* candidates of a window's left events against the handle's current state, on the caller's
// stream; returns the device lists (layout: CornerParams::cand / cand_cnt)
FE_API int esvio_fe_shard_corner_candidates(esvio_fe* fe, const esvio_events* left, void* cuda_stream,
                                            void** cand, void** cand_cnt) {
  if (!fe || !left || !cand || !cand_cnt) return ESVIO_FE_EINVAL;
  if (!left->on_device && left->n) return fail(fe, ESVIO_FE_EINVAL, "device-resident events only", cudaSuccess);
  CU(cudaSetDevice(fe->dev));
  DevEvents ev;
  int rc;
  if ((rc = stage_events(fe, 0, 0, left, &ev)) != ESVIO_FE_OK) return rc;
  launch_corner_flags(corner_params(fe, fe->cur_left, 1, kSlots - 1), ev, fe->flags[kSlots - 1],
                      (cudaStream_t)cuda_stream, &fe->launches);
  CU(cudaGetLastError());
  *cand = fe->cand[kSlots - 1];
  *cand_cnt = fe->cand_cnt[kSlots - 1];
  return ESVIO_FE_OK;
}

FE_API int esvio_fe_shard_sizes(esvio_fe* fe, size_t* image_bytes, size_t* cand_bytes, size_t* cand_cnt_bytes) {
  if (!fe || !image_bytes || !cand_bytes || !cand_cnt_bytes) return ESVIO_FE_EINVAL;
  *image_bytes = fe->pd.bytes;
  *cand_bytes = ((size_t)fe->cap + kCornerBlock) * sizeof(uint32_t);
  *cand_cnt_bytes = ((size_t)fe->cap / kCornerBlock + 2) * sizeof(int);
  return ESVIO_FE_OK;
}

FE_API int esvio_fe_shard_images(esvio_fe* fe, void** left_img, void** right_img) {
  if (!fe || !left_img || !right_img) return ESVIO_FE_EINVAL;
  *left_img = fe->pyr[fe->cur_left];
  *right_img = fe->pyr[fe->cur_right];
  return ESVIO_FE_OK;
}

// the buffers the NEXT submitted window reads its event-stage products from
FE_API int esvio_fe_external_buffers(esvio_fe* fe, void** left_img, void** right_img, void** cand,
                                     void** cand_cnt) {
  if (!fe || !left_img || !right_img || !cand || !cand_cnt) return ESVIO_FE_EINVAL;
  WindowPlan w;
  int rc;
  if ((rc = plan_window(fe, &w)) != ESVIO_FE_OK) return rc;
  *left_img = fe->pyr[w.cur];
  *right_img = fe->pyr[w.rcur];
  *cand = fe->cand[w.slot];
  *cand_cnt = fe->cand_cnt[w.slot];
  return ESVIO_FE_OK;
}

// esvio_fe_track_submit for a window whose event stage ran elsewhere: the image pyramids and
// (publish windows) the corner candidate lists were written into esvio_fe_external_buffers on
// `cuda_stream`; temporal chain, stereo LK and packing run here as usual.
FE_API int esvio_fe_track_submit_external(esvio_fe* fe, double cur_time, int32_t n_left_events,
                                          int32_t pub_this_frame, void* cuda_stream) {
  if (!fe || n_left_events < 0) return ESVIO_FE_EINVAL;
  if (fe->group) return fail(fe, ESVIO_FE_ESTATE, "handle belongs to a group", cudaSuccess);
  CU(cudaSetDevice(fe->dev));
  WindowPlan w;
  int rc;
  if ((rc = plan_window(fe, &w)) != ESVIO_FE_OK) return rc;
  fe->pev_slot = w.slot;
  const int slot = w.slot, M = fe->cfg.max_cnt;
  const TrackBuffers& B = fe->tb;
  cudaStream_t s1 = fe->stream_t1, ss = fe->stream_s[slot & 1], s2 = fe->stream;
  CU(cudaEventRecord(fe->x_ready[slot], (cudaStream_t)cuda_stream));
  CU(cudaStreamWaitEvent(s1, fe->x_ready[slot], 0));
  launch_lk(fe->pd, fe->pyr[w.prev], fe->pyr[w.cur], B.prev_pts, B.cur_pts, B.st_fwd, B.rev_pts, B.st_bwd,
            &B.st->n_prev, M, 3, 0, fe->cfg.flow_back ? 1 : 0, s1, &fe->launches);
  launch_post_temporal(fe->tp, B, pub_this_frame ? -1 : slot, s1, &fe->launches);
  if (pub_this_frame) {
    if (fe->cfg.use_ransac) launch_ransac(fe->tp, B, s1, &fe->launches);
    launch_select(fe->tp, B, n_left_events, fe->cand[slot], fe->cand_cnt[slot], slot, s1, &fe->launches);
  }
  CU(cudaEventRecord(fe->t1_done[slot], s1));
  CU(cudaStreamWaitEvent(ss, fe->t1_done[slot], 0));
  launch_lk(fe->pd, fe->pyr[w.cur], fe->pyr[w.rcur], B.snap_pts + (size_t)slot * M,
            B.right_pts + (size_t)slot * M, B.st_sf + (size_t)slot * M, B.rev_left_pts + (size_t)slot * M,
            B.st_sb + (size_t)slot * M, B.snap_hdr + slot * 16, M, 3, 0, fe->cfg.flow_back ? 2 : 0, ss,
            &fe->launches);
  CU(cudaEventRecord(fe->s_done[slot], ss));
  CU(cudaStreamWaitEvent(s2, fe->s_done[slot], 0));
  if (fe->r_held[slot]) {
    CU(cudaStreamWaitEvent(s2, fe->r_free[slot], 0));
    fe->r_held[slot] = 0;
  }
  launch_finalize(fe->tp, B, slot, cur_time, fe->prev_time, s2, &fe->launches);
  CU(cudaMemcpyAsync(fe->h_result[slot], B.result + (size_t)slot * fe->result_words, fe->result_words * 4,
                     cudaMemcpyDeviceToHost, s2));
  CU(cudaEventRecord(fe->q_done[slot], s2));
  CU(cudaGetLastError());
  fe->q_count++;
  fe->last_slot = slot;
  fe->prev_left = w.prev;
  fe->cur_left = w.cur;
  fe->cur_right = w.rcur;
  fe->windows++;
  fe->prev_time = cur_time;
  fe->pev_valid[slot] = 0;
  return ESVIO_FE_OK;
}

// ---------------------------------------------------------------------------------------
// frame path: FeatureTracker::trackImage (feature_tracker.cpp:164-338)
// ---------------------------------------------------------------------------------------
// One allocation for the scratch of goodFeaturesToTrack (frames.cu), made on the first frame.
static int ensure_gftt(esvio_fe* fe) {
  if (fe->gftt_block) return ESVIO_FE_OK;
  const size_t N = fe->npx, words = (size_t)fe->H * ((fe->W + 31) / 32);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  };
  const size_t o_cov = take(3 * N * 4), o_eig = take(N * 4), o_blk = take(words * 4),
               o_thr = take(256), o_keys = take(N * 8), o_sorted = take(N * 8),
               o_cnt = take(256), o_xy = take(N * 8), o_n = take(256);
  uint8_t* base = nullptr;
  CU(cudaMalloc(&base, off));
  CU(cudaMemset(base, 0, off));
  GfttBuffers& G = fe->gftt;
  for (int c = 0; c < 3; ++c) G.cov[c] = (float*)(base + o_cov) + (size_t)c * N;
  G.eig = (float*)(base + o_eig);
  G.blocked = (uint32_t*)(base + o_blk);
  G.thr = (float*)(base + o_thr);
  G.keys = (unsigned long long*)(base + o_keys);
  G.keys_sorted = (unsigned long long*)(base + o_sorted);
  G.n_cand = (int*)(base + o_cnt);
  G.out_xy = (float2*)(base + o_xy);
  G.out_n = (int*)(base + o_n);
  fe->gftt_block = base;
  return ESVIO_FE_OK;
}

// The frame counterpart of esvio_fe_track_submit.  cfg.max_cnt / cfg.min_dist play MAX_CNT_IMG /
// MIN_DIST_IMG, cfg.width / height COL / ROW.  right == NULL: img_right.empty() (mono).
FE_API int esvio_fe_track_image_submit(esvio_fe* fe, double cur_time, const uint8_t* left,
                                       size_t left_stride, const uint8_t* right,
                                       size_t right_stride, int32_t pub_this_frame) {
  if (!fe || !left || left_stride < (size_t)fe->W || (right && right_stride < (size_t)fe->W))
    return ESVIO_FE_EINVAL;
  if (fe->group) return fail(fe, ESVIO_FE_ESTATE, "handle belongs to a group", cudaSuccess);
  CU(cudaSetDevice(fe->dev));
  int rc;
  if ((rc = ensure_gftt(fe)) != ESVIO_FE_OK) return rc;
  WindowPlan w;
  if ((rc = plan_window(fe, &w)) != ESVIO_FE_OK) return rc;
  fe->pev_slot = w.slot;
  cudaStream_t sc = fe->stream_c, sp = fe->stream_p, s1 = fe->stream_t1, ss = fe->stream_s[w.slot & 1],
               s2 = fe->stream;
  const int slot = w.slot, cur = w.cur, prev = w.prev, rcur = w.rcur, pitch = fe->pd.pitch[0];
  const int M = fe->cfg.max_cnt;
  const TrackBuffers& B = fe->tb;
  const GfttBuffers& G = fe->gftt;
  // ---- image stage: the frames into pyramid level 0 (copy stream), then the pyramids
  prof_mark(fe, kMarkSubmit, sc);
  CU(cudaMemcpy2DAsync(fe->pyr[cur], pitch, left, left_stride, fe->W, fe->H, cudaMemcpyHostToDevice, sc));
  if (right)
    CU(cudaMemcpy2DAsync(fe->pyr[rcur], pitch, right, right_stride, fe->W, fe->H, cudaMemcpyHostToDevice, sc));
  prof_mark(fe, kMarkLanded, sc);
  if ((rc = staging_done(fe, slot)) != ESVIO_FE_OK) return rc;
  CU(cudaStreamWaitEvent(sp, fe->c_done[slot], 0));
  prof_mark(fe, kMarkBinned, sp);
  prof_mark(fe, kMarkK1Start, sp);
  prof_mark(fe, kMarkK1Done, sp);
  uint8_t* imgs[2] = {fe->pyr[cur], fe->pyr[rcur]};
  // EQUALIZE of the image node (stereo_image_tracker_node.cpp:93-97: createCLAHE()->apply on both
  // frames before trackImage), done here when the handle was created with equalize = 1
  if (fe->cfg.equalize)
    launch_clahe_inplace(imgs, right ? 2 : 1, fe->W, fe->H, pitch, fe->clahe_lut, fe->clahe_minmax, sp,
                         &fe->launches);
  launch_pyramids(fe->pd, imgs, right ? 2 : 1, sp, &fe->launches);
  fe->ts_sel[0] = fe->ts_sel[1] = nullptr;
  prof_mark(fe, kMarkPyrDone, sp);
  prof_mark(fe, kMarkFlagsDone, sp);
  CU(cudaEventRecord(fe->p_done[slot], sp));
  // ---- temporal chain (:178-237): forward + full backward LK, Image_setMask, goodFeaturesToTrack
  CU(cudaStreamWaitEvent(s1, fe->p_done[slot], 0));
  prof_mark(fe, kMarkTemporalStart, s1);
  launch_lk(fe->pd, fe->pyr[prev], fe->pyr[cur], B.prev_pts, B.cur_pts, B.st_fwd, B.rev_pts,
            B.st_bwd, &B.st->n_prev, M, 3, 0, fe->cfg.flow_back ? 2 : 0, s1, &fe->launches);
  launch_post_temporal(fe->tp, B, pub_this_frame ? -1 : slot, s1, &fe->launches);
  prof_mark(fe, kMarkTemporalLk, s1);
  if (pub_this_frame) {
    // the scratch of goodFeaturesToTrack is only touched on this stream, so windows in flight
    // cannot collide on it
    launch_image_set_mask(fe->tp, B, G, s1, &fe->launches);
    launch_gftt_eig(G, fe->pyr[cur], pitch, fe->W, fe->H, s1, &fe->launches);
    launch_gftt_thr(G, fe->W, fe->H, true, s1, &fe->launches);
    if (launch_gftt_candidates(G, fe->W, fe->H, true, s1, &fe->launches) != 0) {
      // nothing of this frame has been booked yet: drain what was enqueued and drop the frame
      const cudaError_t ce = cudaGetLastError();
      sync_all(fe);
      return fail(fe, ESVIO_FE_ECUDA, "sort of the corner candidates (frame dropped)", ce);
    }
    launch_gftt_pick_tracks(fe->tp, B, G, slot, s1, &fe->launches);
  }
  prof_mark(fe, kMarkSelect, s1);
  CU(cudaEventRecord(fe->t1_done[slot], s1));
  // ---- stereo LK (:245-322) on one of the two alternating streams
  CU(cudaStreamWaitEvent(ss, fe->t1_done[slot], 0));
  prof_mark(fe, kMarkStereoStart, ss);
  if (right)
    launch_lk(fe->pd, fe->pyr[cur], fe->pyr[rcur], B.snap_pts + (size_t)slot * M,
              B.right_pts + (size_t)slot * M, B.st_sf + (size_t)slot * M,
              B.rev_left_pts + (size_t)slot * M, B.st_sb + (size_t)slot * M, B.snap_hdr + slot * 16, M,
              3, 0, fe->cfg.flow_back ? 2 : 0, ss, &fe->launches);
  else  // no right image: no right points
    CU(cudaMemsetAsync(B.st_sf + (size_t)slot * M, 0, M, ss));
  CU(cudaEventRecord(fe->s_done[slot], ss));
  // ---- packing, in frame order on the result stream
  CU(cudaStreamWaitEvent(s2, fe->s_done[slot], 0));
  if (fe->r_held[slot]) {
    CU(cudaStreamWaitEvent(s2, fe->r_free[slot], 0));
    fe->r_held[slot] = 0;
  }
  if (right) {
    launch_finalize(fe->tp, B, slot, cur_time, fe->prev_time, s2, &fe->launches);
  } else {
    // prev_un_right_pts_map stays as it is (:245: the whole block is skipped)
    launch_right_map_keep(B, 0, s2, &fe->launches);
    launch_finalize(fe->tp, B, slot, cur_time, fe->prev_time, s2, &fe->launches);
    launch_right_map_keep(B, 1, s2, &fe->launches);
  }
  prof_mark(fe, kMarkPacked, s2);
  CU(cudaMemcpyAsync(fe->h_result[slot], B.result + (size_t)slot * fe->result_words, fe->result_words * 4,
                     cudaMemcpyDeviceToHost, s2));
  prof_mark(fe, kMarkResult, s2);
  CU(cudaEventRecord(fe->q_done[slot], s2));
  CU(cudaGetLastError());
  fe->q_count++;
  fe->last_slot = slot;
  fe->prev_left = prev;
  fe->cur_left = cur;
  fe->cur_right = rcur;
  fe->windows++;
  fe->prev_time = cur_time;
  fe->pev_valid[slot] = fe->profiling;
  return ESVIO_FE_OK;
}

FE_API int esvio_fe_track_image(esvio_fe* fe, double cur_time, const uint8_t* left,
                                size_t left_stride, const uint8_t* right, size_t right_stride,
                                int32_t pub_this_frame, esvio_tracks* out) {
  if (!fe || !out) return ESVIO_FE_EINVAL;
  if (fe->q_count != 0) return fail(fe, ESVIO_FE_ESTATE, "track while windows are in flight", cudaSuccess);
  const int rc = esvio_fe_track_image_submit(fe, cur_time, left, left_stride, right, right_stride,
                                             pub_this_frame);
  if (rc != ESVIO_FE_OK) return rc;
  return esvio_fe_track_wait(fe, out);
}

// stage entry: cv::goodFeaturesToTrack(img, max_corners, 0.01, min_distance, mask) on a host
// image (W x H, contiguous); mask NULL or W x H bytes, non-zero = allowed.  out_xy has room for
// `capacity` corners; eig (nullable) receives the cornerMinEigenVal plane.
FE_API int esvio_fe_stage_good_features(esvio_fe* fe, const uint8_t* img, const uint8_t* mask,
                                        int32_t max_corners, double min_distance, float* out_xy,
                                        int32_t capacity, int32_t* out_n, float* eig) {
  if (!fe || !img || !out_xy || !out_n || capacity < 0) return ESVIO_FE_EINVAL;
  if (min_distance > 1.0 && (max_corners <= 0 || max_corners > kMaxCnt))
    return fail(fe, ESVIO_FE_EINVAL, "spaced corners: 1 <= max_corners <= 1024", cudaSuccess);
  if (fe->q_count != 0) return fail(fe, ESVIO_FE_ESTATE, "windows in flight", cudaSuccess);
  CU(cudaSetDevice(fe->dev));
  int rc;
  if ((rc = ensure_gftt(fe)) != ESVIO_FE_OK) return rc;
  if (sync_all(fe) != ESVIO_FE_OK) return ESVIO_FE_ECUDA;
  cudaStream_t s = fe->stream;
  const GfttBuffers& G = fe->gftt;
  const int W = fe->W, H = fe->H, pitch = fe->pd.pitch[0], words = (W + 31) / 32;
  uint8_t* d_img = fe->pyr[kScratchBase];
  CU(cudaMemcpy2DAsync(d_img, pitch, img, W, W, H, cudaMemcpyHostToDevice, s));
  uint32_t* h_blk = nullptr;
  if (mask) {
    h_blk = (uint32_t*)calloc((size_t)H * words, 4);
    if (!h_blk) return fail(fe, ESVIO_FE_ECUDA, "out of host memory", cudaSuccess);
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x)
        if (!mask[(size_t)y * W + x]) h_blk[(size_t)y * words + (x >> 5)] |= 1u << (x & 31);
    const cudaError_t e = cudaMemcpyAsync(G.blocked, h_blk, (size_t)H * words * 4, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) {
      free(h_blk);
      return fail(fe, ESVIO_FE_ECUDA, "mask upload", e);
    }
  }
  launch_gftt_eig(G, d_img, pitch, W, H, s, &fe->launches);
  launch_gftt_thr(G, W, H, mask != nullptr, s, &fe->launches);
  const int sort_rc = launch_gftt_candidates(G, W, H, mask != nullptr, s, &fe->launches);
  launch_gftt_pick_stage(G, W, H, max_corners, min_distance, s, &fe->launches);
  int n = 0;
  cudaError_t e = cudaMemcpyAsync(&n, G.out_n, sizeof(int), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  free(h_blk);
  if (sort_rc != 0 || e != cudaSuccess) return fail(fe, ESVIO_FE_ECUDA, "good features", e);
  *out_n = n;
  const int m = n < capacity ? n : capacity;
  if (m > 0) CU(cudaMemcpy(out_xy, G.out_xy, sizeof(float2) * m, cudaMemcpyDeviceToHost));
  if (eig) CU(cudaMemcpy(eig, G.eig, sizeof(float) * (size_t)W * H, cudaMemcpyDeviceToHost));
  return ESVIO_FE_OK;
}

FE_API int esvio_fe_time_surface(esvio_fe* fe, int32_t cam, uint8_t* dst, size_t stride) {
  if (!fe || !dst || cam < 0 || cam > 1 || stride < (size_t)fe->W) return ESVIO_FE_EINVAL;
  CU(cudaSetDevice(fe->dev));
  if (sync_all(fe) != ESVIO_FE_OK) return ESVIO_FE_ECUDA;
  const uint8_t* src = fe->ts_sel[cam] ? fe->ts_sel[cam] : fe->pyr[cam == 0 ? fe->cur_left : fe->cur_right];
  CU(cudaMemcpy2DAsync(dst, stride, src, fe->pd.pitch[0], fe->W, fe->H, cudaMemcpyDeviceToHost,
                       fe->stream));
  CU(cudaStreamSynchronize(fe->stream));
  return ESVIO_FE_OK;
}

// ---------------------------------------------------------------------------------------
// groups: S independent stereo streams whose event stage runs as ONE batched launch sequence
// ---------------------------------------------------------------------------------------
// SURVEY.md 8e "independent streams" on one GPU (BASELINE configs[4]: 4 stereo pairs): binning,
// SAE update + time surface and the pyramids of all 2S cameras share their launches (one
// k_sae_update_ts per window for the whole group), the per-stream tracking stages then run
// concurrently on the members' own streams.  Results are identical to S separate handles.
struct esvio_fe_group {
  int S, dev;
  esvio_fe* m[kMaxCams / 2];
  double2 *sae, *lat;                                 // [2S][H][W]
  double2 *own_sae[kMaxCams / 2], *own_lat[kMaxCams / 2];  // the members' own planes (unused)
  CUtensorMap map_sae, map_lat;
  EventStageBuffers esb[2];  // even / odd windows, as in esvio_fe
  BinLayout bl;
  cudaStream_t stream_b, stream_e, stream_p;  // batched K0 | K1 | pyramids of all 2S cameras
  cudaEvent_t b_done[kSlots], p_done[kSlots];
  cudaEvent_t k1_beg[kSlots], k1_end[kSlots];  // timed; k1_end doubles as "SAE + time surface done"
  int k1_slot_valid[kSlots];
  float k1_ms;
  int k1_ms_valid;
  int64_t launches;
};

FE_API void esvio_fe_group_destroy(esvio_fe_group* g) {
  if (!g) return;
  cudaSetDevice(g->dev);
  for (cudaStream_t st : {g->stream_b, g->stream_e, g->stream_p})
    if (st) cudaStreamSynchronize(st);
  for (int i = 0; i < g->S; ++i)
    if (g->m[i]) {
      sync_all(g->m[i]);
      g->m[i]->sae = g->own_sae[i];
      g->m[i]->lat = g->own_lat[i];
      g->m[i]->group = nullptr;
      free_all(g->m[i]);
    }
  cudaFree(g->sae);
  cudaFree(g->lat);
  for (int b = 0; b < 2; ++b) event_stage_free(&g->esb[b]);
  for (int k = 0; k < kSlots; ++k)
    for (cudaEvent_t ev : {g->b_done[k], g->p_done[k], g->k1_beg[k], g->k1_end[k]})
      if (ev) cudaEventDestroy(ev);
  for (cudaStream_t st : {g->stream_b, g->stream_e, g->stream_p})
    if (st) cudaStreamDestroy(st);
  free(g);
}

FE_API int esvio_fe_group_create(const esvio_fe_config* cfg, int32_t n_streams,
                                 esvio_fe_group** out) {
  if (!cfg || !out || n_streams < 1 || n_streams > kMaxCams / 2) return ESVIO_FE_EINVAL;
  *out = nullptr;
  if (cfg->equalize || cfg->median_blur_kernel_size || cfg->do_motion_correction)
    return ESVIO_FE_EINVAL;  // the batched event stage covers the plain path only
  esvio_fe_group* g = (esvio_fe_group*)calloc(1, sizeof(esvio_fe_group));
  if (!g) return ESVIO_FE_EINVAL;
  g->S = n_streams;
  g->dev = cfg->device_id;
  for (int i = 0; i < n_streams; ++i) {
    g_creating_group_member = true;
    const int rc = esvio_fe_create(cfg, &g->m[i]);
    g_creating_group_member = false;
    if (rc != ESVIO_FE_OK) {
      g->S = i;
      esvio_fe_group_destroy(g);
      return rc;
    }
    g->own_sae[i] = g->m[i]->sae;
    g->own_lat[i] = g->m[i]->lat;
  }
  esvio_fe* f0 = g->m[0];
  const int NC = 2 * n_streams;
  const size_t npx = f0->npx;
  g->bl = f0->bl;
  cudaError_t ce = cudaSetDevice(g->dev);
#define GC(call) if (ce == cudaSuccess) ce = (call)
  {
    int prio_lo = 0, prio_hi = 0;
    GC(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    GC(cudaStreamCreateWithPriority(&g->stream_b, cudaStreamNonBlocking, prio_hi));
    GC(cudaStreamCreateWithPriority(&g->stream_e, cudaStreamNonBlocking, prio_hi));
    GC(cudaStreamCreateWithPriority(&g->stream_p, cudaStreamNonBlocking, prio_hi));
  }
  GC(cudaMalloc(&g->sae, npx * NC * sizeof(double2)));
  GC(cudaMalloc(&g->lat, npx * NC * sizeof(double2)));
  GC(cudaMemset(g->sae, 0, npx * NC * sizeof(double2)));
  GC(cudaMemset(g->lat, 0, npx * NC * sizeof(double2)));
  for (int b = 0; b < 2; ++b)
    GC(event_stage_alloc(g->bl, NC, f0->cap, &g->esb[b]) == 0 ? cudaSuccess : cudaErrorMemoryAllocation);
  for (int k = 0; k < kSlots; ++k) {
    GC(cudaEventCreateWithFlags(&g->b_done[k], cudaEventDisableTiming));
    GC(cudaEventCreateWithFlags(&g->p_done[k], cudaEventDisableTiming));
    GC(cudaEventCreate(&g->k1_beg[k]));
    GC(cudaEventCreate(&g->k1_end[k]));
  }
  GC(cudaDeviceSynchronize());
#undef GC
  int rc = ce == cudaSuccess ? ESVIO_FE_OK : ESVIO_FE_ECUDA;
  if (rc == ESVIO_FE_OK) rc = make_state_map(f0, g->sae, &g->map_sae, NC);
  if (rc == ESVIO_FE_OK) rc = make_state_map(f0, g->lat, &g->map_lat, NC);
  if (rc != ESVIO_FE_OK) {
    fprintf(stderr, "esvio_fe_group_create: %s\n", ce != cudaSuccess ? cudaGetErrorString(ce) : f0->err);
    esvio_fe_group_destroy(g);
    return rc;
  }
  for (int i = 0; i < n_streams; ++i) {  // the members see their slice of the group's state
    g->m[i]->sae = g->sae + (size_t)2 * i * npx;
    g->m[i]->lat = g->lat + (size_t)2 * i * npx;
    g->m[i]->group = g;
  }
  *out = g;
  return ESVIO_FE_OK;
}

FE_API int esvio_fe_group_reset(esvio_fe_group* g) {
  if (!g) return ESVIO_FE_EINVAL;
  esvio_fe* fe = g->m[0];
  CU(cudaSetDevice(g->dev));
  for (cudaStream_t st : {g->stream_b, g->stream_e, g->stream_p}) CU(cudaStreamSynchronize(st));
  for (int i = 0; i < g->S; ++i) {
    int rc = sync_all(g->m[i]);
    if (rc == ESVIO_FE_OK) rc = reset_state(g->m[i]);  // clears the member's slice of the SAE too
    if (rc != ESVIO_FE_OK) return rc;
  }
  for (int b = 0; b < 2; ++b) event_stage_clear(g->bl, g->esb[b], g->stream_b);
  CU(cudaStreamSynchronize(g->stream_b));
  for (int k = 0; k < kSlots; ++k) g->k1_slot_valid[k] = 0;
  g->k1_ms_valid = 0;
  return ESVIO_FE_OK;
}

FE_API esvio_fe* esvio_fe_group_member(esvio_fe_group* g, int32_t i) {
  return (g && i >= 0 && i < g->S) ? g->m[i] : nullptr;
}

FE_API int esvio_fe_group_track_submit(esvio_fe_group* g, const double* cur_time,
                                       const esvio_events* left, const esvio_events* right,
                                       const int32_t* pub_this_frame) {
  if (!g || !cur_time || !left || !right || !pub_this_frame) return ESVIO_FE_EINVAL;
  esvio_fe* fe = g->m[0];  // error text lands on member 0
  CU(cudaSetDevice(g->dev));
  const int S = g->S;
  WindowPlan w[kMaxCams / 2];
  DevEvents ev[kMaxCams];
  int rc;
  for (int i = 0; i < S; ++i)
    if ((rc = plan_window(g->m[i], &w[i])) != ESVIO_FE_OK) return rc;
  cudaStream_t sb = g->stream_b, se = g->stream_e, s_pyr = g->stream_p;
  const int slot = w[0].slot;  // members are always submitted and waited together
  // ---- K0 of all 2S cameras on the group's binning stream, behind the members' copies and the
  // K1 that last read this set of binned-event buffers (two windows back)
  const EventStageBuffers& esb = g->esb[slot & 1];
  for (int i = 0; i < S; ++i) {
    if ((rc = stage_events(g->m[i], w[i].slot, 0, &left[i], &ev[2 * i])) != ESVIO_FE_OK) return rc;
    if ((rc = stage_events(g->m[i], w[i].slot, 1, &right[i], &ev[2 * i + 1])) != ESVIO_FE_OK) return rc;
    if ((rc = staging_done(g->m[i], w[i].slot)) != ESVIO_FE_OK) return rc;
    CU(cudaStreamWaitEvent(sb, g->m[i]->c_done[w[i].slot], 0));
  }
  CU(cudaStreamWaitEvent(sb, g->k1_end[(slot + kSlots - 2) % kSlots], 0));
  launch_bin_events(g->bl, esb, ev, sb, &g->launches);
  CU(cudaEventRecord(g->b_done[slot], sb));
  // ---- ONE K1 launch for the group, behind the members' corner flags of the window before
  CU(cudaStreamWaitEvent(se, g->b_done[slot], 0));
  for (int i = 0; i < S; ++i)
    if (g->m[i]->f_pending >= 0) {
      CU(cudaStreamWaitEvent(se, g->m[i]->f_done[g->m[i]->f_pending], 0));
      g->m[i]->f_pending = -1;
    }
  SaeTsParams sp;
  sp.W = fe->W;
  sp.H = fe->H;
  sp.tiles_x = g->bl.tiles_x;
  sp.n_tiles = g->bl.n_tiles;
  sp.n_cams = 2 * S;
  sp.decay_sec = fe->cfg.decay_ms / 1000.0;
  sp.inv_decay = 1.0 / sp.decay_sec;
  sp.filter_threshold = fe->cfg.feature_filter_threshold;
  sp.ignore_polarity = fe->cfg.ignore_polarity;
  sp.bin_start = esb.bin_start;
  sp.ts_pitch = fe->pd.pitch[0];
  uint8_t* imgs[kMaxCams];
  for (int c = 0; c < kMaxCams; ++c) {
    const bool on = c < 2 * S;
    const int i = c / 2;
    sp.t_ref[c] = on ? cur_time[i] : 0.0;
    sp.bt[c] = on ? esb.bt[c] : nullptr;
    sp.bk[c] = on ? esb.bk[c] : nullptr;
    imgs[c] = on ? g->m[i]->pyr[(c & 1) ? w[i].rcur : w[i].cur] : nullptr;
    sp.ts[c] = imgs[c];
  }
  CU(cudaEventRecord(g->k1_beg[slot], se));
  launch_sae_update_ts(sp, g->map_sae, g->map_lat, se, &g->launches);
  CU(cudaEventRecord(g->k1_end[slot], se));
  g->k1_slot_valid[slot] = 1;
  // ---- pyramids of all cameras
  CU(cudaStreamWaitEvent(s_pyr, g->k1_end[slot], 0));
  launch_pyramids(fe->pd, imgs, 2 * S, s_pyr, &g->launches);
  CU(cudaGetLastError());
  CU(cudaEventRecord(g->p_done[slot], s_pyr));
  for (int i = 0; i < S; ++i) {
    esvio_fe* m = g->m[i];
    m->ts_sel[0] = m->pyr[w[i].cur];
    m->ts_sel[1] = m->pyr[w[i].rcur];
    m->pev_valid[w[i].slot] = 0;
    if ((rc = submit_tracking(m, w[i], ev[2 * i], cur_time[i], pub_this_frame[i], g->k1_end[slot],
                              g->p_done[slot])) != ESVIO_FE_OK)
      return rc;
  }
  return ESVIO_FE_OK;
}

FE_API int esvio_fe_group_track_wait(esvio_fe_group* g, esvio_tracks* out) {
  if (!g || !out) return ESVIO_FE_EINVAL;
  const int slot = g->m[0]->q_head;
  for (int i = 0; i < g->S; ++i) {
    const int rc = esvio_fe_track_wait(g->m[i], &out[i]);
    if (rc != ESVIO_FE_OK) return rc;
  }
  if (g->k1_slot_valid[slot]) {
    g->k1_ms_valid = cudaEventElapsedTime(&g->k1_ms, g->k1_beg[slot], g->k1_end[slot]) == cudaSuccess;
    g->k1_slot_valid[slot] = 0;
  }
  return ESVIO_FE_OK;
}

FE_API int esvio_fe_group_track(esvio_fe_group* g, const double* cur_time, const esvio_events* left,
                                const esvio_events* right, const int32_t* pub_this_frame,
                                esvio_tracks* out) {
  const int rc = esvio_fe_group_track_submit(g, cur_time, left, right, pub_this_frame);
  if (rc != ESVIO_FE_OK) return rc;
  return esvio_fe_group_track_wait(g, out);
}

FE_API int esvio_fe_group_kernel_launches(esvio_fe_group* g, int64_t* count) {
  if (!g || !count) return ESVIO_FE_EINVAL;
  int64_t n = g->launches;
  for (int i = 0; i < g->S; ++i) n += g->m[i]->launches;
  *count = n;
  return ESVIO_FE_OK;
}

FE_API int esvio_fe_group_sae_ts_ms(esvio_fe_group* g, float* ms) {
  if (!g || !ms) return ESVIO_FE_EINVAL;
  if (!g->k1_ms_valid) return ESVIO_FE_ESTATE;
  *ms = g->k1_ms;
  return ESVIO_FE_OK;
}

// ---------------------------------------------------------------------------------------
// memory helpers / plumbing / profiling
// ---------------------------------------------------------------------------------------
FE_API void* esvio_fe_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}
FE_API void esvio_fe_host_free(void* p) {
  if (p) cudaFreeHost(p);
}
FE_API int esvio_fe_device_alloc(esvio_fe* fe, size_t bytes, void** out) {
  if (!fe || !out) return ESVIO_FE_EINVAL;
  CU(cudaSetDevice(fe->dev));
  CU(cudaMalloc(out, bytes ? bytes : 1));
  return ESVIO_FE_OK;
}
FE_API int esvio_fe_device_free(esvio_fe* fe, void* p) {
  if (!fe) return ESVIO_FE_EINVAL;
  CU(cudaSetDevice(fe->dev));
  CU(cudaFree(p));
  return ESVIO_FE_OK;
}
FE_API int esvio_fe_copy_to_device(esvio_fe* fe, void* dst, const void* src, size_t bytes) {
  if (!fe || !dst || !src) return ESVIO_FE_EINVAL;
  CU(cudaSetDevice(fe->dev));
  CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, fe->stream));
  CU(cudaStreamSynchronize(fe->stream));
  return ESVIO_FE_OK;
}
FE_API int esvio_fe_result_device_ptr(esvio_fe* fe, void** ptr, size_t* bytes) {
  if (!fe || !ptr || !bytes) return ESVIO_FE_EINVAL;
  const int slot = fe->last_slot < 0 ? 0 : fe->last_slot;
  *ptr = fe->tb.result + (size_t)slot * fe->result_words;
  *bytes = fe->result_words * 4;
  return ESVIO_FE_OK;
}

FE_API int esvio_fe_result_acquire(esvio_fe* fe, void* consumer_stream, void** ptr, size_t* bytes) {
  if (!fe || !ptr || !bytes) return ESVIO_FE_EINVAL;
  if (fe->last_slot < 0) return fail(fe, ESVIO_FE_ESTATE, "no window submitted yet", cudaSuccess);
  CU(cudaSetDevice(fe->dev));
  const int slot = fe->last_slot;
  CU(cudaStreamWaitEvent((cudaStream_t)consumer_stream, fe->q_done[slot], 0));
  *ptr = fe->tb.result + (size_t)slot * fe->result_words;
  *bytes = fe->result_words * 4;
  return ESVIO_FE_OK;
}

FE_API int esvio_fe_result_release(esvio_fe* fe, void* consumer_stream) {
  if (!fe) return ESVIO_FE_EINVAL;
  if (fe->last_slot < 0) return fail(fe, ESVIO_FE_ESTATE, "no window submitted yet", cudaSuccess);
  CU(cudaSetDevice(fe->dev));
  CU(cudaEventRecord(fe->r_free[fe->last_slot], (cudaStream_t)consumer_stream));
  fe->r_held[fe->last_slot] = 1;
  return ESVIO_FE_OK;
}
// ---- the replica mode's one collective: every rank's packed track block of
// a publish window, all-gathered so that any rank (or rank 0's adapter) can publish all clouds
FE_API int esvio_fe_nccl_unique_id(void* id128) {
  if (!id128) return ESVIO_FE_EINVAL;
  const NcclApi* n = nccl_api();
  if (!n) return ESVIO_FE_ENODEV;
  return n->GetUniqueId((NcclUniqueIdBytes*)id128) == 0 ? ESVIO_FE_OK : ESVIO_FE_ECUDA;
}

static int comm_setup(esvio_fe* fe, void* comm, int owned, int rank, int world) {
  comm_release(fe);
  fe->nccl_comm = comm;
  fe->comm_owned = owned;
  fe->comm_rank = rank;
  fe->comm_world = world;
  fe->gather_seq = 0;
  int prio_lo = 0, prio_hi = 0;
  CU(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  CU(cudaStreamCreateWithPriority(&fe->stream_g, cudaStreamNonBlocking, prio_hi));
  for (int b = 0; b < 2; ++b) {
    CU(cudaMalloc(&fe->gathered[b], fe->result_words * 4 * (size_t)world));
    CU(cudaMemset(fe->gathered[b], 0, fe->result_words * 4 * (size_t)world));
  }
  return ESVIO_FE_OK;
}

FE_API int esvio_fe_comm_init(esvio_fe* fe, const void* id128, int32_t rank, int32_t world) {
  if (!fe || !id128 || world < 1 || rank < 0 || rank >= world) return ESVIO_FE_EINVAL;
  const NcclApi* n = nccl_api();
  if (!n) return fail(fe, ESVIO_FE_ENODEV, "libnccl.so.2 not found", cudaSuccess);
  CU(cudaSetDevice(fe->dev));
  NcclUniqueIdBytes id;
  memcpy(&id, id128, sizeof(id));
  void* comm = nullptr;
  const int r = n->CommInitRank(&comm, world, id, rank);
  if (r != 0) {
    snprintf(fe->err, sizeof(fe->err), "ncclCommInitRank: %s", n->GetErrorString ? n->GetErrorString(r) : "?");
    return ESVIO_FE_ECUDA;
  }
  return comm_setup(fe, comm, 1, rank, world);
}

FE_API int esvio_fe_comm_attach(esvio_fe* fe, void* nccl_comm, int32_t rank, int32_t world) {
  if (!fe || !nccl_comm || world < 1 || rank < 0 || rank >= world) return ESVIO_FE_EINVAL;
  if (!nccl_api()) return fail(fe, ESVIO_FE_ENODEV, "libnccl.so.2 not found", cudaSuccess);
  CU(cudaSetDevice(fe->dev));
  return comm_setup(fe, nccl_comm, 0, rank, world);
}

FE_API int esvio_fe_allgather_tracks(esvio_fe* fe) {
  if (!fe) return ESVIO_FE_EINVAL;
  if (!fe->nccl_comm) return fail(fe, ESVIO_FE_ESTATE, "no communicator (esvio_fe_comm_init)", cudaSuccess);
  if (fe->last_slot < 0) return fail(fe, ESVIO_FE_ESTATE, "no window submitted yet", cudaSuccess);
  CU(cudaSetDevice(fe->dev));
  const int slot = fe->last_slot;
  // behind the window's packing, off the tracking streams; the slot's block is not rewritten
  // (a pipeline depth later) before the collective has read it
  CU(cudaStreamWaitEvent(fe->stream_g, fe->q_done[slot], 0));
  const int r = nccl_api()->AllGather(fe->tb.result + (size_t)slot * fe->result_words,
                                      fe->gathered[fe->gather_seq & 1], fe->result_words * 4, /*ncclInt8*/ 0,
                                      fe->nccl_comm, fe->stream_g);
  if (r != 0) {
    snprintf(fe->err, sizeof(fe->err), "ncclAllGather: %s",
             nccl_api()->GetErrorString ? nccl_api()->GetErrorString(r) : "?");
    return ESVIO_FE_ECUDA;
  }
  CU(cudaEventRecord(fe->r_free[slot], fe->stream_g));
  fe->r_held[slot] = 1;
  fe->gather_seq++;
  return ESVIO_FE_OK;
}

FE_API int esvio_fe_gathered_tracks(esvio_fe* fe, void** dev_blocks, size_t* bytes_per_rank, void** cuda_stream) {
  if (!fe || !dev_blocks || !bytes_per_rank || !cuda_stream) return ESVIO_FE_EINVAL;
  if (!fe->nccl_comm || fe->gather_seq == 0)
    return fail(fe, ESVIO_FE_ESTATE, "no all-gather enqueued yet", cudaSuccess);
  *dev_blocks = fe->gathered[(fe->gather_seq - 1) & 1];
  *bytes_per_rank = fe->result_words * 4;
  *cuda_stream = (void*)fe->stream_g;
  return ESVIO_FE_OK;
}

FE_API int esvio_fe_stream(esvio_fe* fe, void** cuda_stream) {
  if (!fe || !cuda_stream) return ESVIO_FE_EINVAL;
  *cuda_stream = (void*)fe->stream;
  return ESVIO_FE_OK;
}
FE_API int esvio_fe_set_profiling(esvio_fe* fe, int32_t on) {
  if (!fe) return ESVIO_FE_EINVAL;
  fe->profiling = on != 0;
  fe->stage_ms_valid = 0;
  if (on) {
    CU(cudaSetDevice(fe->dev));
    CU(cudaEventRecord(fe->pev_ref, fe->stream));
  }
  return ESVIO_FE_OK;
}
FE_API int esvio_fe_get_stage_marks(esvio_fe* fe, float* ms) {
  if (!fe || !ms) return ESVIO_FE_EINVAL;
  if (!fe->stage_ms_valid) return fail(fe, ESVIO_FE_ESTATE, "no profiled window", cudaSuccess);
  memcpy(ms, fe->stage_marks, sizeof(fe->stage_marks));
  return ESVIO_FE_OK;
}
FE_API int esvio_fe_get_stage_ms(esvio_fe* fe, float* ms) {
  if (!fe || !ms) return ESVIO_FE_EINVAL;
  if (!fe->stage_ms_valid) return fail(fe, ESVIO_FE_ESTATE, "no profiled window", cudaSuccess);
  memcpy(ms, fe->stage_ms, sizeof(fe->stage_ms));
  return ESVIO_FE_OK;
}
FE_API int esvio_fe_kernel_launches(esvio_fe* fe, int64_t* count) {
  if (!fe || !count) return ESVIO_FE_EINVAL;
  *count = fe->launches;
  return ESVIO_FE_OK;
}

// ---------------------------------------------------------------------------------------
// stage-level entry points (parity tests)
// ---------------------------------------------------------------------------------------
FE_API int esvio_fe_get_sae(esvio_fe* fe, int32_t cam, int32_t plane, double* dst) {
  if (!fe || !dst || cam < 0 || cam > 1 || plane < 0 || plane > 3) return ESVIO_FE_EINVAL;
  CU(cudaSetDevice(fe->dev));
  if (sync_all(fe) != ESVIO_FE_OK) return ESVIO_FE_ECUDA;
  const double2* base = (plane < 2 ? fe->sae : fe->lat) + (size_t)cam * fe->npx;
  const char* src = (const char*)base + (plane & 1) * sizeof(double);
  // strided gather of one double per pixel
  CU(cudaMemcpy2D(dst, sizeof(double), src, sizeof(double2), sizeof(double), fe->npx,
                  cudaMemcpyDeviceToHost));
  return ESVIO_FE_OK;
}

FE_API int esvio_fe_stage_update(esvio_fe* fe, double t_ref, const esvio_events* left,
                                 const esvio_events* right) {
  return esvio_fe_stage_update_mc(fe, t_ref, left, right, nullptr);
}

FE_API int esvio_fe_stage_motion_correct(esvio_fe* fe, const esvio_motion* mc, const float* xy_dt,
                                         int32_t n, int32_t* out_xy) {
  if (!fe || !mc || !xy_dt || !out_xy || n < 0) return ESVIO_FE_EINVAL;
  if (n == 0) return ESVIO_FE_OK;
  CU(cudaSetDevice(fe->dev));
  cudaStream_t s = fe->stream;
  float* d_in = nullptr;
  int* d_out = nullptr;
  CU(cudaMalloc(&d_in, sizeof(float) * 3 * (size_t)n));
  cudaError_t ce = cudaMalloc(&d_out, sizeof(int) * 2 * (size_t)n);
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(d_in, xy_dt, sizeof(float) * 3 * (size_t)n, cudaMemcpyHostToDevice, s);
  if (ce == cudaSuccess) {
    launch_warp_points(mc_params(fe, mc), d_in, n, d_out, s, &fe->launches);
    ce = cudaMemcpyAsync(out_xy, d_out, sizeof(int) * 2 * (size_t)n, cudaMemcpyDeviceToHost, s);
  }
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);
  cudaFree(d_in);
  cudaFree(d_out);
  CU(ce);
  return ESVIO_FE_OK;
}

FE_API int esvio_fe_stage_update_mc(esvio_fe* fe, double t_ref, const esvio_events* left,
                                    const esvio_events* right, const esvio_motion* mc) {
  if (!fe) return ESVIO_FE_EINVAL;
  if (fe->group) return fail(fe, ESVIO_FE_ESTATE, "handle belongs to a group", cudaSuccess);
  if (mc && !fe->cfg.do_motion_correction)
    return fail(fe, ESVIO_FE_ESTATE, "motion compensation needs config.do_motion_correction", cudaSuccess);
  if (fe->q_count != 0) return fail(fe, ESVIO_FE_ESTATE, "windows in flight", cudaSuccess);
  CU(cudaSetDevice(fe->dev));
  DevEvents ev[2];
  int rc;
  if (sync_all(fe) != ESVIO_FE_OK) return ESVIO_FE_ECUDA;
  if ((rc = stage_events(fe, 0, 0, left, &ev[0])) != ESVIO_FE_OK) return rc;
  if ((rc = stage_events(fe, 0, 1, right, &ev[1])) != ESVIO_FE_OK) return rc;
  if ((rc = staging_done(fe, 0)) != ESVIO_FE_OK) return rc;
  if ((rc = run_event_stage(fe, 0, t_ref, ev, fe->cur_left, fe->cur_right, mc)) != ESVIO_FE_OK) return rc;
  return sync_all(fe);
}

FE_API int esvio_fe_stage_corner_flags(esvio_fe* fe, const esvio_events* left, int32_t and_ts_test,
                                       uint8_t* flags) {
  if (!fe || !left || !flags) return ESVIO_FE_EINVAL;
  if (fe->q_count != 0) return fail(fe, ESVIO_FE_ESTATE, "windows in flight", cudaSuccess);
  CU(cudaSetDevice(fe->dev));
  DevEvents ev;
  int rc;
  if (sync_all(fe) != ESVIO_FE_OK) return ESVIO_FE_ECUDA;
  if ((rc = stage_events(fe, 0, 0, left, &ev)) != ESVIO_FE_OK) return rc;
  if ((rc = staging_done(fe, 0)) != ESVIO_FE_OK) return rc;
  CU(cudaStreamWaitEvent(fe->stream_e, fe->c_done[0], 0));
  launch_corner_flags(corner_params(fe, fe->cur_left, and_ts_test, 0), ev, fe->flags[0],
                      fe->stream_e, &fe->launches);
  CU(cudaGetLastError());
  if (ev.n > 0)
    CU(cudaMemcpyAsync(flags, fe->flags[0], ev.n, cudaMemcpyDeviceToHost, fe->stream_e));
  CU(cudaStreamSynchronize(fe->stream_e));
  return ESVIO_FE_OK;
}

FE_API int esvio_fe_get_pyramid_level(esvio_fe* fe, int32_t which, int32_t level, uint8_t* dst,
                                      int32_t* w, int32_t* h) {
  if (!fe || which < 0 || which > 2 || level < 0 || level >= fe->pd.levels) return ESVIO_FE_EINVAL;
  CU(cudaSetDevice(fe->dev));
  if (w) *w = fe->pd.w[level];
  if (h) *h = fe->pd.h[level];
  if (!dst) return ESVIO_FE_OK;
  if (sync_all(fe) != ESVIO_FE_OK) return ESVIO_FE_ECUDA;
  const int idx = which == 0 ? fe->cur_left : (which == 1 ? fe->cur_right : fe->prev_left);
  CU(cudaMemcpy2DAsync(dst, fe->pd.w[level], fe->pyr[idx] + fe->pd.off[level], fe->pd.pitch[level],
                       fe->pd.w[level], fe->pd.h[level], cudaMemcpyDeviceToHost, fe->stream));
  CU(cudaStreamSynchronize(fe->stream));
  return ESVIO_FE_OK;
}

FE_API int esvio_fe_stage_lk(esvio_fe* fe, const uint8_t* prev_img, const uint8_t* next_img,
                             const float* prev_pts, float* next_pts, int32_t n, uint8_t* status,
                             int32_t max_level, int32_t use_initial_flow) {
  if (!fe || !prev_img || !next_img || !prev_pts || !next_pts || !status || n < 0 ||
      n > kMaxCnt || max_level < 0)
    return ESVIO_FE_EINVAL;
  if (n == 0) return ESVIO_FE_OK;
  CU(cudaSetDevice(fe->dev));
  cudaStream_t s = fe->stream;
  CU(cudaMemcpy2DAsync(fe->pyr[kScratchBase], fe->pd.pitch[0], prev_img, fe->W, fe->W, fe->H,
                       cudaMemcpyHostToDevice, s));
  CU(cudaMemcpy2DAsync(fe->pyr[kScratchBase + 1], fe->pd.pitch[0], next_img, fe->W, fe->W, fe->H,
                       cudaMemcpyHostToDevice, s));
  uint8_t* imgs[2] = {fe->pyr[kScratchBase], fe->pyr[kScratchBase + 1]};
  launch_pyramids(fe->pd, imgs, 2, s, &fe->launches);
  CU(cudaMemcpyAsync(fe->d_scratch_n, &n, sizeof(int), cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(fe->d_scratch_p0, prev_pts, sizeof(float2) * n, cudaMemcpyHostToDevice, s));
  if (use_initial_flow)
    CU(cudaMemcpyAsync(fe->d_scratch_p1, next_pts, sizeof(float2) * n, cudaMemcpyHostToDevice, s));
  launch_lk(fe->pd, fe->pyr[kScratchBase], fe->pyr[kScratchBase + 1], fe->d_scratch_p0, fe->d_scratch_p1, fe->d_scratch_st,
            nullptr, nullptr, fe->d_scratch_n, n, max_level, use_initial_flow, 0, s, &fe->launches);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(next_pts, fe->d_scratch_p1, sizeof(float2) * n, cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(status, fe->d_scratch_st, n, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return ESVIO_FE_OK;
}

FE_API int esvio_fe_stage_condition(esvio_fe* fe, const uint8_t* src, int32_t median_ksize,
                                    int32_t equalize, uint8_t* dst) {
  if (!fe || !src || !dst || median_ksize < 0 || median_ksize > 15 ||
      (median_ksize && !(median_ksize & 1)))
    return ESVIO_FE_EINVAL;
  if (!fe->aux[0][0]) return fail(fe, ESVIO_FE_ESTATE, "handle created without equalize / median", cudaSuccess);
  if (fe->q_count != 0) return fail(fe, ESVIO_FE_ESTATE, "windows in flight", cudaSuccess);
  CU(cudaSetDevice(fe->dev));
  if (sync_all(fe) != ESVIO_FE_OK) return ESVIO_FE_ECUDA;
  cudaStream_t s = fe->stream;
  const int pitch = fe->pd.pitch[0];
  CU(cudaMemcpy2DAsync(fe->aux[0][0], pitch, src, fe->W, fe->W, fe->H, cudaMemcpyHostToDevice, s));
  const uint8_t* cur = fe->aux[0][0];
  if (median_ksize > 1) {
    const uint8_t* in[2] = {cur, cur};
    uint8_t* out[2] = {fe->aux[1][0], fe->aux[1][0]};
    launch_median(in, out, 1, fe->W, fe->H, pitch, median_ksize, s, &fe->launches);
    cur = fe->aux[1][0];
  }
  if (equalize) {
    const uint8_t* in[2] = {cur, cur};
    uint8_t* tmp[2] = {fe->aux[2][0], fe->aux[2][0]};
    uint8_t* out[2] = {fe->aux[2][1], fe->aux[2][1]};
    launch_equalize(in, tmp, out, 1, fe->W, fe->H, pitch, fe->clahe_lut, fe->clahe_minmax, s,
                    &fe->launches);
    cur = fe->aux[2][1];
  }
  CU(cudaGetLastError());
  CU(cudaMemcpy2DAsync(dst, fe->W, cur, pitch, fe->W, fe->H, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return ESVIO_FE_OK;
}

FE_API int esvio_fe_stage_fmat_mask(esvio_fe* fe, const float* p1, const float* p2, int32_t n,
                                    double thresh, uint8_t* mask, int32_t* iters) {
  if (!fe || !p1 || !p2 || !mask || n < 0 || n > kMaxCnt) return ESVIO_FE_EINVAL;
  if (n == 0) return ESVIO_FE_OK;
  CU(cudaSetDevice(fe->dev));
  cudaStream_t s = fe->stream;
  CU(cudaMemcpyAsync(fe->d_scratch_p0, p1, sizeof(float2) * n, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(fe->d_scratch_p1, p2, sizeof(float2) * n, cudaMemcpyHostToDevice, s));
  launch_ransac_stage(fe->tp, fe->tb, fe->d_scratch_p0, fe->d_scratch_p1, n, thresh,
                      fe->d_scratch_st, fe->d_scratch_n, s, &fe->launches);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(mask, fe->d_scratch_st, n, cudaMemcpyDeviceToHost, s));
  int it = 0;
  CU(cudaMemcpyAsync(&it, fe->d_scratch_n, sizeof(int), cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  if (iters) *iters = it;
  return ESVIO_FE_OK;
}

FE_API int esvio_fe_stage_select(esvio_fe* fe, const esvio_events* left, int32_t n,
                                 const float* pts, const int32_t* ids, const int32_t* track_cnt,
                                 int32_t* n_out, float* pts_out, int32_t* ids_out,
                                 int32_t* track_cnt_out, int32_t* n_kept) {
  if (!fe || !left || n < 0 || n > fe->cfg.max_cnt || !n_out || !pts_out || !ids_out ||
      !track_cnt_out)
    return ESVIO_FE_EINVAL;
  if (n > 0 && (!pts || !ids || !track_cnt)) return ESVIO_FE_EINVAL;
  if (fe->q_count != 0) return fail(fe, ESVIO_FE_ESTATE, "windows in flight", cudaSuccess);
  CU(cudaSetDevice(fe->dev));
  cudaStream_t s = fe->stream;
  const TrackBuffers& B = fe->tb;
  DevEvents ev;
  int rc;
  if (sync_all(fe) != ESVIO_FE_OK) return ESVIO_FE_ECUDA;
  if ((rc = stage_events(fe, 0, 0, left, &ev)) != ESVIO_FE_OK) return rc;
  CU(cudaStreamSynchronize(fe->stream_c));
  TrackState st;
  CU(cudaMemcpyAsync(&st, B.st, sizeof(st), cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  st.n_cur = n;
  CU(cudaMemcpyAsync(B.st, &st, sizeof(st), cudaMemcpyHostToDevice, s));
  if (n > 0) {
    CU(cudaMemcpyAsync(B.cur_pts, pts, sizeof(float2) * n, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(B.ids, ids, sizeof(int) * n, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(B.cnt, track_cnt, sizeof(int) * n, cudaMemcpyHostToDevice, s));
  }
  launch_corner_flags(corner_params(fe, fe->cur_left, 1, 0), ev, fe->flags[0], s, &fe->launches);
  launch_select(fe->tp, B, ev.n, fe->cand[0], fe->cand_cnt[0], -1, s, &fe->launches);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(&st, B.st, sizeof(st), cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  *n_out = st.n_cur;
  if (n_kept) *n_kept = st.stat_after_mask;
  if (st.n_cur > 0) {
    CU(cudaMemcpyAsync(pts_out, B.cur_pts, sizeof(float2) * st.n_cur, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(ids_out, B.ids, sizeof(int) * st.n_cur, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(track_cnt_out, B.cnt, sizeof(int) * st.n_cur, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
  }
  return ESVIO_FE_OK;
}

FE_API int esvio_fe_stage_set_tracks(esvio_fe* fe, double prev_time, int32_t next_id, int32_t n,
                                     const float* pts, const int32_t* ids, const int32_t* track_cnt,
                                     const float* un, int32_t n_r, const int32_t* ids_r,
                                     const float* un_r) {
  if (!fe || n < 0 || n > fe->cfg.max_cnt || n_r < 0 || n_r > fe->cfg.max_cnt) return ESVIO_FE_EINVAL;
  if (n > 0 && (!pts || !ids || !track_cnt || !un)) return ESVIO_FE_EINVAL;
  if (n_r > 0 && (!ids_r || !un_r)) return ESVIO_FE_EINVAL;
  if (fe->q_count != 0) return fail(fe, ESVIO_FE_ESTATE, "windows in flight", cudaSuccess);
  CU(cudaSetDevice(fe->dev));
  if (sync_all(fe) != ESVIO_FE_OK) return ESVIO_FE_ECUDA;
  cudaStream_t s = fe->stream;
  const TrackBuffers& B = fe->tb;
  TrackState st;
  CU(cudaMemcpyAsync(&st, B.st, sizeof(st), cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  st.n_prev = st.n_cur = n;
  st.n_right = n_r;
  st.next_id = next_id;
  st.n_prev_un = n;
  st.n_prev_un_r = n_r;
  CU(cudaMemcpyAsync(B.st, &st, sizeof(st), cudaMemcpyHostToDevice, s));
  if (n > 0) {
    CU(cudaMemcpyAsync(B.prev_pts, pts, sizeof(float2) * n, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(B.ids, ids, sizeof(int) * n, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(B.cnt, track_cnt, sizeof(int) * n, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(B.prev_un_ids, ids, sizeof(int) * n, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(B.prev_un, un, sizeof(float2) * n, cudaMemcpyHostToDevice, s));
  }
  if (n_r > 0) {
    CU(cudaMemcpyAsync(B.prev_un_r_ids, ids_r, sizeof(int) * n_r, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(B.prev_un_r, un_r, sizeof(float2) * n_r, cudaMemcpyHostToDevice, s));
  }
  CU(cudaStreamSynchronize(s));
  fe->prev_time = prev_time;
  return ESVIO_FE_OK;
}

FE_API int esvio_fe_stage_sort_order(esvio_fe* fe, const int32_t* key, int32_t n, int32_t depth_limit,
                                     int32_t* order) {
  // depth_limit: the replay keeps one pending range per level (32 slots); the library's own budget is <= 20
  if (!fe || n < 0 || n > kMaxCnt || depth_limit > 30 || (n > 0 && (!key || !order))) return ESVIO_FE_EINVAL;
  if (n == 0) return ESVIO_FE_OK;
  CU(cudaSetDevice(fe->dev));
  cudaStream_t s = fe->stream;
  int* d_key = reinterpret_cast<int*>(fe->d_scratch_p0);   // 2 * kMaxCnt float2: room for both
  int* d_order = d_key + kMaxCnt;
  CU(cudaMemcpyAsync(d_key, key, sizeof(int) * n, cudaMemcpyHostToDevice, s));
  launch_sort_order(d_key, n, depth_limit, d_order, s, &fe->launches);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(order, d_order, sizeof(int) * n, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return ESVIO_FE_OK;
}

FE_API int esvio_fe_stage_undistort(esvio_fe* fe, int32_t cam, const float* uv, int32_t n,
                                    float* out) {
  if (!fe || cam < 0 || cam > 1 || !uv || !out || n < 0 || n > kMaxCnt) return ESVIO_FE_EINVAL;
  if (n == 0) return ESVIO_FE_OK;
  CU(cudaSetDevice(fe->dev));
  cudaStream_t s = fe->stream;
  CU(cudaMemcpyAsync(fe->d_scratch_p0, uv, sizeof(float2) * n, cudaMemcpyHostToDevice, s));
  launch_undistort(fe->tp.cam[cam], fe->d_scratch_p0, n, fe->d_scratch_p1, s, &fe->launches);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(out, fe->d_scratch_p1, sizeof(float2) * n, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return ESVIO_FE_OK;
}
