// select.cu -- the per-window bookkeeping of FeatureTracker::trackEvent
// (feature_tracker/src/feature_tracker.cpp:340-603) that sits between the LK launches.
// All of it is tiny and order-dependent, so each step is ONE CTA working on device-resident
// counters (TrackState): nothing returns to the host until the packed result is copied out.
#include "common.cuh"

namespace esvio {

// inBorder_event (feature_tracker.cpp:48-54)
__device__ __forceinline__ bool in_border(int W, int H, float2 p) {
  const int ix = cv_round(p.x), iy = cv_round(p.y);
  return 1 <= ix && ix < W - 1 && 1 <= iy && iy < H - 1;
}

// FeatureTracker::distance (feature_tracker.cpp:1314-1319): float differences, double norm
__device__ __forceinline__ double pt_dist(float2 a, float2 b) {
  const double dx = (double)(a.x - b.x), dy = (double)(a.y - b.y);
  return sqrt(dx * dx + dy * dy);
}

// PinholeCamera::liftProjective (camera_model/src/camera_models/PinholeCamera.cc:450-510)
// with PinholeCamera::distortion (:646-662): 8 fixed-point iterations, fp64.
__device__ __forceinline__ void lift_projective(const Pinhole& c, double u, double v, double& ox,
                                                double& oy) {
  const double inv_fx = 1.0 / c.fx, inv_fy = 1.0 / c.fy;
  const double off_x = -c.cx / c.fx, off_y = -c.cy / c.fy;
  const double xd = inv_fx * u + off_x, yd = inv_fy * v + off_y;
  double xu = xd, yu = yd;
  if (!(c.k1 == 0.0 && c.k2 == 0.0 && c.p1 == 0.0 && c.p2 == 0.0)) {
    for (int it = 0; it < 8; ++it) {
      const double xx = xu * xu, yy = yu * yu, xy = xu * yu;
      const double r2 = xx + yy;
      const double rad = c.k1 * r2 + c.k2 * r2 * r2;
      const double ddx = xu * rad + 2.0 * c.p1 * xy + c.p2 * (r2 + 2.0 * xx);
      const double ddy = yu * rad + 2.0 * c.p2 * xy + c.p1 * (r2 + 2.0 * yy);
      xu = xd - ddx;
      yu = yd - ddy;
    }
  }
  ox = xu;
  oy = yu;
}

// order-preserving compaction offsets for up to blockDim.x flags (one per thread);
// returns the exclusive prefix, total in *total (shared)
__device__ int block_excl_scan_1024(int flag, int* s_warp /*33 ints*/, int* total) {
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const unsigned bal = __ballot_sync(0xffffffffu, flag);
  const int in_warp = __popc(bal & ((1u << lane) - 1u));
  if (lane == 0) s_warp[warp] = __popc(bal);
  __syncthreads();
  if (warp == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    int v = lane < nw ? s_warp[lane] : 0;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    s_warp[lane] = incl - v;
    if (lane == 31) s_warp[32] = incl;
  }
  __syncthreads();
  const int r = s_warp[warp] + in_warp;
  *total = s_warp[32];
  __syncthreads();
  return r;
}

// End of the temporal/selection stage: freeze what the stereo stage needs and roll
// prev_pts = cur_pts (feature_tracker.cpp:586) so the next window's temporal LK can start.
// Called by all 1024 threads of the stage's last kernel, after a __syncthreads() behind the
// writes of cur_pts / ids / cnt and of the counters.
__device__ __forceinline__ void snapshot_tracks(const TrackParams& P, const TrackBuffers& B, int slot,
                                                int n) {
  TrackState* st = B.st;
  const int i = threadIdx.x;
  const int M = P.max_cnt;
  if (i < n) {
    const float2 cp = B.cur_pts[i];
    B.snap_pts[slot * M + i] = cp;
    B.snap_ids[slot * M + i] = B.ids[i];
    B.snap_cnt[slot * M + i] = B.cnt[i];
    B.prev_pts[i] = cp;
  }
  if (i == 0) {
    int* h = B.snap_hdr + slot * 16;
    h[0] = n;
    h[1] = st->stat_n_prev;
    h[2] = st->stat_after_temporal;
    h[3] = st->stat_after_ransac;
    h[4] = st->stat_after_mask;
    h[5] = st->stat_new;
    h[6] = st->stat_corner_flags;
    h[7] = st->stat_ransac_iters;
    h[8] = st->next_id;
    st->n_prev = n;
  }
}

// ------------------------------------------------------------------------------------
// after temporal forward + backward LK (feature_tracker.cpp:419-440)
// ------------------------------------------------------------------------------------
// snap_slot >= 0: nothing else follows in this window's temporal stage (not a publish window),
// so the snapshot is taken here as well
__global__ void __launch_bounds__(1024) k_post_temporal(TrackParams P, TrackBuffers B, int snap_slot) {
  PDL_PROLOGUE();
  __shared__ int s_warp[33];
  TrackState* st = B.st;
  const int n = st->n_prev;
  const int i = threadIdx.x;
  int keep = 0;
  float2 pp, cp;
  int id = 0, cnt = 0;
  if (i < n) {
    pp = B.prev_pts[i];
    cp = B.cur_pts[i];
    id = B.ids[i];
    cnt = B.cnt[i];
    keep = B.st_fwd[i] != 0;
    if (P.flow_back) keep = keep && B.st_bwd[i] && pt_dist(pp, B.rev_pts[i]) <= 0.5;
    if (keep && !in_border(P.W, P.H, cp)) keep = 0;
  }
  int total;
  const int pos = block_excl_scan_1024(keep, s_warp, &total);
  if (keep) {
    B.prev_pts[pos] = pp;
    B.cur_pts[pos] = cp;
    B.ids[pos] = id;
    B.cnt[pos] = cnt + 1;  // for (auto &n : track_cnt) n++;
  }
  if (i == 0) {
    st->stat_n_prev = n;
    st->n_cur = total;
    st->stat_after_temporal = total;
    st->stat_after_ransac = total;
    st->stat_after_mask = total;
    st->stat_new = 0;
    st->stat_ransac_iters = 0;
  }
  if (snap_slot >= 0) {
    __syncthreads();
    snapshot_tracks(P, B, snap_slot, total);
  }
}

void launch_post_temporal(const TrackParams& P, const TrackBuffers& B, int snap_slot,
                          cudaStream_t s, int64_t* launches) {
  launch_pdl(k_post_temporal, dim3(1), dim3(1024), 0, s, P, B, snap_slot);
  ++*launches;
}

// ------------------------------------------------------------------------------------
// Event_setMask + Event_FeaturesToTrack + id assignment (feature_tracker.cpp:123-151,
// 13-38, 446-468).  The W x H mask is one bit per pixel in shared memory.
// ------------------------------------------------------------------------------------
constexpr int kMaxDiscR = 64;

// half widths of OpenCV's filled circle (drawing.cpp Circle, fill): row cy+-k spans
// [cx - hw[k], cx + hw[k]]
__device__ void disc_half_widths(int r, int* hw) {
  for (int k = 0; k <= r; ++k) hw[k] = -1;
  int err = 0, dx = r, dy = 0, plus = 1, minus = (r << 1) - 1;
  while (dx >= dy) {
    if (dx > hw[dy]) hw[dy] = dx;
    if (dy > hw[dx]) hw[dx] = dy;
    dy++;
    err += plus;
    plus += 2;
    const int m = (err <= 0) - 1;
    err -= minus & m;
    dx += m;
    minus -= m & 2;
  }
}

// executed by one full warp: rows are distributed over lanes, so plain ORs do not collide
__device__ __forceinline__ void fill_disc_warp(uint32_t* mask, int words, int W, int H, int cx,
                                               int cy, int r, const int* hw) {
  for (int k = lane_id() - r; k <= r; k += 32) {
    const int yy = cy + k;
    if (yy < 0 || yy >= H) continue;
    const int h = hw[k < 0 ? -k : k];
    if (h < 0) continue;
    int x0 = cx - h, x1 = cx + h;
    if (x0 < 0) x0 = 0;
    if (x1 > W - 1) x1 = W - 1;
    if (x0 > x1) continue;
    uint32_t* row = mask + (size_t)yy * words;
    const int w0 = x0 >> 5, w1 = x1 >> 5;
    for (int w = w0; w <= w1; ++w) {
      uint32_t bits = 0xffffffffu;
      if (w == w0) bits &= 0xffffffffu << (x0 & 31);
      if (w == w1) bits &= 0xffffffffu >> (31 - (x1 & 31));
      row[w] |= bits;
    }
  }
  __syncwarp();
}

__device__ __forceinline__ bool mask_test(const uint32_t* mask, int words, int x, int y) {
  return (mask[(size_t)y * words + (x >> 5)] >> (x & 31)) & 1u;
}

size_t select_smem_bytes(int W, int H) { return (size_t)H * ((W + 31) / 32) * sizeof(uint32_t); }

#ifdef ESVIO_LK_CLOCKS  // scratch builds only: phase clocks of the last k_select launch
__device__ long long g_sel_clk[16];
#define SEL_CLK(i) do { if (threadIdx.x == 0) g_sel_clk[i] = clock64(); } while (0)
#define SEL_ADD(i, v) do { if (threadIdx.x == 0) g_sel_clk[i] += (v); } while (0)
#define SEL_SET(i, v) do { if (threadIdx.x == 0) g_sel_clk[i] = (v); } while (0)
extern "C" __attribute__((visibility("default"))) int esvio_dbg_select_clocks(long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_sel_clk, sizeof(g_sel_clk));
}
#else
#define SEL_CLK(i)
#define SEL_ADD(i, v)
#define SEL_SET(i, v)
#endif

constexpr int kSelThreads = 1024;
constexpr int kSelPerThread = 4;
constexpr int kSelFast = 256;  // tracked points the bit-matrix path of Event_setMask handles

// the pixel (bx, by) lies inside OpenCV's filled circle of radius r around (ax, ay)
__device__ __forceinline__ bool in_disc(int ax, int ay, int bx, int by, int r, const int* hw) {
  const int dy = by > ay ? by - ay : ay - by;
  if (dy > r) return false;
  const int dx = bx > ax ? bx - ax : ax - bx;
  return dx <= hw[dy];
}

// any warp: rows over lanes, atomicOr because several warps fill at once
__device__ __forceinline__ void fill_disc_warp_atomic(uint32_t* mask, int words, int W, int H,
                                                      int cx, int cy, int r, const int* hw) {
  for (int k = lane_id() - r; k <= r; k += 32) {
    const int yy = cy + k;
    if (yy < 0 || yy >= H) continue;
    const int h = hw[k < 0 ? -k : k];
    if (h < 0) continue;
    int x0 = cx - h, x1 = cx + h;
    if (x0 < 0) x0 = 0;
    if (x1 > W - 1) x1 = W - 1;
    if (x0 > x1) continue;
    uint32_t* row = mask + (size_t)yy * words;
    const int w0 = x0 >> 5, w1 = x1 >> 5;
    for (int w = w0; w <= w1; ++w) {
      uint32_t bits = 0xffffffffu;
      if (w == w0) bits &= 0xffffffffu << (x0 & 31);
      if (w == w1) bits &= 0xffffffffu >> (31 - (x1 & 31));
      atomicOr(&row[w], bits);
    }
  }
}

// r <= 15: lane l owns row cy + l - r of the disc (half width my_hw, -1 = no row), which spans at
// most 31 pixels, i.e. one or two mask words.  Executed by one full warp.
__device__ __forceinline__ void fill_disc_rows(uint32_t* mask, int words, int W, int H, int cx,
                                               int cy, int my_k, int my_hw) {
  const int yy = cy + my_k;
  if (my_hw >= 0 && yy >= 0 && yy < H) {
    const int x0 = max(cx - my_hw, 0), x1 = min(cx + my_hw, W - 1);
    if (x0 <= x1) {
      uint32_t* row = mask + (size_t)yy * words;
      const int w0 = x0 >> 5, w1 = x1 >> 5;
      const uint32_t lo = 0xffffffffu << (x0 & 31), hi = 0xffffffffu >> (31 - (x1 & 31));
      if (w0 == w1) {
        row[w0] |= lo & hi;
      } else {
        row[w0] |= lo;
        row[w1] |= hi;
      }
    }
  }
  __syncwarp();
}


// ---- std::sort as libstdc++ runs it, as a permutation ---------------------------------------
// Event_setMask / Image_setMask visit the tracks in the order `sort(cnt_pts_id.begin(),
// cnt_pts_id.end(), a.first > b.first)` leaves them in (feature_tracker.cpp:100-103,132-135).
// Equal track counts are the rule (every corner of a publish window starts at 1 together), and
// std::sort is not stable: which of two close tracks of equal age survives the mask, and the
// order of the published features, depend on where the library's introsort puts the ties.  The
// reference is built with GCC, so its answer is libstdc++'s (bits/stl_algo.h): while a range
// holds more than 16 elements -- median of (first+1, mid, last-1) swapped to the front,
// unguarded Hoare partition around it, right part recursed into, left part continued, heap sort
// when the budget 2*floor(log2 n) is spent -- then ONE insertion sort over everything.
//
// Here: warp 0 replays the partitions on a permutation of indices.  A Hoare partition pairs the
// k-th element from the left that is not before the pivot (key <= pivot: a "left stopper") with
// the k-th element from the right that is not after it (key >= pivot) and swaps them while the
// left one lies left of the right one; the scans only ever read elements no swap has touched
// yet, so both stopper lists can be taken from the range as it is (ballots), all swaps done at
// once, and the cut is min(next left stopper, last swapped right position).  The final
// insertion sort is a STABLE sort of what the partitions left (its comparator is strict), i.e.
// a rank by (key descending, position ascending), one thread per element.
// scratch: 2 * kMaxCnt + 3 * 32 words.  Called by the whole CTA (>= n threads).
constexpr int kSortScratchWords = 2 * kMaxCnt + 96;

__device__ void sort_adjust_heap(const int* key, int* o, int hole, int len, int value) {
  const int top = hole;
  int child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (key[o[child]] > key[o[child - 1]]) child--;
    o[hole] = o[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    o[hole] = o[child - 1];
    hole = child - 1;
  }
  int parent = (hole - 1) / 2;  // __push_heap
  while (hole > top && key[o[parent]] > key[value]) {
    o[hole] = o[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  o[hole] = value;
}

// std::__partial_sort(first, last, last): __make_heap + __sort_heap (one thread; reached only
// when the partitions of a range stay lopsided 2*floor(log2 n) times in a row)
__device__ void sort_heap_range(const int* key, int* o, int len) {
  if (len >= 2)
    for (int parent = (len - 2) / 2;; --parent) {
      sort_adjust_heap(key, o, parent, len, o[parent]);
      if (parent == 0) break;
    }
  for (int last = len; last > 1;) {
    --last;
    const int value = o[last];
    o[last] = o[0];
    sort_adjust_heap(key, o, 0, last, value);
  }
}

__device__ void std_sort_order(const int* __restrict__ s_key, int n, int depth_limit, int* __restrict__ s_order,
                               uint32_t* __restrict__ scratch) {
  int* perm = reinterpret_cast<int*>(scratch);
  uint16_t* lo = reinterpret_cast<uint16_t*>(scratch + kMaxCnt);
  uint16_t* ro = lo + kMaxCnt;
  int* stack = reinterpret_cast<int*>(scratch + 2 * kMaxCnt);
  const int tid = threadIdx.x, lane = lane_id();
  const uint32_t lt_mask = (1u << lane) - 1u;
  if (tid < n) perm[tid] = tid;
  __syncthreads();
  if (n > 16 && tid < 32) {
    int first = 0, last = n, sp = 0;
    int depth = depth_limit >= 0 ? depth_limit : 2 * (31 - __clz(n));
    for (;;) {
      while (last - first > 16) {
        if (depth == 0) {
          if (lane == 0) sort_heap_range(s_key, perm + first, last - first);
          __syncwarp();
          break;
        }
        --depth;
        if (lane == 0) {  // __move_median_to_first(first, first + 1, mid, last - 1)
          const int a = first + 1, b = first + (last - first) / 2, c = last - 1;
          const int ka = s_key[perm[a]], kb = s_key[perm[b]], kc = s_key[perm[c]];
          int m;
          if (ka > kb) m = kb > kc ? b : (ka > kc ? c : a);
          else m = ka > kc ? a : (kb > kc ? c : b);
          const int t = perm[first];
          perm[first] = perm[m];
          perm[m] = t;
        }
        __syncwarp();
        const int pk = s_key[perm[first]];
        const int lo0 = first + 1, m = last - lo0;
        int nL = 0, nR = 0;
        for (int base = 0; base < m; base += 32) {
          const bool v = base + lane < m;
          const int i = lo0 + base + lane, j = last - 1 - base - lane;
          const bool is_l = v && s_key[perm[i]] <= pk;
          const bool is_r = v && s_key[perm[j]] >= pk;
          const uint32_t bl = __ballot_sync(0xffffffffu, is_l), br = __ballot_sync(0xffffffffu, is_r);
          if (is_l) lo[nL + __popc(bl & lt_mask)] = (uint16_t)i;
          if (is_r) ro[nR + __popc(br & lt_mask)] = (uint16_t)j;
          nL += __popc(bl);
          nR += __popc(br);
        }
        __syncwarp();
        const int kmax = min(nL, nR);
        int ks = 0;  // swaps: pairs k < ks
        for (int base = 0; base < kmax; base += 32) {
          const int k = base + lane;
          const uint32_t b = __ballot_sync(0xffffffffu, k < kmax && lo[k] < ro[k]);
          ks += __popc(b);
          if (b != 0xffffffffu) break;
        }
        int cut = ks < nL ? (int)lo[ks] : last;
        if (ks > 0) cut = min(cut, (int)ro[ks - 1]);
        for (int k = lane; k < ks; k += 32) {
          const int a = lo[k], b = ro[k];
          const int t = perm[a];
          perm[a] = perm[b];
          perm[b] = t;
        }
        if (lane == 0) {  // __introsort_loop(cut, last, depth) later; [first, cut) now
          stack[3 * sp] = cut;
          stack[3 * sp + 1] = last;
          stack[3 * sp + 2] = depth;
        }
        ++sp;
        last = cut;
        __syncwarp();
      }
      if (sp == 0) break;
      --sp;
      first = stack[3 * sp];
      last = stack[3 * sp + 1];
      depth = stack[3 * sp + 2];
      __syncwarp();
    }
  }
  __syncthreads();
  if (tid < n) {  // __final_insertion_sort
    const int c = s_key[perm[tid]];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const int cj = s_key[perm[j]];
      rank += (cj > c) || (cj == c && j < tid);
    }
    s_order[rank] = perm[tid];
  }
  __syncthreads();
}

// test entry (esvio_fe_stage_sort_order): the permutation alone
__global__ void __launch_bounds__(kSelThreads) k_sort_order(const int* __restrict__ key, int n, int depth_limit,
                                                            int* __restrict__ order) {
  __shared__ int s_key[kMaxCnt], s_order[kMaxCnt];
  __shared__ uint32_t s_scratch[kSortScratchWords];
  if (threadIdx.x < n) s_key[threadIdx.x] = key[threadIdx.x];
  __syncthreads();
  std_sort_order(s_key, n, depth_limit, s_order, s_scratch);
  if (threadIdx.x < n) order[threadIdx.x] = s_order[threadIdx.x];
}

void launch_sort_order(const int* key, int n, int depth_limit, int* order, cudaStream_t s, int64_t* launches) {
  k_sort_order<<<1, kSelThreads, 0, s>>>(key, n, depth_limit, order);
  if (launches) ++*launches;
}

// Event_setMask + Event_FeaturesToTrack + id assignment.
//  (1) Event_setMask (feature_tracker.cpp:123-151): points are visited by track_cnt descending,
//      ties where libstdc++'s std::sort puts them (std_sort_order); a point survives iff no surviving earlier point's filled circle
//      covers its rounded pixel.  "Covers" is evaluated pairwise for all pairs in parallel (bit
//      matrix, <= 256 points), the greedy pass then only ANDs bit rows, and all surviving discs
//      are rastered into the bit mask at once.  More than 256 points: the serial mask walk.
//  (2) Event_FeaturesToTrack (:13-38): k_corner_flags left the flagged events (Arc* corner on
//      a live time-surface pixel) as one short list per 128 events, in stream order.  The
//      lists of up to 1024 consecutive event blocks are gathered into shared memory (those
//      already masked by a track are dropped on the way) and served first come, first served by
//      the whole CTA, up to 32 free candidates per round; the walk stops as soon as MAX_CNT is
//      reached.
__global__ void __launch_bounds__(kSelThreads)
k_select(TrackParams P, TrackBuffers B, int n_events, const uint32_t* __restrict__ cand,
         const int* __restrict__ cand_cnt, int snap_slot) {
  PDL_PROLOGUE();
  extern __shared__ uint32_t s_mask[];
  __shared__ int s_hw[kMaxDiscR + 1];
  __shared__ int s_warp[33];
  __shared__ int s_order[kMaxCnt];
  __shared__ float2 s_pts[kMaxCnt];
  __shared__ int s_ids[kMaxCnt], s_cnt[kMaxCnt];
  __shared__ uint32_t s_conf[kSelFast][kSelFast / 32];  // [b][a/32]: a (earlier) covers b
  __shared__ short2 s_px[kSelFast];
  __shared__ uint32_t s_keptbits[kSelFast / 32];
  __shared__ uint32_t s_cand[kSelThreads * kSelPerThread];
  __shared__ int s_kept, s_found;
  __shared__ uint32_t s_first[32];  // the walk: the first free candidates of a round ...
  __shared__ int s_first_idx[32];   // ... and their positions in s_cand
  __shared__ int s_nacc;
  __shared__ uint32_t s_rows[32];        // the walk: s_rows[i] = lanes whose pixel candidate i's disc covers

  TrackState* st = B.st;
  const int tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
  const int W = P.W, H = P.H, words = (W + 31) / 32;
  const int n = st->n_cur;
  SEL_CLK(0);
  for (int i = tid; i < H * words; i += blockDim.x) s_mask[i] = 0;
  if (tid == 0) {
    disc_half_widths(P.min_dist, s_hw);
    s_kept = 0;
    s_found = 0;
  }
  if (tid < n) {
    s_pts[tid] = B.cur_pts[tid];
    s_ids[tid] = B.ids[tid];
    s_cnt[tid] = B.cnt[tid];
  }
  __syncthreads();
  static_assert(sizeof(s_cand) >= sizeof(uint32_t) * kSortScratchWords, "sort scratch");
  std_sort_order(s_cnt, n, -1, s_order, s_cand);  // s_cand is filled further down
  SEL_CLK(1);
  if (n <= kSelFast) {
    // ---- (1) bit-matrix path
    if (tid < n) {
      const float2 p = s_pts[s_order[tid]];
      s_px[tid] = make_short2((short)max(-30000, min(30000, cv_round(p.x))),
                              (short)max(-30000, min(30000, cv_round(p.y))));
    }
    __syncthreads();
    const int nw = (n + 31) >> 5;
    for (int i = tid; i < n * nw; i += blockDim.x) {
      const int b = i / nw, aw = i - b * nw;
      const short2 pb = s_px[b];
      uint32_t bits = 0;
      for (int k = 0; k < 32; ++k) {
        const int a = aw * 32 + k;
        if (a < b) {
          const short2 pa = s_px[a];
          if (in_disc(pa.x, pa.y, pb.x, pb.y, P.min_dist, s_hw)) bits |= 1u << k;
        }
      }
      s_conf[b][aw] = bits;
    }
    __syncthreads();
    SEL_CLK(2);
    if (warp == 0) {
      // The greedy pass: a point survives iff no surviving earlier point covers it.  32 order
      // positions at a time, lane = point: survivors of the blocks before are final, so whether
      // one of them covers the point is a parallel test; inside the block the lowest live lane
      // survives and the lanes it covers die (one ballot on its column of the conflict matrix),
      // so the pass costs one step per SURVIVOR instead of one vote per point.
      int kept = 0;
      for (int w = 0; w < nw; ++w) {
        const int b = 32 * w + lane;
        bool ok = false;
        uint32_t my_row = 0;  // bit a: point 32 w + a (earlier in the order) covers me
        if (b < n) {
          const short2 pb = s_px[b];
          ok = pb.x >= 0 && pb.x < W && pb.y >= 0 && pb.y < H;
          for (int w2 = 0; w2 < w; ++w2) ok = ok && (s_conf[b][w2] & s_keptbits[w2]) == 0;
          my_row = s_conf[b][w];
        }
        uint32_t alive = __ballot_sync(0xffffffffu, ok), acc = 0;
        while (alive) {
          const int l = __ffs(alive) - 1;
          acc |= 1u << l;
          alive &= ~__ballot_sync(0xffffffffu, (my_row >> l) & 1u) & ~((2u << l) - 1u);
        }
        if (lane == 0) s_keptbits[w] = acc;
        kept += __popc(acc);
        __syncwarp();
      }
      if (lane >= nw && lane < kSelFast / 32) s_keptbits[lane] = 0;
      if (lane == 0) s_kept = kept;
    }
    __syncthreads();
    SEL_CLK(3);
    if (tid < n && ((s_keptbits[tid >> 5] >> (tid & 31)) & 1u)) {
      // survivors keep their visiting order: position = survivors before me
      int pos = __popc(s_keptbits[tid >> 5] & ((1u << (tid & 31)) - 1u));
      for (int w = 0; w < (tid >> 5); ++w) pos += __popc(s_keptbits[w]);
      const int i = s_order[tid];
      B.cur_pts[pos] = s_pts[i];
      B.ids[pos] = s_ids[i];
      B.cnt[pos] = s_cnt[i];
    }
    for (int b = warp; b < n; b += kSelThreads / 32)
      if ((s_keptbits[b >> 5] >> (b & 31)) & 1u)
        fill_disc_warp_atomic(s_mask, words, W, H, s_px[b].x, s_px[b].y, P.min_dist, s_hw);
  } else if (warp == 0) {
    // ---- (1) serial walk over the mask
    int kept = 0;
    for (int k = 0; k < n; ++k) {
      const int i = s_order[k];
      const float2 p = s_pts[i];
      const int cx = cv_round(p.x), cy = cv_round(p.y);
      if (cx < 0 || cx >= W || cy < 0 || cy >= H) continue;
      if (!mask_test(s_mask, words, cx, cy)) {
        if (lane == 0) {
          B.cur_pts[kept] = p;
          B.ids[kept] = s_ids[i];
          B.cnt[kept] = s_cnt[i];
        }
        ++kept;
        __syncwarp();
        fill_disc_warp(s_mask, words, W, H, cx, cy, P.min_dist, s_hw);
      }
    }
    if (lane == 0) s_kept = kept;
  }
  __syncthreads();
  SEL_CLK(4);
  const int kept = s_kept;
  const int want = P.max_cnt - kept;
  // the bit matrix is done with: its first half takes the accepted corners (x | y << 16), its
  // second half the ends of the candidate lists of a chunk
  uint32_t* s_new = &s_conf[0][0];
  int* s_off = reinterpret_cast<int*>(&s_conf[kSelFast / 2][0]);
  static_assert(kMaxCnt <= kSelFast * (kSelFast / 32) / 2 && kSelThreads <= kSelFast * (kSelFast / 32) / 2,
                "accepted corners and list ends fit in the two halves of the conflict matrix");

  // ---- Event_FeaturesToTrack: first come, first served in stream order
  if (want > 0 && n_events > 0) {
    constexpr uint32_t kNone = 0xffffffffu;
    constexpr int kCap = kSelThreads * kSelPerThread;
    const int n_blk = (n_events + kCornerBlock - 1) / kCornerBlock;
    int b0 = 0;
    while (b0 < n_blk) {
      // thread t takes event block b0 + t: inclusive scan of the list lengths
      const int blk = b0 + tid;
      const bool real = blk < n_blk;
      const int c = real ? __ldg(cand_cnt + blk) : 0;
      int incl = c;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
      }
      if (lane == 31) s_warp[warp] = incl;
      __syncthreads();
      if (warp == 0) {
        const int v = s_warp[lane];
        int w = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int o = __shfl_up_sync(0xffffffffu, w, d);
          if (lane >= d) w += o;
        }
        s_warp[lane] = w - v;
      }
      __syncthreads();
      incl += s_warp[warp];
      // the lists that fit the shared-memory buffer are a prefix of the blocks (a list holds at
      // most kCornerBlock entries, so at least one always fits)
      const bool fits = real && incl <= kCap;
      const int n_fit = __syncthreads_count(fits);
      s_off[tid] = fits ? incl : 0x7fffffff;  // inclusive end of block tid's list in s_cand
      if (fits && tid == n_fit - 1) s_warp[32] = incl;
      __syncthreads();
      const int total = s_warp[32];
      // entry e of the chunk belongs to the first block whose list ends behind e (binary search
      // over the n_fit ends); a thread copies entries tid, tid + 1024, ... -- a thread per block
      // walking its own list one entry at a time took five times as long
      {
        uint32_t xy[kSelPerThread];
#pragma unroll
        for (int q = 0; q < kSelPerThread; ++q) {  // all loads of a thread in flight together
          const int e = tid + q * kSelThreads;
          xy[q] = kNone;
          if (e < total) {
            int lo = 0, hi = n_fit - 1;
            while (lo < hi) {
              const int mid = (lo + hi) >> 1;
              if (s_off[mid] > e) hi = mid;
              else lo = mid + 1;
            }
            const int first = lo ? s_off[lo - 1] : 0;
            xy[q] = __ldg(cand + (size_t)(b0 + lo) * kCornerBlock + (e - first));
          }
        }
#pragma unroll
        for (int q = 0; q < kSelPerThread; ++q) {
          const int e = tid + q * kSelThreads;
          if (e < total) s_cand[e] = mask_test(s_mask, words, xy[q] & 0xffff, xy[q] >> 16) ? kNone : xy[q];
        }
      }
      __syncthreads();
      // The walk over the chunk, by the whole CTA in rounds.  A candidate becomes a corner iff its
      // pixel is free when it is visited, i.e. free under the mask of the rounds before AND outside
      // the discs of the corners accepted earlier in its own round.  Round: (a) every thread tests
      // one candidate of the window [pos, pos + 1024) against the mask; (b) the first 32 free ones,
      // in stream order, are compacted; (c) warp i computes which of them candidate i's disc covers (in_disc = the
      // raster of cv::circle), warp 0 then settles them with bit operations: the lowest live lane is
      // a corner, the lanes its disc covers die; (d) the new discs are rastered, one per warp;
      // (e) the next window starts behind the last settled candidate.  The mask only grows, so
      // what a round skipped as masked stays masked.
      if (total > 0) {
        const int r = P.min_dist;
        const uint32_t lt_mask = (1u << lane) - 1u;
        int pos = 0, found = s_found;  // uniform over the CTA
        while (pos < total && found < want) {
          const int ci = pos + tid;
          const uint32_t xy = ci < total ? s_cand[ci] : kNone;
          const bool free_px = xy != kNone && !mask_test(s_mask, words, xy & 0xffff, xy >> 16);
          const uint32_t fr = __ballot_sync(0xffffffffu, free_px);
          if (lane == 0) s_warp[warp] = __popc(fr);
          __syncthreads();
          const int cnt_l = s_warp[lane];
          int incl = cnt_l;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
          }
          const int n_free = __shfl_sync(0xffffffffu, incl, 31);
          const int my_rank = __shfl_sync(0xffffffffu, incl - cnt_l, warp) + __popc(fr & lt_mask);
          if (free_px && my_rank < 32) {
            s_first[my_rank] = xy;
            s_first_idx[my_rank] = ci;
          }
          __syncthreads();
          if (n_free == 0) {
            pos += kSelThreads;
            continue;
          }
          // who covers whom among the <= 32: warp i tests candidate i's disc against all of them
          const int m = min(n_free, 32);
          {
            const uint32_t mine = lane < m ? s_first[lane] : kNone;
            const uint32_t ci_xy = warp < m ? s_first[warp] : kNone;
            const uint32_t row = __ballot_sync(0xffffffffu, lane < m && warp < m &&
                                               in_disc(ci_xy & 0xffff, ci_xy >> 16, mine & 0xffff, mine >> 16, r, s_hw));
            if (lane == 0) s_rows[warp] = row;
          }
          __syncthreads();
          if (warp == 0) {
            const uint32_t cxy = lane < m ? s_first[lane] : kNone;
            const uint32_t my_row = s_rows[lane];
            uint32_t alive = m == 32 ? 0xffffffffu : ((1u << m) - 1u), acc = 0;
            int nacc = 0;
            while (alive && found + nacc < want) {
              const int l = __ffs(alive) - 1;
              acc |= 1u << l;
              ++nacc;
              alive &= ~__shfl_sync(0xffffffffu, my_row, l) & ~((2u << l) - 1u);  // lanes up to l are settled
            }
            if ((acc >> lane) & 1u) s_new[found + __popc(acc & lt_mask)] = cxy;
            if (lane == 0) s_nacc = nacc;
          }
          __syncthreads();
          const int nacc = s_nacc;
          for (int k = warp; k < nacc; k += kSelThreads / 32) {
            const uint32_t a = s_new[found + k];
            fill_disc_warp_atomic(s_mask, words, W, H, a & 0xffff, a >> 16, r, s_hw);
          }
          found += nacc;
          pos = n_free > 32 ? s_first_idx[31] + 1 : pos + kSelThreads;
          __syncthreads();  // the discs are in the mask; s_first / s_warp may be rewritten
        }
        if (tid == 0) s_found = found;
      }
      __syncthreads();
      if (s_found >= want) break;
      b0 += n_fit;
    }
  }
  __syncthreads();
  SEL_CLK(5);
  for (int i = tid; i < s_found; i += blockDim.x) {  // new points: ids from n_id++, track_cnt 1
    const uint32_t xy = s_new[i];
    B.cur_pts[kept + i] = make_float2((float)(xy & 0xffff), (float)(xy >> 16));
    B.ids[kept + i] = st->next_id + i;
    B.cnt[kept + i] = 1;
  }
  __syncthreads();
  if (tid == 0) {
    const int found = s_found;
    st->stat_after_mask = kept;
    st->stat_new = found;
    st->n_cur = kept + found;
    st->next_id += found;
  }
  if (snap_slot >= 0) {  // the publish window's temporal stage ends here
    __syncthreads();
    snapshot_tracks(P, B, snap_slot, s_kept + s_found);
  }
  SEL_CLK(6);
}

__global__ void k_image_set_mask(TrackParams P, TrackBuffers B, uint32_t* __restrict__ blocked);

// called by esvio_fe_create on the handle's device: both single-CTA kernels that keep the
// W x H bit mask in dynamic shared memory get room for this handle's frame size
int select_configure(int W, int H) {
  static SmemLimit lim_select = {}, lim_image = {};
  const size_t bytes = select_smem_bytes(W, H);
  cudaFuncAttributes fa;
  if (cudaFuncGetAttributes(&fa, k_select) != cudaSuccess) return -1;
  if (bytes + fa.sharedSizeBytes > 227 * 1024) return -1;
  if (raise_dyn_smem(k_select, bytes, &lim_select) != 0) return -1;
  return raise_dyn_smem(k_image_set_mask, bytes, &lim_image);
}

void launch_select(const TrackParams& P, const TrackBuffers& B, int n_events, const uint32_t* cand,
                   const int* cand_cnt, int snap_slot, cudaStream_t s, int64_t* launches) {
  launch_pdl(k_select, dim3(1), dim3(kSelThreads), select_smem_bytes(P.W, P.H), s, P, B, n_events, cand,
             cand_cnt, snap_slot);
  ++*launches;
}

// ------------------------------------------------------------------------------------
// left undistort + velocity, stereo forward/backward check, right undistort + velocity,
// result packing and the state roll (feature_tracker.cpp:470-473, 496-510, 570-574, 585-590)
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
k_finalize(TrackParams P, TrackBuffers B, int slot, double cur_time, double prev_time) {
  PDL_PROLOGUE();
  __shared__ int s_warp[33];
  TrackState* st = B.st;
  const int i = threadIdx.x;
  const int M = P.max_cnt;
  const int* hdr = B.snap_hdr + slot * 16;
  const float2* s_cur = B.snap_pts + slot * M;
  const int* s_ids = B.snap_ids + slot * M;
  const int* s_cnt = B.snap_cnt + slot * M;
  const int n = hdr[0];
  int32_t* res = B.result + (size_t)slot * B.result_words;
  int32_t* r_id = res + kResultHdr;
  int32_t* r_cnt = r_id + M;
  float* r_u = reinterpret_cast<float*>(r_cnt + M);
  float *r_v = r_u + M, *r_unx = r_v + M, *r_uny = r_unx + M, *r_vx = r_uny + M, *r_vy = r_vx + M;
  int32_t* r_idr = reinterpret_cast<int32_t*>(r_vy + M);
  float* r_ru = reinterpret_cast<float*>(r_idr + M);
  float *r_rv = r_ru + M, *r_runx = r_rv + M, *r_runy = r_runx + M, *r_rvx = r_runy + M,
        *r_rvy = r_rvx + M;
  const double dt = cur_time - prev_time;
  const int np_un = st->n_prev_un, np_un_r = st->n_prev_un_r;
  // prev_un_pts_map / prev_un_right_pts_map (id -> point) staged in shared memory
  __shared__ int s_pid[kMaxCnt], s_pid_r[kMaxCnt];
  __shared__ float2 s_pun[kMaxCnt], s_pun_r[kMaxCnt];
  if (i < np_un) {
    s_pid[i] = B.prev_un_ids[i];
    s_pun[i] = B.prev_un[i];
  }
  if (i < np_un_r) {
    s_pid_r[i] = B.prev_un_r_ids[i];
    s_pun_r[i] = B.prev_un_r[i];
  }
  __syncthreads();

  float2 cp = make_float2(0.f, 0.f), un = make_float2(0.f, 0.f);
  int id = -1;
  int keep = 0;
  float2 rp = make_float2(0.f, 0.f);
  if (i < n) {
    cp = s_cur[i];
    id = s_ids[i];
    double x, y;
    lift_projective(P.cam[0], (double)cp.x, (double)cp.y, x, y);
    un = make_float2((float)x, (float)y);
    // ptsVelocity (feature_tracker.cpp:1004-1045)
    float vx = 0.f, vy = 0.f;
    if (np_un > 0 && id != -1) {
      for (int j = 0; j < np_un; ++j)
        if (s_pid[j] == id) {
          const float2 q = s_pun[j];
          vx = (float)((double)(un.x - q.x) / dt);
          vy = (float)((double)(un.y - q.y) / dt);
          break;
        }
    }
    r_id[i] = id;
    r_cnt[i] = s_cnt[i];
    r_u[i] = cp.x;
    r_v[i] = cp.y;
    r_unx[i] = un.x;
    r_uny[i] = un.y;
    r_vx[i] = vx;
    r_vy[i] = vy;
    // stereo check
    rp = B.right_pts[slot * M + i];
    keep = B.st_sf[slot * M + i] != 0;
    if (P.flow_back)
      keep = keep && B.st_sb[slot * M + i] && in_border(P.W, P.H, rp) &&
             pt_dist(cp, B.rev_left_pts[slot * M + i]) <= 0.5;
  }
  int total;
  const int pos = block_excl_scan_1024(keep, s_warp, &total);
  // all reads of prev_un / prev_un_ids are done (barriers inside the scan); roll the state
  if (i < n) {
    B.prev_un_ids[i] = id;
    B.prev_un[i] = un;
  }
  float2 unr = make_float2(0.f, 0.f);
  float rvx = 0.f, rvy = 0.f;
  if (keep) {
    double x, y;
    lift_projective(P.cam[1], (double)rp.x, (double)rp.y, x, y);
    unr = make_float2((float)x, (float)y);
    if (np_un_r > 0 && id != -1) {
      for (int j = 0; j < np_un_r; ++j)
        if (s_pid_r[j] == id) {
          const float2 q = s_pun_r[j];
          rvx = (float)((double)(unr.x - q.x) / dt);
          rvy = (float)((double)(unr.y - q.y) / dt);
          break;
        }
    }
    r_idr[pos] = id;
    r_ru[pos] = rp.x;
    r_rv[pos] = rp.y;
    r_runx[pos] = unr.x;
    r_runy[pos] = unr.y;
    r_rvx[pos] = rvx;
    r_rvy[pos] = rvy;
  }
  __syncthreads();
  if (keep) {
    B.prev_un_r_ids[pos] = id;
    B.prev_un_r[pos] = unr;
  }
  if (i == 0) {
    st->n_right = total;
    st->n_prev_un = n;
    st->n_prev_un_r = total;
    res[0] = n;
    res[1] = total;
    for (int q = 1; q <= 8; ++q) res[1 + q] = hdr[q];
  }
}

void launch_finalize(const TrackParams& P, const TrackBuffers& B, int slot, double cur_time,
                     double prev_time, cudaStream_t s, int64_t* launches) {
  launch_pdl(k_finalize, dim3(1), dim3(1024), 0, s, P, B, slot, cur_time, prev_time);
  ++*launches;
}

__global__ void k_undistort(Pinhole cam, const float2* __restrict__ uv, int n,
                            float2* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x, y;
  lift_projective(cam, (double)uv[i].x, (double)uv[i].y, x, y);
  out[i] = make_float2((float)x, (float)y);
}

void launch_undistort(const Pinhole& cam, const float2* uv, int n, float2* out, cudaStream_t s,
                      int64_t* launches) {
  if (n <= 0) return;
  k_undistort<<<(n + 127) / 128, 128, 0, s>>>(cam, uv, n, out);
  ++*launches;
}

// =====================================================================================
// frame path (FeatureTracker::trackImage, feature_tracker.cpp:164-338)
// =====================================================================================
// Image_setMask (feature_tracker.cpp:91-121): points are visited by track_cnt descending (ties
// where std::sort puts them, as in Event_setMask above); a point survives iff its rounded pixel is
// still free, and then blocks the filled circle of radius MIN_DIST_IMG around it.  Survivors
// are compacted in place; the mask (1 bit per pixel, 1 = blocked) goes to global memory for
// goodFeaturesToTrack.
__global__ void __launch_bounds__(kSelThreads)
k_image_set_mask(TrackParams P, TrackBuffers B, uint32_t* __restrict__ blocked) {
  PDL_PROLOGUE();
  extern __shared__ uint32_t s_mask[];
  __shared__ int s_hw[kMaxDiscR + 1];
  __shared__ int s_order[kMaxCnt];
  __shared__ float2 s_pts[kMaxCnt];
  __shared__ int s_ids[kMaxCnt], s_cnt[kMaxCnt];
  __shared__ uint32_t s_sort[kSortScratchWords];
  __shared__ int s_kept;
  TrackState* st = B.st;
  const int tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
  const int W = P.W, H = P.H, words = (W + 31) / 32;
  const int n = st->n_cur;
  for (int i = tid; i < H * words; i += blockDim.x) s_mask[i] = 0;
  if (tid == 0) {
    disc_half_widths(P.min_dist, s_hw);
    s_kept = 0;
  }
  if (tid < n) {
    s_pts[tid] = B.cur_pts[tid];
    s_ids[tid] = B.ids[tid];
    s_cnt[tid] = B.cnt[tid];
  }
  __syncthreads();
  std_sort_order(s_cnt, n, -1, s_order, s_sort);
  if (warp == 0) {
    int kept = 0;
    for (int k = 0; k < n; ++k) {
      const int i = s_order[k];
      const float2 p = s_pts[i];
      const int cx = cv_round(p.x), cy = cv_round(p.y);
      if (cx < 0 || cx >= W || cy < 0 || cy >= H) continue;  // cannot happen after inBorder
      if (!mask_test(s_mask, words, cx, cy)) {
        if (lane == 0) {
          B.cur_pts[kept] = p;
          B.ids[kept] = s_ids[i];
          B.cnt[kept] = s_cnt[i];
        }
        ++kept;
        __syncwarp();
        fill_disc_warp(s_mask, words, W, H, cx, cy, P.min_dist, s_hw);
      }
    }
    if (lane == 0) s_kept = kept;
  }
  __syncthreads();
  for (int i = tid; i < H * words; i += blockDim.x) blocked[i] = s_mask[i];
  if (tid == 0) {
    st->n_cur = s_kept;
    st->stat_after_mask = s_kept;
  }
}

void launch_image_set_mask(const TrackParams& P, const TrackBuffers& B, const GfttBuffers& G,
                           cudaStream_t s, int64_t* launches) {
  // the W x H bit mask lives in dynamic shared memory (38 KB at 640x480); the limit was raised
  // for this device when the handle was created (select_configure)
  const size_t bytes = select_smem_bytes(P.W, P.H);
  launch_pdl(k_image_set_mask, dim3(1), dim3(kSelThreads), bytes, s, P, B, G.blocked);
  ++*launches;
}

// goodFeaturesToTrack's minimum-distance pass (featureselect.cpp): the candidates come best
// first (keys sorted descending, key 0 = end); one is accepted iff no corner accepted before it
// lies closer than minDistance (Euclidean, on integer pixel coordinates), until maxCorners.
// 1024 candidates at a time: all threads test theirs against the corners accepted so far, the
// ones still alive are compacted in order and warp 0 settles them 32 at a time (the lowest
// alive lane is accepted, the others are tested against it, and so on).
// TRACKS: append to the track arrays (ids from next_id, track_cnt 1), update the counters and
// take the snapshot; otherwise write the corners to out_xy / out_n.
template <bool TRACKS>
__global__ void __launch_bounds__(1024)
k_gftt_pick(TrackParams P, TrackBuffers B, int snap_slot, const unsigned long long* __restrict__ keys,
            const int* __restrict__ n_cand, int capacity, int W, int max_corners, float md2, int spaced,
            float2* __restrict__ out_xy,
            int* __restrict__ out_n) {
  PDL_PROLOGUE();
  __shared__ int s_warp[33];
  __shared__ float2 s_acc[kMaxCnt];       // accepted corners (TRACKS: at most max_cnt)
  __shared__ uint32_t s_cand[1024];
  __shared__ int s_found, s_end;
  TrackState* st = B.st;
  const int tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
  const int kept = TRACKS ? st->n_cur : 0;
  // stage form: max_corners <= 0 = unlimited, which the spaced pass serves up to kMaxCnt
  // the frame's candidates (k_gftt_rank sorted exactly these).  __ldcg, not a plain load through
  // the const __restrict__ pointer: nvcc turns that into ld.global.nc and is then free to hoist it
  // above griddepcontrol.wait, i.e. to read the count before k_gftt_keys has produced it
  // (tests/test_sass_pdl.py checks every kernel for this)
  const int n_keys = min(__ldcg(n_cand), capacity);
  int want = TRACKS ? P.max_cnt - kept : (max_corners > 0 ? max_corners : n_keys);
  if (spaced && want > kMaxCnt) want = kMaxCnt;  // capacity of s_acc (the ABI refuses more)
  if (tid == 0) s_found = 0, s_end = 0;
  __syncthreads();
  if (want > 0) {
    for (int base = 0; base < n_keys; base += 1024) {
      const int found0 = s_found;
      const unsigned long long key = base + tid < n_keys ? keys[base + tid] : 0ull;
      const bool valid = key != 0ull;
      const int idx = (int)(key & 0xffffffffull);
      const int yi = idx / W, xi = idx - yi * W;
      const float x = (float)xi, y = (float)yi;
      bool alive = valid;
      if (alive && spaced) {
        const int lim = found0 < kMaxCnt ? found0 : kMaxCnt;
        for (int j = 0; j < lim; ++j) {
          const float dx = x - s_acc[j].x, dy = y - s_acc[j].y;
          if (dx * dx + dy * dy < md2) {
            alive = false;
            break;
          }
        }
      }
      int total;
      const int pos = block_excl_scan_1024(alive ? 1 : 0, s_warp, &total);
      if (alive) s_cand[pos] = (uint32_t)xi | ((uint32_t)yi << 16);
      if (!valid && tid == 0) s_end = 1;  // keys are sorted: a zero key ends the candidates
      if (tid == 1023 && !valid) s_end = 1;
      __syncthreads();
      if (warp == 0) {
        int found = found0;
        for (int c0 = 0; c0 < total && found < want; c0 += 32) {
          const bool have = c0 + lane < total;
          const uint32_t xy = have ? s_cand[c0 + lane] : 0u;
          const float cx = (float)(xy & 0xffff), cy = (float)(xy >> 16);
          bool ok = have;
          if (ok && spaced)  // against the corners accepted earlier in this batch
            for (int j = found0; j < found && j < kMaxCnt; ++j) {
              const float dx = cx - s_acc[j].x, dy = cy - s_acc[j].y;
              if (dx * dx + dy * dy < md2) {
                ok = false;
                break;
              }
            }
          uint32_t live = __ballot_sync(0xffffffffu, ok);
          while (live && found < want) {
            const int l = __ffs(live) - 1;
            const float ax = __shfl_sync(0xffffffffu, cx, l), ay = __shfl_sync(0xffffffffu, cy, l);
            if (lane == 0) {
              if (found < kMaxCnt) s_acc[found] = make_float2(ax, ay);
              if (TRACKS) {
                B.cur_pts[kept + found] = make_float2(ax, ay);
                B.ids[kept + found] = st->next_id + found;
                B.cnt[kept + found] = 1;
              } else {
                out_xy[found] = make_float2(ax, ay);
              }
            }
            ++found;
            __syncwarp();
            bool still = ((live >> lane) & 1u) && lane > l;
            if (still && spaced) {
              const float dx = cx - ax, dy = cy - ay;
              if (dx * dx + dy * dy < md2) still = false;
            }
            live = __ballot_sync(0xffffffffu, still);
          }
        }
        if (lane == 0) s_found = found;
      }
      __syncthreads();
      if (s_found >= want || s_end) break;
    }
  }
  __syncthreads();
  const int found = s_found;
  if (TRACKS) {
    if (tid == 0) {
      st->stat_new = found;
      st->n_cur = kept + found;
      st->next_id += found;
    }
    if (snap_slot >= 0) {
      __syncthreads();
      snapshot_tracks(P, B, snap_slot, kept + found);
    }
  } else if (tid == 0) {
    *out_n = found;
  }
}

void launch_gftt_pick_tracks(const TrackParams& P, const TrackBuffers& B, const GfttBuffers& G,
                             int snap_slot, cudaStream_t s, int64_t* launches) {
  const float md = (float)P.min_dist;
  launch_pdl(k_gftt_pick<true>, dim3(1), dim3(1024), 0, s, P, B, snap_slot,
             (const unsigned long long*)G.keys_sorted, (const int*)G.n_cand, P.W * P.H, P.W, 0, md * md,
             P.min_dist > 1 ? 1 : 0, (float2*)nullptr, (int*)nullptr);  // distance 1: distinct
                                                                          // pixels never clash
  ++*launches;
}

void launch_gftt_pick_stage(const GfttBuffers& G, int W, int H, int max_corners,
                            double min_distance, cudaStream_t s, int64_t* launches) {
  TrackParams P{};
  TrackBuffers B{};
  launch_pdl(k_gftt_pick<false>, dim3(1), dim3(1024), 0, s, P, B, -1,
             (const unsigned long long*)G.keys_sorted, (const int*)G.n_cand, W * H, W, max_corners,
             (float)(min_distance * min_distance), min_distance > 1.0 ? 1 : 0, G.out_xy, G.out_n);
  ++*launches;
}

__global__ void k_right_map_keep(TrackState* st, int restore) {
  if (restore) st->n_prev_un_r = st->pad[0];
  else st->pad[0] = st->n_prev_un_r;
}

void launch_right_map_keep(const TrackBuffers& B, int restore, cudaStream_t s, int64_t* launches) {
  k_right_map_keep<<<1, 1, 0, s>>>(B.st, restore);
  ++*launches;
}

}  // namespace esvio
