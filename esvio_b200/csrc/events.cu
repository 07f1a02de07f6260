// events.cu -- per-event stages of the front-end as sm_100a kernels:
//   K0 bin_events      stable counting sort of a window's events into 16x8-pixel fine tiles
//   K1 sae_update_ts   createSAE_left/right + SAEtoTimeSurface_left/right, fused
//                      (feature_tracker/src/event_detector/event_detector.cc:149-166,
//                       212-228, 230-305), one CTA per 32x8 tile, tile state TMA-staged in smem
//   K2 corner_flags    EventDetector::isCorner (Arc*) for every left event
//                      (event_detector.cc:308-544)
#include <stdlib.h>

#include "common.cuh"

namespace esvio {

// =====================================================================================
// Motion compensation: EventDetector::motioncorrection (event_detector.cc:547-591)
// =====================================================================================
// The reference evaluates this with Eigen fixed-size float types; the order of the float
// operations below is the oracle's restatement of Eigen 3.3's kernels (oracle/esvio_oracle.c,
// "Motion-compensated SAE update"), so the two agree bit for bit (-fmad=false).
__device__ __forceinline__ float dot3f(float a0, float b0, float a1, float b1, float a2, float b2) {
  return a0 * b0 + (a1 * b1 + a2 * b2);
}
__device__ __forceinline__ void mat3_mul_f(const float* A, const float* B, float* C) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      C[i * 3 + j] = dot3f(A[i * 3], B[j], A[i * 3 + 1], B[3 + j], A[i * 3 + 2], B[6 + j]);
}
__device__ __forceinline__ void mat3_vec_f(const float* A, const float* v, float* o) {
#pragma unroll
  for (int i = 0; i < 3; ++i) o[i] = dot3f(A[i * 3], v[0], A[i * 3 + 1], v[1], A[i * 3 + 2], v[2]);
}
__device__ __forceinline__ float cof3f(const float* m, int i, int j) {
  const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
  return m[i1 * 3 + j1] * m[i2 * 3 + j2] - m[i1 * 3 + j2] * m[i2 * 3 + j1];
}
__device__ __forceinline__ void mat3_inv_f(const float* m, float* r) {
  const float c0 = cof3f(m, 0, 0), c1 = cof3f(m, 1, 0), c2 = cof3f(m, 2, 0);
  const float det = dot3f(c0, m[0], c1, m[3], c2, m[6]);
  const float invdet = 1.0f / det;
  r[0] = c0 * invdet;
  r[1] = c1 * invdet;
  r[2] = c2 * invdet;
  r[3] = cof3f(m, 0, 1) * invdet;
  r[4] = cof3f(m, 1, 1) * invdet;
  r[5] = cof3f(m, 2, 1) * invdet;
  r[6] = cof3f(m, 0, 2) * invdet;
  r[7] = cof3f(m, 1, 2) * invdet;
  r[8] = cof3f(m, 2, 2) * invdet;
}

// Matrix3f::exp(): Pade 3/5/7 by L1 norm, scaling and squaring, partial-pivot LU solve
__device__ void mat3_exp_f(const float* A_in, float* R) {
  float A[9], A2[9], A4[9], A6[9], tmp[9], U[9], V[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) A[i] = A_in[i];
  float l1 = 0.f;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const float cs = fabsf(A[j]) + (fabsf(A[3 + j]) + fabsf(A[6 + j]));
    if (cs > l1) l1 = cs;
  }
  int squarings = 0;
  if (l1 < 4.258730016922831e-001f) {
    mat3_mul_f(A, A, A2);
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const float id = (i % 4 == 0) ? 1.f : 0.f;
      tmp[i] = 1.f * A2[i] + 60.f * id;
      V[i] = 12.f * A2[i] + 120.f * id;
    }
    mat3_mul_f(A, tmp, U);
  } else if (l1 < 1.880152677804762e+000f) {
    mat3_mul_f(A, A, A2);
    mat3_mul_f(A2, A2, A4);
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const float id = (i % 4 == 0) ? 1.f : 0.f;
      tmp[i] = 1.f * A4[i] + 420.f * A2[i] + 15120.f * id;
      V[i] = 30.f * A4[i] + 3360.f * A2[i] + 30240.f * id;
    }
    mat3_mul_f(A, tmp, U);
  } else {
    frexpf(l1 / 3.925724783138660f, &squarings);
    if (squarings < 0) squarings = 0;
#pragma unroll
    for (int i = 0; i < 9; ++i) A[i] = ldexpf(A[i], -squarings);
    mat3_mul_f(A, A, A2);
    mat3_mul_f(A2, A2, A4);
    mat3_mul_f(A4, A2, A6);
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const float id = (i % 4 == 0) ? 1.f : 0.f;
      tmp[i] = 1.f * A6[i] + 1512.f * A4[i] + 277200.f * A2[i] + 8648640.f * id;
      V[i] = 56.f * A6[i] + 25200.f * A4[i] + 1995840.f * A2[i] + 17297280.f * id;
    }
    mat3_mul_f(A, tmp, U);
  }
  float lu[9], x[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    x[i] = U[i] + V[i];
    lu[i] = -U[i] + V[i];
  }
  int piv[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    int best = k;
    float score = fabsf(lu[k * 3 + k]);
#pragma unroll
    for (int i = k + 1; i < 3; ++i)
      if (fabsf(lu[i * 3 + k]) > score) {
        score = fabsf(lu[i * 3 + k]);
        best = i;
      }
    piv[k] = best;
    if (score != 0.f) {
#pragma unroll
      for (int i = k + 1; i < 3; ++i)
        if (best == i) {
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const float t = lu[k * 3 + j];
            lu[k * 3 + j] = lu[i * 3 + j];
            lu[i * 3 + j] = t;
          }
        }
#pragma unroll
      for (int i = k + 1; i < 3; ++i) lu[i * 3 + k] /= lu[k * 3 + k];
    }
#pragma unroll
    for (int i = k + 1; i < 3; ++i)
#pragma unroll
      for (int j = k + 1; j < 3; ++j) lu[i * 3 + j] -= lu[i * 3 + k] * lu[k * 3 + j];
  }
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int i = k + 1; i < 3; ++i)
      if (piv[k] == i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const float t = x[k * 3 + j];
          x[k * 3 + j] = x[i * 3 + j];
          x[i * 3 + j] = t;
        }
      }
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float b = x[k * 3 + j];
#pragma unroll
      for (int i = k + 1; i < 3; ++i) x[i * 3 + j] -= b * lu[i * 3 + k];
    }
#pragma unroll
  for (int k = 2; k >= 0; --k) {
    const float a = 1.0f / lu[k * 3 + k];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float b = (x[k * 3 + j] *= a);
#pragma unroll
      for (int i = 0; i < k; ++i) x[i * 3 + j] -= b * lu[i * 3 + k];
    }
  }
  for (int s = 0; s < squarings; ++s) {
    float t[9];
    mat3_mul_f(x, x, t);
#pragma unroll
    for (int i = 0; i < 9; ++i) x[i] = t[i];
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) R[i] = x[i];
}

__device__ void motion_correct(const McParams& m, double ex, double ey, double dt, int* ox, int* oy) {
  *ox = (int)ex;
  *oy = (int)ey;
  const int border = 6;
  if (!(ex > border && ex <= (m.W - border) && ey > border && ey <= (m.H - border))) return;
  const float fdt = (float)dt;
  const float rv[3] = {m.omega[0] * fdt, m.omega[1] * fdt, m.omega[2] * fdt};
  const float skew[9] = {0.f, -rv[2], rv[1], rv[2], 0.f, -rv[0], -rv[1], rv[0], 0.f};
  const float K[9] = {m.K[0], 0.f, m.K[2], 0.f, m.K[1], m.K[3], 0.f, 0.f, 1.f};
  float R[9], Rt[9], Kinv[9], KR[9], rotK[9];
  mat3_exp_f(skew, R);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) Rt[i * 3 + j] = R[j * 3 + i];
  mat3_inv_f(K, Kinv);
  mat3_mul_f(K, Rt, KR);
  mat3_mul_f(KR, Kinv, rotK);
  const float c = (float)(0.5 * dt);
  float tr[3], w[3], nrotK[9], transK[3], o[3];
  const float ev[3] = {(float)ex, (float)ey, 1.f};
#pragma unroll
  for (int i = 0; i < 3; ++i) tr[i] = c * (m.v_cur[i] + m.v_pre[i]);
  mat3_vec_f(Kinv, tr, w);
#pragma unroll
  for (int i = 0; i < 9; ++i) nrotK[i] = -rotK[i];
  mat3_vec_f(nrotK, w, transK);
  mat3_vec_f(rotK, ev, o);
#pragma unroll
  for (int i = 0; i < 3; ++i) o[i] = o[i] + transK[i];
  o[0] = o[0] / o[2];
  o[1] = o[1] / o[2];
  const int x = (int)floorf(o[0]), y = (int)floorf(o[1]);
  if (x > 0 && x < m.W - 1 && y > 0 && y < m.H - 1) {
    *ox = x;
    *oy = y;
  }
}

// trackEvent(..., measurements), feature_tracker.cpp:628-642: an event is warped iff
// dt_window > 0 and (t - t0) / dt_window < 1 (the |accel| > 5 gate is applied by the caller,
// which only launches this kernel when it holds)
__global__ void __launch_bounds__(256)
k_warp_events(McParams P, DevEvents ev0, DevEvents ev1, uint16_t* __restrict__ wx0,
              uint16_t* __restrict__ wy0, uint16_t* __restrict__ wx1, uint16_t* __restrict__ wy1) {
  const int cam = blockIdx.y;
  const DevEvents& ev = cam ? ev1 : ev0;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ev.n) return;
  const double t0 = load_event(ev0, 0).t;  // first LEFT event, for both cameras
  const double dtw = P.t1 - t0;
  const Ev e = load_event(ev, i);
  int x = e.x, y = e.y;
  if (dtw > 0 && (e.t - t0) / dtw < 1 && x < P.W && y < P.H) motion_correct(P, x, y, e.t - t0, &x, &y);
  (cam ? wx1 : wx0)[i] = (uint16_t)x;
  (cam ? wy1 : wy0)[i] = (uint16_t)y;
}

void launch_warp_events(const McParams& P, const DevEvents ev[2], uint16_t* const wx[2],
                        uint16_t* const wy[2], cudaStream_t s, int64_t* launches) {
  const int n = ev[0].n > ev[1].n ? ev[0].n : ev[1].n;
  if (n <= 0 || ev[0].n <= 0) return;
  k_warp_events<<<dim3((n + 255) / 256, 2), 256, 0, s>>>(P, ev[0], ev[1], wx[0], wy[0], wx[1], wy[1]);
  ++*launches;
}

__global__ void __launch_bounds__(128)
k_warp_points(McParams P, const float* __restrict__ xy_dt, int n, int* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int x, y;
  motion_correct(P, (double)xy_dt[3 * i], (double)xy_dt[3 * i + 1], (double)xy_dt[3 * i + 2], &x, &y);
  out[2 * i] = x;
  out[2 * i + 1] = y;
}

void launch_warp_points(const McParams& P, const float* xy_dt, int n, int* out_xy, cudaStream_t s,
                        int64_t* launches) {
  if (n <= 0) return;
  k_warp_points<<<(n + 127) / 128, 128, 0, s>>>(P, xy_dt, n, out_xy);
  ++*launches;
}

// =====================================================================================
// K0: stable sort of a window's events by fine tile, as a two-digit LSD radix sort
// =====================================================================================
// Same-pixel events must be applied in stream order (acceptance of an event depends on
// the previous same- and opposite-polarity event at its pixel, event_detector.cc:157),
// so the sort is stable.  Bins are 16x8-pixel fine tiles, numbered
//     bin = (y / 8) * (2 * tiles_x) + x / 16          (the two fine tiles of one 32x8 tile, one
//                                                      TMA box of the SAE state, are adjacent)
// i.e. a two-digit key: minor = x / 16, major = y / 8, each < kDigit = 128 (W <= 2048,
// H <= 1024).  A counting sort over the 2 400 bins of a 640x480 sensor needs tables as large as
// the window itself (2 048-event chunks x 2 401 counters) and 86 KB of shared memory per
// CTA; two stable passes over <= 128 bins need 4 KB and leave nothing to scan:
//   k_bin_hist   per 2 048-event chunk: histogram of the minor digit -> cnt_a[cam][chunk][128];
//                events per fine tile accumulated with global reductions -> bin_total
//   k_bin_pass1  stable scatter by minor digit into (it, ik, im) = (time, key, major digit); a
//                CTA derives its write bases from the cnt_a rows itself; the major-digit
//                histogram of pass 2's chunks is accumulated on the way (cnt_b); one extra CTA
//                per camera turns bin_total into bin_start and clears it for the next window
//   k_bin_pass2  stable scatter by major digit into (bt, bk), the runs K1 replays
// CTA c owns events [c*2048, (c+1)*2048), warp w of it the 256 consecutive events
// [w*256, (w+1)*256), walked 32 at a time; events beyond the sensor are dropped in pass 1 (the
// reference would index out of range).
constexpr int kDigitBits = 7, kDigit = 1 << kDigitBits;
constexpr int kBinWarps = kChunkThreads / 32;

struct BinTables {
  uint32_t* cnt_a;  // [n_cams][max_chunks][kDigit]
  uint32_t* cnt_b;  // [n_cams][max_chunks][kDigit]
};

__device__ __forceinline__ void load_xy(const DevEvents& ev, int i, int& x, int& y) {
  if (ev.wx) {
    x = __ldg(ev.wx + i);
    y = __ldg(ev.wy + i);
  } else if (ev.aos) {
    const uint32_t xy = __ldg(reinterpret_cast<const uint32_t*>(ev.aos + i));
    x = xy & 0xffffu;
    y = xy >> 16;
  } else {
    x = __ldg(ev.x + i);
    y = __ldg(ev.y + i);
  }
}

// Global reductions on a few thousand hot counters are slow (events arrive in bursts on the same
// tile: 333 k per-event reductions cost 6-12 us), so the CTAs first count in shared memory and
// then send one reduction per counter they touched.  Sensors with more fine tiles than
// kAggBins (> 1024x512) fall back to per-event reductions.
constexpr int kAggBins = 4096;

__global__ void __launch_bounds__(kChunkThreads)
k_bin_hist(BinLayout L, const __grid_constant__ CamBatch B, BinTables T, uint32_t* __restrict__ bin_total) {
  PDL_PROLOGUE();
  __shared__ uint32_t s_hist[kDigit];
  extern __shared__ uint32_t s_fine[];  // [n_bins] when n_bins <= kAggBins
  const int cam = blockIdx.y, chunk = blockIdx.x;
  const DevEvents& ev = B.ev[cam];
  if (chunk >= B.n_chunks[cam]) return;
  const bool agg = L.n_bins <= kAggBins;
  if (threadIdx.x < kDigit) s_hist[threadIdx.x] = 0;
  if (agg)
    for (int b = threadIdx.x; b < L.n_bins; b += kChunkThreads) s_fine[b] = 0;
  __syncthreads();
  const int base = chunk * kChunk;
  const int minor_n = L.tiles_x * kFine;
  uint32_t* __restrict__ tot = bin_total + (size_t)cam * (L.n_bins + 1);
#pragma unroll
  for (int k = 0; k < kChunkSteps; ++k) {
    const int i = base + k * kChunkThreads + threadIdx.x;
    if (i < ev.n) {
      int x, y;
      load_xy(ev, i, x, y);
      if (x < L.W && y < L.H) {
        const int mi = x / kFineW, ma = y / kTileH;
        atomicAdd(&s_hist[mi], 1u);
        if (agg) atomicAdd(&s_fine[ma * minor_n + mi], 1u);
        else atomicAdd(tot + ma * minor_n + mi, 1u);
      }
    }
  }
  __syncthreads();
  const size_t row = ((size_t)cam * L.max_chunks + chunk) * kDigit;
  if (threadIdx.x < kDigit) {
    T.cnt_a[row + threadIdx.x] = s_hist[threadIdx.x];
    T.cnt_b[row + threadIdx.x] = 0;
  }
  if (agg)
    for (int b = threadIdx.x; b < L.n_bins; b += kChunkThreads) {
      const uint32_t c = s_fine[b];
      if (c) atomicAdd(tot + b, c);
    }
}

// Write bases of this chunk for every digit value: s_base[d] = (events of smaller digits, all
// chunks) + (events of digit d in earlier chunks), from the per-chunk histograms `cnt` (rows
// 0..n_rows-1 of this camera).  Returns the total number of events in the table.
__device__ __forceinline__ uint32_t bin_bases(const uint32_t* __restrict__ cnt, int n_rows, int chunk,
                                              uint32_t* s_base, uint32_t (*s_red)[2][kDigit]) {
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  uint4 tot = make_uint4(0, 0, 0, 0), pre = make_uint4(0, 0, 0, 0);
  // rows warp, warp + 8, ...: eight loads in flight per trip to L2
  constexpr int kBatch = 8;
  for (int c0 = warp; c0 < n_rows; c0 += kBinWarps * kBatch) {
    uint4 v[kBatch];
#pragma unroll
    for (int j = 0; j < kBatch; ++j) {
      const int c = c0 + j * kBinWarps;
      v[j] = make_uint4(0, 0, 0, 0);
      if (c < n_rows) v[j] = __ldcg(reinterpret_cast<const uint4*>(cnt + (size_t)c * kDigit) + lane);
    }
#pragma unroll
    for (int j = 0; j < kBatch; ++j) {
      const int c = c0 + j * kBinWarps;
      tot.x += v[j].x, tot.y += v[j].y, tot.z += v[j].z, tot.w += v[j].w;
      if (c < chunk) pre.x += v[j].x, pre.y += v[j].y, pre.z += v[j].z, pre.w += v[j].w;
    }
  }
  reinterpret_cast<uint4*>(s_red[warp][0])[lane] = tot;
  reinterpret_cast<uint4*>(s_red[warp][1])[lane] = pre;
  __syncthreads();
  uint32_t total = 0;
  if (warp == 0) {
    // lane owns digits 4*lane .. 4*lane+3
    uint32_t t[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
#pragma unroll
    for (int w = 0; w < kBinWarps; ++w)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        t[j] += s_red[w][0][4 * lane + j];
        q[j] += s_red[w][1][4 * lane + j];
      }
    const uint32_t mine = t[0] + t[1] + t[2] + t[3];
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    uint32_t run = incl - mine;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      s_base[4 * lane + j] = run + q[j];
      run += t[j];
    }
    total = __shfl_sync(0xffffffffu, incl, 31);
    if (lane == 0) s_base[kDigit] = total;
  }
  __syncthreads();
  return s_base[kDigit];
}

// Stable rank of every lane's event among the events of the same digit in this warp's slice:
// the lanes of a 32-event step that share a digit find each other with one ballot per digit
// bit; the running per-digit count of the slice lives in the warp's own row of s_wc.
__device__ __forceinline__ uint32_t bin_rank_step(uint32_t* my_wc, int d, bool valid, uint32_t lt_mask) {
  uint32_t peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
  for (int b = 0; b < kDigitBits; ++b) {
    const uint32_t m = __ballot_sync(0xffffffffu, (d >> b) & 1);
    peers &= ((d >> b) & 1) ? m : ~m;
  }
  uint32_t before = 0;
  if (valid) before = my_wc[d];
  __syncwarp();
  if (valid && (peers & lt_mask) == 0) my_wc[d] = before + __popc(peers);
  __syncwarp();
  return before + __popc(peers & lt_mask);
}

// exclusive prefix of every digit's slice counts over the warps of the CTA, in place
__device__ __forceinline__ void bin_warp_prefix(uint32_t (*s_wc)[kDigit]) {
  if (threadIdx.x < kDigit) {
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < kBinWarps; ++w) {
      const uint32_t c = s_wc[w][threadIdx.x];
      s_wc[w][threadIdx.x] = run;
      run += c;
    }
  }
}

__global__ void __launch_bounds__(kChunkThreads)
k_bin_pass1(BinLayout L, const __grid_constant__ CamBatch B, BinTables T, uint32_t* __restrict__ bin_total,
            uint32_t* __restrict__ bin_start) {
  PDL_PROLOGUE();
  __shared__ __align__(16) uint32_t s_red[kBinWarps][2][kDigit];
  __shared__ uint32_t s_wc[kBinWarps][kDigit];
  __shared__ uint32_t s_base[kDigit + 1];
  __shared__ uint32_t s_scan[kBinWarps + 1];
  // [2][n_bins] when n_bins <= kAggBins: this chunk's events per (minor, major) that land below
  // / at or above the one 2 048-boundary of pass 2's chunks their minor run can straddle
  extern __shared__ uint32_t s_side[];
  const int cam = blockIdx.y, chunk = blockIdx.x;
  const DevEvents& ev = B.ev[cam];
  const bool agg = L.n_bins <= kAggBins;
  const int minor_n = L.tiles_x * kFine;
  const int n_rows = B.n_chunks[cam];
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  if (chunk < n_rows) {
    for (int b = threadIdx.x; b < kBinWarps * kDigit; b += blockDim.x) (&s_wc[0][0])[b] = 0;
    if (agg)
      for (int b = threadIdx.x; b < 2 * L.n_bins; b += kChunkThreads) s_side[b] = 0;
    // this thread's events (in flight while the bases are summed up)
    int dmi[kChunkSteps], dma[kChunkSteps];
    uint32_t key[kChunkSteps], rank[kChunkSteps];
    double tt[kChunkSteps];
    const int wbase = chunk * kChunk + warp * (32 * kChunkSteps);
#pragma unroll
    for (int k = 0; k < kChunkSteps; ++k) {
      const int i = wbase + k * 32 + lane;
      dmi[k] = -1;
      dma[k] = 0;
      key[k] = 0;
      tt[k] = 0.0;
      if (i < ev.n) {
        const Ev e = load_event(ev, i);
        if (e.x < L.W && e.y < L.H) {
          dmi[k] = e.x / kFineW;
          dma[k] = e.y / kTileH;
          key[k] = (uint32_t)((e.y % kTileH) * kTileW + (e.x % kTileW)) | ((uint32_t)e.p << kPolShift);
          tt[k] = e.t;
        }
      }
    }
    bin_bases(T.cnt_a + (size_t)cam * L.max_chunks * kDigit, n_rows, chunk, s_base, s_red);
#pragma unroll
    for (int k = 0; k < kChunkSteps; ++k) rank[k] = bin_rank_step(s_wc[warp], dmi[k], dmi[k] >= 0, lt_mask);
    __syncthreads();
    bin_warp_prefix(s_wc);
    __syncthreads();
    double* __restrict__ it = B.it[cam];
    uint16_t* __restrict__ ik = B.ik[cam];
    uint8_t* __restrict__ im = B.im[cam];
    uint32_t* __restrict__ cb = T.cnt_b + (size_t)cam * L.max_chunks * kDigit;
#pragma unroll
    for (int k = 0; k < kChunkSteps; ++k) {
      if (dmi[k] >= 0) {
        const uint32_t pos = s_base[dmi[k]] + s_wc[warp][dmi[k]] + rank[k];
        it[pos] = tt[k];
        ik[pos] = (uint16_t)key[k];
        im[pos] = (uint8_t)dma[k];
        if (agg) {
          const int side = (int)(pos / kChunk) - (int)(s_base[dmi[k]] / kChunk);  // 0 or 1
          atomicAdd(&s_side[side * L.n_bins + dma[k] * minor_n + dmi[k]], 1u);
        } else {
          atomicAdd(cb + (size_t)(pos / kChunk) * kDigit + dma[k], 1u);
        }
      }
    }
    if (agg) {
      __syncthreads();
      for (int b = threadIdx.x; b < 2 * L.n_bins; b += kChunkThreads) {
        const uint32_t c = s_side[b];
        if (c) {
          const int side = b >= L.n_bins, bin = b - side * L.n_bins;
          const int ma = bin / minor_n, mi = bin - ma * minor_n;
          atomicAdd(cb + (size_t)(s_base[mi] / kChunk + side) * kDigit + ma, c);
        }
      }
    }
  }
  if (chunk != (int)gridDim.x - 1) return;
  // ---- the extra CTA of the camera: bin_start = exclusive scan of the fine-tile totals
  // (complete since k_bin_hist); the totals are cleared for the next window that uses this table
  const int nb = L.n_bins + 1;
  uint32_t* __restrict__ tot = bin_total + (size_t)cam * nb;
  uint32_t* __restrict__ bs = bin_start + (size_t)cam * (nb + 1);
  // thread t owns bins [t * per, (t + 1) * per): all of its totals are loaded in one batch
  constexpr int kRegPer = 16;  // kept in registers up to 4 096 bins (640x480: 10 per thread)
  const int per = (L.n_bins + kChunkThreads - 1) / kChunkThreads;
  const int b_lo = threadIdx.x * per;
  uint32_t held[kRegPer];
  uint32_t mine = 0;
#pragma unroll
  for (int j = 0; j < kRegPer; ++j) {
    const int b = b_lo + j;
    held[j] = (j < per && b < L.n_bins) ? __ldcg(tot + b) : 0u;
  }
#pragma unroll
  for (int j = 0; j < kRegPer; ++j) mine += held[j];
  for (int j = kRegPer; j < per; ++j)
    if (b_lo + j < L.n_bins) mine += __ldcg(tot + b_lo + j);
  uint32_t incl = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += o;
  }
  if (lane == 31) s_scan[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < kBinWarps ? s_scan[lane] : 0u;
#pragma unroll
    for (int d = 1; d < kBinWarps; d <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, w, d);
      if (lane >= d) w += o;
    }
    if (lane < kBinWarps) s_scan[lane] = w;  // inclusive over warps
  }
  __syncthreads();
  uint32_t run = (warp ? s_scan[warp - 1] : 0u) + incl - mine;
#pragma unroll
  for (int j = 0; j < kRegPer; ++j) {
    const int b = b_lo + j;
    if (j < per && b < L.n_bins) {
      bs[b] = run;
      tot[b] = 0;
      run += held[j];
    }
  }
  for (int j = kRegPer; j < per; ++j) {
    const int b = b_lo + j;
    if (b < L.n_bins) {
      const uint32_t v = __ldcg(tot + b);
      bs[b] = run;
      tot[b] = 0;
      run += v;
    }
  }
  if (threadIdx.x == kChunkThreads - 1) {
    // bins beyond the last: the out-of-range bin and the end marker both start at the total
    bs[L.n_bins] = run;
    bs[nb] = run;
  }
}

__global__ void __launch_bounds__(kChunkThreads)
k_bin_pass2(BinLayout L, const __grid_constant__ CamBatch B, BinTables T) {
  PDL_PROLOGUE();
  __shared__ __align__(16) uint32_t s_red[kBinWarps][2][kDigit];
  __shared__ uint32_t s_wc[kBinWarps][kDigit];
  __shared__ uint32_t s_base[kDigit + 1];
  const int cam = blockIdx.y, chunk = blockIdx.x;
  const int n_rows = B.n_chunks[cam];
  if (chunk >= n_rows) return;
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  for (int b = threadIdx.x; b < kBinWarps * kDigit; b += blockDim.x) (&s_wc[0][0])[b] = 0;
  // the in-range events of the camera, sorted by minor digit: n = sum of the table
  const int n = (int)bin_bases(T.cnt_b + (size_t)cam * L.max_chunks * kDigit, n_rows, chunk, s_base, s_red);
  if (chunk * kChunk >= n) return;
  const double* __restrict__ it = B.it[cam];
  const uint16_t* __restrict__ ik = B.ik[cam];
  const uint8_t* __restrict__ im = B.im[cam];
  int dma[kChunkSteps];
  uint32_t key[kChunkSteps], rank[kChunkSteps];
  double tt[kChunkSteps];
  const int wbase = chunk * kChunk + warp * (32 * kChunkSteps);
#pragma unroll
  for (int k = 0; k < kChunkSteps; ++k) {
    const int i = wbase + k * 32 + lane;
    dma[k] = -1;
    key[k] = 0;
    tt[k] = 0.0;
    if (i < n) {
      dma[k] = im[i];
      key[k] = ik[i];
      tt[k] = it[i];
    }
  }
#pragma unroll
  for (int k = 0; k < kChunkSteps; ++k) rank[k] = bin_rank_step(s_wc[warp], dma[k], dma[k] >= 0, lt_mask);
  __syncthreads();
  bin_warp_prefix(s_wc);
  __syncthreads();
  double* __restrict__ bt = B.bt[cam];
  uint16_t* __restrict__ bk = B.bk[cam];
#pragma unroll
  for (int k = 0; k < kChunkSteps; ++k) {
    if (dma[k] >= 0) {
      const uint32_t pos = s_base[dma[k]] + s_wc[warp][dma[k]] + rank[k];
      bt[pos] = tt[k];
      bk[pos] = (uint16_t)key[k];
    }
  }
}

// sensor sizes the two 7-bit digits cover
int bin_configure(const BinLayout& L) {
  return (L.tiles_x * kFine <= kDigit && L.tiles_y <= kDigit) ? 0 : -1;
}

int event_stage_alloc(const BinLayout& L, int n_cams, int cap, EventStageBuffers* E) {
  const size_t nb = (size_t)L.n_bins + 1;
  E->n_cams = n_cams;
  const size_t tab = (size_t)n_cams * L.max_chunks * kDigit;
#define ESA(call) do { if ((call) != cudaSuccess) return -1; } while (0)
  ESA(cudaMalloc(&E->counts, 2 * tab * sizeof(uint32_t)));
  ESA(cudaMemset(E->counts, 0, 2 * tab * sizeof(uint32_t)));
  ESA(cudaMalloc(&E->bin_total, n_cams * nb * sizeof(uint32_t)));
  ESA(cudaMemset(E->bin_total, 0, n_cams * nb * sizeof(uint32_t)));
  ESA(cudaMalloc(&E->bin_start, n_cams * (nb + 1) * sizeof(uint32_t)));
  ESA(cudaMemset(E->bin_start, 0, n_cams * (nb + 1) * sizeof(uint32_t)));
  for (int c = 0; c < n_cams; ++c) {
    ESA(cudaMalloc(&E->bt[c], (size_t)cap * sizeof(double)));
    ESA(cudaMalloc(&E->bk[c], (size_t)cap * sizeof(uint16_t)));
    ESA(cudaMalloc(&E->it[c], (size_t)cap * sizeof(double)));
    ESA(cudaMalloc(&E->ik[c], (size_t)cap * sizeof(uint16_t)));
    ESA(cudaMalloc(&E->im[c], (size_t)cap));
  }
#undef ESA
  return 0;
}

// The part of a two-camera buffer set that belongs to ONE camera, as a one-camera set of its own
// (left-first windows run the cameras' event stages as separate launches): the per-camera arrays,
// the camera's rows of bin_total / bin_start, and one half of the histogram table (its two pass
// tables laid out for n_cams = 1; the table is scratch, rebuilt by every k_bin_hist).
EventStageBuffers event_stage_cam_view(const BinLayout& L, const EventStageBuffers& E, int cam) {
  EventStageBuffers v = E;
  v.n_cams = 1;
  v.counts = E.counts + (size_t)cam * 2 * L.max_chunks * kDigit;
  v.bin_total = E.bin_total + (size_t)cam * (L.n_bins + 1);
  v.bin_start = E.bin_start + (size_t)cam * (L.n_bins + 2);
  v.bt[0] = E.bt[cam], v.bk[0] = E.bk[cam], v.it[0] = E.it[cam], v.ik[0] = E.ik[cam], v.im[0] = E.im[cam];
  for (int c = 1; c < kMaxCams; ++c) v.bt[c] = nullptr, v.bk[c] = nullptr, v.it[c] = nullptr, v.ik[c] = nullptr, v.im[c] = nullptr;
  return v;
}

void event_stage_free(EventStageBuffers* E) {
  for (int c = 0; c < kMaxCams; ++c) {
    cudaFree(E->bt[c]), cudaFree(E->bk[c]), cudaFree(E->it[c]), cudaFree(E->ik[c]), cudaFree(E->im[c]);
    E->bt[c] = nullptr, E->bk[c] = nullptr, E->it[c] = nullptr, E->ik[c] = nullptr, E->im[c] = nullptr;
  }
  cudaFree(E->counts), cudaFree(E->bin_total), cudaFree(E->bin_start);
  E->counts = E->bin_total = E->bin_start = nullptr;
}

// a window that was abandoned half way (reset) may have left totals behind
void event_stage_clear(const BinLayout& L, const EventStageBuffers& E, cudaStream_t s) {
  cudaMemsetAsync(E.bin_total, 0, (size_t)E.n_cams * (L.n_bins + 1) * sizeof(uint32_t), s);
}

void launch_bin_events(const BinLayout& L, const EventStageBuffers& B, const DevEvents* ev,
                       cudaStream_t s, int64_t* launches) {
  CamBatch cb;
  cb.n_cams = B.n_cams;
  int nc = 0;
  for (int c = 0; c < kMaxCams; ++c) {
    const bool on = c < B.n_cams;
    cb.ev[c] = on ? ev[c] : DevEvents{};
    cb.bt[c] = on ? B.bt[c] : nullptr;
    cb.bk[c] = on ? B.bk[c] : nullptr;
    cb.it[c] = on ? B.it[c] : nullptr;
    cb.ik[c] = on ? B.ik[c] : nullptr;
    cb.im[c] = on ? B.im[c] : nullptr;
    cb.n_chunks[c] = on ? (ev[c].n + kChunk - 1) / kChunk : 0;
    if (cb.n_chunks[c] > nc) nc = cb.n_chunks[c];
  }
  BinTables T;
  T.cnt_a = B.counts;
  T.cnt_b = B.counts + (size_t)B.n_cams * L.max_chunks * kDigit;
  const size_t agg = L.n_bins <= kAggBins ? (size_t)L.n_bins * sizeof(uint32_t) : 0;
  if (nc > 0) {
    launch_pdl(k_bin_hist, dim3(nc, B.n_cams), dim3(kChunkThreads), agg, s, L, cb, T, B.bin_total);
    ++*launches;
  }
  // one more CTA per camera than there are chunks: it writes bin_start
  launch_pdl(k_bin_pass1, dim3(nc + 1, B.n_cams), dim3(kChunkThreads), 2 * agg, s, L, cb, T, B.bin_total,
             B.bin_start);
  ++*launches;
  if (nc > 0) {
    launch_pdl(k_bin_pass2, dim3(nc, B.n_cams), dim3(kChunkThreads), 0, s, L, cb, T);
    ++*launches;
  }
}

// =====================================================================================
// K1: fused SAE update + time surface
// =====================================================================================
// One CTA (kFine warps) per 32x8 tile of one camera.  The tile's `sae` and `lat` state
// (2 x 4 KB) arrives by TMA; warp w applies the events of fine tile w (its own 16x8 pixel
// block, so the warps never touch the same pixel) in stream order, 32 events per step: the
// lanes of one step that hit the same pixel find their predecessors with match.any / ballot
// and decide acceptance from them instead of replaying the events one by one.  The time
// surface is computed from shared memory (4 pixels per thread) and dirty tiles go back by
// TMA store.
constexpr int kSaeThreads = 32 * kFine;
constexpr int kSaeAhead = 4;  // 32-event steps loaded per round trip to L2/HBM
constexpr int kSaePrefetchDist = 1036;  // CTAs: half a wave ahead (sweep: profiles/r2_k1_prefetch.txt)

// convertTo(CV_8U) of a double: cvRound (half to even) then saturate
__device__ __forceinline__ uint8_t sat_u8(double v) {
  int r = __double2int_rn(v);
  r = r < 0 ? 0 : (r > 255 ? 255 : r);
  return (uint8_t)r;
}

struct SaeMaps {
  CUtensorMap sae, lat;  // f64 [cams][H][2W], box {2 * kTileW, kTileH, 1}
};

// 2^(j/32), j = 0..31, rounded to nearest
__constant__ double c_exp2_32[32] = {
    1.0, 1.0218971486541166, 1.0442737824274138, 1.0671404006768237, 1.0905077326652577,
    1.1143867425958924, 1.1387886347566916, 1.1637248587775775, 1.189207115002721,
    1.215247359980469, 1.241857812073484, 1.2690509571917332, 1.2968395546510096,
    1.3252366431597413, 1.3542555469368927, 1.383909881963832, 1.4142135623730951,
    1.4451808069770467, 1.4768261459394993, 1.5091644275934228, 1.5422108254079407,
    1.5759808451078865, 1.6104903319492543, 1.645755478153965, 1.681792830507429,
    1.718619298122478, 1.7562521603732995, 1.7947090750031072, 1.8340080864093424,
    1.8741676341103, 1.9152065613971474, 1.9571441241754002};

__device__ __forceinline__ float exp2f_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// exp(a) for a in [-40, 8] to < 3 ulp: a = (32 m + j) ln2/32 + r, |r| <= ln2/64;
// exp(a) = 2^m * 2^(j/32) * P6(r).  Straight-line (no branches) so that the four pixels of a
// thread interleave.  The CV_8U value derived from it equals the one derived from a
// correctly rounded exp unless 127.5 * e falls within ~1e-14 of a rounding boundary.
__device__ __forceinline__ double exp_small(double a, const double* __restrict__ tab) {
  const double kMagic = 6755399441055744.0;  // 1.5 * 2^52: the sum's low word is rint(a * 32/ln2)
  const double kf = fma(a, 46.16624130844683, kMagic);
  const int k = __double2loint(kf);
  const double kd = kf - kMagic;
  double r = fma(-kd, 0x1.62e42fefa3000p-6, a);  // ln2/32 hi (40 bits: kd * hi is exact)
  r = fma(-kd, 0x1.3de6af278ece6p-47, r);         // ln2/32 lo
  double p = fma(r, 1.0 / 720.0, 1.0 / 120.0);
  p = fma(p, r, 1.0 / 24.0);
  p = fma(p, r, 1.0 / 6.0);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const double scale = __hiloint2double((1023 + (k >> 5)) << 20, 0);
  return tab[k & 31] * p * scale;
}

// SAEtoTimeSurface_* (event_detector.cc:230-267).  The reference evaluates 127.5 * (+-e) + 127.5
// with e = exp(-dt / decay) in double (cv::MatExpr folds 255 * (m + 1) / 2 into one scale +
// shift; 255 * e when polarity is ignored) and rounds half to even into CV_8U.
//
// ts_exact: that computation for one pixel, in double.  Three regimes of a = -dt / decay:
//   a >= -7      e = exp(a) decides the value: computed (< 3 ulp, table + polynomial);
//   a <  -7      127.5 * e < 0.12: the sum rounds to 128 (positive) / 127 (negative) whatever
//                the last bits of e are -- until 127.5 * e drops below half an ulp of 127.5
//                (2^-47, at a = -37.43, i.e. 0.75 s of silence): then the double sum IS 127.5 and
//                rounds to the even 128 for BOTH polarities.  Around that edge (a in
//                [-38.5, -36.5]) the sum is evaluated for real again.
constexpr double kTsExpFrom = -7.0, kTsEdgeLo = -38.5, kTsEdgeHi = -36.5;
__device__ __noinline__ uint32_t ts_exact(double stamp, bool pos, double t_ref, double decay_sec,
                                          double inv_decay, int ignore_polarity,
                                          const double* __restrict__ tab) {
  // -dt / decay_sec, correctly rounded (Markstein: q + (n - q*d) * RN(1/d) with FMAs)
  const double n = -(t_ref - stamp);
  const double q = n * inv_decay;
  const double a = fma(fma(-q, decay_sec, n), inv_decay, q);
  if (!(a >= kTsExpFrom || (a >= kTsEdgeLo && a <= kTsEdgeHi)))
    return ignore_polarity ? 0u : ((pos || a < kTsEdgeLo) ? 128u : 127u);
  double e = exp_small(fmin(fmax(a, -40.0), 8.0), tab);
  if (!ignore_polarity && !pos) e = -e;
  return sat_u8(ignore_polarity ? e * 255.0 : e * 127.5 + 127.5);
}

// Four pixels of one column (rows r, r+1, r+2, r+3 of the tile: a warp reads 32 adjacent
// double2 per row, free of bank conflicts).  Almost every pixel is settled by a float estimate: v ~ shift +
// scale * 2^x, x = (float)dt * karg, karg = -log2(e) / decay.  Its error against the double
// value is < 4e-4 grey levels (three float roundings of an exponent <= 10.2 in magnitude: 1.3e-6
// relative in e, ex2.approx 2.4e-7, times <= 255; the FFMA rounds to 2^-17), so the rounded
// value is the reference's unless the estimate lies within kTsGuard of a rounding boundary --
// those pixels (0.3 % of the recently hit ones) take ts_exact.  Regimes of x:
//   x >= -10.2 (a >= -7.07)      the estimate;
//   -53.97 > x > -10.2           the constants 128 / 127 (valid for any a < -5.6);
//   x in [-54.02, -53.97]        the far edge: 127.5 * e crosses half an ulp of 127.5 at
//                                x* = -(47 + log2(127.5)) = -53.9943; ts_exact decides;
//   x < -54.02                   128 for both polarities (the double sum is exactly 127.5).
constexpr float kTsGuard = 1.5e-3f;
constexpr float kTsXFast = -10.2f, kTsXEdgeHi = -53.97f, kTsXEdgeLo = -54.02f;
__device__ __forceinline__ uchar4 ts_pixel4(const double2* __restrict__ px, const SaeTsParams& P,
                                            const double t_ref, const double* __restrict__ tab) {
  // px[i * kTileW]: the pixel of row i
  const float scale = P.ignore_polarity ? 255.f : 127.5f, shift = P.ignore_polarity ? 0.f : 127.5f;
  const uint32_t none = P.ignore_polarity ? 0u : 128u;  // never hit, or silent beyond the far edge
  uint32_t o[4];
  double stamp[4];
  bool pos[4];
  uint32_t slow = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double2 v = px[i * kTileW];
    pos[i] = v.y > v.x;
    stamp[i] = pos[i] ? v.y : v.x;
    const bool hit = stamp[i] > 0.0;
    const float x = fminf((float)(t_ref - stamp[i]) * P.ts_karg, 1.f);
    const float e = exp2f_approx(x);
    const float est = fmaf((P.ignore_polarity || pos[i]) ? e : -e, scale, shift);
    const float r = rintf(est);
    int vi = (int)r;
    vi = vi < 0 ? 0 : (vi > 255 ? 255 : vi);
    const bool fast = x >= kTsXFast;
    const uint32_t konst = (P.ignore_polarity || x < kTsXEdgeLo) ? none : (pos[i] ? 128u : 127u);
    o[i] = !hit ? none : (fast ? (uint32_t)vi : konst);
    const bool edge = !P.ignore_polarity && x <= kTsXEdgeHi && x >= kTsXEdgeLo;
    if (hit && ((fast && fabsf(est - r) > 0.5f - kTsGuard) || edge)) slow |= 1u << i;
  }
  if (slow) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (slow & (1u << i))
        o[i] = ts_exact(stamp[i], pos[i], t_ref, P.decay_sec, P.inv_decay, P.ignore_polarity, tab);
  }
  return make_uchar4((uint8_t)o[0], (uint8_t)o[1], (uint8_t)o[2], (uint8_t)o[3]);
}

// DBG: perf-experiment switches (0 in the product): 1 skip events, 2 skip TS, 4 skip store
template <int DBG>
__global__ void __launch_bounds__(kSaeThreads)
k_sae_update_ts(const __grid_constant__ SaeMaps maps, const __grid_constant__ SaeTsParams P) {
  PDL_PROLOGUE();
  __shared__ __align__(128) double2 s_sae[kTilePx];
  __shared__ __align__(128) double2 s_lat[kTilePx];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ double s_exp2[32];

  const int warp = threadIdx.x >> 5, lane = lane_id();
  const int cam = blockIdx.z, tile = blockIdx.y * P.tiles_x + blockIdx.x;
  const int x0 = blockIdx.x * kTileW, y0 = blockIdx.y * kTileH;

  if (threadIdx.x == 0) {
    mbar_init(&s_bar, 1);
    mbar_fence_init();
  }
  const uint32_t* bs = P.bin_start + cam * (P.n_tiles * kFine + 2) + tile * kFine;
  const int seg_begin = (int)bs[warp], seg_end = (DBG & 1) ? seg_begin : (int)bs[warp + 1];
  const bool dirty = bs[kFine] > bs[0];
  if (threadIdx.x == 0) {
    mbar_expect_tx(&s_bar, (dirty ? 2u : 1u) * (uint32_t)(kTilePx * sizeof(double2)));
    tma_load_3d(s_sae, &maps.sae, &s_bar, 2 * x0, y0, cam);
    if (dirty) tma_load_3d(s_lat, &maps.lat, &s_bar, 2 * x0, y0, cam);
  }
  // The grid is several waves deep and CTAs start in blockIdx order: pull the state of the tile
  // a CTA `prefetch_dist` launches later will want into L2 now, so that its TMA load is an L2
  // hit instead of a trip to HBM (the launch is latency-bound per CTA: load -> replay -> store).
  if (P.prefetch_dist > 0 && threadIdx.x == 32) {
    int fx = (int)blockIdx.x + P.pf_dx, fy = (int)blockIdx.y + P.pf_dy, fc = cam + P.pf_dz;
    if (fx >= (int)gridDim.x) fx -= gridDim.x, ++fy;
    if (fy >= (int)gridDim.y) fy -= gridDim.y, ++fc;
    if (fc < P.n_cams) {
      const uint32_t* fbs = P.bin_start + fc * (P.n_tiles * kFine + 2) + (fy * P.tiles_x + fx) * kFine;
      tma_prefetch_3d(&maps.sae, 2 * fx * kTileW, fy * kTileH, fc);
      if (fbs[kFine] > fbs[0]) tma_prefetch_3d(&maps.lat, 2 * fx * kTileW, fy * kTileH, fc);
    }
  }
  // the first events of the run travel while the tile loads
  const double* __restrict__ bt = P.bt[cam];
  const uint16_t* __restrict__ bk = P.bk[cam];
  double t_nx[kSaeAhead];
  uint32_t k_nx[kSaeAhead];
#pragma unroll
  for (int j = 0; j < kSaeAhead; ++j) {
    const int i = seg_begin + j * 32 + lane;
    t_nx[j] = 0.0;
    k_nx[j] = 0;
    if (i < seg_end) {
      t_nx[j] = __ldg(bt + i);
      k_nx[j] = __ldg(bk + i);
    }
  }
  if (threadIdx.x < 32) s_exp2[threadIdx.x] = c_exp2_32[threadIdx.x];
  __syncthreads();  // barrier initialised before anybody polls it
  mbar_wait(&s_bar, 0);

  const uint32_t lt_mask = (1u << lane) - 1u;
  const uint32_t gt_mask = ~lt_mask & ~(1u << lane);
  for (int base = seg_begin; base < seg_end; base += 32 * kSaeAhead) {
    double t_cu[kSaeAhead];
    uint32_t k_cu[kSaeAhead];
#pragma unroll
    for (int j = 0; j < kSaeAhead; ++j) {
      t_cu[j] = t_nx[j];
      k_cu[j] = k_nx[j];
      const int i = base + (kSaeAhead + j) * 32 + lane;
      if (i < seg_end) {
        t_nx[j] = __ldg(bt + i);
        k_nx[j] = __ldg(bk + i);
      }
    }
#pragma unroll
    for (int j = 0; j < kSaeAhead; ++j) {
      if (base + j * 32 < seg_end) {  // warp-uniform
        const bool valid = base + j * 32 + lane < seg_end;
        const double t = t_cu[j];
        const int pix = k_cu[j] & 0xff, pol = (k_cu[j] >> kPolShift) & 1;
        // lanes of this step on my pixel, split by polarity (createSAE_*,
        // event_detector.cc:157: t_last = latest[pol], t_last_inv = latest[!pol] as left by
        // the events before me)
        const uint32_t grp = __match_any_sync(0xffffffffu, valid ? pix : (0x100 + lane));
        const uint32_t pos_lanes = __ballot_sync(0xffffffffu, valid && pol);
        const uint32_t same = grp & (pol ? pos_lanes : ~pos_lanes);
        const uint32_t before_same = same & lt_mask, before_opp = grp & ~same & lt_mask;
        double2 st = make_double2(0.0, 0.0);
        if (valid) st = s_lat[pix];
        const double t_bs = __shfl_sync(0xffffffffu, t, before_same ? 31 - __clz(before_same) : lane);
        const double t_bo = __shfl_sync(0xffffffffu, t, before_opp ? 31 - __clz(before_opp) : lane);
        const double prev_same = before_same ? t_bs : (pol ? st.y : st.x);
        const double prev_opp = before_opp ? t_bo : (pol ? st.x : st.y);
        const bool accept = valid && (t > prev_same + P.filter_threshold || prev_opp > prev_same);
        const uint32_t acc = __ballot_sync(0xffffffffu, accept);
        __syncwarp();
        if (valid && (same & gt_mask) == 0) reinterpret_cast<double*>(&s_lat[pix])[pol] = t;
        if (accept && (same & acc & gt_mask) == 0) reinterpret_cast<double*>(&s_sae[pix])[pol] = t;
        __syncwarp();
      }
    }
  }
  __syncthreads();

  // time surface of the tile straight from shared memory: lane = column, warp w = rows 4w..4w+3.
  // Columns >= W of the last tile (zero-filled by TMA) land in the row padding of the image.
  if (!(DBG & 2)) {
    const int r0 = warp * (kTileH / kFine);
    const uchar4 v = ts_pixel4(&s_sae[r0 * kTileW + lane], P, P.t_ref[cam], s_exp2);
    uint8_t* __restrict__ out = P.ts[cam] + (size_t)(y0 + r0) * P.ts_pitch + x0 + lane;
    const uint8_t b[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (y0 + r0 + i < P.H) out[(size_t)i * P.ts_pitch] = b[i];
  }

  if (dirty && !(DBG & 4)) {
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
      tma_store_3d(&maps.sae, s_sae, 2 * x0, y0, cam);
      tma_store_3d(&maps.lat, s_lat, 2 * x0, y0, cam);
      tma_store_commit();
      tma_store_wait_read0();
    }
  }
}

void launch_sae_update_ts(const SaeTsParams& P_in, const CUtensorMap& map_sae,
                          const CUtensorMap& map_lat, cudaStream_t s, int64_t* launches) {
  SaeTsParams P = P_in;
  {
    // half a wave of CTAs ahead (148 SMs x ~14 resident CTAs); experiments: ESVIO_K1_PREFETCH=<n>
    static const int dist = getenv("ESVIO_K1_PREFETCH") ? atoi(getenv("ESVIO_K1_PREFETCH")) : kSaePrefetchDist;
    P.prefetch_dist = dist;
    const int gx = P.tiles_x, gy = P.n_tiles / P.tiles_x;
    P.pf_dx = dist % gx;
    P.pf_dy = (dist / gx) % gy;
    P.pf_dz = dist / (gx * gy);
    P.ts_karg = (float)(-1.4426950408889634 / P.decay_sec);
  }
  SaeMaps maps;
  maps.sae = map_sae;
  maps.lat = map_lat;
  const dim3 grid(P.tiles_x, P.n_tiles / P.tiles_x, P.n_cams);
#ifdef ESVIO_K1_EXPERIMENTS  // perf experiments only (scratch/stage_times.py); never in the product build
  static const int dbg = getenv("ESVIO_K1_DBG") ? atoi(getenv("ESVIO_K1_DBG")) : 0;
  switch (dbg) {
    case 1: k_sae_update_ts<1><<<grid, kSaeThreads, 0, s>>>(maps, P); break;
    case 2: k_sae_update_ts<2><<<grid, kSaeThreads, 0, s>>>(maps, P); break;
    case 3: k_sae_update_ts<3><<<grid, kSaeThreads, 0, s>>>(maps, P); break;
    case 4: k_sae_update_ts<4><<<grid, kSaeThreads, 0, s>>>(maps, P); break;
    case 7: k_sae_update_ts<7><<<grid, kSaeThreads, 0, s>>>(maps, P); break;
    default: k_sae_update_ts<0><<<grid, kSaeThreads, 0, s>>>(maps, P); break;
  }
#else
  launch_pdl(k_sae_update_ts<0>, grid, dim3(kSaeThreads), 0, s, maps, P);
#endif
  ++*launches;
}

// =====================================================================================
// K2: Arc* corner flags
// =====================================================================================
__constant__ int8_t c_ring3[16][2] = {{0, 3},  {1, 3},   {2, 2},   {3, 1},  {3, 0},  {3, -1},
                                      {2, -2}, {1, -3},  {0, -3},  {-1, -3}, {-2, -2}, {-3, -1},
                                      {-3, 0}, {-3, 1},  {-2, 2},  {-1, 3}};
__constant__ int8_t c_ring4[20][2] = {{0, 4},   {1, 4},   {2, 3},   {3, 2},  {4, 1},
                                      {4, 0},   {4, -1},  {3, -2},  {2, -3}, {1, -4},
                                      {0, -4},  {-1, -4}, {-2, -3}, {-3, -2}, {-4, -1},
                                      {-4, 0},  {-4, 1},  {-3, 2},  {-2, 3}, {-1, 4}};

// One circle of the Arc* test (event_detector.cc:337-435 / 441-540): start at the newest
// ring element, repeatedly extend the arm (clockwise or counter-clockwise) whose next
// element is newer, and remember the longest prefix whose elements are all newer than
// everything outside it.
// STRIDE: distance between consecutive ring elements (1 = a thread's own array; > 1 = a column of
// a shared-memory array [element][thread], which keeps the dynamically indexed ring on chip)
template <int N, int LO, int HI, int STRIDE = 1>
__device__ __forceinline__ bool arc_ring_valid(const double* ring_base) {
  auto R = [&](int i) { return ring_base[i * STRIDE]; };
  int newest = 0;
#pragma unroll
  for (int i = 1; i < N; ++i)
    if (R(i) > R(newest)) newest = i;
  double seg_min = R(newest);
  int cw = (newest + 1) % N, ccw = (newest + N - 1) % N;
  double cw_v = R(cw), ccw_v = R(ccw), cw_min = cw_v, ccw_min = ccw_v;
  int seg_len = LO;
  for (int it = 1; it < N; ++it) {
    const bool take_cw = cw_v > ccw_v;
    const double v = take_cw ? cw_v : ccw_v;
    const double vmin = take_cw ? cw_min : ccw_min;
    if (it < LO) {
      seg_min = fmin(seg_min, vmin);
    } else if (v >= seg_min) {
      seg_len = it + 1;
      seg_min = fmin(seg_min, vmin);
    }
    if (take_cw) {
      cw = (cw + 1) % N;
      cw_v = R(cw);
      cw_min = fmin(cw_min, cw_v);
    } else {
      ccw = (ccw + N - 1) % N;
      ccw_v = R(ccw);
      ccw_min = fmin(ccw_min, ccw_v);
    }
  }
  return seg_len <= HI || (seg_len >= N - HI && seg_len <= N - LO);
}

// Everything of EventDetector::isCorner (event_detector.cc:308-544) that depends on the pixel and
// the polarity only -- the whole test except "t > latest[p] + threshold", which needs the
// event's own time: not a live time-surface pixel -> no; the opposite polarity fired later ->
// no; within MIN_DIST + 1 of the border -> no; then the two Arc* circles.  `ring(dx, dy)` returns
// sae[pol] at (x + dx, y + dy).
__device__ __forceinline__ bool corner_pixel_cheap(const CornerParams& P, int x, int y, int pol, double2 lat) {
  if (P.and_ts_test && (double)P.ts[(size_t)y * P.ts_pitch + x] == P.ts_lk_threshold) return false;
  const double last_same = pol ? lat.y : lat.x, last_opp = pol ? lat.x : lat.y;
  if (last_opp > last_same) return false;
  const int border = P.min_dist + 1;
  return !(x < border || x >= P.W - border || y < border || y >= P.H - border);
}
template <class Ring>
__device__ __forceinline__ bool corner_pixel_rings(Ring ring) {
  double r[20];
#pragma unroll
  for (int k = 0; k < 16; ++k) r[k] = ring(c_ring3[k][0], c_ring3[k][1]);
  if (!arc_ring_valid<16, 4, 6>(r)) return false;
#pragma unroll
  for (int k = 0; k < 20; ++k) r[k] = ring(c_ring4[k][0], c_ring4[k][1]);
  return arc_ring_valid<20, 5, 8>(r);
}
template <class Ring>
__device__ __forceinline__ bool corner_pixel_test(const CornerParams& P, int x, int y, int pol, double2 lat,
                                                  Ring ring) {
  return corner_pixel_cheap(P, x, y, pol, lat) && corner_pixel_rings(ring);
}

// k_corner_plane: the pixel part of the Arc* test ONCE per (pixel, polarity) that an event of the
// window can ask about, instead of once per event.  One CTA per 32x8 tile: the tile's sae planes
// plus a 4-pixel halo go to shared memory (10 KB; the per-event kernel pulled 36 scattered
// 32-byte sectors per event through L2, 190 MB per 640x480 window at 5 Mev/s, and took 40 us),
// the (pixel, polarity) pairs whose last event is not older than the window's first event are
// settled -- the cheap conditions per pixel, the circles compacted and tested densely -- and the
// verdicts land in the byte plane: bit p = verdict, bit 2 + p = "tested".
constexpr int kCpHalo = 4, kCpW = kTileW + 2 * kCpHalo, kCpH = kTileH + 2 * kCpHalo;
__global__ void __launch_bounds__(kTileW* kTileH) k_corner_plane(CornerParams P, DevEvents ev) {
  PDL_PROLOGUE();
  __shared__ double2 s_sae[kCpH][kCpW];
  __shared__ uint16_t s_item[2 * kTileW * kTileH];
  __shared__ uint32_t s_bits[kTileW * kTileH];
  __shared__ int s_wc[kTileW * kTileH / 32 + 1];
  const int tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
  const int x0 = blockIdx.x * kTileW, y0 = blockIdx.y * kTileH;
  for (int i = tid; i < kCpW * kCpH; i += kTileW * kTileH) {
    const int r = i / kCpW, c = i - r * kCpW;
    const int gx = x0 - kCpHalo + c, gy = y0 - kCpHalo + r;
    double2 v = make_double2(0.0, 0.0);
    if (gx >= 0 && gx < P.W && gy >= 0 && gy < P.H) v = P.sae[(size_t)gy * P.W + gx];
    s_sae[r][c] = v;
  }
  const int tx = tid % kTileW, ty = tid / kTileW;
  const int x = x0 + tx, y = y0 + ty;
  const bool inside = x < P.W && y < P.H;
  double2 l = make_double2(0.0, 0.0);
  if (inside) l = P.lat[(size_t)y * P.W + x];
  // The pairs an event of this window can ask about: those whose last event is not older than
  // the window's first event (minus the filter threshold: an event passes "t <= latest + thr").
  // Events are time-ascending in the reference's windows; for any other order k_corner_flags
  // tests what is missing here.  The cheap parts of the test are settled right away, only the
  // pairs that reach the two circles are compacted over the CTA.
  const double t_hint = load_event(ev, 0).t - P.filter_threshold - 1e-6;
  int want = 0, tested = 0;
  if (inside) {
#pragma unroll
    for (int pol = 0; pol < 2; ++pol) {
      const double last = pol ? l.y : l.x;
      if (last >= t_hint && last > 0.0) {
        tested |= 4 << pol;
        if (corner_pixel_cheap(P, x, y, pol, l)) want |= 1 << pol;
      }
    }
  }
  s_bits[tid] = (uint32_t)tested;
  const int cnt = __popc(want);
  int incl = cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += o;
  }
  if (lane == 31) s_wc[warp] = incl;
  __syncthreads();
  int base = 0, n_items = 0;
#pragma unroll
  for (int w = 0; w < kTileW * kTileH / 32; ++w) {
    const int c = s_wc[w];
    base += w < warp ? c : 0;
    n_items += c;
  }
  int pos = base + incl - cnt;
  if (want & 1) s_item[pos++] = (uint16_t)(tid << 1);
  if (want & 2) s_item[pos] = (uint16_t)(tid << 1 | 1);
  __syncthreads();
  for (int i = tid; i < n_items; i += kTileW * kTileH) {
    const int it = s_item[i], pol = it & 1, px = it >> 1;
    const int ix = px % kTileW, iy = px / kTileW;
    const double* S = reinterpret_cast<const double*>(&s_sae[iy + kCpHalo][ix + kCpHalo]) + pol;
    if (corner_pixel_rings([&](int dx, int dy) { return S[2 * (dy * kCpW + dx)]; })) atomicOr(&s_bits[px], 1u << pol);
  }
  __syncthreads();
  if (inside) P.plane[(size_t)y * P.W + x] = (uint8_t)s_bits[tid];
}

// k_corner_flags: the flag of every event.  With the pixel plane of k_corner_plane an event only
// checks its own time against latest[p] and picks up its pixel's verdict; a (pixel, polarity)
// the plane kernel did not test (an event older than its hint) is tested here, so the result
// never depends on the hint.  Besides the flags the kernel leaves, per CTA of kCornerBlock
// consecutive events, the flagged events' pixels compacted in stream order
// (cand[block * kCornerBlock + j], x | y << 16) and their number: the selection kernel then
// walks a few thousand candidates instead of re-reading the whole window.
__global__ void __launch_bounds__(kCornerBlock)
k_corner_flags(CornerParams P, DevEvents ev, uint8_t* __restrict__ flags) {
  PDL_PROLOGUE();
  __shared__ int s_wc[kCornerBlock / 32];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  uint8_t out = 0;
  Ev e;
  e.x = e.y = e.p = 0;
  e.t = 0.0;
  if (i < ev.n) {
    e = load_event(ev, i);
    do {
      if (e.x >= P.W || e.y >= P.H) break;
      const size_t px = (size_t)e.x + (size_t)e.y * P.W;
      const double2 l = P.lat[px];
      const double last_same = e.p ? l.y : l.x;
      if (e.t > last_same + P.filter_threshold) break;
      if (P.plane) {
        const uint32_t b = __ldcg(P.plane + px);  // written by the kernel before
        if ((b >> (2 + e.p)) & 1u) {
          out = (b >> e.p) & 1u;
          break;
        }
      }
      const double* S = reinterpret_cast<const double*>(P.sae) + e.p;
      const int ex = e.x, ey = e.y, W = P.W;
      out = corner_pixel_test(P, ex, ey, e.p, l, [&](int dx, int dy) {
        return S[2 * ((size_t)(ex + dx) + (size_t)(ey + dy) * W)];
      });
    } while (0);
    flags[i] = out;
  }
  if (P.cand == nullptr) return;
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const uint32_t m = __ballot_sync(0xffffffffu, out != 0);
  if (lane == 0) s_wc[warp] = __popc(m);
  __syncthreads();
  int off = 0, total = 0;
#pragma unroll
  for (int w = 0; w < kCornerBlock / 32; ++w) {
    if (w < warp) off += s_wc[w];
    total += s_wc[w];
  }
  if (out)
    P.cand[(size_t)blockIdx.x * kCornerBlock + off + __popc(m & ((1u << lane) - 1u))] =
        (uint32_t)e.x | ((uint32_t)e.y << 16);
  if (threadIdx.x == 0) P.cand_cnt[blockIdx.x] = total;
}

void launch_corner_flags(const CornerParams& P, const DevEvents& ev, uint8_t* flags,
                         cudaStream_t s, int64_t* launches) {
  if (ev.n <= 0) return;
  if (P.plane) {
    launch_pdl(k_corner_plane, dim3((P.W + kTileW - 1) / kTileW, (P.H + kTileH - 1) / kTileH),
               dim3(kTileW * kTileH), 0, s, P, ev);
    ++*launches;
  }
  launch_pdl(k_corner_flags, dim3((ev.n + kCornerBlock - 1) / kCornerBlock), dim3(kCornerBlock), 0, s, P, ev,
             flags);
  ++*launches;
}

// =====================================================================================
// time-window shard: element-wise maximum over state planes
// =====================================================================================
struct MergeSrc {
  const double2* p[kMaxMergeSrc];
};
__global__ void __launch_bounds__(256)
k_merge_max(double2* __restrict__ dst, const __grid_constant__ MergeSrc src, int n_src, size_t n2) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
    double2 m = src.p[0][i];
    for (int k = 1; k < n_src; ++k) {
      const double2 v = src.p[k][i];
      m.x = fmax(m.x, v.x);
      m.y = fmax(m.y, v.y);
    }
    dst[i] = m;
  }
}

void launch_merge_max(double* dst, const double* const* srcs, int n_src, size_t n, cudaStream_t s,
                      int64_t* launches) {
  if (n == 0 || n_src < 1) return;
  MergeSrc src;
  for (int k = 0; k < kMaxMergeSrc; ++k) src.p[k] = reinterpret_cast<const double2*>(srcs[k < n_src ? k : 0]);
  const size_t n2 = n / 2;  // planes are double2 arrays
  const int blocks = (int)((n2 + 255) / 256 < 148 * 8 ? (n2 + 255) / 256 : 148 * 8);
  k_merge_max<<<blocks, 256, 0, s>>>(reinterpret_cast<double2*>(dst), src, n_src, n2);
  ++*launches;
}

// see prefer_shared_lk (lk.cu): kernels that are to share an SM have to ask for the same
// shared-memory carve-out
void prefer_shared_events() {
  static const int pct = getenv("ESVIO_CARVEOUT_K1") ? atoi(getenv("ESVIO_CARVEOUT_K1"))
                         : getenv("ESVIO_CARVEOUT") ? atoi(getenv("ESVIO_CARVEOUT")) : 50;  // experiments
  cudaFuncSetAttribute(k_sae_update_ts<0>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
}

}  // namespace esvio
