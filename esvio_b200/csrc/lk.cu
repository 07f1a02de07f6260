// lk.cu -- pyramidal Lucas-Kanade, arithmetic-compatible with cv::calcOpticalFlowPyrLK as the
// reference calls it (feature_tracker/src/feature_tracker.cpp:410 temporal forward,
// :417-418 temporal backward with OPTFLOW_USE_INITIAL_FLOW and maxLevel 1, :490/:495 stereo
// forward/backward): 21x21 window, 30 iterations / eps 0.01, minEigThreshold 1e-4, 14-bit
// bilinear weights, 5-bit up-scaled intensities, int16 Scharr derivatives (OpenCV
// modules/video/src/lkpyramid.cpp -- not part of the reference tree; arithmetic spec in
// SURVEY.md section 8a).
//
// One CTA of 128 threads tracks one point through all pyramid levels, forward and (when
// asked) straight on into the backward check, inside a single launch.  Per level the
// template neighbourhood (24x24 intensities, 22x22 Scharr derivatives) and a 32x32 search
// region of J are staged in shared memory, so the <= 30 Newton iterations touch no global
// memory; the region is re-staged only if the window walks out of it.  The template
// (Iw, Ixw, Iyw) lives in registers, <= 4 pixels per thread; the normal-equation sums
// A11,A12,A22 and b1,b2 are accumulated as exact integers (redux.sync inside a warp,
// shared memory across the 4 warps) and rounded to float once -- OpenCV accumulates the
// same integers in float, so results agree to float rounding (~1e-4 px).
#include <float.h>

#include "common.cuh"

namespace esvio {

constexpr int kWBits = 14;
constexpr int kLkThreads = 128;
constexpr int kPxPerLane = (kWin * kWin + 31) / 32;  // 14
constexpr int kIP = 24;  // staged intensity patch (window + bilinear tap + Scharr ring)
constexpr int kDP = 22;  // derivative patch (window + bilinear tap)
constexpr int kJM = 5;   // margin of the staged search region
constexpr int kJR = kWin + 1 + 2 * kJM;  // 32

__device__ __forceinline__ int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

__device__ __forceinline__ void bilinear_weights(float a, float b, int& w00, int& w01, int& w10,
                                                 int& w11) {
  w00 = cv_round((1.f - a) * (1.f - b) * (float)(1 << kWBits));
  w01 = cv_round(a * (1.f - b) * (float)(1 << kWBits));
  w10 = cv_round((1.f - a) * b * (float)(1 << kWBits));
  w11 = (1 << kWBits) - w00 - w01 - w10;
}

struct LkShared {
  uint8_t I[kMaxLevels][kIP][kIP];      // template neighbourhoods of all levels
  short2 D[kMaxLevels][kDP][kDP];       // their Scharr derivatives
  short Tw[kMaxLevels][kPxPerLane * 32];  // templates: Iw, Ixw, Iyw per window pixel
  short Tx[kMaxLevels][kPxPerLane * 32];
  short Ty[kMaxLevels][kPxPerLane * 32];
  float A[kMaxLevels][3];
  int flag[kMaxLevels];  // 0 ok, 1 template window outside the image, 2 minEig / det test failed
  // search region of the level being iterated, as packed 2x2 neighbourhoods:
  // Q[r][c] = J(r,c) | J(r,c+1) << 8 | J(r+1,c) << 16 | J(r+1,c+1) << 24
  uint32_t Q[kJR][kJR];
  float2 np;
  int st;
};

// exact sum over the warp of one int32 per lane, as int64
__device__ __forceinline__ long long warp_sum_exact(int v) {
  const unsigned lo = __reduce_add_sync(0xffffffffu, (unsigned)v & 0xffffu);
  const int hi = __reduce_add_sync(0xffffffffu, v >> 16);
  return ((long long)hi << 16) + (long long)lo;
}

// exact warp sum of one int32 per lane, rounded to float once: the total is hi * 65536 + lo
// with |hi| < 2^24 and lo < 2^24, so both terms are exact floats and their sum rounds once
__device__ __forceinline__ float warp_sum_to_float(int v) {
  const unsigned lo = __reduce_add_sync(0xffffffffu, (unsigned)v & 0xffffu);
  const int hi = __reduce_add_sync(0xffffffffu, v >> 16);
  return (float)hi * 65536.f + (float)lo;
}

// one warp: lane = column; the column bytes are loaded with all 32 row loads in flight, paired
// with the right-hand neighbour's by a shuffle and packed with the row below
__device__ __forceinline__ void stage_J(LkShared& S, const uint8_t* __restrict__ Jl, int w, int h,
                                        int pitch, int rx0, int ry0) {
  const int lane = lane_id();
  const int gx = reflect101(rx0 + lane, w);
  uint32_t col[kJR];
#pragma unroll
  for (int r = 0; r < kJR; ++r) col[r] = Jl[(size_t)reflect101(ry0 + r, h) * pitch + gx];
#pragma unroll
  for (int r = 0; r < kJR; ++r) col[r] |= __shfl_down_sync(0xffffffffu, col[r], 1) << 8;
#pragma unroll
  for (int r = 0; r + 1 < kJR; ++r) S.Q[r][lane] = col[r] | (col[r + 1] << 16);
  __syncwarp();
}

// One calcOpticalFlowPyrLK call for one point, run by the whole CTA:
//   phase 1-2 (all threads)  stage the 24x24 intensity patches and 22x22 Scharr patches of ALL
//                            levels at once (they depend only on the point, not on the flow):
//                            intensities reflect-101 outside the image like the border
//                            buildOpticalFlowPyramid adds, derivatives zero outside the image
//                            like the constant border of the derivative buffer
//   phase 3 (warp L)         template + normal matrix of level L
//   phase 4 (warp 0)         coarse-to-fine Newton iterations; <= 14 pixels per lane in
//                            registers, the J search region in shared memory
// Every thread returns the same (np, status).
__device__ void lk_point(LkShared& S, const PyrDesc& pd, const uint8_t* __restrict__ I,
                         const uint8_t* __restrict__ J, float2 p0, float2 init, int use_init,
                         int top, float2& np_out, int& st_out) {
  const int tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
  const float half = (float)kHalfWin;
  const float flt_scale = 1.f / (float)(1 << 20);
  const int nlev = top + 1;
  __syncthreads();  // previous use of S is over
  // ---- phase 1: intensity patches of all levels.  All of a thread's (<= 18) global loads are
  // issued before the first store, so they share one round trip to L2 instead of queueing
  // one behind the other.
  {
    constexpr int kPerThread = (kMaxLevels * kIP * kIP + kLkThreads - 1) / kLkThreads;  // 18
    uint8_t v[kPerThread];
#pragma unroll
    for (int q = 0; q < kPerThread; ++q) {
      const int i = tid + q * kLkThreads;
      v[q] = 0;
      if (i < nlev * kIP * kIP) {
        const int L = i / (kIP * kIP), rem = i - L * (kIP * kIP);
        const int r = rem / kIP, c = rem - r * kIP;
        const float sc = 1.f / (float)(1 << L);
        const int ipx = (int)floorf(p0.x * sc - half), ipy = (int)floorf(p0.y * sc - half);
        const int w = pd.w[L], h = pd.h[L];
        if (!(ipx < -kWin || ipx >= w || ipy < -kWin || ipy >= h))
          v[q] = __ldg(I + pd.off[L] + (size_t)reflect101(ipy - 1 + r, h) * pd.pitch[L] +
                       reflect101(ipx - 1 + c, w));
      }
    }
#pragma unroll
    for (int q = 0; q < kPerThread; ++q) {
      const int i = tid + q * kLkThreads;
      if (i < nlev * kIP * kIP) (&S.I[0][0][0])[i] = v[q];
    }
  }
  __syncthreads();
  // ---- phase 2: Scharr derivatives
  for (int i = tid; i < nlev * kDP * kDP; i += kLkThreads) {
    const int L = i / (kDP * kDP), rem = i - L * (kDP * kDP);
    const int r = rem / kDP, c = rem - r * kDP;
    const float sc = 1.f / (float)(1 << L);
    const int ipx = (int)floorf(p0.x * sc - half), ipy = (int)floorf(p0.y * sc - half);
    const int gx = ipx + c, gy = ipy + r;
    short2 d = make_short2(0, 0);
    if (gx >= 0 && gx < pd.w[L] && gy >= 0 && gy < pd.h[L]) {
      const uint8_t *up = S.I[L][r], *mid = S.I[L][r + 1], *dn = S.I[L][r + 2];
      const int t0l = (up[c] + dn[c]) * 3 + mid[c] * 10;
      const int t0r = (up[c + 2] + dn[c + 2]) * 3 + mid[c + 2] * 10;
      const int t1l = dn[c] - up[c], t1m = dn[c + 1] - up[c + 1], t1r = dn[c + 2] - up[c + 2];
      d.x = (short)(t0r - t0l);
      d.y = (short)((t1r + t1l) * 3 + t1m * 10);
    }
    S.D[L][r][c] = d;
  }
  __syncthreads();
  // ---- phase 3: warp L builds the template of level L
  if (warp < nlev) {
    const int L = warp;
    const float sc = 1.f / (float)(1 << L);
    const float ppx = p0.x * sc - half, ppy = p0.y * sc - half;
    const int ipx = (int)floorf(ppx), ipy = (int)floorf(ppy);
    int flag = 0;
    if (ipx < -kWin || ipx >= pd.w[L] || ipy < -kWin || ipy >= pd.h[L]) {
      flag = 1;
    } else {
      const float a = ppx - (float)ipx, b = ppy - (float)ipy;
      int iw00, iw01, iw10, iw11;
      bilinear_weights(a, b, iw00, iw01, iw10, iw11);
      int s11 = 0, s12 = 0, s22 = 0;
#pragma unroll
      for (int j = 0; j < kPxPerLane; ++j) {
        const int kk = lane + 32 * j;
        int iv = 0, ix = 0, iy = 0;
        if (kk < kWin * kWin) {
          const int y = kk / kWin, x = kk - y * kWin;
          iv = descale(S.I[L][y + 1][x + 1] * iw00 + S.I[L][y + 1][x + 2] * iw01 +
                           S.I[L][y + 2][x + 1] * iw10 + S.I[L][y + 2][x + 2] * iw11,
                       kWBits - 5);
          const short2 d00 = S.D[L][y][x], d01 = S.D[L][y][x + 1], d10 = S.D[L][y + 1][x],
                       d11 = S.D[L][y + 1][x + 1];
          ix = descale(d00.x * iw00 + d01.x * iw01 + d10.x * iw10 + d11.x * iw11, kWBits);
          iy = descale(d00.y * iw00 + d01.y * iw01 + d10.y * iw10 + d11.y * iw11, kWBits);
          s11 += ix * ix;
          s12 += ix * iy;
          s22 += iy * iy;
        }
        S.Tw[L][kk] = (short)iv;
        S.Tx[L][kk] = (short)ix;
        S.Ty[L][kk] = (short)iy;
      }
      const float A11 = (float)warp_sum_exact(s11) * flt_scale;
      const float A12 = (float)warp_sum_exact(s12) * flt_scale;
      const float A22 = (float)warp_sum_exact(s22) * flt_scale;
      const float D = A11 * A22 - A12 * A12;
      const float min_eig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) /
                            (float)(2 * kWin * kWin);
      if ((double)min_eig < 1e-4 || D < FLT_EPSILON) flag = 2;
      if (lane == 0) {
        S.A[L][0] = A11;
        S.A[L][1] = A12;
        S.A[L][2] = A22;
      }
    }
    if (lane == 0) S.flag[L] = flag;
  }
  __syncthreads();
  // ---- phase 4: warp 0 iterates, coarse to fine
  if (warp == 0) {
    float2 np = use_init ? init : make_float2(0.f, 0.f);
    int st = 1;
    const double eps2 = 0.01 * 0.01;
    int joff[kPxPerLane];
#pragma unroll
    for (int j = 0; j < kPxPerLane; ++j) {
      const int kk = lane + 32 * j;
      const int y = kk / kWin, x = kk - y * kWin;
      joff[j] = kk < kWin * kWin ? y * kJR + x : 0;  // in 32-bit words
    }
    for (int level = top; level >= 0; --level) {
      const int w = pd.w[level], h = pd.h[level], pitch = pd.pitch[level];
      const uint8_t* __restrict__ Jl = J + pd.off[level];
      const float sc = 1.f / (float)(1 << level);
      if (level == top) {
        if (use_init) {
          np.x *= sc;
          np.y *= sc;
        } else {
          np.x = p0.x * sc;
          np.y = p0.y * sc;
        }
      } else {
        np.x *= 2.f;
        np.y *= 2.f;
      }
      if (S.flag[level] != 0) {
        if (level == 0) st = 0;
        continue;
      }
      const float A11 = S.A[level][0], A12 = S.A[level][1], A22 = S.A[level][2];
      const float D = 1.f / (A11 * A22 - A12 * A12);
      int Iw[kPxPerLane], Dx[kPxPerLane], Dy[kPxPerLane];
#pragma unroll
      for (int j = 0; j < kPxPerLane; ++j) {
        const int kk = lane + 32 * j;
        Iw[j] = S.Tw[level][kk];
        Dx[j] = S.Tx[level][kk];  // zero beyond the 441st pixel: those lanes add nothing
        Dy[j] = S.Ty[level][kk];
      }
      float npx = np.x - half, npy = np.y - half;
      int rx0 = 0, ry0 = 0;
      bool staged = false;
      float pdx = 0.f, pdy = 0.f;
      for (int it = 0; it < 30; ++it) {
        const int inx = (int)floorf(npx), iny = (int)floorf(npy);
        if (inx < -kWin || inx >= w || iny < -kWin || iny >= h) {
          if (level == 0) st = 0;
          break;
        }
        if (!staged || inx < rx0 || iny < ry0 || inx + kWin > rx0 + kJR - 1 ||
            iny + kWin > ry0 + kJR - 1) {
          rx0 = inx - kJM;
          ry0 = iny - kJM;
          __syncwarp();
          stage_J(S, Jl, w, h, pitch, rx0, ry0);
          staged = true;
        }
        const float a = npx - (float)inx, b = npy - (float)iny;
        int iw00, iw01, iw10, iw11;
        bilinear_weights(a, b, iw00, iw01, iw10, iw11);
        // sum_i pix_i * w_i = 128 * dp4a(pix, w >> 7) + dp4a(pix, w & 127); w <= 2^14
        const unsigned wh = (unsigned)(iw00 >> 7) | ((unsigned)(iw01 >> 7) << 8) |
                            ((unsigned)(iw10 >> 7) << 16) | ((unsigned)(iw11 >> 7) << 24);
        const unsigned wl = (unsigned)(iw00 & 127) | ((unsigned)(iw01 & 127) << 8) |
                            ((unsigned)(iw10 & 127) << 16) | ((unsigned)(iw11 & 127) << 24);
        const uint32_t* base = &S.Q[iny - ry0][inx - rx0];
        int sb1 = 0, sb2 = 0, sc1 = 0, sc2 = 0;  // two accumulation chains each
#pragma unroll
        for (int j = 0; j < kPxPerLane; ++j) {
          const unsigned q = base[joff[j]];
          const unsigned v = (__dp4a(q, wh, 0u) << 7) + __dp4a(q, wl, 1u << (kWBits - 5 - 1));
          const int diff = (int)(v >> (kWBits - 5)) - Iw[j];
          if (j & 1) {
            sc1 += diff * Dx[j];
            sc2 += diff * Dy[j];
          } else {
            sb1 += diff * Dx[j];
            sb2 += diff * Dy[j];
          }
        }
        const float b1 = warp_sum_to_float(sb1 + sc1) * flt_scale;
        const float b2 = warp_sum_to_float(sb2 + sc2) * flt_scale;
        const float dx = (A12 * b2 - A22 * b1) * D;
        const float dy = (A12 * b1 - A11 * b2) * D;
        npx += dx;
        npy += dy;
        np.x = npx + half;
        np.y = npy + half;
        // OpenCV evaluates both stopping rules in double.  Far from the thresholds a float
        // estimate decides the same way (its error is ~1e-7 relative, the margins below are a
        // factor 2 / 10 %), so the double arithmetic only runs in the rare close calls.
        const float d2 = dx * dx + dy * dy;
        if (d2 <= 2e-4f) {
          if (d2 < 5e-5f || (double)dx * (double)dx + (double)dy * (double)dy <= eps2) break;
        }
        if (it > 0) {
          const float sx = fabsf(dx + pdx), sy = fabsf(dy + pdy);
          if (sx < 0.011f && sy < 0.011f &&
              ((sx < 0.009f && sy < 0.009f) || ((double)sx < 0.01 && (double)sy < 0.01))) {
            np.x -= dx * 0.5f;
            np.y -= dy * 0.5f;
            break;
          }
        }
        pdx = dx;
        pdy = dy;
      }
      if (st && level == 0) {
        // the reference passes an `err` vector, so OpenCV re-validates the final window
        const int inx = (int)floorf(np.x - half), iny = (int)floorf(np.y - half);
        if (inx < -kWin || inx >= w || iny < -kWin || iny >= h) st = 0;
      }
    }
    if (lane == 0) {
      S.np = np;
      S.st = st;
    }
  }
  __syncthreads();
  np_out = S.np;
  st_out = S.st;
}

// mode 0: forward only (next = LK(I->J, prev[, init = next]))
// mode 1: temporal pair (feature_tracker.cpp:410,417-418): forward maxLevel `top`, then
//         backward J->I from the forward result with initial flow = prev, maxLevel 1
// mode 2: stereo pair (:490,495): forward, then backward J->I, both maxLevel `top`, no init
__global__ void __launch_bounds__(kLkThreads)
k_lk(const __grid_constant__ PyrDesc pd, const uint8_t* __restrict__ I, const uint8_t* __restrict__ J,
     const float2* __restrict__ prev_pts, float2* __restrict__ next_pts,
     uint8_t* __restrict__ status, float2* __restrict__ rev_pts, uint8_t* __restrict__ rev_status,
     const int* __restrict__ n_ptr, int top, int use_init, int mode) {
  PDL_PROLOGUE();
  __shared__ LkShared S;
  const int k = blockIdx.x;
  if (k >= *n_ptr) return;
  const float2 p0 = prev_pts[k];
  float2 init = make_float2(0.f, 0.f);
  if (use_init) init = next_pts[k];
  float2 np;
  int st;
  lk_point(S, pd, I, J, p0, init, use_init, top, np, st);
  if (threadIdx.x == 0) {
    next_pts[k] = np;
    status[k] = (uint8_t)st;
  }
  if (mode == 0) return;
  float2 rp;
  int rst;
  if (mode == 1) {
    const int top_b = top < 1 ? top : 1;
    lk_point(S, pd, J, I, np, p0, 1, top_b, rp, rst);
  } else {
    lk_point(S, pd, J, I, np, init, 0, top, rp, rst);
  }
  if (threadIdx.x == 0) {
    rev_pts[k] = rp;
    rev_status[k] = (uint8_t)rst;
  }
}

void launch_lk(const PyrDesc& pd, const uint8_t* I, const uint8_t* J, const float2* prev_pts,
               float2* next_pts, uint8_t* status, float2* rev_pts, uint8_t* rev_status,
               const int* n_ptr, int n_max, int max_level, int use_initial_flow, int mode,
               cudaStream_t s, int64_t* launches) {
  if (n_max <= 0) return;
  int top = pd.levels - 1;
  if (top > max_level) top = max_level;
  launch_pdl(k_lk, dim3(n_max), dim3(kLkThreads), 0, s, pd, I, J, prev_pts, next_pts, status, rev_pts,
             rev_status, n_ptr, top, use_initial_flow, mode);
  ++*launches;
}

// Up to three windows run concurrently on three streams, and the LK kernels sit on every SM for
// most of a step.  Kernels that are to share an SM have to ask for the same L1/shared-memory
// carve-out: with the driver's per-kernel defaults (small for k_lk, large for the many small
// CTAs of k_sae_update_ts) the SAE kernel of the next window gets one CTA per SM next to two LK
// kernels and takes twice as long (measured at 346x260: 23 us instead of 12 us; MaxShared on
// both fixes that too but starves the gather-heavy kernels on those SMs of L1: -5..9 %
// throughput at 640x480).  Both ask for half of the array: room for 2-3 LK CTAs plus 6 SAE
// CTAs per SM, 114 KB of L1 left.
void prefer_shared_lk() {
  cudaFuncSetAttribute(k_lk, cudaFuncAttributePreferredSharedMemoryCarveout, 50);
}

}  // namespace esvio
