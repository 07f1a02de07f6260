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
constexpr int kLkWarps = kLkThreads / 32;
constexpr int kPxPerThread = (kWin * kWin + kLkThreads - 1) / kLkThreads;  // 4
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
  uint8_t I[kIP][kIP];
  short2 D[kDP][kDP];
  uint8_t J[kJR][kJR];
  int red[2][kLkWarps][6];  // double-buffered per-warp (hi, lo) partials of up to 3 sums
};

// exact CTA-wide sums of NV int32 values per thread; one __syncthreads per call
template <int NV>
__device__ __forceinline__ void cta_sum_exact(LkShared& S, int buf, const int (&v)[NV],
                                              long long (&out)[NV]) {
  const int lane = lane_id(), warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    const unsigned lo = __reduce_add_sync(0xffffffffu, (unsigned)v[q] & 0xffffu);
    const int hi = __reduce_add_sync(0xffffffffu, v[q] >> 16);
    if (lane == 0) {
      S.red[buf][warp][2 * q] = hi;
      S.red[buf][warp][2 * q + 1] = (int)lo;
    }
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    long long acc = 0;
#pragma unroll
    for (int w = 0; w < kLkWarps; ++w)
      acc += ((long long)S.red[buf][w][2 * q] << 16) + (long long)S.red[buf][w][2 * q + 1];
    out[q] = acc;
  }
}

// One calcOpticalFlowPyrLK call for one point; every thread of the CTA runs it with the same
// scalars and returns the same (np, status).
__device__ void lk_point(LkShared& S, const PyrDesc& pd, const uint8_t* __restrict__ I,
                         const uint8_t* __restrict__ J, float2 p0, float2 init, int use_init,
                         int top, float2& np_out, int& st_out) {
  const int tid = threadIdx.x;
  float2 np = use_init ? init : make_float2(0.f, 0.f);
  int st = 1;
  int buf = 0;
  const float half = (float)kHalfWin;
  const float flt_scale = 1.f / (float)(1 << 20);
  const double eps2 = 0.01 * 0.01;

  int wx[kPxPerThread], wy[kPxPerThread];
#pragma unroll
  for (int j = 0; j < kPxPerThread; ++j) {
    const int kk = tid + kLkThreads * j;
    wy[j] = kk / kWin;
    wx[j] = kk - wy[j] * kWin;
  }

  for (int level = top; level >= 0; --level) {
    const int w = pd.w[level], h = pd.h[level], pitch = pd.pitch[level];
    const uint8_t* __restrict__ Il = I + pd.off[level];
    const uint8_t* __restrict__ Jl = J + pd.off[level];
    const float sc = 1.f / (float)(1 << level);
    float ppx = p0.x * sc, ppy = p0.y * sc;
    if (level == top) {
      if (use_init) {
        np.x *= sc;
        np.y *= sc;
      } else {
        np.x = ppx;
        np.y = ppy;
      }
    } else {
      np.x *= 2.f;
      np.y *= 2.f;
    }
    ppx -= half;
    ppy -= half;
    const int ipx = (int)floorf(ppx), ipy = (int)floorf(ppy);
    if (ipx < -kWin || ipx >= w || ipy < -kWin || ipy >= h) {
      if (level == 0) st = 0;
      continue;
    }
    // ---- stage the template neighbourhood (intensities: reflect-101 outside the image like
    //      the border buildOpticalFlowPyramid adds; derivatives: zero outside the image like
    //      the constant border of the derivative buffer) and the J search region
    float npx = np.x - half, npy = np.y - half;
    int rx0 = (int)floorf(npx) - kJM, ry0 = (int)floorf(npy) - kJM;
    __syncthreads();
    for (int i = tid; i < kIP * kIP; i += kLkThreads) {
      const int r = i / kIP, c = i - r * kIP;
      S.I[r][c] = Il[(size_t)reflect101(ipy - 1 + r, h) * pitch + reflect101(ipx - 1 + c, w)];
    }
    if (rx0 >= -kWin - kJM && rx0 < w && ry0 >= -kWin - kJM && ry0 < h) {
      for (int i = tid; i < kJR * kJR; i += kLkThreads) {
        const int r = i / kJR, c = i - r * kJR;
        S.J[r][c] = Jl[(size_t)reflect101(ry0 + r, h) * pitch + reflect101(rx0 + c, w)];
      }
    }
    __syncthreads();
    for (int i = tid; i < kDP * kDP; i += kLkThreads) {
      const int r = i / kDP, c = i - r * kDP;
      const int gx = ipx + c, gy = ipy + r;
      short2 d = make_short2(0, 0);
      if (gx >= 0 && gx < w && gy >= 0 && gy < h) {
        const uint8_t *up = S.I[r], *mid = S.I[r + 1], *dn = S.I[r + 2];
        const int t0l = (up[c] + dn[c]) * 3 + mid[c] * 10;
        const int t0r = (up[c + 2] + dn[c + 2]) * 3 + mid[c + 2] * 10;
        const int t1l = dn[c] - up[c], t1m = dn[c + 1] - up[c + 1], t1r = dn[c + 2] - up[c + 2];
        d.x = (short)(t0r - t0l);
        d.y = (short)((t1r + t1l) * 3 + t1m * 10);
      }
      S.D[r][c] = d;
    }
    __syncthreads();

    float a = ppx - (float)ipx, b = ppy - (float)ipy;
    int iw00, iw01, iw10, iw11;
    bilinear_weights(a, b, iw00, iw01, iw10, iw11);
    int Iw[kPxPerThread], Dx[kPxPerThread], Dy[kPxPerThread];
    int sA[3] = {0, 0, 0};
#pragma unroll
    for (int j = 0; j < kPxPerThread; ++j) {
      Iw[j] = Dx[j] = Dy[j] = 0;
      if (tid + kLkThreads * j < kWin * kWin) {
        const int y = wy[j], x = wx[j];
        const int iv = descale(S.I[y + 1][x + 1] * iw00 + S.I[y + 1][x + 2] * iw01 +
                                   S.I[y + 2][x + 1] * iw10 + S.I[y + 2][x + 2] * iw11,
                               kWBits - 5);
        const short2 d00 = S.D[y][x], d01 = S.D[y][x + 1], d10 = S.D[y + 1][x],
                     d11 = S.D[y + 1][x + 1];
        const int ix = descale(d00.x * iw00 + d01.x * iw01 + d10.x * iw10 + d11.x * iw11, kWBits);
        const int iy = descale(d00.y * iw00 + d01.y * iw01 + d10.y * iw10 + d11.y * iw11, kWBits);
        Iw[j] = iv;
        Dx[j] = ix;
        Dy[j] = iy;
        sA[0] += ix * ix;
        sA[1] += ix * iy;
        sA[2] += iy * iy;
      }
    }
    long long tA[3];
    cta_sum_exact<3>(S, buf, sA, tA);
    buf ^= 1;
    const float A11 = (float)tA[0] * flt_scale;
    const float A12 = (float)tA[1] * flt_scale;
    const float A22 = (float)tA[2] * flt_scale;
    float D = A11 * A22 - A12 * A12;
    const float min_eig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) /
                          (float)(2 * kWin * kWin);
    if ((double)min_eig < 1e-4 || D < FLT_EPSILON) {
      if (level == 0) st = 0;
      continue;
    }
    D = 1.f / D;
    float pdx = 0.f, pdy = 0.f;
    for (int it = 0; it < 30; ++it) {
      const int inx = (int)floorf(npx), iny = (int)floorf(npy);
      if (inx < -kWin || inx >= w || iny < -kWin || iny >= h) {
        if (level == 0) st = 0;
        break;
      }
      if (inx < rx0 || iny < ry0 || inx + kWin > rx0 + kJR - 1 || iny + kWin > ry0 + kJR - 1) {
        // the window left the staged region: re-stage around the current position (all
        // threads are past the barrier that followed their last read of S.J)
        rx0 = inx - kJM;
        ry0 = iny - kJM;
        for (int i = tid; i < kJR * kJR; i += kLkThreads) {
          const int r = i / kJR, c = i - r * kJR;
          S.J[r][c] = Jl[(size_t)reflect101(ry0 + r, h) * pitch + reflect101(rx0 + c, w)];
        }
        __syncthreads();
      }
      a = npx - (float)inx;
      b = npy - (float)iny;
      bilinear_weights(a, b, iw00, iw01, iw10, iw11);
      int sb[2] = {0, 0};
      const int ox = inx - rx0, oy = iny - ry0;
#pragma unroll
      for (int j = 0; j < kPxPerThread; ++j) {
        if (tid + kLkThreads * j < kWin * kWin) {
          const uint8_t* jp = &S.J[oy + wy[j]][ox + wx[j]];
          const int v = jp[0] * iw00 + jp[1] * iw01 + jp[kJR] * iw10 + jp[kJR + 1] * iw11;
          const int diff = descale(v, kWBits - 5) - Iw[j];
          sb[0] += diff * Dx[j];
          sb[1] += diff * Dy[j];
        }
      }
      long long tb[2];
      cta_sum_exact<2>(S, buf, sb, tb);
      buf ^= 1;
      const float b1 = (float)tb[0] * flt_scale;
      const float b2 = (float)tb[1] * flt_scale;
      const float dx = (A12 * b2 - A22 * b1) * D;
      const float dy = (A12 * b1 - A11 * b2) * D;
      npx += dx;
      npy += dy;
      np.x = npx + half;
      np.y = npy + half;
      if ((double)dx * (double)dx + (double)dy * (double)dy <= eps2) break;
      if (it > 0 && (double)fabsf(dx + pdx) < 0.01 && (double)fabsf(dy + pdy) < 0.01) {
        np.x -= dx * 0.5f;
        np.y -= dy * 0.5f;
        break;
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
  np_out = np;
  st_out = st;
}

// mode 0: forward only (next = LK(I->J, prev[, init = next]))
// mode 1: temporal pair (feature_tracker.cpp:410,417-418): forward maxLevel `top`, then
//         backward J->I from the forward result with initial flow = prev, maxLevel 1
// mode 2: stereo pair (:490,495): forward, then backward J->I, both maxLevel `top`, no init
__global__ void __launch_bounds__(kLkThreads)
k_lk(PyrDesc pd, const uint8_t* __restrict__ I, const uint8_t* __restrict__ J,
     const float2* __restrict__ prev_pts, float2* __restrict__ next_pts,
     uint8_t* __restrict__ status, float2* __restrict__ rev_pts, uint8_t* __restrict__ rev_status,
     const int* __restrict__ n_ptr, int top, int use_init, int mode) {
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
  k_lk<<<n_max, kLkThreads, 0, s>>>(pd, I, J, prev_pts, next_pts, status, rev_pts, rev_status,
                                    n_ptr, top, use_initial_flow, mode);
  ++*launches;
}

}  // namespace esvio
