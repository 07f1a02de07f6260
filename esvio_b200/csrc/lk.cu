// lk.cu -- pyramidal Lucas-Kanade, arithmetic-compatible with cv::calcOpticalFlowPyrLK as the
// reference calls it (feature_tracker/src/feature_tracker.cpp:410 temporal forward,
// :417-418 temporal backward with OPTFLOW_USE_INITIAL_FLOW and maxLevel 1, :490/:495 stereo
// forward/backward): 21x21 window, 30 iterations / eps 0.01, minEigThreshold 1e-4, 14-bit
// bilinear weights, 5-bit up-scaled intensities, int16 Scharr derivatives (OpenCV
// modules/video/src/lkpyramid.cpp -- not part of the reference tree; arithmetic spec in
// SURVEY.md section 8a).
//
// One warp tracks one point through all pyramid levels inside a single launch.  The
// template (Iw, Ixw, Iyw of the 441 window pixels) lives in registers, 14 pixels per
// lane; the 2x2 normal-equation sums A11,A12,A22 and the per-iteration b1,b2 are
// accumulated as exact integers per lane and combined with warp reductions
// (redux.sync), then rounded to float once -- OpenCV accumulates the same integers in
// float, so results agree to float rounding (~1e-4 px).
#include <float.h>

#include "common.cuh"

namespace esvio {

constexpr int kWBits = 14;
constexpr int kPxPerLane = 14;  // ceil(441 / 32)
constexpr int kIP = 24;         // staged intensity patch (window + bilinear tap + Scharr ring)
constexpr int kDP = 22;         // derivative patch (window + bilinear tap)

__device__ __forceinline__ int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

// exact sum over the warp of one int32 per lane, as int64
__device__ __forceinline__ long long warp_sum_exact(int v) {
  const unsigned lo = __reduce_add_sync(0xffffffffu, (unsigned)v & 0xffffu);
  const int hi = __reduce_add_sync(0xffffffffu, v >> 16);
  return ((long long)hi << 16) + (long long)lo;
}

__device__ __forceinline__ void bilinear_weights(float a, float b, int& w00, int& w01, int& w10,
                                                 int& w11) {
  w00 = cv_round((1.f - a) * (1.f - b) * (float)(1 << kWBits));
  w01 = cv_round(a * (1.f - b) * (float)(1 << kWBits));
  w10 = cv_round((1.f - a) * b * (float)(1 << kWBits));
  w11 = (1 << kWBits) - w00 - w01 - w10;
}

__global__ void __launch_bounds__(32)
k_lk(PyrDesc pd, const uint8_t* __restrict__ I, const uint8_t* __restrict__ J,
     const float2* __restrict__ prev_pts, float2* __restrict__ next_pts,
     uint8_t* __restrict__ status, const int* __restrict__ n_ptr, int top, int use_init) {
  __shared__ uint8_t s_I[kIP][kIP];
  __shared__ short2 s_D[kDP][kDP];
  const int k = blockIdx.x;
  if (k >= *n_ptr) return;
  const int lane = lane_id();
  const float2 p0 = prev_pts[k];
  float2 np = use_init ? next_pts[k] : make_float2(0.f, 0.f);
  int st = 1;
  const float half = (float)kHalfWin;
  const float flt_scale = 1.f / (float)(1 << 20);
  const double eps2 = 0.01 * 0.01;

  int wx[kPxPerLane], wy[kPxPerLane];
#pragma unroll
  for (int j = 0; j < kPxPerLane; ++j) {
    const int kk = lane + 32 * j;
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
    // ---- stage the template neighbourhood: intensities (reflect-101 outside the image,
    //      like the border buildOpticalFlowPyramid adds) and Scharr derivatives (zero
    //      outside the image, like the constant border of the derivative buffer)
    __syncwarp();
    for (int i = lane; i < kIP * kIP; i += 32) {
      const int r = i / kIP, c = i - r * kIP;
      s_I[r][c] = Il[(size_t)reflect101(ipy - 1 + r, h) * pitch + reflect101(ipx - 1 + c, w)];
    }
    __syncwarp();
    for (int i = lane; i < kDP * kDP; i += 32) {
      const int r = i / kDP, c = i - r * kDP;
      const int gx = ipx + c, gy = ipy + r;
      short2 d = make_short2(0, 0);
      if (gx >= 0 && gx < w && gy >= 0 && gy < h) {
        const uint8_t *up = s_I[r], *mid = s_I[r + 1], *dn = s_I[r + 2];
        const int t0l = (up[c] + dn[c]) * 3 + mid[c] * 10;
        const int t0r = (up[c + 2] + dn[c + 2]) * 3 + mid[c + 2] * 10;
        const int t1l = dn[c] - up[c], t1m = dn[c + 1] - up[c + 1], t1r = dn[c + 2] - up[c + 2];
        d.x = (short)(t0r - t0l);
        d.y = (short)((t1r + t1l) * 3 + t1m * 10);
      }
      s_D[r][c] = d;
    }
    __syncwarp();

    float a = ppx - (float)ipx, b = ppy - (float)ipy;
    int iw00, iw01, iw10, iw11;
    bilinear_weights(a, b, iw00, iw01, iw10, iw11);
    int Iw[kPxPerLane], Dx[kPxPerLane], Dy[kPxPerLane];
    int s11 = 0, s12 = 0, s22 = 0;
#pragma unroll
    for (int j = 0; j < kPxPerLane; ++j) {
      Iw[j] = Dx[j] = Dy[j] = 0;
      if (lane + 32 * j < kWin * kWin) {
        const int y = wy[j], x = wx[j];
        const int iv = descale(s_I[y + 1][x + 1] * iw00 + s_I[y + 1][x + 2] * iw01 +
                                   s_I[y + 2][x + 1] * iw10 + s_I[y + 2][x + 2] * iw11,
                               kWBits - 5);
        const short2 d00 = s_D[y][x], d01 = s_D[y][x + 1], d10 = s_D[y + 1][x],
                     d11 = s_D[y + 1][x + 1];
        const int ix = descale(d00.x * iw00 + d01.x * iw01 + d10.x * iw10 + d11.x * iw11, kWBits);
        const int iy = descale(d00.y * iw00 + d01.y * iw01 + d10.y * iw10 + d11.y * iw11, kWBits);
        Iw[j] = iv;
        Dx[j] = ix;
        Dy[j] = iy;
        s11 += ix * ix;
        s12 += ix * iy;
        s22 += iy * iy;
      }
    }
    const float A11 = (float)warp_sum_exact(s11) * flt_scale;
    const float A12 = (float)warp_sum_exact(s12) * flt_scale;
    const float A22 = (float)warp_sum_exact(s22) * flt_scale;
    float D = A11 * A22 - A12 * A12;
    const float min_eig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) /
                          (float)(2 * kWin * kWin);
    if ((double)min_eig < 1e-4 || D < FLT_EPSILON) {
      if (level == 0) st = 0;
      continue;
    }
    D = 1.f / D;
    float npx = np.x - half, npy = np.y - half;
    float pdx = 0.f, pdy = 0.f;
    for (int it = 0; it < 30; ++it) {
      const int inx = (int)floorf(npx), iny = (int)floorf(npy);
      if (inx < -kWin || inx >= w || iny < -kWin || iny >= h) {
        if (level == 0) st = 0;
        break;
      }
      a = npx - (float)inx;
      b = npy - (float)iny;
      bilinear_weights(a, b, iw00, iw01, iw10, iw11);
      int sb1 = 0, sb2 = 0;
      const bool inside = inx >= 0 && iny >= 0 && inx + kWin <= w - 1 && iny + kWin <= h - 1;
      if (inside) {
        const uint8_t* __restrict__ base = Jl + (size_t)iny * pitch + inx;
#pragma unroll
        for (int j = 0; j < kPxPerLane; ++j) {
          if (lane + 32 * j < kWin * kWin) {
            const uint8_t* jp = base + wy[j] * pitch + wx[j];
            const int v = jp[0] * iw00 + jp[1] * iw01 + jp[pitch] * iw10 + jp[pitch + 1] * iw11;
            const int diff = descale(v, kWBits - 5) - Iw[j];
            sb1 += diff * Dx[j];
            sb2 += diff * Dy[j];
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < kPxPerLane; ++j) {
          if (lane + 32 * j < kWin * kWin) {
            const int x0 = reflect101(inx + wx[j], w), x1 = reflect101(inx + wx[j] + 1, w);
            const uint8_t* r0 = Jl + (size_t)reflect101(iny + wy[j], h) * pitch;
            const uint8_t* r1 = Jl + (size_t)reflect101(iny + wy[j] + 1, h) * pitch;
            const int v = r0[x0] * iw00 + r0[x1] * iw01 + r1[x0] * iw10 + r1[x1] * iw11;
            const int diff = descale(v, kWBits - 5) - Iw[j];
            sb1 += diff * Dx[j];
            sb2 += diff * Dy[j];
          }
        }
      }
      const float b1 = (float)warp_sum_exact(sb1) * flt_scale;
      const float b2 = (float)warp_sum_exact(sb2) * flt_scale;
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
  if (lane == 0) {
    next_pts[k] = np;
    status[k] = (uint8_t)st;
  }
}

void launch_lk(const PyrDesc& pd, const uint8_t* I, const uint8_t* J, const float2* prev_pts,
               float2* next_pts, uint8_t* status, const int* n_ptr, int n_max, int max_level,
               int use_initial_flow, cudaStream_t s, int64_t* launches) {
  if (n_max <= 0) return;
  int top = pd.levels - 1;
  if (top > max_level) top = max_level;
  k_lk<<<n_max, 32, 0, s>>>(pd, I, J, prev_pts, next_pts, status, n_ptr, top, use_initial_flow);
  ++*launches;
}

}  // namespace esvio
