// lk.cu -- pyramidal Lucas-Kanade, arithmetic-compatible with cv::calcOpticalFlowPyrLK as the
// reference calls it (feature_tracker/src/feature_tracker.cpp:410 temporal forward,
// :417-418 temporal backward with OPTFLOW_USE_INITIAL_FLOW and maxLevel 1, :490/:495 stereo
// forward/backward): 21x21 window, 30 iterations / eps 0.01, minEigThreshold 1e-4, 14-bit
// bilinear weights, 5-bit up-scaled intensities, int16 Scharr derivatives (OpenCV
// modules/video/src/lkpyramid.cpp -- not part of the reference tree; arithmetic spec in
// SURVEY.md section 8a).
//
// One CTA of 256 threads tracks one point through all pyramid levels, forward and (when
// asked) straight on into the backward check, inside a single launch.  The kernel is
// latency-bound (a point is one dependent chain of <= 4 levels x 30 Newton iterations, and a
// launch is as slow as its slowest point), so everything is organised around the length of
// that chain:
//   * set-up (all 8 warps): the 24x24 intensity patches of ALL levels and the search region of
//     the top level are fetched in ONE batch of independent loads (addresses are branch-free:
//     reflect101_nb), then the Scharr patches and the templates of all levels are built, two
//     warps per level;
//   * Newton iterations (warps 0-3, the "N group"): the 441 window pixels are split over 128
//     threads (<= 4 each, template in registers, search region in shared memory as packed 2x2
//     neighbourhoods + dp4a); the sums are exact integers: redux.sync inside a warp, one
//     16-byte shared-memory slot per warp, ONE named barrier per iteration, after which every
//     thread adds the four partials and runs the same 2x2 solve and stopping rules;
//   * staging (warps 4-7, the "S group"): while the N group iterates on level L, the S group
//     fetches the search region of level L-1 around the predicted position 2*p into the other
//     buffer, so the trip to L2 hides behind the iterations; a level only re-stages when the
//     prediction misses by more than the 5-pixel margin.
// OpenCV accumulates the same integers in float, so results agree to float rounding
// (~1e-4 px); against the round-1 kernel (one warp iterating) they are bit-identical.
#include <float.h>
#include <stdlib.h>

#include "common.cuh"

namespace esvio {

#ifdef ESVIO_LK_CLOCKS  // scratch builds only (make EXTRA=-DESVIO_LK_CLOCKS): per-phase clock64 of every CTA
__device__ long long g_lk_clk[1024][2][24];
#define LK_CLK(i) do { if (threadIdx.x == 0 && blockIdx.x < 1024) g_lk_clk[blockIdx.x][clk_call][i] = clock64(); } while (0)
#define LK_VAL(i, v) do { if (threadIdx.x == 0 && blockIdx.x < 1024) g_lk_clk[blockIdx.x][clk_call][i] = (v); } while (0)
extern "C" __attribute__((visibility("default"))) int esvio_dbg_lk_clocks(long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_lk_clk, sizeof(g_lk_clk));
}
#else
#define LK_CLK(i)
#define LK_VAL(i, v)
#endif

#ifndef ESVIO_LK_MINB
#define ESVIO_LK_MINB 3
#endif
constexpr int kWBits = 14;
constexpr int kLkThreads = 256;
constexpr int kLkWarps = kLkThreads / 32;
constexpr int kMaxNWarps = 4;                      // Newton group: warps 0 .. NW-1, NW = 1, 2 or 4
constexpr int kSFirst = 4, kSWarps = 4;            // staging group: warps 4-7
constexpr int kTplLen = 448;                       // template slots per level (441 used)
constexpr int kIP = 24;  // staged intensity patch (window + bilinear tap + Scharr ring)
constexpr int kDP = 22;  // derivative patch (window + bilinear tap)
constexpr int kJM = 5;   // margin of the staged search region
constexpr int kJR = kWin + 1 + 2 * kJM;  // 32
// Row stride of the search region in 32-bit words: 21 (mod 32), so that the word of window
// pixel k = 21 * y + x sits in bank (k + const) % 32 and 32 consecutive pixels never collide.
constexpr int kQS = 53;

__device__ __forceinline__ int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

// d = c + a.s16[0] * b.u8[0 | 2] + a.s16[1] * b.u8[1 | 3]
__device__ __forceinline__ int dp2a_lo_su(unsigned a, unsigned b, int c) {
  int d;
  asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ int dp2a_hi_su(unsigned a, unsigned b, int c) {
  int d;
  asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// cvRound of a float in [0, 2^22): adding 1.5 * 2^23 leaves the rounded integer (half to even,
// the FADD's own rounding) in the low mantissa bits -- two 4-cycle ALU instructions instead of
// an F2I on the conversion pipe, which sits on the Newton iteration's dependent chain
__device__ __forceinline__ int cv_round_small(float v) {
  return __float_as_int(v + 12582912.f) - 0x4B400000;
}

__device__ __forceinline__ void bilinear_weights(float a, float b, int& w00, int& w01, int& w10,
                                                 int& w11) {
  w00 = cv_round_small((1.f - a) * (1.f - b) * (float)(1 << kWBits));
  w01 = cv_round_small(a * (1.f - b) * (float)(1 << kWBits));
  w10 = cv_round_small((1.f - a) * b * (float)(1 << kWBits));
  w11 = (1 << kWBits) - w00 - w01 - w10;
}

struct LkShared {
  // search regions (current level / prefetched next level) as packed 2x2 neighbourhoods:
  // Q[r][c] = J(r,c) | J(r,c+1) << 8 | J(r+1,c) << 16 | J(r+1,c+1) << 24
  uint32_t Q[2][kJR][kQS];
  // The template neighbourhoods of all levels arrive in one batch of loads; only the top
  // level's template is built before the iterations start, the others one level ahead by the
  // staging group while the Newton group iterates.  So what is live at any time is: the
  // patches, ONE level's Scharr derivatives, and the templates of two levels (the one being
  // iterated and the one being built) -- 23 KB per CTA, like the round-1 layout that built
  // all four templates up front.
  uint8_t I[kMaxLevels][kIP][kIP];
  short2 D[kDP][kDP];
  short Tw[2][kTplLen];                 // templates of level L in slot L & 1: Iw, Ixw, Iyw per window pixel
  short Tx[2][kTplLen];
  short Ty[2][kTplLen];
  long long Apart[2][kLkWarps][3];      // per-warp sums of Ixw^2, Ixw*Iyw, Iyw^2 (unused warps: 0)
  int flag_win[kMaxLevels];             // 1: template window outside the image
  int4 part[2][kMaxNWarps];             // per-iteration partial sums (b1, b2 as low 16 bits / rest) of the N-group warps
  float2 np[2];                         // result of a level, slot = Newton levels run so far & 1
  int st[2];
};

// exact sum over the warp of one int32 per lane, as int64
__device__ __forceinline__ long long warp_sum_exact(int v) {
  const unsigned lo = __reduce_add_sync(0xffffffffu, (unsigned)v & 0xffffu);
  const int hi = __reduce_add_sync(0xffffffffu, v >> 16);
  return ((long long)hi << 16) + (long long)lo;
}

// shared memory by 32-bit address (computed once per level: through a generic pointer the
// compiler re-derives the shared window base inside the Newton loop, an S2R on the chain)
__device__ __forceinline__ uint32_t keep_in_register(uint32_t v) {
  uint32_t r;
  asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
  return r;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ int4 lds_v4(uint32_t addr) {
  int4 v;
  asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_v4(uint32_t addr, int a, int b, int c, int d) {
  asm volatile("st.shared.v4.s32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// barrier of the NW warps of the N group (barrier 0 is __syncthreads of the whole CTA)
template <int NW>
__device__ __forceinline__ void bar_newton() {
  if (NW == 1) __syncwarp();
  else asm volatile("bar.sync 1, %0;" ::"n"(32 * NW) : "memory");
}

// Stage the kJR x kJR search region with origin (rx0, ry0) of one level of J into Q.  Warp
// `wi` of NW cooperating warps owns kJR / NW rows: lane = column; every lane issues its row
// loads back to back (one trip to L2 for the lot), pairs each byte with the right-hand
// neighbour's by a shuffle and packs it with the row below.
template <int NW>
__device__ __forceinline__ void stage_J_load(uint32_t (&col)[kJR / NW + 1], const uint8_t* __restrict__ Jl,
                                             int w, int h, int pitch, int rx0, int ry0, int wi) {
  constexpr int kRows = kJR / NW;
  const int gx = reflect101_nb(rx0 + lane_id(), w);
  const int r0 = wi * kRows;
#pragma unroll
  for (int r = 0; r <= kRows; ++r)
    col[r] = __ldg(Jl + (size_t)reflect101_nb(ry0 + r0 + r, h) * pitch + gx);
}
template <int NW>
__device__ __forceinline__ void stage_J_store(uint32_t (*Q)[kQS], uint32_t (&col)[kJR / NW + 1], int wi) {
  constexpr int kRows = kJR / NW;
  const int lane = lane_id();
  const int r0 = wi * kRows;
#pragma unroll
  for (int r = 0; r <= kRows; ++r) col[r] |= __shfl_down_sync(0xffffffffu, col[r], 1) << 8;
#pragma unroll
  for (int r = 0; r < kRows; ++r) Q[r0 + r][lane] = col[r] | (col[r + 1] << 16);
}
template <int NW>
__device__ __forceinline__ void stage_J(uint32_t (*Q)[kQS], const uint8_t* __restrict__ Jl, int w,
                                        int h, int pitch, int rx0, int ry0, int wi) {
  uint32_t col[kJR / NW + 1];
  stage_J_load<NW>(col, Jl, w, h, pitch, rx0, ry0, wi);
  stage_J_store<NW>(Q, col, wi);
}

// Scharr derivatives of level L's patch into S.D, zero outside the image like the constant
// border of OpenCV's derivative buffer; NWB warps, this one is number wi.
template <int NWB>
__device__ __forceinline__ void lk_scharr_level(LkShared& S, const PyrDesc& pd, float2 p0, int L, int wi) {
  const float sc = 1.f / (float)(1 << L);
  const int ipx = (int)floorf(p0.x * sc - (float)kHalfWin), ipy = (int)floorf(p0.y * sc - (float)kHalfWin);
  const int w = pd.w[L], h = pd.h[L];
  for (int i = wi * 32 + lane_id(); i < kDP * kDP; i += 32 * NWB) {
    const int r = i / kDP, c = i - r * kDP;
    const int gx = ipx + c, gy = ipy + r;
    short2 d = make_short2(0, 0);
    if (gx >= 0 && gx < w && gy >= 0 && gy < h) {
      const uint8_t *up = S.I[L][r], *mid = S.I[L][r + 1], *dn = S.I[L][r + 2];
      const int t0l = (up[c] + dn[c]) * 3 + mid[c] * 10;
      const int t0r = (up[c + 2] + dn[c + 2]) * 3 + mid[c + 2] * 10;
      const int t1l = dn[c] - up[c], t1m = dn[c + 1] - up[c + 1], t1r = dn[c + 2] - up[c + 2];
      d.x = (short)(t0r - t0l);
      d.y = (short)((t1r + t1l) * 3 + t1m * 10);
    }
    S.D[r][c] = d;
  }
}

// Template of level L (Iw, Ixw, Iyw at the 21x21 window pixels, bilinear at the sub-pixel
// position of p0) into slot L & 1, the normal matrix's exact integer sums per warp; NWB warps.
template <int NWB>
__device__ __forceinline__ void lk_template_level(LkShared& S, const PyrDesc& pd, float2 p0, int L, int wi) {
  constexpr int kPx = (kWin * kWin + 32 * NWB - 1) / (32 * NWB);
  const int lane = lane_id(), tb = L & 1;
  const float sc = 1.f / (float)(1 << L);
  const float ppx = p0.x * sc - (float)kHalfWin, ppy = p0.y * sc - (float)kHalfWin;
  const int ipx = (int)floorf(ppx), ipy = (int)floorf(ppy);
  const bool outside = ipx < -kWin || ipx >= pd.w[L] || ipy < -kWin || ipy >= pd.h[L];
  if (wi == 0 && lane == 0) S.flag_win[L] = outside ? 1 : 0;
  if (NWB < kLkWarps && wi == 0 && lane < 3 * (kLkWarps - NWB)) (&S.Apart[tb][NWB][0])[lane] = 0;
  if (outside) return;
  const float a = ppx - (float)ipx, b = ppy - (float)ipy;
  int iw00, iw01, iw10, iw11;
  bilinear_weights(a, b, iw00, iw01, iw10, iw11);
  int s11 = 0, s12 = 0, s22 = 0;
#pragma unroll
  for (int j = 0; j < kPx; ++j) {
    const int kk = wi * 32 + lane + 32 * NWB * j;
    if (kk < kWin * kWin) {
      const int y = kk / kWin, x = kk - y * kWin;
      const int iv = descale(S.I[L][y + 1][x + 1] * iw00 + S.I[L][y + 1][x + 2] * iw01 +
                                 S.I[L][y + 2][x + 1] * iw10 + S.I[L][y + 2][x + 2] * iw11,
                             kWBits - 5);
      const short2 d00 = S.D[y][x], d01 = S.D[y][x + 1], d10 = S.D[y + 1][x], d11 = S.D[y + 1][x + 1];
      const int ix = descale(d00.x * iw00 + d01.x * iw01 + d10.x * iw10 + d11.x * iw11, kWBits);
      const int iy = descale(d00.y * iw00 + d01.y * iw01 + d10.y * iw10 + d11.y * iw11, kWBits);
      s11 += ix * ix;
      s12 += ix * iy;
      s22 += iy * iy;
      S.Tw[tb][kk] = (short)iv;
      S.Tx[tb][kk] = (short)ix;
      S.Ty[tb][kk] = (short)iy;
    }
  }
  const long long t11 = warp_sum_exact(s11), t12 = warp_sum_exact(s12), t22 = warp_sum_exact(s22);
  if (lane == 0) {
    S.Apart[tb][wi][0] = t11;
    S.Apart[tb][wi][1] = t12;
    S.Apart[tb][wi][2] = t22;
  }
}

// barrier of the four warps of the staging group
__device__ __forceinline__ void bar_staging() { asm volatile("bar.sync 2, %0;" ::"n"(32 * kSWarps) : "memory"); }

struct Region {
  int level, rx0, ry0;  // level -1: nothing staged
};

__device__ __forceinline__ bool region_covers(const Region& g, int level, int inx, int iny) {
  return g.level == level && inx >= g.rx0 && iny >= g.ry0 && inx + kWin <= g.rx0 + kJR - 1 &&
         iny + kWin <= g.ry0 + kJR - 1;
}

// the region of level `level` (image w x h) around window origin floor(p - half); level -1
// if that window lies outside the image (OpenCV then gives up on the level)
__device__ __forceinline__ Region region_around(int level, int w, int h, float px, float py) {
  const int inx = (int)floorf(px - (float)kHalfWin), iny = (int)floorf(py - (float)kHalfWin);
  Region g;
  g.level = (inx < -kWin || inx >= w || iny < -kWin || iny >= h) ? -1 : level;
  g.rx0 = inx - kJM;
  g.ry0 = iny - kJM;
  return g;
}

// One calcOpticalFlowPyrLK call for one point, run by the whole CTA.  Every thread returns
// the same (np, status).
template <int NW>
__device__ void lk_point(LkShared& S, const PyrDesc& pd, const uint8_t* __restrict__ I,
                         const uint8_t* __restrict__ J, float2 p0, float2 init, int use_init,
                         int top, float2& np_out, int& st_out, int clk_call = 0) {
  const int tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
  const float half = (float)kHalfWin;
  const float flt_scale = 1.f / (float)(1 << 20);
  const int nlev = top + 1;
  constexpr int kNThreads = 32 * NW;
  constexpr int kPxN = (kWin * kWin + kNThreads - 1) / kNThreads;  // window pixels per thread: 14, 7 or 4
  const bool n_group = tid < kNThreads;

  // position at the top level (OpenCV: nextPt = prevPt, or the initial flow, scaled)
  float2 np;
  {
    const float sc = 1.f / (float)(1 << top);
    np = use_init ? make_float2(init.x * sc, init.y * sc) : make_float2(p0.x * sc, p0.y * sc);
  }
  // what the two search-region buffers hold: rc = S.Q[cur] (the level being iterated), rn =
  // S.Q[cur ^ 1] (the prefetched next level)
  Region rc = region_around(top, pd.w[top], pd.h[top], np.x, np.y), rn;
  rn.level = -1;
  __syncthreads();  // previous use of S is over
  LK_CLK(0);

  // ---- phase 1: ONE batch of loads: the intensity patches of all levels (they depend on the
  // point only, not on the flow) and the search region of the top level around `np`.
  // Intensities reflect-101 outside the image like the border buildOpticalFlowPyramid adds.
  {
    constexpr int kPerThread = (kMaxLevels * kIP * kIP + kLkThreads - 1) / kLkThreads;  // 9
    uint8_t v[kPerThread];
    uint32_t top_col[kJR / kLkWarps + 1];
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
          v[q] = __ldg(I + pd.off[L] + (size_t)reflect101_nb(ipy - 1 + r, h) * pd.pitch[L] +
                       reflect101_nb(ipx - 1 + c, w));
      }
    }
    if (rc.level >= 0)
      stage_J_load<kLkWarps>(top_col, J + pd.off[top], pd.w[top], pd.h[top], pd.pitch[top], rc.rx0,
                             rc.ry0, warp);
#pragma unroll
    for (int q = 0; q < kPerThread; ++q) {
      const int i = tid + q * kLkThreads;
      if (i < nlev * kIP * kIP) (&S.I[0][0][0])[i] = v[q];
    }
    if (rc.level >= 0) stage_J_store<kLkWarps>(S.Q[0], top_col, warp);
  }
  __syncthreads();
  LK_CLK(1);
  // ---- phases 2 and 3, for the TOP level only, by all eight warps: Scharr derivatives, then
  // the template.  The lower levels' are built by the staging group during the iterations.
  lk_scharr_level<kLkWarps>(S, pd, p0, top, warp);
  __syncthreads();
  LK_CLK(2);
  lk_template_level<kLkWarps>(S, pd, p0, top, warp);
  __syncthreads();
  LK_CLK(3);

  // ---- phase 4: coarse to fine.  The N group iterates, the S group prefetches the next level.
  int st = 1;
  int cur = 0;       // which Q buffer holds the level being iterated
  int n_newton = 0;  // levels iterated so far (slot of S.np / S.st)
  const double eps2 = 0.01 * 0.01;
  int joff[kPxN];
#pragma unroll
  for (int j = 0; j < kPxN; ++j) {
    const int kk = tid + kNThreads * j;
    const int y = kk / kWin, x = kk - y * kWin;
    joff[j] = (n_group && kk < kWin * kWin) ? 4 * (y * kQS + x) : 0;  // in bytes
  }
  for (int level = top; level >= 0; --level) {
    const int w = pd.w[level], h = pd.h[level], pitch = pd.pitch[level];
    const uint8_t* __restrict__ Jl = J + pd.off[level];
    if (level != top) {
      np.x *= 2.f;
      np.y *= 2.f;
    }
    // normal matrix of the level: every thread, from the exact integer sums
    const int tb = level & 1;  // slot of this level's template
    float A11 = 0.f, A12 = 0.f, A22 = 0.f;
    int flag = S.flag_win[level];
    if (!flag) {
      long long a11 = 0, a12 = 0, a22 = 0;
#pragma unroll
      for (int q = 0; q < kLkWarps; ++q) {
        a11 += S.Apart[tb][q][0];
        a12 += S.Apart[tb][q][1];
        a22 += S.Apart[tb][q][2];
      }
      A11 = (float)a11 * flt_scale;
      A12 = (float)a12 * flt_scale;
      A22 = (float)a22 * flt_scale;
      const float det = A11 * A22 - A12 * A12;
      const float min_eig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) /
                            (float)(2 * kWin * kWin);
      if ((double)min_eig < 1e-4 || det < FLT_EPSILON) flag = 2;
    }
    // flag != 0: template window outside the image, or minEig / determinant test failed: OpenCV
    // gives up on the level (the staging group still builds the next level's template)
    const bool skip = flag != 0;
    if (skip && level == 0) st = 0;
    const float D = skip ? 0.f : 1.f / (A11 * A22 - A12 * A12);
    LK_CLK(4 + 4 * level);
    // the current buffer must hold this level around the start position (it does when the
    // prefetch of the level above predicted well, and for the top level from phase 1)
    if (!skip) {
      const Region want = region_around(level, w, h, np.x, np.y);
      if (want.level >= 0 && !region_covers(rc, level, want.rx0 + kJM, want.ry0 + kJM)) {
        rc = want;
        stage_J<kLkWarps>(S.Q[cur], Jl, w, h, pitch, want.rx0, want.ry0, warp);
        __syncthreads();
      }
    }
    LK_CLK(5 + 4 * level);
    // the next level's region around the predicted position 2 * np
    rn.level = -1;
    if (!skip && level > 0)
      rn = region_around(level - 1, pd.w[level - 1], pd.h[level - 1], 2.f * np.x, 2.f * np.y);
    const int slot = n_newton & 1;
    if (warp >= kSFirst) {
      if (rn.level >= 0)
        stage_J<kSWarps>(S.Q[cur ^ 1], J + pd.off[level - 1], pd.w[level - 1], pd.h[level - 1],
                         pd.pitch[level - 1], rn.rx0, rn.ry0, warp - kSFirst);
      if (level > 0) {  // next level's derivatives and template, into the other template slot
        lk_scharr_level<kSWarps>(S, pd, p0, level - 1, warp - kSFirst);
        bar_staging();
        lk_template_level<kSWarps>(S, pd, p0, level - 1, warp - kSFirst);
      }
    } else if (n_group && !skip) {
      int Iw[kPxN], Dx[kPxN], Dy[kPxN];
#pragma unroll
      for (int j = 0; j < kPxN; ++j) {
        const int kk = tid + kNThreads * j;
        const bool on = kk < kWin * kWin;
        Iw[j] = on ? S.Tw[tb][kk] : 0;
        Dx[j] = on ? S.Tx[tb][kk] : 0;  // zero in the unused slots: they add nothing
        Dy[j] = on ? S.Ty[tb][kk] : 0;
      }
      float npx = np.x - half, npy = np.y - half;
      int rx0 = rc.rx0, ry0 = rc.ry0;
      // An iteration is ONE dependent chain (position -> weights -> pixel sums -> reduction ->
      // solve -> stopping rules) of which the 21x21 pixels are a small part: whatever the width
      // of the N group, it costs what its ~150 dependent instructions cost (measured: 620-700
      // cycles with 4, 2 or 1 warps).  So the chain itself is kept short: floor() and cvRound()
      // as magic-number additions on the ALU pipe, weights packed with PRMT, the four bounds tests
      // on the float position, shared memory addressed with 32-bit addresses computed once, the
      // exact sums converted with one FFMA, one test in front of both stopping rules.
      const float kMagic = 12582912.f;  // 1.5 * 2^23: x + kMagic has ulp 1 for |x| < 2^22
      const float wf = (float)w, hf = (float)h;
      // (passed through an opaque mov: otherwise the compiler re-derives them inside the loop)
      const uint32_t q_base = keep_in_register((uint32_t)__cvta_generic_to_shared(&S.Q[cur][0][0]));
      const uint32_t part_base = keep_in_register((uint32_t)__cvta_generic_to_shared(&S.part[0][0]));
      // b = T * 2^-20 only ever enters products that are scaled by D afterwards: a power of two
      // moves through every rounding unchanged (no intermediate comes near the denormals), so
      // 2^-20 is applied to D once instead of to both sums in every iteration
      const float Ds = D * flt_scale;
      float pdx = 1e30f, pdy = 1e30f;  // no previous step: the oscillation rule cannot fire at it = 0
      int n_it = 0;
      for (int it = 0; it < 30; ++it) {
        ++n_it;
        // inx = floor(npx) < -kWin  <=>  npx < -kWin;  inx >= w  <=>  npx >= w (also catches NaN)
        if (!(npx >= -(float)kWin && npx < wf && npy >= -(float)kWin && npy < hf)) {
          if (level == 0) st = 0;
          break;
        }
        // floor: round-toward-minus-infinity addition leaves kMagic + floor(x), exactly
        const float rx = __fadd_rd(npx, kMagic), ry = __fadd_rd(npy, kMagic);
        const int inx = __float_as_int(rx) - 0x4B400000, iny = __float_as_int(ry) - 0x4B400000;
        int ox = inx - rx0, oy = iny - ry0;
        if ((unsigned)ox > (unsigned)(kJR - 1 - kWin) || (unsigned)oy > (unsigned)(kJR - 1 - kWin)) {
          // the window walked out of the staged region (rare): the N group re-stages it
          rx0 = inx - kJM;
          ry0 = iny - kJM;
          ox = kJM;
          oy = kJM;
          bar_newton<NW>();  // everybody is done reading the old region
          stage_J<NW>(S.Q[cur], Jl, w, h, pitch, rx0, ry0, warp);
          bar_newton<NW>();
        }
        const float a = npx - (rx - kMagic), b = npy - (ry - kMagic);  // rx - kMagic = (float)inx
        const float a1 = 1.f - a, b1 = 1.f - b;
        // cvRound(v * 2^14) for v in [0, 1]: the FFMA rounds v * 16384 + kMagic once (half to
        // even), the product being exact, and leaves the integer in the low mantissa bits
        const unsigned u00 = __float_as_uint(fmaf(__fmul_rn(a1, b1), 16384.f, kMagic));
        const unsigned u01 = __float_as_uint(fmaf(__fmul_rn(a, b1), 16384.f, kMagic));
        const unsigned u10 = __float_as_uint(fmaf(__fmul_rn(a1, b), 16384.f, kMagic));
        // iw11 = 2^14 - iw00 - iw01 - iw10 (all three carry the 0x4B400000 of kMagic); it is
        // what is left after three roundings and is -1 when all three round up (a * b < 3e-5,
        // about 5 iterations in 10^5): OpenCV multiplies by that -1, int in its scalar code and
        // int16 in its SIMD code, and so does the signed 16-bit lane of dp2a
        const unsigned u11 = (0xE1C00000u + (1u << kWBits)) - (u00 + u01 + u10);
        // sum_i pix_i * w_i over the packed 2x2 neighbourhood: two dot products of signed 16-bit
        // weights with unsigned 8-bit pixels (top row, then bottom row)
        const unsigned wa = __byte_perm(u00, u01, 0x5410);
        const unsigned wb = __byte_perm(u10, u11, 0x5410);
        const uint32_t qa = q_base + 4u * (unsigned)(oy * kQS + ox);
        int sb1 = 0, sb2 = 0;
#pragma unroll
        for (int j = 0; j < kPxN; ++j) {
          const unsigned q = lds_u32(qa + joff[j]);
          const int v = dp2a_hi_su(wb, q, dp2a_lo_su(wa, q, 1 << (kWBits - 5 - 1)));
          const int diff = (v >> (kWBits - 5)) - Iw[j];
          sb1 += diff * Dx[j];
          sb2 += diff * Dy[j];
        }
        // exact totals as (low 16 bits, the rest): per thread |sb| < 2^29, so over 441 pixels the
        // low parts sum to < 2^23 and the high parts to < 2^19 in magnitude -- both exact in float,
        // and hi * 65536 + lo rounded once by the FFMA is the exact integer rounded to float once
        int lo1 = (int)__reduce_add_sync(0xffffffffu, (unsigned)sb1 & 0xffffu);
        int hi1 = __reduce_add_sync(0xffffffffu, sb1 >> 16);
        int lo2 = (int)__reduce_add_sync(0xffffffffu, (unsigned)sb2 & 0xffffu);
        int hi2 = __reduce_add_sync(0xffffffffu, sb2 >> 16);
        if (NW > 1) {
          const uint32_t pa = part_base + (uint32_t)((it & 1) * kMaxNWarps * 16);
          sts_v4(pa + 16u * (unsigned)warp, lo1, hi1, lo2, hi2);  // every lane, the same value
          bar_newton<NW>();
          lo1 = hi1 = lo2 = hi2 = 0;
#pragma unroll
          for (int q = 0; q < NW; ++q) {
            const int4 pq = lds_v4(pa + 16u * q);
            lo1 += pq.x;
            hi1 += pq.y;
            lo2 += pq.z;
            hi2 += pq.w;
          }
        }
        const float b1s = fmaf((float)hi1, 65536.f, (float)lo1);
        const float b2s = fmaf((float)hi2, 65536.f, (float)lo2);
        const float dx = (A12 * b2s - A22 * b1s) * Ds;
        const float dy = (A12 * b1s - A11 * b2s) * Ds;
        npx += dx;
        npy += dy;
        np.x = npx + half;
        np.y = npy + half;
        // OpenCV evaluates both stopping rules in double.  Far from the thresholds a float
        // estimate decides the same way (its error is ~1e-7 relative, the margins below are a
        // factor 2 / 10 %), so one float test guards both rules and the double arithmetic only
        // runs in the rare close calls.
        const float d2 = dx * dx + dy * dy;
        const float sx = fabsf(dx + pdx), sy = fabsf(dy + pdy);
        const bool near_eps = d2 <= 2e-4f, near_osc = sx < 0.011f && sy < 0.011f;
        if (near_eps || near_osc) {
          if (near_eps && (d2 < 5e-5f || (double)dx * (double)dx + (double)dy * (double)dy <= eps2)) break;
          if (near_osc && ((sx < 0.009f && sy < 0.009f) || ((double)sx < 0.01 && (double)sy < 0.01))) {
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
      LK_CLK(6 + 4 * level);
      LK_VAL(7 + 4 * level, n_it);
      if (tid == 0) {
        S.np[slot] = np;
        S.st[slot] = st;
      }
    }
    __syncthreads();  // level done: its result, the prefetched region and the next template are visible to all
    if (!skip) {
      np = S.np[slot];
      st = S.st[slot];
      ++n_newton;
      cur ^= 1;
      rc = rn;
    }
  }
  LK_CLK(20);
  np_out = np;
  st_out = st;
}

// mode 0: forward only (next = LK(I->J, prev[, init = next]))
// mode 1: temporal pair (feature_tracker.cpp:410,417-418): forward maxLevel `top`, then
//         backward J->I from the forward result with initial flow = prev, maxLevel 1
// mode 2: stereo pair (:490,495): forward, then backward J->I, both maxLevel `top`, no init
template <int NW>
__global__ void __launch_bounds__(kLkThreads, ESVIO_LK_MINB)
k_lk(const __grid_constant__ PyrDesc pd, const uint8_t* __restrict__ I, const uint8_t* __restrict__ J,
     const float2* __restrict__ prev_pts, float2* __restrict__ next_pts,
     uint8_t* __restrict__ status, float2* __restrict__ rev_pts, uint8_t* __restrict__ rev_status,
     const int* __restrict__ n_ptr, int top, int use_init, int mode) {
  PDL_PROLOGUE();
  __shared__ __align__(16) LkShared S;
  const int k = blockIdx.x;
  if (k >= __ldcg(n_ptr)) return;  // produced by the kernel before: never as ld.global.nc (see k_gftt_pick)
  const float2 p0 = prev_pts[k];
  float2 init = make_float2(0.f, 0.f);
  if (use_init) init = next_pts[k];
  float2 np;
  int st;
  lk_point<NW>(S, pd, I, J, p0, init, use_init, top, np, st);
  if (threadIdx.x == 0) {
    next_pts[k] = np;
    status[k] = (uint8_t)st;
  }
  if (mode == 0) return;
  float2 rp;
  int rst;
  if (mode == 1) {
    const int top_b = top < 1 ? top : 1;
    lk_point<NW>(S, pd, J, I, np, p0, 1, top_b, rp, rst, 1);
  } else {
    lk_point<NW>(S, pd, J, I, np, init, 0, top, rp, rst, 1);
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
  static const int nw = getenv("ESVIO_LK_NW") ? atoi(getenv("ESVIO_LK_NW")) : 4;  // experiments
  auto go = [&](auto kern) {
    launch_pdl(kern, dim3(n_max), dim3(kLkThreads), 0, s, pd, I, J, prev_pts, next_pts, status, rev_pts,
               rev_status, n_ptr, top, use_initial_flow, mode);
  };
  if (nw == 1) go(k_lk<1>);
  else if (nw == 2) go(k_lk<2>);
  else go(k_lk<4>);
  ++*launches;
}

// Up to three windows run concurrently on three streams, and the LK kernels sit on every SM for
// most of a step.  Kernels that are to share an SM have to ask for the same L1/shared-memory
// carve-out: with the driver's per-kernel defaults (small for k_lk, large for the many small
// CTAs of k_sae_update_ts) the SAE kernel of the next window gets one CTA per SM next to two LK
// kernels and takes twice as long (measured at 346x260: 23 us instead of 12 us; MaxShared on
// both fixes that too but starves the gather-heavy kernels on those SMs of L1: -5..9 %
// throughput at 640x480).  Both ask for half of the array: room for 3 LK CTAs plus 6 SAE
// CTAs per SM, 114 KB of L1 left.
void prefer_shared_lk() {
  static const int pct = getenv("ESVIO_CARVEOUT") ? atoi(getenv("ESVIO_CARVEOUT")) : 50;  // experiments
  cudaFuncSetAttribute(k_lk<1>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
  cudaFuncSetAttribute(k_lk<2>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
  cudaFuncSetAttribute(k_lk<4>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
}

}  // namespace esvio
