// Frame front-end pieces: cv::goodFeaturesToTrack as
// FeatureTracker::trackImage calls it (feature_tracker/src/feature_tracker.cpp:228) --
// minimum-eigenvalue corners, blockSize 3, Sobel aperture 3, quality 0.01, a mask and a
// minimum distance.  OpenCV's sources are not part of the reference tree; the arithmetic
// below follows the CPU restatement in oracle/esvio_oracle.c (ora_corner_min_eigen_val_u8,
// ora_good_features_to_track), which is pinned bit for bit against cv2.
//
//   k_gftt_cov   Sobel derivatives + their three products per pixel            (W x H threads)
//   k_gftt_eig   3x3 box sums as RUNNING f64 column sums + minimum eigenvalue    (one thread per column)
//   k_gftt_thr   max of the eigenvalues under the mask -> threshold              (one CTA)
//   k_gftt_keys  thresholded 3x3 local maxima -> compacted 64-bit sort keys       (W x H threads)
//   k_gftt_rank  descending order (value, then pixel index) by counting: the rank of a
//                candidate is the number of keys above it                        (one thread per candidate)
//   k_gftt_pick  greedy minimum-distance pick in that order                      (select.cu)
#include "common.cuh"

namespace esvio {

// ---------------------------------------------------------------------------------------
// Sobel(1,0) / Sobel(0,1), ksize 3, scale 1/(255*12), CV_8U -> CV_32F, BORDER_REFLECT_101, in
// the operation order of OpenCV's AVX2/FMA3 build:
//   Dx: rows [-1 0 1] (exact), columns fma(S0 + S2, k, S1 * 2k)
//   Dy: rows k*A (+) 2k*B (+) k*C -- fused multiply-adds in the 32-pixel vector body, separate
//       multiplies and adds in the row tail (x >= 32 * (W / 32)) --, columns S2 - S0
// (the file is compiled with -fmad=false: only the fmaf() calls fuse)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_gftt_cov(const uint8_t* __restrict__ img, int pitch, int W, int H, float* __restrict__ cxx,
           float* __restrict__ cxy, float* __restrict__ cyy) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= W || y >= H) return;
  const float k1 = (float)(1.0 / 3060.0), k0 = (float)(2.0 * (1.0 / 3060.0));
  const int xm = reflect101_nb(x - 1, W), xp = reflect101_nb(x + 1, W);
  const bool body = x < (W / 32) * 32;
  float d[3], s[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const uint8_t* row = img + (size_t)reflect101_nb(y - 1 + r, H) * pitch;
    const float a = (float)row[xm], b = (float)row[x], c = (float)row[xp];
    d[r] = c - a;
    s[r] = body ? fmaf(c, k1, fmaf(b, k0, a * k1)) : (a * k1 + b * k0) + c * k1;
  }
  const float gx = fmaf(d[0] + d[2], k1, d[1] * k0), gy = s[2] - s[0];
  const size_t i = (size_t)y * W + x;
  cxx[i] = gx * gx;
  cxy[i] = gx * gy;
  cyy[i] = gy * gy;
}

// ---------------------------------------------------------------------------------------
// boxFilter(3x3, normalize = false) of the three products and calcMinEigenVal.  OpenCV sums
// rows as (S0 + S1) + S2 in f64 and then keeps ONE running f64 sum per pixel column down the
// whole image (s = SUM + row[y+1]; out = (float)s; SUM = s - row[y-1]); the rounding history of
// that sum is part of the result, so every column is walked top to bottom by one thread.
// ---------------------------------------------------------------------------------------
constexpr int kEigThreads = 64;

__device__ __forceinline__ double row_sum3(const float* __restrict__ p, size_t row, int xm, int x,
                                           int xp) {
  return ((double)p[row + xm] + (double)p[row + x]) + (double)p[row + xp];
}

__global__ void __launch_bounds__(kEigThreads)
k_gftt_eig(const float* __restrict__ cxx, const float* __restrict__ cxy,
           const float* __restrict__ cyy, int W, int H, float* __restrict__ eig) {
  const int x = blockIdx.x * kEigThreads + threadIdx.x;
  if (x >= W) return;
  const int xm = reflect101_nb(x - 1, W), xp = reflect101_nb(x + 1, W);
  const float* pl[3] = {cxx, cxy, cyy};
  double r0[3], r1[3], sum[3];
  {
    const size_t ra = (size_t)reflect101_nb(-1, H) * W, rb = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      r0[c] = row_sum3(pl[c], ra, xm, x, xp);
      r1[c] = row_sum3(pl[c], rb, xm, x, xp);
      sum[c] = (0.0 + r0[c]) + r1[c];
    }
  }
#pragma unroll 4
  for (int y = 0; y < H; ++y) {
    const size_t rn = (size_t)reflect101_nb(y + 1, H) * W;
    float cv[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double r2 = row_sum3(pl[c], rn, xm, x, xp);
      const double s0 = sum[c] + r2;
      cv[c] = (float)s0;
      sum[c] = s0 - r0[c];
      r0[c] = r1[c];
      r1[c] = r2;
    }
    const float a = cv[0] * 0.5f, b = cv[1], c2 = cv[2] * 0.5f;
    const float t = a - c2;
    eig[(size_t)y * W + x] = (a + c2) - sqrtf(t * t + b * b);
  }
}

// order-preserving map of a float onto an unsigned integer
__device__ __forceinline__ uint32_t float_order(float v) {
  const uint32_t b = __float_as_uint(v);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float float_unorder(uint32_t o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// blocked: one bit per pixel, [H][words], 1 = the mask of goodFeaturesToTrack is zero there
__device__ __forceinline__ bool gftt_blocked(const uint32_t* __restrict__ blocked, int words, int x,
                                             int y) {
  return blocked && ((blocked[(size_t)y * words + (x >> 5)] >> (x & 31)) & 1u);
}

// minMaxLoc(eig, 0, &maxVal, 0, 0, mask); threshold = (float)(maxVal * qualityLevel)
__global__ void __launch_bounds__(1024)
k_gftt_thr(const float* __restrict__ eig, int W, int H, const uint32_t* __restrict__ blocked,
           double quality, float* __restrict__ thr_out, int* __restrict__ n_cand) {
  PDL_PROLOGUE();
  if (threadIdx.x == 0) *n_cand = 0;  // k_gftt_keys of this frame counts from here
  __shared__ uint32_t s_max[32];
  const int words = (W + 31) / 32;
  uint32_t best = 0;  // below every float_order() value: "no pixel seen"
  for (int i = threadIdx.x; i < W * H; i += blockDim.x) {
    const int y = i / W, x = i - y * W;
    if (!gftt_blocked(blocked, words, x, y)) best = max(best, float_order(eig[i]));
  }
  best = __reduce_max_sync(0xffffffffu, best);
  if (lane_id() == 0) s_max[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x < 32) {
    best = __reduce_max_sync(0xffffffffu, s_max[threadIdx.x]);
    if (threadIdx.x == 0) {
      const double max_val = best ? (double)float_unorder(best) : 0.0;
      *thr_out = (float)(max_val * quality);
    }
  }
}

// threshold(THRESH_TOZERO) + dilate(3x3) + "val != 0 && val == dilated && mask": every such
// pixel becomes the key (float_order(val) << 32 | pixel index); the keys in descending order are
// featureselect.cpp's greaterThanPtr order (value, then address).  A frame has a few thousand
// candidates among its W x H pixels, so they are compacted here (a CTA reserves its slice of
// the list with one atomic; the order inside the list does not matter) and only they are ordered.
__global__ void __launch_bounds__(256)
k_gftt_keys(const float* __restrict__ eig, int W, int H, const uint32_t* __restrict__ blocked,
            const float* __restrict__ thr_ptr, unsigned long long* __restrict__ keys,
            int* __restrict__ n_cand) {
  PDL_PROLOGUE();
  __shared__ int s_wc[8], s_base;
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const int x = blockIdx.x * 32 + lane, y = blockIdx.y * 8 + warp;
  unsigned long long key = 0;
  if (x >= 1 && x < W - 1 && y >= 1 && y < H - 1) {
    const size_t i = (size_t)y * W + x;
    const float thr = __ldcg(thr_ptr);
    const float v = eig[i];
    if (v > thr && v != 0.f && !gftt_blocked(blocked, (W + 31) / 32, x, y)) {
      bool is_max = true;
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
          const float n = eig[i + (ptrdiff_t)dy * W + dx];
          const float nz = n > thr ? n : 0.f;
          if (nz > v) is_max = false;
        }
      if (is_max) key = ((unsigned long long)float_order(v) << 32) | (unsigned long long)i;
    }
  }
  const uint32_t m = __ballot_sync(0xffffffffu, key != 0);
  if (lane == 0) s_wc[warp] = __popc(m);
  __syncthreads();
  if (threadIdx.x == 0) {
    int total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const int c = s_wc[w];
      s_wc[w] = total;
      total += c;
    }
    s_base = total ? atomicAdd(n_cand, total) : 0;
  }
  __syncthreads();
  if (key) keys[s_base + s_wc[warp] + __popc(m & ((1u << lane) - 1u))] = key;
}

// keys[0 .. *n_cand) -> sorted[rank], rank = number of keys greater than mine (the keys are
// distinct: the pixel index is part of them); sorted[*n_cand] = 0 ends the list for k_gftt_pick.
// All keys pass through shared memory in tiles; n^2 / 2 comparisons are ~10 us for the few
// thousand candidates of a frame (a 640x480 noise image with 77 000 maxima takes ~1 ms).
constexpr int kRankTile = 2048;
__global__ void __launch_bounds__(256)
k_gftt_rank(const unsigned long long* __restrict__ keys, const int* __restrict__ n_cand,
            unsigned long long* __restrict__ sorted, int capacity) {
  PDL_PROLOGUE();
  __shared__ unsigned long long s_tile[kRankTile];
  const int n = min(__ldcg(n_cand), capacity);  // not ld.global.nc: see k_gftt_pick
  if (blockIdx.x == 0 && threadIdx.x == 0 && n < capacity) sorted[n] = 0ull;
  const int first = blockIdx.x * 256;
  if (first >= n) return;
  const int i = first + threadIdx.x;
  const unsigned long long mine = i < n ? keys[i] : ~0ull;
  int rank = 0;
  for (int t0 = 0; t0 < n; t0 += kRankTile) {
    const int cnt = min(kRankTile, n - t0);
    __syncthreads();
    for (int j = threadIdx.x; j < cnt; j += 256) s_tile[j] = keys[t0 + j];
    __syncthreads();
    int j = 0;
    for (; j + 4 <= cnt; j += 4)
      rank += (s_tile[j] > mine) + (s_tile[j + 1] > mine) + (s_tile[j + 2] > mine) + (s_tile[j + 3] > mine);
    for (; j < cnt; ++j) rank += s_tile[j] > mine;
  }
  if (i < n) sorted[rank] = mine;
}

// eig + thr (+ mask) -> keys_sorted: the candidate corners, best first, then a zero key
int launch_gftt_candidates(const GfttBuffers& G, int W, int H, bool use_mask, cudaStream_t s,
                           int64_t* launches) {
  const dim3 grid((W + 31) / 32, (H + 7) / 8);
  launch_pdl(k_gftt_keys, grid, dim3(256), 0, s, (const float*)G.eig, W, H,
             use_mask ? (const uint32_t*)G.blocked : (const uint32_t*)nullptr,
             (const float*)G.thr, G.keys, G.n_cand);
  launch_pdl(k_gftt_rank, dim3((W * H + 255) / 256), dim3(256), 0, s, (const unsigned long long*)G.keys,
             (const int*)G.n_cand, G.keys_sorted, W * H);
  *launches += 2;
  return cudaPeekAtLastError() == cudaSuccess ? 0 : -1;
}

void launch_gftt_eig(const GfttBuffers& G, const uint8_t* img, int pitch, int W, int H,
                     cudaStream_t s, int64_t* launches) {
  const dim3 grid((W + 31) / 32, (H + 7) / 8);
  k_gftt_cov<<<grid, 256, 0, s>>>(img, pitch, W, H, G.cov[0], G.cov[1], G.cov[2]);
  k_gftt_eig<<<(W + kEigThreads - 1) / kEigThreads, kEigThreads, 0, s>>>(G.cov[0], G.cov[1],
                                                                         G.cov[2], W, H, G.eig);
  *launches += 2;
}

void launch_gftt_thr(const GfttBuffers& G, int W, int H, bool use_mask, cudaStream_t s,
                     int64_t* launches) {
  launch_pdl(k_gftt_thr, dim3(1), dim3(1024), 0, s, (const float*)G.eig, W, H,
             use_mask ? (const uint32_t*)G.blocked : (const uint32_t*)nullptr, 0.01, G.thr, G.n_cand);
  ++*launches;
}

}  // namespace esvio
