// pyramid.cu -- image pyramid of the time surface, as cv::buildOpticalFlowPyramid builds it
// inside cv::calcOpticalFlowPyrLK (called at feature_tracker/src/feature_tracker.cpp:410,
// 417,490,495): level l+1 = 5x5 separable [1 4 6 4 1] of level l, BORDER_REFLECT_101,
// decimated by 2, size ((w+1)/2,(h+1)/2), value (sum + 128) >> 8.
#include "common.cuh"

namespace esvio {

// All levels in one launch: a CTA owns a 4x4 tile of level 3 (8x8 of level 2, 16x16 of
// level 1) and cascades through shared memory: level-0 patch 53x53 -> level-1 patch 25x25
// -> level-2 patch 11x11 -> level-3 tile 4x4 (every patch carries the halo the next level's
// 5x5 taps need; the halo pixels are recomputed by the neighbouring CTAs).  Patches are
// addressed in the image coordinates of their level, so BORDER_REFLECT_101 is applied per
// level exactly as a level-by-level build would.  Small tiles on purpose: the kernel moves
// ~0.4 MB per image and is bound by the length of one CTA's chain of passes, so it wants many
// short CTAs (600 for a 640x480 stereo pair, 4 per SM) rather than few long ones.  The level-0
// patch of tile bx spans columns [32 bx - 14, 32 bx + 39): four aligned 16-byte loads per row.
constexpr int kPyrT3 = 4;
constexpr int kPyrN2 = 2 * kPyrT3 + 3;   // 11
constexpr int kPyrN1 = 2 * kPyrN2 + 3;   // 25
constexpr int kPyrN0 = 2 * kPyrN1 + 3;   // 53
constexpr int kPyrThreads = 128;
constexpr int kPyrChunks = 4;            // 16-byte chunks per level-0 patch row
constexpr int kPyrLead = 2;              // patch column 0 sits at byte 2 of the first chunk
static_assert(8 * kPyrT3 == 32 && kPyrLead + kPyrN0 <= 16 * kPyrChunks, "level-0 patch = 4 aligned chunks");

// dst patch (nd x nd at image origin (dx0, dy0), image dw x dh) from src patch (ns x ns at
// (sx0, sy0), image sw x sh); pixels of the owned rectangle [ox0, ox0+on) x [oy0, oy0+on)
// also go to global memory.
template <int NS, int ND>
__device__ __forceinline__ void pyr_stage(const uint8_t* __restrict__ s_src, int ssp, int sx0,
                                          int sy0, int sw, int sh, uint16_t* __restrict__ s_row,
                                          uint8_t* __restrict__ s_dst, int dsp, int dx0, int dy0,
                                          int dw, int dh, uint8_t* __restrict__ g_dst, int gpitch,
                                          int ox0, int oy0, int on) {
  // horizontal [1 4 6 4 1] on every source row of the patch
  for (int i = threadIdx.x; i < NS * ND; i += kPyrThreads) {
    const int r = i / ND, c = i - r * ND;
    const int gx = dx0 + c;
    uint32_t v = 0;
    if (gx >= 0 && gx < dw) {
      const uint8_t* row = s_src + r * ssp;
      const int x0 = reflect101_nb(2 * gx - 2, sw) - sx0, x1 = reflect101_nb(2 * gx - 1, sw) - sx0,
                x2 = 2 * gx - sx0, x3 = reflect101_nb(2 * gx + 1, sw) - sx0,
                x4 = reflect101_nb(2 * gx + 2, sw) - sx0;
      v = row[x2] * 6 + (row[x1] + row[x3]) * 4 + row[x0] + row[x4];
    }
    s_row[r * ND + c] = (uint16_t)v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ND * ND; i += kPyrThreads) {
    const int r = i / ND, c = i - r * ND;
    const int gx = dx0 + c, gy = dy0 + r;
    if (gx >= 0 && gx < dw && gy >= 0 && gy < dh) {
      const int y0 = reflect101_nb(2 * gy - 2, sh) - sy0, y1 = reflect101_nb(2 * gy - 1, sh) - sy0,
                y2 = 2 * gy - sy0, y3 = reflect101_nb(2 * gy + 1, sh) - sy0,
                y4 = reflect101_nb(2 * gy + 2, sh) - sy0;
      const int v = s_row[y2 * ND + c] * 6 + (s_row[y1 * ND + c] + s_row[y3 * ND + c]) * 4 +
                    s_row[y0 * ND + c] + s_row[y4 * ND + c];
      const uint8_t o = (uint8_t)((v + 128) >> 8);
      s_dst[r * dsp + c] = o;
      if (gx >= ox0 && gx < ox0 + on && gy >= oy0 && gy < oy0 + on)
        g_dst[(size_t)gy * gpitch + gx] = o;
    }
  }
  __syncthreads();
}

struct PyrImages {
  uint8_t* img[kMaxCams];
};

__global__ void __launch_bounds__(kPyrThreads)
k_pyr_build(PyrDesc pd, const __grid_constant__ PyrImages imgs) {
  PDL_PROLOGUE();
  constexpr int SP0 = 16 * kPyrChunks, SP1 = kPyrN1 + 3, SP2 = kPyrN2 + 1;
  __shared__ __align__(16) uint8_t s0_raw[kPyrN0 * SP0];
  __shared__ uint8_t s1[kPyrN1 * SP1];
  __shared__ uint8_t s2[kPyrN2 * SP2];
  __shared__ uint8_t s3[kPyrT3 * kPyrT3];
  __shared__ uint16_t s_row[kPyrN0 * kPyrN1];
  uint8_t* __restrict__ img = imgs.img[blockIdx.z];
  const int o3x = blockIdx.x * kPyrT3, o3y = blockIdx.y * kPyrT3;
  const int p2x = 2 * o3x - 2, p2y = 2 * o3y - 2;
  const int p1x = 2 * p2x - 2, p1y = 2 * p2y - 2;
  const int p0x = 2 * p1x - 2, p0y = 2 * p1y - 2;
  const uint8_t* s0 = s0_raw + kPyrLead;  // s0[r * SP0 + c] = patch pixel (r, c)
  {
    // level-0 patch by 16-byte vector loads (rows are 32-byte aligned; bytes beyond the image
    // width inside the pitch are never used: reflect-101 folds every tap back inside)
    const uint8_t* __restrict__ src = img + pd.off[0];
    const int h = pd.h[0], pitch = pd.pitch[0];
    const int xa = p0x - kPyrLead;  // 32 * blockIdx.x - 16
    for (int i = threadIdx.x; i < kPyrN0 * kPyrChunks; i += kPyrThreads) {
      const int r = i / kPyrChunks, j = i - r * kPyrChunks;
      const int gy = p0y + r, gx = xa + 16 * j;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (gy >= 0 && gy < h && gx >= 0 && gx + 16 <= pitch)
        v = __ldg(reinterpret_cast<const uint4*>(src + (size_t)gy * pitch + gx));
      *reinterpret_cast<uint4*>(s0_raw + r * SP0 + 16 * j) = v;
    }
  }
  __syncthreads();
  pyr_stage<kPyrN0, kPyrN1>(s0, SP0, p0x, p0y, pd.w[0], pd.h[0], s_row, s1, SP1, p1x, p1y, pd.w[1],
                            pd.h[1], img + pd.off[1], pd.pitch[1], 4 * o3x, 4 * o3y, 4 * kPyrT3);
  if (pd.levels > 2)
    pyr_stage<kPyrN1, kPyrN2>(s1, SP1, p1x, p1y, pd.w[1], pd.h[1], s_row, s2, SP2, p2x, p2y, pd.w[2],
                              pd.h[2], img + pd.off[2], pd.pitch[2], 2 * o3x, 2 * o3y, 2 * kPyrT3);
  if (pd.levels > 3)
    pyr_stage<kPyrN2, kPyrT3>(s2, SP2, p2x, p2y, pd.w[2], pd.h[2], s_row, s3, kPyrT3, o3x, o3y,
                              pd.w[3], pd.h[3], img + pd.off[3], pd.pitch[3], o3x, o3y, kPyrT3);
}

void launch_pyramids(const PyrDesc& pd, uint8_t* const* pyr, int n_img, cudaStream_t s,
                     int64_t* launches) {
  if (pd.levels < 2 || n_img < 1 || n_img > kMaxCams) return;
  PyrImages imgs;
  for (int i = 0; i < kMaxCams; ++i) imgs.img[i] = i < n_img ? pyr[i] : nullptr;
  const dim3 grid((pd.w[1] + 4 * kPyrT3 - 1) / (4 * kPyrT3), (pd.h[1] + 4 * kPyrT3 - 1) / (4 * kPyrT3),
                  n_img);
  launch_pdl(k_pyr_build, grid, dim3(kPyrThreads), 0, s, pd, imgs);
  ++*launches;
}

}  // namespace esvio
