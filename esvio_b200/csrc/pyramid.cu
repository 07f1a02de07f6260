// pyramid.cu -- image pyramid of the time surface, as cv::buildOpticalFlowPyramid builds it
// inside cv::calcOpticalFlowPyrLK (called at feature_tracker/src/feature_tracker.cpp:410,
// 417,490,495): level l+1 = 5x5 separable [1 4 6 4 1] of level l, BORDER_REFLECT_101,
// decimated by 2, size ((w+1)/2,(h+1)/2), value (sum + 128) >> 8.
#include "common.cuh"

namespace esvio {

constexpr int kPyrTW = 32, kPyrTH = 8;  // output tile of one CTA

// One CTA produces a 32x8 output tile: the (2*32+3) x (2*8+3) input patch is staged in
// shared memory (reflect-101 applied while staging), filtered horizontally into int rows,
// then vertically.
__global__ void __launch_bounds__(kPyrTW* kPyrTH)
k_pyr_down(const uint8_t* __restrict__ src0, const uint8_t* __restrict__ src1, int sw, int sh,
           int spitch, uint8_t* __restrict__ dst0, uint8_t* __restrict__ dst1, int dw, int dh,
           int dpitch) {
  constexpr int PW = 2 * kPyrTW + 3, PH = 2 * kPyrTH + 3;
  __shared__ uint8_t s_in[PH][PW + 1];
  __shared__ int s_row[PH][kPyrTW];
  const uint8_t* __restrict__ src = blockIdx.z ? src1 : src0;
  uint8_t* __restrict__ dst = blockIdx.z ? dst1 : dst0;
  const int ox0 = blockIdx.x * kPyrTW, oy0 = blockIdx.y * kPyrTH;
  const int ix0 = 2 * ox0 - 2, iy0 = 2 * oy0 - 2;
  const int tid = threadIdx.y * kPyrTW + threadIdx.x;
  for (int i = tid; i < PW * PH; i += kPyrTW * kPyrTH) {
    const int py = i / PW, px = i - py * PW;
    const int gx = reflect101(ix0 + px, sw), gy = reflect101(iy0 + py, sh);
    s_in[py][px] = src[(size_t)gy * spitch + gx];
  }
  __syncthreads();
  for (int i = tid; i < PH * kPyrTW; i += kPyrTW * kPyrTH) {
    const int py = i / kPyrTW, ox = i - py * kPyrTW;
    const uint8_t* r = &s_in[py][2 * ox];
    s_row[py][ox] = r[2] * 6 + (r[1] + r[3]) * 4 + r[0] + r[4];
  }
  __syncthreads();
  const int ox = ox0 + threadIdx.x, oy = oy0 + threadIdx.y;
  if (ox < dw && oy < dh) {
    const int ly = 2 * threadIdx.y, lx = threadIdx.x;
    const int v = s_row[ly + 2][lx] * 6 + (s_row[ly + 1][lx] + s_row[ly + 3][lx]) * 4 +
                  s_row[ly][lx] + s_row[ly + 4][lx];
    dst[(size_t)oy * dpitch + ox] = (uint8_t)((v + 128) >> 8);
  }
}

void launch_pyramids(const PyrDesc& pd, uint8_t* const pyr[2], int n_img, cudaStream_t s,
                     int64_t* launches) {
  for (int l = 0; l + 1 < pd.levels; ++l) {
    dim3 grid((pd.w[l + 1] + kPyrTW - 1) / kPyrTW, (pd.h[l + 1] + kPyrTH - 1) / kPyrTH, n_img);
    k_pyr_down<<<grid, dim3(kPyrTW, kPyrTH), 0, s>>>(
        pyr[0] + pd.off[l], pyr[n_img > 1 ? 1 : 0] + pd.off[l], pd.w[l], pd.h[l], pd.pitch[l],
        pyr[0] + pd.off[l + 1], pyr[n_img > 1 ? 1 : 0] + pd.off[l + 1], pd.w[l + 1], pd.h[l + 1],
        pd.pitch[l + 1]);
    ++*launches;
  }
}

}  // namespace esvio
