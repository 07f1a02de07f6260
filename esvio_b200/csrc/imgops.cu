// imgops.cu -- optional conditioning of the CV_8U time surface:
//   cv::medianBlur(ts, ts, 2k+1)                    event_detector.cc:262-264
//   cv::createCLAHE()->apply + cv::normalize(MINMAX) feature_tracker.cpp:375-382
// restated from OpenCV's modules/imgproc/src/{median_blur,clahe}.cpp and core convertTo;
// arithmetic order follows the CPU oracle (oracle/esvio_oracle.c), which is pinned against
// cv2 outputs in tests/golden/imgops_golden.npz.
#include "common.cuh"

namespace esvio {

// ------------------------------------------------------------------------------------
// median: exact median of the ksize x ksize window, BORDER_REPLICATE
// ------------------------------------------------------------------------------------
constexpr int kMedTW = 32, kMedTH = 8, kMedMaxR = 7;

__global__ void __launch_bounds__(kMedTW* kMedTH)
k_median(const uint8_t* __restrict__ src0, const uint8_t* __restrict__ src1,
         uint8_t* __restrict__ dst0, uint8_t* __restrict__ dst1, int W, int H, int pitch, int r) {
  __shared__ uint8_t s_in[(kMedTH + 2 * kMedMaxR) * (kMedTW + 2 * kMedMaxR)];
  const uint8_t* __restrict__ src = blockIdx.z ? src1 : src0;
  uint8_t* __restrict__ dst = blockIdx.z ? dst1 : dst0;
  const int pw = kMedTW + 2 * r, ph = kMedTH + 2 * r;
  const int x0 = blockIdx.x * kMedTW - r, y0 = blockIdx.y * kMedTH - r;
  const int tid = threadIdx.y * kMedTW + threadIdx.x;
  for (int i = tid; i < pw * ph; i += kMedTW * kMedTH) {
    const int py = i / pw, px = i - py * pw;
    const int gx = min(max(x0 + px, 0), W - 1), gy = min(max(y0 + py, 0), H - 1);
    s_in[py * pw + px] = src[(size_t)gy * pitch + gx];
  }
  __syncthreads();
  const int x = blockIdx.x * kMedTW + threadIdx.x, y = blockIdx.y * kMedTH + threadIdx.y;
  if (x >= W || y >= H) return;
  const int k = 2 * r + 1, half = (k * k) / 2;
  const uint8_t* win = s_in + threadIdx.y * pw + threadIdx.x;
  int lo = 0, hi = 255;  // smallest v with #(window <= v) > half
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    int cnt = 0;
    for (int dy = 0; dy < k; ++dy)
      for (int dx = 0; dx < k; ++dx) cnt += win[dy * pw + dx] <= mid;
    if (cnt > half) hi = mid;
    else lo = mid + 1;
  }
  dst[(size_t)y * pitch + x] = (uint8_t)lo;
}

void launch_median(const uint8_t* const src[2], uint8_t* const dst[2], int n_img, int W, int H,
                   int pitch, int ksize, cudaStream_t s, int64_t* launches) {
  const dim3 grid((W + kMedTW - 1) / kMedTW, (H + kMedTH - 1) / kMedTH, n_img);
  k_median<<<grid, dim3(kMedTW, kMedTH), 0, s>>>(src[0], src[n_img > 1 ? 1 : 0], dst[0],
                                                dst[n_img > 1 ? 1 : 0], W, H, pitch, ksize / 2);
  ++*launches;
}

// ------------------------------------------------------------------------------------
// CLAHE (clip limit 40, 8x8 tiles) + normalize(0, 255, MINMAX)
// ------------------------------------------------------------------------------------
constexpr int kClaheTiles = 8;
constexpr double kClaheClip = 40.0;

struct ClaheGeom {
  int W, H, pitch;
  int tw, th;  // tile size of the (padded) image
  int clip;
  float lut_scale;
};

static ClaheGeom clahe_geom(int W, int H, int pitch) {
  ClaheGeom g;
  g.W = W;
  g.H = H;
  g.pitch = pitch;
  int EW = W, EH = H;
  if (W % kClaheTiles != 0 || H % kClaheTiles != 0) {  // clahe.cpp pads BOTH dimensions then
    EW = W + (kClaheTiles - W % kClaheTiles);
    EH = H + (kClaheTiles - H % kClaheTiles);
  }
  g.tw = EW / kClaheTiles;
  g.th = EH / kClaheTiles;
  const int area = g.tw * g.th;
  g.lut_scale = (float)255 / area;
  g.clip = (int)(kClaheClip * area / 256);
  if (g.clip < 1) g.clip = 1;
  return g;
}

__device__ __forceinline__ uint8_t sat_u8_f(float v) {  // saturate_cast<uchar>(float)
  int r = __float2int_rn(v);
  r = r < 0 ? 0 : (r > 255 ? 255 : r);
  return (uint8_t)r;
}

// one CTA (256 threads = 256 bins) per tile and image: clipped histogram -> LUT
__global__ void __launch_bounds__(256)
k_clahe_lut(ClaheGeom g, const uint8_t* __restrict__ src0, const uint8_t* __restrict__ src1,
            uint8_t* __restrict__ lut, int* __restrict__ minmax) {
  __shared__ int s_hist[256];
  __shared__ int s_warp[8];
  __shared__ int s_clipped;
  const int cam = blockIdx.z, tx = blockIdx.x, ty = blockIdx.y, tid = threadIdx.x;
  const uint8_t* __restrict__ src = cam ? src1 : src0;
  if (tx == 0 && ty == 0 && tid == 0) {
    minmax[2 * cam] = 255;
    minmax[2 * cam + 1] = 0;
  }
  s_hist[tid] = 0;
  if (tid == 0) s_clipped = 0;
  __syncthreads();
  for (int i = tid; i < g.tw * g.th; i += 256) {
    const int py = i / g.tw, px = i - py * g.tw;
    const int sx = reflect101_nb(tx * g.tw + px, g.W), sy = reflect101_nb(ty * g.th + py, g.H);
    atomicAdd(&s_hist[src[(size_t)sy * g.pitch + sx]], 1);
  }
  __syncthreads();
  int h = s_hist[tid];
  if (h > g.clip) {
    atomicAdd(&s_clipped, h - g.clip);
    h = g.clip;
  }
  __syncthreads();
  const int clipped = s_clipped;
  const int batch = clipped / 256, residual = clipped - batch * 256;
  h += batch;
  if (residual != 0) {
    const int step = max(256 / residual, 1);
    // bins 0, step, 2*step, ... get one more, `residual` of them at most
    if (tid % step == 0 && tid / step < residual) ++h;
  }
  // inclusive scan over the 256 bins
  const int lane = tid & 31, warp = tid >> 5;
  int incl = h;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += o;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  int base = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w)
    if (w < warp) base += s_warp[w];
  const int sum = base + incl;
  lut[(((size_t)cam * kClaheTiles + ty) * kClaheTiles + tx) * 256 + tid] =
      sat_u8_f((float)sum * g.lut_scale);
}

// bilinear blend of the four nearest tile LUTs + min/max of the result
__global__ void __launch_bounds__(256)
k_clahe_apply(ClaheGeom g, const uint8_t* src0, const uint8_t* src1, const uint8_t* __restrict__ lut,
              uint8_t* dst0, uint8_t* dst1, int* __restrict__ minmax) {
  const int cam = blockIdx.z;
  const uint8_t* src = cam ? src1 : src0;  // may alias dst (launch_clahe_inplace): a thread
  uint8_t* dst = cam ? dst1 : dst0;        // reads only the pixel it writes
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  int v_out = -1;
  if (x < g.W && y < g.H) {
    const float inv_tw = 1.0f / g.tw, inv_th = 1.0f / g.th;
    const float tyf = y * inv_th - 0.5f;
    int ty1 = (int)floorf(tyf), ty2 = ty1 + 1;
    const float ya = tyf - ty1, ya1 = 1.0f - ya;
    ty1 = max(ty1, 0);
    ty2 = min(ty2, kClaheTiles - 1);
    const float txf = x * inv_tw - 0.5f;
    int tx1 = (int)floorf(txf), tx2 = tx1 + 1;
    const float xa = txf - tx1, xa1 = 1.0f - xa;
    tx1 = max(tx1, 0);
    tx2 = min(tx2, kClaheTiles - 1);
    const int v = src[(size_t)y * g.pitch + x];
    const uint8_t* p1 = lut + (((size_t)cam * kClaheTiles + ty1) * kClaheTiles) * 256;
    const uint8_t* p2 = lut + (((size_t)cam * kClaheTiles + ty2) * kClaheTiles) * 256;
    const float res = (p1[tx1 * 256 + v] * xa1 + p1[tx2 * 256 + v] * xa) * ya1 +
                      (p2[tx1 * 256 + v] * xa1 + p2[tx2 * 256 + v] * xa) * ya;
    v_out = sat_u8_f(res);
    dst[(size_t)y * g.pitch + x] = (uint8_t)v_out;
  }
  int mn = v_out < 0 ? 255 : v_out, mx = v_out < 0 ? 0 : v_out;
  mn = __reduce_min_sync(0xffffffffu, mn);
  mx = __reduce_max_sync(0xffffffffu, mx);
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&minmax[2 * cam], mn);
    atomicMax(&minmax[2 * cam + 1], mx);
  }
}

// cv::normalize(0, 255, NORM_MINMAX) -> convertTo(CV_8U, scale, shift) in float
__global__ void __launch_bounds__(256)
k_normalize(int W, int H, int pitch, const uint8_t* __restrict__ src0,
            const uint8_t* __restrict__ src1, uint8_t* __restrict__ dst0,
            uint8_t* __restrict__ dst1, const int* __restrict__ minmax) {
  const int cam = blockIdx.z;
  const uint8_t* __restrict__ src = cam ? src1 : src0;
  uint8_t* __restrict__ dst = cam ? dst1 : dst0;
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= W || y >= H) return;
  const int mn = minmax[2 * cam], mx = minmax[2 * cam + 1];
  const double scale = 255.0 * ((double)(mx - mn) > 2.220446049250313e-16 ? 1.0 / (double)(mx - mn) : 0.0);
  const double shift = 0.0 - (double)mn * scale;
  dst[(size_t)y * pitch + x] =
      sat_u8_f(fmaf((float)src[(size_t)y * pitch + x], (float)scale, (float)shift));
}

size_t clahe_lut_bytes() { return (size_t)2 * kClaheTiles * kClaheTiles * 256; }

// src -> CLAHE -> tmp -> normalize -> dst (src, tmp, dst: [n_img] images of the same pitch)
void launch_equalize(const uint8_t* const src[2], uint8_t* const tmp[2], uint8_t* const dst[2],
                     int n_img, int W, int H, int pitch, uint8_t* lut, int* minmax,
                     cudaStream_t s, int64_t* launches) {
  const ClaheGeom g = clahe_geom(W, H, pitch);
  const int o = n_img > 1 ? 1 : 0;
  k_clahe_lut<<<dim3(kClaheTiles, kClaheTiles, n_img), 256, 0, s>>>(g, src[0], src[o], lut, minmax);
  const dim3 grid((W + 31) / 32, (H + 7) / 8, n_img);
  k_clahe_apply<<<grid, 256, 0, s>>>(g, src[0], src[o], lut, tmp[0], tmp[o], minmax);
  k_normalize<<<grid, 256, 0, s>>>(W, H, pitch, tmp[0], tmp[o], dst[0], dst[o], minmax);
  *launches += 3;
}

// cv::createCLAHE()->apply(img, img) alone, in place (the image node's EQUALIZE,
// stereo_image_tracker_node.cpp:93-97): the LUTs are complete before the first pixel changes,
// and every thread rewrites only the pixel it read
void launch_clahe_inplace(uint8_t* const img[2], int n_img, int W, int H, int pitch, uint8_t* lut,
                          int* minmax, cudaStream_t s, int64_t* launches) {
  const ClaheGeom g = clahe_geom(W, H, pitch);
  const int o = n_img > 1 ? 1 : 0;
  k_clahe_lut<<<dim3(kClaheTiles, kClaheTiles, n_img), 256, 0, s>>>(g, img[0], img[o], lut, minmax);
  const dim3 grid((W + 31) / 32, (H + 7) / 8, n_img);
  k_clahe_apply<<<grid, 256, 0, s>>>(g, img[0], img[o], lut, img[0], img[o], minmax);
  *launches += 2;
}

}  // namespace esvio
