// Latency of ONE isolated host->device transfer of a window's events (4.33 MB, pinned), the way the
// synchronous esvio_fe_track call sees it: the GPU is idle before, the host waits for the end.
//   (a) one cudaMemcpyAsync            (b) the same bytes as k pieces on k streams
//   (c) a copy kernel on the SMs reading the pinned block through its device alias (UVA)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scratch/h2d_latency scratch/h2d_latency.cu
#include <cuda_runtime.h>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

template <int U>
__global__ void __launch_bounds__(256) k_copy(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n16) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (U - 1) * stride < n16; i += U * stride) {
    uint4 v[U];
#pragma unroll
    for (int j = 0; j < U; ++j) v[j] = __ldcs(src + i + j * stride);
#pragma unroll
    for (int j = 0; j < U; ++j) dst[i + j * stride] = v[j];
  }
  for (; i < n16; i += stride) dst[i] = __ldcs(src + i);
}

static double now_us() {
  return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main() {
  const size_t bytes = 4333344;  // two cameras x 166 667 events x 13 B, rounded to 16
  uint8_t *h, *d;
  CK(cudaHostAlloc(&h, bytes, cudaHostAllocDefault));
  CK(cudaMalloc(&d, bytes));
  memset(h, 7, bytes);
  cudaStream_t s[8];
  for (auto& x : s) CK(cudaStreamCreateWithFlags(&x, cudaStreamNonBlocking));
  cudaEvent_t ev[8];
  for (auto& x : ev) CK(cudaEventCreateWithFlags(&x, cudaEventDisableTiming));
  const int reps = 30;
  auto idle = [] { std::this_thread::sleep_for(std::chrono::microseconds(300)); };
  auto report = [&](const char* name, std::vector<double>& t) {
    std::sort(t.begin(), t.end());
    printf("%-44s median %6.1f us  min %6.1f  max %6.1f  (%.1f GB/s at the median)\n", name, t[t.size() / 2], t[0],
           t.back(), bytes / t[t.size() / 2] / 1e3);
  };
  for (int pieces : {1, 2, 4, 8}) {
    std::vector<double> t;
    for (int r = 0; r < reps + 3; ++r) {
      idle();
      const double t0 = now_us();
      const size_t per = (bytes / pieces + 15) & ~(size_t)15;
      for (int p = 0; p < pieces; ++p) {
        const size_t off = p * per, len = std::min(per, bytes - off);
        CK(cudaMemcpyAsync(d + off, h + off, len, cudaMemcpyHostToDevice, s[p]));
        if (p) { CK(cudaEventRecord(ev[p], s[p])); CK(cudaStreamWaitEvent(s[0], ev[p], 0)); }
      }
      CK(cudaStreamSynchronize(s[0]));
      if (r >= 3) t.push_back(now_us() - t0);
    }
    char nm[64];
    snprintf(nm, sizeof nm, "cudaMemcpyAsync, %d piece(s) on %d stream(s)", pieces, pieces);
    report(nm, t);
  }
  // back to back on one stream, no idle (what scratch/h2d_bw.py measures)
  {
    const double t0 = now_us();
    for (int r = 0; r < 100; ++r) CK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s[0]));
    CK(cudaStreamSynchronize(s[0]));
    printf("%-44s %6.1f us per copy\n", "cudaMemcpyAsync back to back", (now_us() - t0) / 100);
  }
  const size_t n16 = bytes / 16;
  for (int grid : {148, 296, 592, 1184}) {
    for (int u : {4, 8}) {
      std::vector<double> t;
      for (int r = 0; r < reps + 3; ++r) {
        idle();
        const double t0 = now_us();
        if (u == 4) k_copy<4><<<grid, 256, 0, s[0]>>>((const uint4*)h, (uint4*)d, n16);
        else k_copy<8><<<grid, 256, 0, s[0]>>>((const uint4*)h, (uint4*)d, n16);
        CK(cudaStreamSynchronize(s[0]));
        if (r >= 3) t.push_back(now_us() - t0);
      }
      char nm[64];
      snprintf(nm, sizeof nm, "SM copy kernel, %d CTAs x 256, unroll %d", grid, u);
      report(nm, t);
    }
  }
  // only the (x, y) columns, as k_bin_hist would read them straight from the host block
  {
    std::vector<double> t;
    const size_t part = (bytes * 4 / 13) / 16;
    for (int r = 0; r < reps + 3; ++r) {
      idle();
      const double t0 = now_us();
      k_copy<8><<<592, 256, 0, s[0]>>>((const uint4*)h, (uint4*)d, part);
      CK(cudaStreamSynchronize(s[0]));
      if (r >= 3) t.push_back(now_us() - t0);
    }
    report("SM copy kernel, 4/13 of the bytes", t);
  }
  // empty kernel + sync: the floor of any launch-and-wait on this box
  {
    std::vector<double> t;
    for (int r = 0; r < reps + 3; ++r) {
      idle();
      const double t0 = now_us();
      k_copy<4><<<1, 32, 0, s[0]>>>((const uint4*)h, (uint4*)d, 0);
      CK(cudaStreamSynchronize(s[0]));
      if (r >= 3) t.push_back(now_us() - t0);
    }
    report("empty kernel + synchronize", t);
  }
  // verify the SM copy
  std::vector<uint8_t> back(bytes);
  CK(cudaMemset(d, 0, bytes));
  k_copy<8><<<592, 256, 0, s[0]>>>((const uint4*)h, (uint4*)d, n16);
  CK(cudaMemcpyAsync(back.data(), d, bytes, cudaMemcpyDeviceToHost, s[0]));
  CK(cudaStreamSynchronize(s[0]));
  printf("SM copy %s\n", memcmp(back.data(), h, n16 * 16) == 0 ? "verified" : "MISMATCH");
  return 0;
}
