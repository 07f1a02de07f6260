// events.cu -- per-event stages of the front-end as sm_100a kernels:
//   K0 bin_events      stable counting sort of a window's events into 32x8-pixel tiles
//   K1 sae_update_ts   createSAE_left/right + SAEtoTimeSurface_left/right, fused
//                      (feature_tracker/src/event_detector/event_detector.cc:149-166,
//                       212-228, 230-305), one warp per tile, tile state TMA-staged in smem
//   K2 corner_flags    EventDetector::isCorner (Arc*) for every left event
//                      (event_detector.cc:308-544)
#include "common.cuh"

namespace esvio {

// =====================================================================================
// K0: stable counting sort by tile
// =====================================================================================
// Same-pixel events must be applied in stream order (acceptance of an event depends on
// the previous same- and opposite-polarity event at its pixel, event_detector.cc:157),
// so the sort is stable: CTA c owns events [c*2048, (c+1)*2048), warp w of it owns the
// 256 consecutive events [w*256, (w+1)*256) and walks them 32 at a time.

__device__ __forceinline__ int tile_of(const BinLayout& L, int x, int y) {
  if (x >= L.W || y >= L.H) return L.n_tiles;  // dropped (the reference would index out of range)
  return (y / kTileH) * L.tiles_x + (x / kTileW);
}

__global__ void __launch_bounds__(kChunkThreads)
k_bin_hist(BinLayout L, DevEvents ev0, DevEvents ev1, uint32_t* __restrict__ counts) {
  extern __shared__ uint32_t s_hist[];
  const int cam = blockIdx.y;
  const DevEvents& ev = cam ? ev1 : ev0;
  const int chunk = blockIdx.x;
  const int nb = L.n_tiles + 1;
  if ((long long)chunk * kChunk >= ev.n) return;
  for (int b = threadIdx.x; b < nb; b += blockDim.x) s_hist[b] = 0;
  __syncthreads();
  const int base = chunk * kChunk;
#pragma unroll
  for (int k = 0; k < kChunkSteps; ++k) {
    const int i = base + k * kChunkThreads + threadIdx.x;
    if (i < ev.n) {
      int x, y;
      if (ev.aos) {
        const uint32_t xy = __ldg(reinterpret_cast<const uint32_t*>(ev.aos + i));
        x = xy & 0xffffu;
        y = xy >> 16;
      } else {
        x = __ldg(ev.x + i);
        y = __ldg(ev.y + i);
      }
      atomicAdd(&s_hist[tile_of(L, x, y)], 1u);
    }
  }
  __syncthreads();
  uint32_t* out = counts + (size_t)cam * nb * L.max_chunks + chunk;
  for (int b = threadIdx.x; b < nb; b += blockDim.x) out[(size_t)b * L.max_chunks] = s_hist[b];
}

// one warp per bin: exclusive scan of the bin's per-chunk counts, in place
__global__ void __launch_bounds__(256)
k_bin_scan_chunks(BinLayout L, int n_chunks0, int n_chunks1, uint32_t* __restrict__ counts,
                  uint32_t* __restrict__ bin_total) {
  const int cam = blockIdx.y;
  const int nb = L.n_tiles + 1;
  const int bin = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (bin >= nb) return;
  const int n_chunks = cam ? n_chunks1 : n_chunks0;
  const int lane = lane_id();
  uint32_t* row = counts + ((size_t)cam * nb + bin) * L.max_chunks;
  uint32_t carry = 0;
  for (int c0 = 0; c0 < n_chunks; c0 += 32) {
    const int c = c0 + lane;
    const uint32_t v = c < n_chunks ? row[c] : 0u;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    if (c < n_chunks) row[c] = carry + incl - v;
    carry += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (lane == 0) bin_total[cam * nb + bin] = carry;
}

// one CTA per camera: exclusive scan of the bin totals -> bin_start[0..nb]
__global__ void __launch_bounds__(1024)
k_bin_scan_bins(BinLayout L, const uint32_t* __restrict__ bin_total,
                uint32_t* __restrict__ bin_start) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_carry;
  const int cam = blockIdx.x;
  const int nb = L.n_tiles + 1;
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int b0 = 0; b0 < nb; b0 += blockDim.x) {
    const int b = b0 + threadIdx.x;
    const uint32_t v = b < nb ? bin_total[cam * nb + b] : 0u;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = s_warp[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, w, d);
        if (lane >= d) w += o;
      }
      s_warp[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const uint32_t warp_off = warp ? s_warp[warp - 1] : 0u;
    const uint32_t carry = s_carry;
    if (b < nb) bin_start[cam * (nb + 1) + b] = carry + warp_off + incl - v;
    __syncthreads();
    if (threadIdx.x == 0) s_carry = carry + s_warp[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) bin_start[cam * (nb + 1) + nb] = s_carry;
}

__global__ void __launch_bounds__(kChunkThreads)
k_bin_scatter(BinLayout L, DevEvents ev0, DevEvents ev1, const uint32_t* __restrict__ counts,
              const uint32_t* __restrict__ bin_start, double* __restrict__ bt0,
              uint16_t* __restrict__ bk0, double* __restrict__ bt1, uint16_t* __restrict__ bk1) {
  extern __shared__ uint32_t s_wh[];  // [8 warps][nb]
  const int cam = blockIdx.y;
  const DevEvents& ev = cam ? ev1 : ev0;
  double* __restrict__ bt = cam ? bt1 : bt0;
  uint16_t* __restrict__ bk = cam ? bk1 : bk0;
  const int chunk = blockIdx.x;
  const int nb = L.n_tiles + 1;
  if ((long long)chunk * kChunk >= ev.n) return;
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  for (int b = threadIdx.x; b < nb * 8; b += blockDim.x) s_wh[b] = 0;
  __syncthreads();

  uint32_t* my = s_wh + warp * nb;
  int bin[kChunkSteps];
  uint32_t key[kChunkSteps];
  double tt[kChunkSteps];
  const int wbase = chunk * kChunk + warp * (32 * kChunkSteps);
#pragma unroll
  for (int k = 0; k < kChunkSteps; ++k) {
    const int i = wbase + k * 32 + lane;
    bin[k] = -1 - lane;  // unique per lane: matches nothing
    key[k] = 0;
    tt[k] = 0.0;
    if (i < ev.n) {
      const Ev e = load_event(ev, i);
      bin[k] = tile_of(L, e.x, e.y);
      key[k] = (uint32_t)((e.y % kTileH) * kTileW + (e.x % kTileW)) | ((uint32_t)e.p << 8);
      tt[k] = e.t;
    }
    const uint32_t m = __match_any_sync(0xffffffffu, bin[k]);
    if (bin[k] >= 0 && (m & lt_mask) == 0) my[bin[k]] += __popc(m);
    __syncwarp();
  }
  __syncthreads();
  // per-bin exclusive prefix over the 8 warps, on top of the global offsets
  const uint32_t* cnt_row = counts + (size_t)cam * nb * L.max_chunks + chunk;
  for (int b = threadIdx.x; b < nb; b += blockDim.x) {
    uint32_t run = bin_start[cam * (nb + 1) + b] + cnt_row[(size_t)b * L.max_chunks];
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const uint32_t c = s_wh[w * nb + b];
      s_wh[w * nb + b] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kChunkSteps; ++k) {
    const uint32_t m = __match_any_sync(0xffffffffu, bin[k]);
    uint32_t pos = 0;
    if (bin[k] >= 0) pos = my[bin[k]] + __popc(m & lt_mask);
    __syncwarp();
    if (bin[k] >= 0 && (m & lt_mask) == 0) my[bin[k]] += __popc(m);
    __syncwarp();
    if (bin[k] >= 0 && bin[k] < L.n_tiles) {
      bt[pos] = tt[k];
      bk[pos] = (uint16_t)key[k];
    }
  }
}

int bin_configure(int n_tiles) {
  const size_t bytes = (size_t)(n_tiles + 1) * 8 * sizeof(uint32_t);
  if (bytes > 200 * 1024) return -1;
  return cudaFuncSetAttribute(k_bin_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)bytes) == cudaSuccess ? 0 : -1;
}

void launch_bin_events(const BinLayout& L, const EventStageBuffers& B, const DevEvents ev[2],
                       cudaStream_t s, int64_t* launches) {
  const int nb = L.n_tiles + 1;
  const int c0 = (ev[0].n + kChunk - 1) / kChunk, c1 = (ev[1].n + kChunk - 1) / kChunk;
  const int nc = c0 > c1 ? c0 : c1;
  if (nc > 0) {
    k_bin_hist<<<dim3(nc, 2), kChunkThreads, nb * sizeof(uint32_t), s>>>(L, ev[0], ev[1],
                                                                          B.counts);
    ++*launches;
  }
  k_bin_scan_chunks<<<dim3((nb + 7) / 8, 2), 256, 0, s>>>(L, c0, c1, B.counts, B.bin_total);
  k_bin_scan_bins<<<2, 1024, 0, s>>>(L, B.bin_total, B.bin_start);
  *launches += 2;
  if (nc > 0) {
    k_bin_scatter<<<dim3(nc, 2), kChunkThreads, (size_t)nb * 8 * sizeof(uint32_t), s>>>(
        L, ev[0], ev[1], B.counts, B.bin_start, B.bt[0], B.bk[0], B.bt[1], B.bk[1]);
    ++*launches;
  }
}

// =====================================================================================
// K1: fused SAE update + time surface
// =====================================================================================
constexpr int kSaeWarps = 4;

// convertTo(CV_8U) of a double: cvRound (half to even) then saturate
__device__ __forceinline__ uint8_t sat_u8(double v) {
  int r = __double2int_rn(v);
  r = r < 0 ? 0 : (r > 255 ? 255 : r);
  return (uint8_t)r;
}

struct SaeMaps {
  CUtensorMap sae, lat;  // f64 [cams][H][2W], box {64, 8, 1}
};

__global__ void __launch_bounds__(kSaeWarps * 32)
k_sae_update_ts(const __grid_constant__ SaeMaps maps, SaeTsParams P) {
  // per warp: sae tile [8][32] double2 (4 KB) + lat tile (4 KB)
  __shared__ __align__(128) double2 s_tiles[kSaeWarps][2][kTilePx];
  __shared__ __align__(8) uint64_t s_bar[kSaeWarps];

  const int warp = threadIdx.x >> 5, lane = lane_id();
  const int task = blockIdx.x * kSaeWarps + warp;
  if (task >= 2 * P.n_tiles) return;
  const int cam = task / P.n_tiles, tile = task - cam * P.n_tiles;
  const int ty = tile / P.tiles_x, tx = tile - ty * P.tiles_x;
  const int x0 = tx * kTileW, y0 = ty * kTileH;

  const uint32_t* bs = P.bin_start + cam * (P.n_tiles + 2);
  const int seg_begin = (int)bs[tile], seg_end = (int)bs[tile + 1];
  const bool dirty = seg_end > seg_begin;

  double2* s_sae = s_tiles[warp][0];
  double2* s_lat = s_tiles[warp][1];
  uint64_t* bar = &s_bar[warp];
  if (lane == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  __syncwarp();
  if (lane == 0) {
    mbar_expect_tx(bar, (dirty ? 2u : 1u) * (uint32_t)(kTilePx * sizeof(double2)));
    tma_load_3d(s_sae, &maps.sae, bar, 2 * x0, y0, cam);
    if (dirty) tma_load_3d(s_lat, &maps.lat, bar, 2 * x0, y0, cam);
  }
  // overlap the first event batch with the tile load
  const double* __restrict__ bt = P.bt[cam];
  const uint16_t* __restrict__ bk = P.bk[cam];
  double t_next = 0.0;
  uint32_t k_next = 0;
  if (seg_begin + lane < seg_end) {
    t_next = __ldg(bt + seg_begin + lane);
    k_next = __ldg(bk + seg_begin + lane);
  }
  mbar_wait(bar, 0);

  const uint32_t lt_mask = (1u << lane) - 1u;
  for (int base = seg_begin; base < seg_end; base += 32) {
    const bool valid = base + lane < seg_end;
    const double t = t_next;
    const uint32_t key = k_next;
    if (base + 32 + lane < seg_end) {
      t_next = __ldg(bt + base + 32 + lane);
      k_next = __ldg(bk + base + 32 + lane);
    }
    const int pix = key & 0xff, pol = (key >> 8) & 1;
    // events of one pixel inside this batch are applied in stream order
    const uint32_t m = __match_any_sync(0xffffffffu, valid ? pix : (0x100 + lane));
    const int rank = __popc(m & lt_mask);
    const int max_rank = __reduce_max_sync(0xffffffffu, rank);
    for (int r = 0; r <= max_rank; ++r) {
      if (valid && rank == r) {
        double* lat = reinterpret_cast<double*>(&s_lat[pix]);
        const double prev_same = lat[pol], prev_opp = lat[1 - pol];
        if (t > prev_same + P.filter_threshold || prev_opp > prev_same)
          reinterpret_cast<double*>(&s_sae[pix])[pol] = t;
        lat[pol] = t;
      }
      __syncwarp();
    }
  }

  // time surface of the tile straight from shared memory
  uint8_t* __restrict__ ts = P.ts[cam];
  const int x = x0 + lane;
#pragma unroll
  for (int r = 0; r < kTileH; ++r) {
    const double2 v = s_sae[r * kTileW + lane];
    const bool pos_newer = v.y > v.x;
    const double stamp = pos_newer ? v.y : v.x;
    double e = 0.0;
    if (stamp > 0.0) {
      const double dt = P.t_ref - stamp;
      e = exp(-dt / P.decay_sec);
      if (!P.ignore_polarity && !pos_newer) e = -e;
    }
    const double scaled = P.ignore_polarity ? e * 255.0 : e * 127.5 + 127.5;
    const int y = y0 + r;
    if (x < P.W && y < P.H) ts[(size_t)y * P.ts_pitch + x] = sat_u8(scaled);
  }

  if (dirty) {
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_3d(&maps.sae, s_sae, 2 * x0, y0, cam);
      tma_store_3d(&maps.lat, s_lat, 2 * x0, y0, cam);
      tma_store_commit();
      tma_store_wait_read0();
    }
    __syncwarp();
  }
}

void launch_sae_update_ts(const SaeTsParams& P, const CUtensorMap& map_sae,
                          const CUtensorMap& map_lat, cudaStream_t s, int64_t* launches) {
  SaeMaps maps;
  maps.sae = map_sae;
  maps.lat = map_lat;
  const int tasks = 2 * P.n_tiles;
  k_sae_update_ts<<<(tasks + kSaeWarps - 1) / kSaeWarps, kSaeWarps * 32, 0, s>>>(maps, P);
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
template <int N, int LO, int HI>
__device__ __forceinline__ bool arc_ring_valid(const double* ring) {
  int newest = 0;
#pragma unroll
  for (int i = 1; i < N; ++i)
    if (ring[i] > ring[newest]) newest = i;
  double seg_min = ring[newest];
  int cw = (newest + 1) % N, ccw = (newest + N - 1) % N;
  double cw_v = ring[cw], ccw_v = ring[ccw], cw_min = cw_v, ccw_min = ccw_v;
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
      cw_v = ring[cw];
      cw_min = fmin(cw_min, cw_v);
    } else {
      ccw = (ccw + N - 1) % N;
      ccw_v = ring[ccw];
      ccw_min = fmin(ccw_min, ccw_v);
    }
  }
  return seg_len <= HI || (seg_len >= N - HI && seg_len <= N - LO);
}

__global__ void __launch_bounds__(128)
k_corner_flags(CornerParams P, DevEvents ev, uint8_t* __restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ev.n) return;
  const Ev e = load_event(ev, i);
  uint8_t out = 0;
  do {
    if (e.x >= P.W || e.y >= P.H) break;
    const size_t px = (size_t)e.x + (size_t)e.y * P.W;
    if (P.and_ts_test && (double)P.ts[(size_t)e.y * P.ts_pitch + e.x] == P.ts_lk_threshold) break;
    const double2 l = P.lat[px];
    const double last_same = e.p ? l.y : l.x, last_opp = e.p ? l.x : l.y;
    if (e.t > last_same + P.filter_threshold || last_opp > last_same) break;
    const int border = P.min_dist + 1;
    if (e.x < border || e.x >= P.W - border || e.y < border || e.y >= P.H - border) break;
    const double* S = reinterpret_cast<const double*>(P.sae) + e.p;
    double ring[20];
#pragma unroll
    for (int k = 0; k < 16; ++k)
      ring[k] = S[2 * ((size_t)(e.x + c_ring3[k][0]) + (size_t)(e.y + c_ring3[k][1]) * P.W)];
    if (!arc_ring_valid<16, 4, 6>(ring)) break;
#pragma unroll
    for (int k = 0; k < 20; ++k)
      ring[k] = S[2 * ((size_t)(e.x + c_ring4[k][0]) + (size_t)(e.y + c_ring4[k][1]) * P.W)];
    if (!arc_ring_valid<20, 5, 8>(ring)) break;
    out = 1;
  } while (0);
  flags[i] = out;
}

void launch_corner_flags(const CornerParams& P, const DevEvents& ev, uint8_t* flags,
                         cudaStream_t s, int64_t* launches) {
  if (ev.n <= 0) return;
  k_corner_flags<<<(ev.n + 127) / 128, 128, 0, s>>>(P, ev, flags);
  ++*launches;
}

}  // namespace esvio
