// ransac.cu -- FeatureTracker::rejectWithF_event (feature_tracker/src/feature_tracker.cpp:
// 910-947): lift prev/cur points through the pinhole model onto a virtual f=460 camera and
// keep the inliers of cv::findFundamentalMat(FM_RANSAC, F_THRESHOLD, 0.99).
//
// OpenCV (modules/calib3d/src/fundam.cpp + ptsetreg.cpp; not in the reference tree) runs a
// sequential, adaptively shortened loop of 7-point hypotheses driven by cv::RNG(-1); for
// fewer than 15 points it switches to LMedS.  That loop is replayed here in parallel stages
// (four launches) whose result is the model the sequential loop would have picked:
//   k_ransac_len      the RNG stream of cv::RNG(-1) is a constant, so its first 15360 raw
//                     draws are a table in HBM; reduce them mod n and find for every stream
//                     offset how many draws the "7 distinct indices" rule consumes
//   k_ransac_chase    one CTA chases the offsets to the start of every sample (attempt)
//   k_ransac_hyp      8 attempts per CTA, one warp each: degeneracy test, 7-point solve with
//                     the elimination spread over the lanes, scores of its <= 3 models against
//                     all points
//   k_ransac_fold     one CTA: iteration index = prefix count of valid samples, running best
//                     = prefix max of inlier counts, iteration budget = RANSACUpdateNumIters
//                     of that prefix max; the first sample past the budget ends the loop; the
//                     winning model's inlier mask compacts the tracks (reduceVector)
// The 2-D null space of the 7x9 system comes from Gauss-Jordan elimination with complete
// pivoting instead of an SVD: any basis of the same null space gives the same F matrices.
#include <float.h>

#include "common.cuh"

namespace esvio {

constexpr int kModelPts = 7;
constexpr int kDrawsPerThread = 15;
constexpr int kNumDraws = 1024 * kDrawsPerThread;  // 15360 raw draws of cv::RNG(-1)
constexpr int kMaxLen = 48;                        // cap on draws consumed by one sample
constexpr int kMaxAttempts = 1280;
constexpr int kHyp = 8;   // attempts (= warps) per CTA in k_ransac_hyp: 160 CTAs for the 1280 attempts, one per
                          // SM (scoring 3 models x 150 points per attempt in f64 is the bulk of the kernel:
                          // with 32 attempts per CTA only 40 SMs worked, 31 us instead of 10)
// haveCollinearPoints: is the last point collinear with any earlier pair
__device__ bool collinear_with_last(const float2* m, int count) {
  const int i = count - 1;
  for (int j = 0; j < i; ++j) {
    const double dx1 = (double)m[j].x - (double)m[i].x, dy1 = (double)m[j].y - (double)m[i].y;
    for (int k = 0; k < j; ++k) {
      const double dx2 = (double)m[k].x - (double)m[i].x, dy2 = (double)m[k].y - (double)m[i].y;
      if (fabs(dx2 * dy1 - dy2 * dx1) <=
          (double)FLT_EPSILON * (fabs(dx1) + fabs(dy1) + fabs(dx2) + fabs(dy2)))
        return true;
    }
  }
  return false;
}

__device__ int solve_cubic(const double* c, double* roots) {
  double a0 = c[0], a1 = c[1], a2 = c[2], a3 = c[3];
  double x0 = 0, x1 = 0, x2 = 0;
  int n = 0;
  const double kPi = 3.14159265358979323846;
  if (a0 == 0) {
    if (a1 == 0) {
      if (a2 == 0) n = a3 == 0 ? -1 : 0;
      else {
        x0 = -a3 / a2;
        n = 1;
      }
    } else {
      double d = a2 * a2 - 4 * a1 * a3;
      if (d >= 0) {
        d = sqrt(d);
        const double q1 = (-a2 + d) * 0.5, q2 = (a2 + d) * -0.5;
        if (fabs(q1) > fabs(q2)) {
          x0 = q1 / a1;
          x1 = a3 / q1;
        } else {
          x0 = q2 / a1;
          x1 = a3 / q2;
        }
        n = d > 0 ? 2 : 1;
      }
    }
  } else {
    a0 = 1. / a0;
    a1 *= a0;
    a2 *= a0;
    a3 *= a0;
    const double Q = (a1 * a1 - 3 * a2) * (1. / 9);
    const double R = (2 * a1 * a1 * a1 - 9 * a1 * a2 + 27 * a3) * (1. / 54);
    const double Qc = Q * Q * Q;
    double d = Qc - R * R;
    if (d > 0) {
      const double theta = acos(R / sqrt(Qc));
      const double sq = sqrt(Q);
      const double t0 = -2 * sq, t1 = theta * (1. / 3), t2 = a1 * (1. / 3);
      x0 = t0 * cos(t1) - t2;
      x1 = t0 * cos(t1 + (2. * kPi / 3)) - t2;
      x2 = t0 * cos(t1 + (4. * kPi / 3)) - t2;
      n = 3;
    } else if (d == 0) {
      if (R >= 0) {
        x0 = -2 * pow(R, 1. / 3) - a1 / 3;
        x1 = pow(R, 1. / 3) - a1 / 3;
      } else {
        x0 = 2 * pow(-R, 1. / 3) - a1 / 3;
        x1 = -pow(-R, 1. / 3) - a1 / 3;
      }
      x2 = 0;
      n = x0 == x1 ? 1 : 2;
      x1 = x0 == x1 ? 0 : x1;
    } else {
      d = sqrt(-d);
      double e = pow(d + fabs(R), 1. / 3);
      if (R > 0) e = -e;
      x0 = (e + Q / e) - a1 * (1. / 3);
      n = 1;
    }
  }
  roots[0] = x0;
  roots[1] = x1;
  roots[2] = x2;
  return n;
}

// FMEstimatorCallback::runKernel for 7 points (run7Point): up to 3 matrices, row-major.
// One WARP per sample.  The 7x9 system sits in shared memory (sA, 63 doubles; sperm, 9 ints):
// Gauss-Jordan with complete pivoting, every step spread over the lanes -- the pivot is the
// first maximum of |A[i][j]|, i, j >= k, in row-major order (warp arg-max, ties to the lower
// index, like the serial scan with its strict >), row / column swaps, the scaling of the
// pivot row and the elimination touch each element exactly once per step with the same two
// operations (multiply, subtract) the serial code applies, so the result is bit-identical to
// one thread doing it alone.  That serial version -- 32 samples on the 32 lanes of one warp,
// the matrix in local memory because the pivots index it dynamically -- took ~40 000 cycles per
// CTA while the other 31 warps waited; this one ~4 000, all 32 warps busy.
// m1 / m2: the sample (the same values in every lane).  Returns the number of models
// (warp-uniform); the models land in Fout (shared memory, 27 doubles).
__device__ int run_7point_warp(const float2* m1, const float2* m2, double* sA, int* sperm, double* Fout) {
  const int lane = lane_id();
  double c1x = 0, c1y = 0, c2x = 0, c2y = 0;
  for (int i = 0; i < 7; ++i) {
    c1x += m1[i].x;
    c1y += m1[i].y;
    c2x += m2[i].x;
    c2y += m2[i].y;
  }
  const double t = 1. / 7;
  c1x *= t, c1y *= t, c2x *= t, c2y *= t;
  double s1 = 0, s2 = 0;
  for (int i = 0; i < 7; ++i) {
    const double ax = m1[i].x - c1x, ay = m1[i].y - c1y;
    const double bx = m2[i].x - c2x, by = m2[i].y - c2y;
    s1 += sqrt(ax * ax + ay * ay);
    s2 += sqrt(bx * bx + by * by);
  }
  s1 *= t;
  s2 *= t;
  if (s1 < FLT_EPSILON || s2 < FLT_EPSILON) return 0;
  s1 = sqrt(2.) / s1;
  s2 = sqrt(2.) / s2;

  // row i of the system by lane i
#pragma unroll
  for (int i = 0; i < 7; ++i)
    if (lane == i) {
      const double x0 = (m1[i].x - c1x) * s1, y0 = (m1[i].y - c1y) * s1;
      const double x1 = (m2[i].x - c2x) * s2, y1 = (m2[i].y - c2y) * s2;
      double* r = sA + 9 * i;
      r[0] = x1 * x0, r[1] = x1 * y0, r[2] = x1;
      r[3] = y1 * x0, r[4] = y1 * y0, r[5] = y1;
      r[6] = x0, r[7] = y0, r[8] = 1;
    }
  if (lane < 9) sperm[lane] = lane;
  __syncwarp();
  // Gauss-Jordan with complete pivoting -> [I | C] in permuted columns
  for (int k = 0; k < 7; ++k) {
    // lane owns elements e = lane and lane + 32 (e = 9 i + j)
    double best = -1;
    int best_e = 63;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int e = lane + 32 * q;
      if (e < 63) {
        const int i = e / 9, j = e - 9 * i;
        if (i >= k && j >= k) {
          const double v = fabs(sA[e]);
          if (v > best) best = v, best_e = e;  // e grows with q: the earlier element wins a tie
        }
      }
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, d);
      const int oe = __shfl_xor_sync(0xffffffffu, best_e, d);
      if (ob > best || (ob == best && oe < best_e)) best = ob, best_e = oe;
    }
    if (!(best > 1e-14)) return 0;  // rank deficient sample
    const int pi = best_e / 9, pj = best_e - 9 * pi;
    if (pi != k && lane < 9) {
      const double tmp = sA[9 * k + lane];
      sA[9 * k + lane] = sA[9 * pi + lane];
      sA[9 * pi + lane] = tmp;
    }
    __syncwarp();
    if (pj != k) {
      if (lane < 7) {
        const double tmp = sA[9 * lane + k];
        sA[9 * lane + k] = sA[9 * lane + pj];
        sA[9 * lane + pj] = tmp;
      } else if (lane == 7) {
        const int tp = sperm[k];
        sperm[k] = sperm[pj];
        sperm[pj] = tp;
      }
    }
    __syncwarp();
    const double inv = 1.0 / sA[9 * k + k];
    __syncwarp();
    if (lane >= k && lane < 9) sA[9 * k + lane] *= inv;
    __syncwarp();
    // A[i][j] -= A[i][k] * A[k][j] for i != k, j >= k: read everything, then write
    double f[2], akj[2];
    int ee[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int e = lane + 32 * q;
      ee[q] = -1;
      f[q] = 0, akj[q] = 0;
      if (e < 63) {
        const int i = e / 9, j = e - 9 * i;
        if (i != k && j >= k) {
          ee[q] = e;
          f[q] = sA[9 * i + k];
          akj[q] = sA[9 * k + j];
        }
      }
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 2; ++q)
      if (ee[q] >= 0 && f[q] != 0.0) sA[ee[q]] -= f[q] * akj[q];
    __syncwarp();
  }
  double f1[9], f2[9];
#pragma unroll
  for (int j = 0; j < 9; ++j) f1[j] = 0, f2[j] = 0;
  for (int j = 0; j < 7; ++j) {
    const int pj = sperm[j];
    const double a7 = -sA[9 * j + 7], a8 = -sA[9 * j + 8];
#pragma unroll
    for (int q = 0; q < 9; ++q)
      if (q == pj) f1[q] = a7, f2[q] = a8;
  }
  {
    const int p7 = sperm[7], p8 = sperm[8];
#pragma unroll
    for (int q = 0; q < 9; ++q) {
      if (q == p7) f1[q] = 1, f2[q] = 0;
      if (q == p8) f1[q] = 0, f2[q] = 1;
    }
  }
  __syncwarp();

  for (int i = 0; i < 9; ++i) f1[i] -= f2[i];
  double c[4], r[3] = {0, 0, 0};
  double t0 = f2[4] * f2[8] - f2[5] * f2[7];
  double t1 = f2[3] * f2[8] - f2[5] * f2[6];
  double t2 = f2[3] * f2[7] - f2[4] * f2[6];
  c[3] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2;
  c[2] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2 - f1[3] * (f2[1] * f2[8] - f2[2] * f2[7]) +
         f1[4] * (f2[0] * f2[8] - f2[2] * f2[6]) - f1[5] * (f2[0] * f2[7] - f2[1] * f2[6]) +
         f1[6] * (f2[1] * f2[5] - f2[2] * f2[4]) - f1[7] * (f2[0] * f2[5] - f2[2] * f2[3]) +
         f1[8] * (f2[0] * f2[4] - f2[1] * f2[3]);
  t0 = f1[4] * f1[8] - f1[5] * f1[7];
  t1 = f1[3] * f1[8] - f1[5] * f1[6];
  t2 = f1[3] * f1[7] - f1[4] * f1[6];
  c[0] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2;
  c[1] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2 - f2[3] * (f1[1] * f1[8] - f1[2] * f1[7]) +
         f2[4] * (f1[0] * f1[8] - f1[2] * f1[6]) - f2[5] * (f1[0] * f1[7] - f1[1] * f1[6]) +
         f2[6] * (f1[1] * f1[5] - f1[2] * f1[4]) - f2[7] * (f1[0] * f1[5] - f1[2] * f1[3]) +
         f2[8] * (f1[0] * f1[4] - f1[1] * f1[3]);
  const int n = solve_cubic(c, r);
  if (n < 1 || n > 3) return n;
  const double T1[9] = {s1, 0, -s1 * c1x, 0, s1, -s1 * c1y, 0, 0, 1};
  const double T2[9] = {s2, 0, -s2 * c2x, 0, s2, -s2 * c2y, 0, 0, 1};
  for (int k = 0; k < n; ++k) {
    double F[9];
    double lambda = r[k], mu = 1.;
    const double s = f1[8] * r[k] + f2[8];
    double G[9];
    if (fabs(s) > DBL_EPSILON) {
      mu = 1. / s;
      lambda *= mu;
      G[8] = 1.;
    } else
      G[8] = 0.;
    for (int i = 0; i < 8; ++i) G[i] = f1[i] * lambda + f2[i] * mu;
    double M[9];
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) {
        double acc = 0;
        for (int m = 0; m < 3; ++m) acc += T2[m * 3 + a] * G[m * 3 + b];
        M[a * 3 + b] = acc;
      }
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) {
        double acc = 0;
        for (int m = 0; m < 3; ++m) acc += M[a * 3 + m] * T1[m * 3 + b];
        F[a * 3 + b] = acc;
      }
    if (fabs(F[8]) > FLT_EPSILON) {
      const double inv = 1. / F[8];
      for (int i = 0; i < 9; ++i) F[i] *= inv;
    }
#pragma unroll
    for (int i = 0; i < 9; ++i)
      if (lane == i) Fout[9 * k + i] = F[i];
  }
  __syncwarp();
  return n;
}


// FMEstimatorCallback::computeError for one correspondence
__device__ __forceinline__ float fm_error(const double* F, float2 p1, float2 p2) {
  const double x1 = p1.x, y1 = p1.y, x2 = p2.x, y2 = p2.y;
  double a = F[0] * x1 + F[1] * y1 + F[2];
  double b = F[3] * x1 + F[4] * y1 + F[5];
  double c = F[6] * x1 + F[7] * y1 + F[8];
  const double s2 = 1. / (a * a + b * b);
  const double d2 = x2 * a + y2 * b + c;
  a = F[0] * x2 + F[3] * y2 + F[6];
  b = F[1] * x2 + F[4] * y2 + F[7];
  c = F[2] * x2 + F[5] * y2 + F[8];
  const double s1 = 1. / (a * a + b * b);
  const double d1 = x1 * a + y1 * b + c;
  const double e1 = d1 * d1 * s1, e2 = d2 * d2 * s2;
  return (float)(e1 > e2 ? e1 : e2);
}

// RANSACUpdateNumIters
__device__ int ransac_update_iters(double p, double ep, int model_points, int max_iters) {
  p = fmax(p, 0.);
  p = fmin(p, 1.);
  ep = fmax(ep, 0.);
  ep = fmin(ep, 1.);
  double num = fmax(1. - p, DBL_MIN);
  double denom = 1. - pow(1. - ep, (double)model_points);
  if (denom < DBL_MIN) return 0;
  num = log(num);
  denom = log(denom);
  return (denom >= 0 || -num >= max_iters * (-denom)) ? max_iters : (int)llrint(num / denom);
}


// ------------------------------------------------------------------------------------------
// scratch shared by the three launches (HBM, one per handle)
// ------------------------------------------------------------------------------------------
enum { kModeSkip = 0, kModeRansac = 1, kModeLmeds = 2, kModeSeven = 3, kModeFail = 4 };

struct RansacScratch {
  int n, mode, n_attempts, pad;
  double thresh, confidence;
  int max_iters, pad2;
  float2 p1[kMaxCnt], p2[kMaxCnt];
  alignas(16) uint16_t idx[kMaxAttempts][8];
  alignas(16) uint16_t v[kNumDraws];  // draws reduced mod n
  alignas(16) uint8_t len[kNumDraws];  // draws consumed by a sample starting at that offset
  int valid[kMaxAttempts];
  int nmodels[kMaxAttempts];
  int good[kMaxAttempts][3];      // inlier counts (RANSAC)
  float median[kMaxAttempts][3];  // median residuals (LMedS)
  double F[kMaxAttempts][27];
};

size_t ransac_scratch_bytes() { return sizeof(RansacScratch); }

// cv::RNG(-1): state = (uint64)-1; next(): state = (unsigned)state * 4164903690 + (state >> 32)
void ransac_fill_draw_table(uint32_t* host_table) {
  uint64_t state = (uint64_t)-1;
  for (int i = 0; i < kNumDraws; ++i) {
    state = (uint64_t)(unsigned)state * 4164903690ULL + (unsigned)(state >> 32);
    host_table[i] = (unsigned)state;
  }
}
int ransac_num_draws() { return kNumDraws; }

__device__ __forceinline__ void lift_pinhole(const Pinhole& c, double u, double v, double& ox,
                                             double& oy) {
  const double inv_fx = 1.0 / c.fx, inv_fy = 1.0 / c.fy;
  const double off_x = -c.cx / c.fx, off_y = -c.cy / c.fy;
  const double xd = inv_fx * u + off_x, yd = inv_fy * v + off_y;
  double xu = xd, yu = yd;
  if (!(c.k1 == 0.0 && c.k2 == 0.0 && c.p1 == 0.0 && c.p2 == 0.0)) {
    for (int it = 0; it < 8; ++it) {
      const double xx = xu * xu, yy = yu * yu, xy = xu * yu;
      const double r2 = xx + yy;
      const double rad = c.k1 * r2 + c.k2 * r2 * r2;
      const double ddx = xu * rad + 2.0 * c.p1 * xy + c.p2 * (r2 + 2.0 * xx);
      const double ddy = yu * rad + 2.0 * c.p2 * xy + c.p1 * (r2 + 2.0 * yy);
      xu = xd - ddx;
      yu = yd - ddy;
    }
  }
  ox = xu;
  oy = yu;
}

// ------------------------------------------------------------------------------------------
// k_ransac_prepare: points + the start of every sample in the RNG stream
// ------------------------------------------------------------------------------------------
struct PrepareArgs {
  int from_tracks;          // 1: lift B.prev_pts/B.cur_pts (rejectWithF_event); 0: stage points
  const float2 *p1, *p2;    // stage mode
  int n;                    // stage mode
  double thresh;
  int min_points;           // 8 in the tracker (feature_tracker.cpp:912), 7 for the stage entry
};

// (a) every CTA owns 1024 stream offsets: reduce the draws mod n and find, for every offset, how
//     many draws the "7 distinct indices" rule of PointSetRegistrator::getSubset consumes
__global__ void __launch_bounds__(1024)
k_ransac_len(TrackBuffers B, PrepareArgs A, const uint32_t* __restrict__ draws,
             RansacScratch* __restrict__ R) {
  PDL_PROLOGUE();
  __shared__ uint16_t s_v[1024 + kMaxLen];
  const int tid = threadIdx.x;
  const int n = A.from_tracks ? B.st->n_cur : A.n;
  if (n < A.min_points || n <= 7) return;
  const int base = blockIdx.x * 1024;
  for (int i = tid; i < 1024 + kMaxLen; i += blockDim.x) {
    const int g = base + i;
    s_v[i] = g < kNumDraws ? (uint16_t)(draws[g] % (unsigned)n) : (uint16_t)0xffff;
  }
  __syncthreads();
  int got = 0, k = 0;
  uint16_t sel[kModelPts];
  while (got < kModelPts && k < kMaxLen) {
    const uint16_t c = s_v[tid + k++];
    if (c == 0xffff) break;
    bool dup = false;
#pragma unroll
    for (int q = 0; q < kModelPts; ++q) dup |= (q < got && sel[q] == c);
    if (!dup) {
#pragma unroll
      for (int q = 0; q < kModelPts; ++q)
        if (q == got) sel[q] = c;
      ++got;
    }
  }
  R->v[base + tid] = s_v[tid];
  R->len[base + tid] = (uint8_t)(got == kModelPts ? k : 0);  // 0: stream exhausted / capped
}

// (b) one CTA: lift the points, then chase the offsets to the start of every sample.  Thread 0
//     cannot afford 1280 dependent hops, so 8-hop jumps are computed for all offsets in
//     parallel, thread 0 follows those, and one thread per group of 8 fills in the rest.
__global__ void __launch_bounds__(1024)
k_ransac_chase(TrackParams P, TrackBuffers B, PrepareArgs A, RansacScratch* __restrict__ R) {
  PDL_PROLOGUE();
  extern __shared__ __align__(16) unsigned char s_dyn[];
  uint16_t* s_v = reinterpret_cast<uint16_t*>(s_dyn);  // draws reduced mod n
  uint16_t* s_hop8 = s_v + kNumDraws;                  // stream offset 8 samples further on
  uint8_t* s_len = reinterpret_cast<uint8_t*>(s_hop8 + kNumDraws);  // draws one sample consumes
  __shared__ uint16_t s_start8[kMaxAttempts / 8 + 1];
  __shared__ int s_groups, s_natt;
  const int tid = threadIdx.x;
  const int n = A.from_tracks ? B.st->n_cur : A.n;
  if (tid == 0) {
    R->n = n;
    R->thresh = A.thresh <= 0 ? 3.0 : A.thresh;
    R->confidence = 0.99;
    R->max_iters = 1000;
    R->n_attempts = 0;
    R->mode = n < A.min_points ? (A.from_tracks ? kModeSkip : kModeFail)
                               : (n == 7 ? kModeSeven : (n >= 15 ? kModeRansac : kModeLmeds));
  }
  if (n < A.min_points) return;
  for (int i = tid; i < n; i += blockDim.x) {
    if (A.from_tracks) {
      const float2 pp = B.prev_pts[i], cp = B.cur_pts[i];
      double x, y;
      lift_pinhole(P.cam[0], (double)pp.x, (double)pp.y, x, y);
      R->p1[i] = make_float2((float)(P.focal_length * x + P.W / 2.0),
                             (float)(P.focal_length * y + P.H / 2.0));
      lift_pinhole(P.cam[0], (double)cp.x, (double)cp.y, x, y);
      R->p2[i] = make_float2((float)(P.focal_length * x + P.W / 2.0),
                             (float)(P.focal_length * y + P.H / 2.0));
    } else {
      R->p1[i] = A.p1[i];
      R->p2[i] = A.p2[i];
    }
  }
  if (n == 7) {
    if (tid < 8) R->idx[0][tid] = (uint16_t)(tid < 7 ? tid : 0);
    if (tid == 0) R->n_attempts = 1;
    return;
  }
  for (int i = tid; i < kNumDraws / 8; i += blockDim.x)
    reinterpret_cast<uint4*>(s_v)[i] = reinterpret_cast<const uint4*>(R->v)[i];
  for (int i = tid; i < kNumDraws / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(s_len)[i] = reinterpret_cast<const uint4*>(R->len)[i];
  __syncthreads();
  for (int o = tid; o < kNumDraws; o += blockDim.x) {
    int p = o;
    bool ok = true;
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      const int l = (ok && p < kNumDraws) ? s_len[p] : 0;
      ok = ok && l != 0;
      p += l;
    }
    s_hop8[o] = (uint16_t)(ok && p < kNumDraws ? p : 0xffff);
  }
  if (tid == 0) s_natt = 0;
  __syncthreads();
  if (tid == 0) {
    int p = 0, g = 0;
    while (g < kMaxAttempts / 8 && p != 0xffff) {
      s_start8[g++] = (uint16_t)p;
      p = s_hop8[p];
    }
    s_groups = g;  // groups of 8 attempts start at s_start8[0..g-1]; only the last may be partial
  }
  __syncthreads();
  const int groups = s_groups;
  if (tid < groups) {
    int p = s_start8[tid];
    int cnt = 0;
    for (int h = 0; h < 8; ++h) {
      const int l = p < kNumDraws ? s_len[p] : 0;
      if (l == 0) break;
      // re-derive the 7 distinct indices of the sample starting at p
      int got = 0, k = 0;
      const int a = tid * 8 + h;
      __align__(16) uint16_t sel[8];
      while (got < kModelPts) {
        const uint16_t c = s_v[p + k++];
        bool dup = false;
#pragma unroll
        for (int q = 0; q < kModelPts; ++q) dup |= (q < got && sel[q] == c);
        if (!dup) {
#pragma unroll
          for (int q = 0; q < kModelPts; ++q)
            if (q == got) sel[q] = c;
          ++got;
        }
      }
      sel[7] = 0;
      *reinterpret_cast<uint4*>(R->idx[a]) = *reinterpret_cast<const uint4*>(sel);
      p += l;
      ++cnt;
    }
    if (cnt < 8 || tid == groups - 1) atomicMax(&s_natt, tid * 8 + cnt);
  }
  __syncthreads();
  if (tid == 0) R->n_attempts = s_natt;
}

// (c) The 7 indices of every attempt depend on the point count n and on nothing else: the
//     draws of cv::RNG(-1) are a constant and getSubset only rejects duplicates.  A handle
//     therefore runs (a) + (b) once per n when it is created (ransac_build_cache) and a window
//     just copies the row of its n -- 20 KB -- next to lifting its points.
__global__ void __launch_bounds__(1024)
k_ransac_prepare_cached(TrackParams P, TrackBuffers B, PrepareArgs A,
                        RansacScratch* __restrict__ R) {
  PDL_PROLOGUE();
  const int tid = threadIdx.x;
  const int n = A.from_tracks ? B.st->n_cur : A.n;
  if (tid == 0) {
    R->n = n;
    R->thresh = A.thresh <= 0 ? 3.0 : A.thresh;
    R->confidence = 0.99;
    R->max_iters = 1000;
    R->n_attempts = 0;
    R->mode = n < A.min_points ? (A.from_tracks ? kModeSkip : kModeFail)
                               : (n == 7 ? kModeSeven : (n >= 15 ? kModeRansac : kModeLmeds));
  }
  if (n < A.min_points) return;
  for (int i = tid; i < n; i += blockDim.x) {
    if (A.from_tracks) {
      const float2 pp = B.prev_pts[i], cp = B.cur_pts[i];
      double x, y;
      lift_pinhole(P.cam[0], (double)pp.x, (double)pp.y, x, y);
      R->p1[i] = make_float2((float)(P.focal_length * x + P.W / 2.0),
                             (float)(P.focal_length * y + P.H / 2.0));
      lift_pinhole(P.cam[0], (double)cp.x, (double)cp.y, x, y);
      R->p2[i] = make_float2((float)(P.focal_length * x + P.W / 2.0),
                             (float)(P.focal_length * y + P.H / 2.0));
    } else {
      R->p1[i] = A.p1[i];
      R->p2[i] = A.p2[i];
    }
  }
  if (n == 7) {
    if (tid < 8) R->idx[0][tid] = (uint16_t)(tid < 7 ? tid : 0);
    if (tid == 0) R->n_attempts = 1;
    return;
  }
  const uint4* __restrict__ src =
      reinterpret_cast<const uint4*>(B.rs_idx_cache) + (size_t)(n - B.rs_cache_lo) * kMaxAttempts;
  uint4* dst = reinterpret_cast<uint4*>(&R->idx[0][0]);
  for (int a = tid; a < kMaxAttempts; a += blockDim.x) dst[a] = __ldg(src + a);
  __syncthreads();
  if (tid == 0) R->n_attempts = B.rs_natt_cache[n - B.rs_cache_lo];
}

__global__ void k_ransac_cache_row(const RansacScratch* __restrict__ R, uint4* __restrict__ idx_row,
                                   int* __restrict__ natt) {
  const uint4* src = reinterpret_cast<const uint4*>(&R->idx[0][0]);
  for (int a = threadIdx.x; a < kMaxAttempts; a += blockDim.x) idx_row[a] = src[a];
  if (threadIdx.x == 0) *natt = R->n_attempts;
}

// ------------------------------------------------------------------------------------------
// k_ransac_hyp: solve and score kHyp attempts per CTA, one warp each
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * kHyp) k_ransac_hyp(RansacScratch* __restrict__ R) {
  PDL_PROLOGUE();
  __shared__ float2 s_p1[kMaxCnt], s_p2[kMaxCnt];
  __shared__ double s_F[kHyp][27];
  __shared__ double s_A[kHyp][63];
  __shared__ int s_perm[kHyp][9];
  const int mode = R->mode;
  if (mode == kModeSkip || mode == kModeFail) return;
  const int n = R->n, n_att = R->n_attempts;
  const int a0 = blockIdx.x * kHyp;
  if (a0 >= n_att) return;
  const int tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
  for (int i = tid; i < n; i += blockDim.x) {
    s_p1[i] = R->p1[i];
    s_p2[i] = R->p2[i];
  }
  __syncthreads();
  // one warp per attempt: degeneracy test, 7-point solve, then the scores of its <= 3 models
  const int a = a0 + warp;
  if (a >= n_att) return;
  float2 m1[kModelPts], m2[kModelPts];
#pragma unroll
  for (int q = 0; q < kModelPts; ++q) {
    const int id = R->idx[a][q];
    m1[q] = s_p1[id];
    m2[q] = s_p2[id];
  }
  // FMEstimatorCallback::checkSubset
  const int valid = (mode == kModeSeven) ||
                    !(collinear_with_last(m1, kModelPts) || collinear_with_last(m2, kModelPts));
  int nm = 0;
  if (valid) {
    nm = run_7point_warp(m1, m2, s_A[warp], s_perm[warp], s_F[warp]);
    nm = nm < 0 ? 0 : (nm > 3 ? 3 : nm);
  }
  if (lane == 0) {
    R->valid[a] = valid;
    R->nmodels[a] = nm;
  }
  const float t2 = (float)(R->thresh * R->thresh);
  for (int m = 0; m < nm; ++m) {
    const double* F = s_F[warp] + 9 * m;
    if (lane < 9) R->F[a][9 * m + lane] = F[lane];
    if (mode == kModeRansac) {
      int good = 0;
      for (int i = lane; i < n; i += 32) good += fm_error(F, s_p1[i], s_p2[i]) <= t2;
      good = __reduce_add_sync(0xffffffffu, good);
      if (lane == 0) R->good[a][m] = good;
    } else if (mode == kModeLmeds) {
      // n < 15: median of the residuals = element n/2 of the sorted list
      const float e = lane < n ? fm_error(F, s_p1[lane], s_p2[lane]) : FLT_MAX;
      int rank = 0;
      for (int j = 0; j < n; ++j) {
        const float o = __shfl_sync(0xffffffffu, e, j);
        rank += (o < e) || (o == e && j < lane);
      }
      if (lane < n && rank == n / 2) R->median[a][m] = e;
    }
  }
}

// ------------------------------------------------------------------------------------------
// k_ransac_fold: replay the sequential loop's decisions, build the mask, compact the tracks
// ------------------------------------------------------------------------------------------
struct FoldArgs {
  int to_tracks;
  uint8_t* mask;  // stage mode
  int* iters;     // stage mode
};

__global__ void __launch_bounds__(1024)
k_ransac_fold(TrackParams P, TrackBuffers B, FoldArgs A, RansacScratch* __restrict__ R) {
  PDL_PROLOGUE();
  __shared__ int s_warp[33];
  __shared__ int s_pm[kMaxAttempts];   // per attempt: best inlier count among its models, -1 none
  __shared__ int s_it[kMaxAttempts];   // iteration index (exclusive prefix count of valid)
  __shared__ uint8_t s_mask[kMaxCnt];
  __shared__ double s_bestF[9];
  __shared__ int s_best_a, s_best_m, s_iters, s_result, s_first_dead;
  __shared__ float s_thr2;
  const int tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
  const int mode = R->mode, n = R->n;
  if (mode == kModeSkip) return;
  const int n_att = R->n_attempts;
  if (tid == 0) {
    s_best_a = -1;
    s_best_m = 0;
    s_iters = 0;
    s_result = 0;
    s_first_dead = 0x7fffffff;
    s_thr2 = (float)(R->thresh * R->thresh);
  }
  for (int i = tid; i < kMaxCnt; i += blockDim.x) s_mask[i] = 0;
  __syncthreads();
  if (mode == kModeSeven) {
    if (tid == 0) {
      s_result = R->nmodels[0] > 0;
      s_iters = 1;
    }
    for (int i = tid; i < n; i += blockDim.x) s_mask[i] = 1;
    __syncthreads();
  } else if (mode == kModeRansac || mode == kModeLmeds) {
    // iteration index of every attempt: exclusive prefix count of valid samples.  Done by one
    // warp over <= 1280 entries (40 strides); the rest is embarrassingly parallel.
    if (warp == 0 && mode == kModeLmeds) {
      int carry = 0;
      for (int base = 0; base < n_att; base += 32) {
        const int a = base + lane;
        const int v = a < n_att ? R->valid[a] : 0;
        const unsigned bal = __ballot_sync(0xffffffffu, v);
        if (a < n_att) s_it[a] = carry + __popc(bal & ((1u << lane) - 1u));
        carry += __popc(bal);
      }
    }
    __syncthreads();
    if (mode == kModeRansac) {
      // Sequential semantics: the budget after an attempt is RANSACUpdateNumIters of the best
      // inlier count so far (only counts > 6 update it); an attempt runs iff its iteration
      // index is below the budget left by its predecessors.  The budget never grows and the
      // index never shrinks, so the loop ends at the first valid attempt that fails the test.
      // Every thread owns two consecutive attempts; prefix count / prefix max are block scans.
      const int a0 = 2 * tid;
      int v[2], g[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int a = a0 + e;
        v[e] = a < n_att ? R->valid[a] : 0;
        g[e] = -1;
        if (v[e])
          for (int m = 0; m < R->nmodels[a]; ++m) g[e] = max(g[e], R->good[a][m]);
      }
      // exclusive block scan of (count of valid, max of g) over threads
      int isum = v[0] + v[1], imax = max(g[0], g[1]);
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int os = __shfl_up_sync(0xffffffffu, isum, d);
        const int om = __shfl_up_sync(0xffffffffu, imax, d);
        if (lane >= d) {
          isum += os;
          imax = max(imax, om);
        }
      }
      if (lane == 31) {
        s_it[warp] = isum;
        s_pm[warp] = imax;
      }
      __syncthreads();
      if (warp == 0) {
        int ws = s_it[lane], wm = s_pm[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int os = __shfl_up_sync(0xffffffffu, ws, d);
          const int om = __shfl_up_sync(0xffffffffu, wm, d);
          if (lane >= d) {
            ws += os;
            wm = max(wm, om);
          }
        }
        s_it[32 + lane] = ws;  // inclusive over warps
        s_pm[32 + lane] = wm;
      }
      __syncthreads();
      int esum = __shfl_up_sync(0xffffffffu, isum, 1), emax = __shfl_up_sync(0xffffffffu, imax, 1);
      if (lane == 0) {
        esum = 0;
        emax = -1;
      }
      if (warp > 0) {
        esum += s_it[32 + warp - 1];
        emax = max(emax, s_pm[32 + warp - 1]);
      }
      const double conf = R->confidence;
      const int max_it = R->max_iters;
      int it[2] = {esum, esum + v[0]};
      int bb[2];
      bb[0] = max(kModelPts - 1, emax);
      bb[1] = max(bb[0], g[0]);
      bool runs[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int budget = bb[e] > kModelPts - 1
                               ? ransac_update_iters(conf, (double)(n - bb[e]) / n, kModelPts, max_it)
                               : max_it;
        runs[e] = v[e] && it[e] < budget;
        if (v[e] && !runs[e]) atomicMin(&s_first_dead, a0 + e);
      }
      __syncthreads();
      const int first_dead = s_first_dead;
      int my_iters = 0, my_best = -1;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const bool counted = runs[e] && (a0 + e) < first_dead;
        my_iters += counted;
        if (counted && g[e] > bb[e]) my_best = a0 + e;  // strictly better than all before it
      }
      my_iters = __reduce_add_sync(0xffffffffu, my_iters);
      my_best = __reduce_max_sync(0xffffffffu, my_best);
      if (lane == 0) {
        if (my_iters) atomicAdd(&s_iters, my_iters);
        if (my_best >= 0) atomicMax(&s_best_a, my_best);
      }
      __syncthreads();
      if (tid == 0 && s_best_a >= 0) {
        const int ba = s_best_a;
        int bm = 0;
        for (int m = 1; m < R->nmodels[ba]; ++m)
          if (R->good[ba][m] > R->good[ba][bm]) bm = m;
        s_best_m = bm;
        s_result = 1;
      }
    } else {
      // LMedS: fixed budget; the smallest median wins, earliest on ties
      if (warp == 0) {
        int budget = ransac_update_iters(R->confidence, 0.45, kModelPts, R->max_iters);
        budget = budget < 3 ? 3 : budget;
        float best = FLT_MAX;
        int best_a = -1, best_m = 0, iters = 0;
        for (int base = 0; base < n_att; base += 32) {
          const int a = base + lane;
          const bool runs = a < n_att && R->valid[a] && s_it[a] < budget;
          iters += __popc(__ballot_sync(0xffffffffu, runs));
          float mine = FLT_MAX;
          int mine_m = 0;
          if (runs)
            for (int m = 0; m < R->nmodels[a]; ++m)
              if (R->median[a][m] < mine) mine = R->median[a][m], mine_m = m;
          // earliest lane holding the stride minimum
          float mn = mine;
#pragma unroll
          for (int d = 16; d; d >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, d));
          const unsigned who = __ballot_sync(0xffffffffu, mine == mn && mine < FLT_MAX);
          if (who && mn < best) {
            const int l = __ffs(who) - 1;
            best = mn;
            best_a = base + l;
            best_m = __shfl_sync(0xffffffffu, mine_m, l);
          }
        }
        if (lane == 0) {
          s_iters = iters;
          s_best_a = best_a;
          s_best_m = best_m;
          if (best_a >= 0) {
            double sigma = 2.5 * 1.4826 * (1 + 5. / (n - kModelPts)) * sqrt((double)best);
            sigma = fmax(sigma, 0.001);
            s_thr2 = (float)(sigma * sigma);
          }
        }
      }
    }
    __syncthreads();
    if (s_best_a >= 0) {
      if (tid < 9) s_bestF[tid] = R->F[s_best_a][9 * s_best_m + tid];
      __syncthreads();
      for (int i = tid; i < n; i += blockDim.x)
        s_mask[i] = (uint8_t)(fm_error(s_bestF, R->p1[i], R->p2[i]) <= s_thr2);
      __syncthreads();
      if (mode == kModeLmeds && tid == 0) {
        int cnt = 0;
        for (int i = 0; i < n; ++i) cnt += s_mask[i];
        s_result = cnt >= kModelPts;
      }
      __syncthreads();
    }
  }
  const int ok = s_result;
  if (!A.to_tracks) {
    for (int i = tid; i < n; i += blockDim.x) A.mask[i] = ok ? s_mask[i] : 0;
    if (tid == 0) *A.iters = s_iters;
    return;
  }
  // reduceVector(prev_pts / cur_pts / ids / track_cnt, status) (feature_tracker.cpp:938-942)
  TrackState* st = B.st;
  float2 pp = make_float2(0, 0), cp = make_float2(0, 0);
  int id = 0, cnt = 0;
  const int keep = tid < n ? (ok ? s_mask[tid] : 0) : 0;
  if (tid < n) {
    pp = B.prev_pts[tid];
    cp = B.cur_pts[tid];
    id = B.ids[tid];
    cnt = B.cnt[tid];
  }
  const unsigned bal = __ballot_sync(0xffffffffu, keep);
  if (lane == 0) s_warp[warp] = __popc(bal);
  __syncthreads();
  if (warp == 0) {
    const int v = s_warp[lane];
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    s_warp[lane] = incl - v;
    if (lane == 31) s_warp[32] = incl;
  }
  __syncthreads();
  if (keep) {
    const int pos = s_warp[warp] + __popc(bal & ((1u << lane) - 1u));
    B.prev_pts[pos] = pp;
    B.cur_pts[pos] = cp;
    B.ids[pos] = id;
    B.cnt[pos] = cnt;
  }
  if (tid == 0) {
    st->n_cur = s_warp[32];
    st->stat_after_ransac = s_warp[32];
    st->stat_after_mask = s_warp[32];
    st->stat_ransac_iters = s_iters;
  }
}

constexpr size_t kPrepareSmem = (size_t)kNumDraws * 5;

static void ransac_configure() {
  static bool done[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && done[dev]) return;
  cudaFuncSetAttribute(k_ransac_chase, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)kPrepareSmem);
  if (dev >= 0 && dev < 64) done[dev] = true;
}

void launch_ransac(const TrackParams& P, const TrackBuffers& B, cudaStream_t s,
                   int64_t* launches) {
  ransac_configure();
  RansacScratch* R = reinterpret_cast<RansacScratch*>(B.rs);
  PrepareArgs pa;
  pa.from_tracks = 1;
  pa.p1 = pa.p2 = nullptr;
  pa.n = 0;
  pa.thresh = P.f_threshold;
  pa.min_points = 8;
  launch_pdl(k_ransac_prepare_cached, dim3(1), dim3(1024), 0, s, P, B, pa, R);
  launch_pdl(k_ransac_hyp, dim3(kMaxAttempts / kHyp), dim3(32 * kHyp), 0, s, R);
  FoldArgs fa;
  fa.to_tracks = 1;
  fa.mask = nullptr;
  fa.iters = nullptr;
  launch_pdl(k_ransac_fold, dim3(1), dim3(1024), 0, s, P, B, fa, R);
  *launches += 3;
}

size_t ransac_cache_idx_bytes(int n_lo, int n_hi) {
  return (size_t)(n_hi - n_lo + 1) * kMaxAttempts * 8 * sizeof(uint16_t);
}

// attempt indices of every point count n_lo..n_hi (see k_ransac_prepare_cached); `dummy` =
// any n_hi float2 on the device
int ransac_build_cache(const TrackParams& P, const TrackBuffers& B, const float2* dummy, int n_lo,
                       int n_hi, uint16_t* cache_idx, int* cache_natt, cudaStream_t s) {
  RansacScratch* R = reinterpret_cast<RansacScratch*>(B.rs);
  ransac_configure();
  for (int n = n_lo; n <= n_hi; ++n) {
    PrepareArgs pa;
    pa.from_tracks = 0;
    pa.p1 = pa.p2 = dummy;
    pa.n = n;
    pa.thresh = 1.0;
    pa.min_points = 7;
    k_ransac_len<<<kNumDraws / 1024, 1024, 0, s>>>(B, pa, B.rng_draws, R);
    k_ransac_chase<<<1, 1024, kPrepareSmem, s>>>(P, B, pa, R);
    k_ransac_cache_row<<<1, 256, 0, s>>>(
        R, reinterpret_cast<uint4*>(cache_idx) + (size_t)(n - n_lo) * kMaxAttempts,
        cache_natt + (n - n_lo));
  }
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

void launch_ransac_stage(const TrackParams& P, const TrackBuffers& B, const float2* p1,
                         const float2* p2, int n, double thresh, uint8_t* mask, int* iters,
                         cudaStream_t s, int64_t* launches) {
  RansacScratch* R = reinterpret_cast<RansacScratch*>(B.rs);
  PrepareArgs pa;
  pa.from_tracks = 0;
  pa.p1 = p1;
  pa.p2 = p2;
  pa.n = n;
  pa.thresh = thresh;
  pa.min_points = 7;
  ransac_configure();
  launch_pdl(k_ransac_len, dim3(kNumDraws / 1024), dim3(1024), 0, s, B, pa, B.rng_draws, R);
  launch_pdl(k_ransac_chase, dim3(1), dim3(1024), kPrepareSmem, s, P, B, pa, R);
  launch_pdl(k_ransac_hyp, dim3(kMaxAttempts / kHyp), dim3(32 * kHyp), 0, s, R);
  FoldArgs fa;
  fa.to_tracks = 0;
  fa.mask = mask;
  fa.iters = iters;
  launch_pdl(k_ransac_fold, dim3(1), dim3(1024), 0, s, P, B, fa, R);
  *launches += 4;
}

}  // namespace esvio
