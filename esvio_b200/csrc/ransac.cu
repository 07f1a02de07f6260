// ransac.cu -- FeatureTracker::rejectWithF_event (feature_tracker/src/feature_tracker.cpp:
// 910-947): lift prev/cur points through the pinhole model onto a virtual f=460 camera and
// keep the inliers of cv::findFundamentalMat(FM_RANSAC, F_THRESHOLD, 0.99).
//
// OpenCV (modules/calib3d/src/fundam.cpp + ptsetreg.cpp; not in the reference tree) runs a
// sequential, adaptively shortened loop of 7-point hypotheses driven by cv::RNG(-1); for
// fewer than 15 points it switches to LMedS.  Here one CTA replays exactly that sequence in
// rounds of 32 hypotheses: thread 0 draws the 32 subsets from the same RNG stream, 32 lanes
// solve the 7-point problems, 32 warps score one hypothesis each, and thread 0 folds the
// scores in iteration order with the same "better than best" and iteration-count update
// rules, so the surviving model (and hence the inlier mask) is the one the sequential loop
// picks.  The 2-D null space comes from Gauss-Jordan elimination with complete pivoting
// instead of an SVD: any basis of the same null space gives the same F matrices.
#include <float.h>

#include "common.cuh"

namespace esvio {

constexpr int kRansacThreads = 1024;
constexpr int kHyp = 32;  // hypotheses per round
constexpr int kModelPts = 7;

struct CvRng {
  uint64_t state;
  __device__ unsigned next() {
    state = (uint64_t)(unsigned)state * 4164903690ULL + (unsigned)(state >> 32);
    return (unsigned)state;
  }
  __device__ int uniform(int a, int b) { return a == b ? a : (int)(next() % (unsigned)(b - a) + a); }
};

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

// PointSetRegistrator::getSubset
__device__ bool get_subset7(const float2* p1, const float2* p2, int count, CvRng& rng,
                            int max_attempts, float2* s1, float2* s2) {
  int idx[kModelPts];
  for (int iters = 0; iters < max_attempts; ++iters) {
    for (int i = 0; i < kModelPts; ++i) {
      int cand;
      for (;;) {
        cand = rng.uniform(0, count);
        bool dup = false;
        for (int q = 0; q < i; ++q) dup |= (idx[q] == cand);
        if (!dup) break;
      }
      idx[i] = cand;
      s1[i] = p1[cand];
      s2[i] = p2[cand];
    }
    if (!collinear_with_last(s1, kModelPts) && !collinear_with_last(s2, kModelPts)) return true;
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

// FMEstimatorCallback::runKernel for 7 points (run7Point): up to 3 matrices, row-major
__device__ int run_7point(const float2* m1, const float2* m2, double* Fout) {
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

  double A[7][9];
  for (int i = 0; i < 7; ++i) {
    const double x0 = (m1[i].x - c1x) * s1, y0 = (m1[i].y - c1y) * s1;
    const double x1 = (m2[i].x - c2x) * s2, y1 = (m2[i].y - c2y) * s2;
    A[i][0] = x1 * x0, A[i][1] = x1 * y0, A[i][2] = x1;
    A[i][3] = y1 * x0, A[i][4] = y1 * y0, A[i][5] = y1;
    A[i][6] = x0, A[i][7] = y0, A[i][8] = 1;
  }
  // Gauss-Jordan with complete pivoting -> [I | C] in permuted columns
  int perm[9];
  for (int j = 0; j < 9; ++j) perm[j] = j;
  for (int k = 0; k < 7; ++k) {
    int pi = k, pj = k;
    double best = -1;
    for (int i = k; i < 7; ++i)
      for (int j = k; j < 9; ++j)
        if (fabs(A[i][j]) > best) best = fabs(A[i][j]), pi = i, pj = j;
    if (!(best > 1e-14)) return 0;  // rank deficient sample
    if (pi != k)
      for (int j = 0; j < 9; ++j) {
        const double tmp = A[k][j];
        A[k][j] = A[pi][j];
        A[pi][j] = tmp;
      }
    if (pj != k) {
      for (int i = 0; i < 7; ++i) {
        const double tmp = A[i][k];
        A[i][k] = A[i][pj];
        A[i][pj] = tmp;
      }
      const int tp = perm[k];
      perm[k] = perm[pj];
      perm[pj] = tp;
    }
    const double inv = 1.0 / A[k][k];
    for (int j = k; j < 9; ++j) A[k][j] *= inv;
    for (int i = 0; i < 7; ++i) {
      if (i == k) continue;
      const double f = A[i][k];
      if (f != 0.0)
        for (int j = k; j < 9; ++j) A[i][j] -= f * A[k][j];
    }
  }
  double f1[9], f2[9];
  for (int j = 0; j < 7; ++j) {
    f1[perm[j]] = -A[j][7];
    f2[perm[j]] = -A[j][8];
  }
  f1[perm[7]] = 1, f1[perm[8]] = 0;
  f2[perm[7]] = 0, f2[perm[8]] = 1;

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
    double* F = Fout + 9 * k;
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
  }
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

struct RansacShared {
  float2 p1[kMaxCnt], p2[kMaxCnt];
  float2 s1[kHyp][kModelPts], s2[kHyp][kModelPts];
  double F[kHyp][27];
  int nmodels[kHyp];
  int sub_ok[kHyp];
  int good[kHyp][3];
  float median[kHyp][3];
  double bestF[9];
  int have_best, stop, iters_done;
  uint8_t mask[kMaxCnt];
};

// Computes the inlier mask of findFundamentalMat(p1, p2, FM_RANSAC, thresh, 0.99) into
// S.mask; returns (uniformly) 1 when a model was found.  Called by the whole CTA.
__device__ int fundamental_mask(RansacShared& S, int n, double thresh, double confidence,
                                int max_iters, int* iters_out) {
  const int tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
  if (thresh <= 0) thresh = 3;
  if (confidence < DBL_EPSILON || confidence > 1 - DBL_EPSILON) confidence = 0.99;
  __shared__ CvRng s_rng;
  __shared__ int s_niters, s_iter, s_best;
  __shared__ double s_min_median;
  if (tid == 0) {
    s_rng.state = (uint64_t)-1;
    s_iter = 0;
    s_best = 0;
    S.have_best = 0;
    S.stop = 0;
    s_min_median = DBL_MAX;
  }
  for (int i = tid; i < n; i += blockDim.x) S.mask[i] = 0;
  __syncthreads();
  if (n < 7) return 0;
  if (n == 7) {  // direct 7-point, every point is an inlier
    if (tid == 0) S.nmodels[0] = run_7point(S.p1, S.p2, S.F[0]);
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) S.mask[i] = 1;
    __syncthreads();
    if (iters_out && tid == 0) *iters_out = 1;
    return S.nmodels[0] > 0;
  }
  const bool ransac = n >= 15;
  const float t2 = (float)(thresh * thresh);
  if (tid == 0) {
    if (ransac) s_niters = max_iters > 1 ? max_iters : 1;
    else {
      int it = ransac_update_iters(confidence, 0.45, kModelPts, max_iters);
      s_niters = it < 3 ? 3 : it;
    }
  }
  __syncthreads();

  while (true) {
    // ---- draw the next kHyp subsets from the RNG stream
    if (tid == 0) {
      for (int h = 0; h < kHyp; ++h) {
        S.sub_ok[h] = 0;
        if (s_iter + h >= s_niters) break;  // cannot be needed
        S.sub_ok[h] = get_subset7(S.p1, S.p2, n, s_rng, ransac ? 10000 : 1000, S.s1[h], S.s2[h]) ? 1 : -1;
        if (S.sub_ok[h] < 0) break;
      }
    }
    __syncthreads();
    // ---- solve
    if (tid < kHyp) {
      S.nmodels[tid] = 0;
      if (S.sub_ok[tid] > 0) {
        const int nm = run_7point(S.s1[tid], S.s2[tid], S.F[tid]);
        S.nmodels[tid] = nm < 0 ? 0 : (nm > 3 ? 3 : nm);
      }
    }
    __syncthreads();
    // ---- score: warp h scores hypothesis h
    if (warp < kHyp) {
      const int nm = S.nmodels[warp];
      for (int m = 0; m < nm; ++m) {
        const double* F = S.F[warp] + 9 * m;
        if (ransac) {
          int good = 0;
          for (int i = lane; i < n; i += 32) good += fm_error(F, S.p1[i], S.p2[i]) <= t2;
          good = __reduce_add_sync(0xffffffffu, good);
          if (lane == 0) S.good[warp][m] = good;
        } else {
          // n < 15: median of the errors = element n/2 of the sorted list
          const float e = lane < n ? fm_error(F, S.p1[lane], S.p2[lane]) : FLT_MAX;
          int rank = 0;
          for (int j = 0; j < n; ++j) {
            const float o = __shfl_sync(0xffffffffu, e, j);
            rank += (o < e) || (o == e && j < lane);
          }
          if (lane < n && rank == n / 2) S.median[warp][m] = e;
        }
      }
    }
    __syncthreads();
    // ---- fold in iteration order
    if (tid == 0) {
      int h = 0;
      for (; h < kHyp; ++h) {
        if (s_iter >= s_niters) break;
        if (S.sub_ok[h] <= 0) {  // getSubset failed
          S.stop = 1;
          break;
        }
        for (int m = 0; m < S.nmodels[h]; ++m) {
          if (ransac) {
            const int good = S.good[h][m];
            if (good > (s_best > kModelPts - 1 ? s_best : kModelPts - 1)) {
              s_best = good;
              for (int q = 0; q < 9; ++q) S.bestF[q] = S.F[h][9 * m + q];
              S.have_best = 1;
              s_niters = ransac_update_iters(confidence, (double)(n - good) / n, kModelPts, s_niters);
            }
          } else {
            const double med = (double)S.median[h][m];
            if (med < s_min_median) {
              s_min_median = med;
              for (int q = 0; q < 9; ++q) S.bestF[q] = S.F[h][9 * m + q];
              S.have_best = 1;
            }
          }
        }
        ++s_iter;
      }
      if (s_iter >= s_niters) S.stop = 1;
    }
    __syncthreads();
    if (S.stop) break;
  }
  int result = 0;
  if (S.have_best) {
    float thr2 = t2;
    if (!ransac) {
      double sigma = 2.5 * 1.4826 * (1 + 5. / (n - kModelPts)) * sqrt(s_min_median);
      sigma = fmax(sigma, 0.001);
      thr2 = (float)(sigma * sigma);
    }
    for (int i = tid; i < n; i += blockDim.x)
      S.mask[i] = (uint8_t)(fm_error(S.bestF, S.p1[i], S.p2[i]) <= thr2);
    __syncthreads();
    if (ransac) result = 1;
    else {
      int cnt = 0;
      for (int i = 0; i < n; ++i) cnt += S.mask[i];
      result = cnt >= kModelPts;
    }
  }
  __syncthreads();
  if (iters_out && tid == 0) *iters_out = s_iter;
  return result;
}

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

__global__ void __launch_bounds__(kRansacThreads) k_ransac_tracks(TrackParams P, TrackBuffers B) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  RansacShared& S = *reinterpret_cast<RansacShared*>(s_raw);
  __shared__ int s_warp[33];
  TrackState* st = B.st;
  const int n = st->n_cur;
  const int tid = threadIdx.x;
  if (n < 8) return;  // cur_pts.size() >= 8 guard (feature_tracker.cpp:912)
  float2 pp = make_float2(0, 0), cp = make_float2(0, 0);
  int id = 0, cnt = 0;
  if (tid < n) {
    pp = B.prev_pts[tid];
    cp = B.cur_pts[tid];
    id = B.ids[tid];
    cnt = B.cnt[tid];
    double x, y;
    lift_pinhole(P.cam[0], (double)pp.x, (double)pp.y, x, y);
    S.p1[tid] = make_float2((float)(P.focal_length * x + P.W / 2.0),
                            (float)(P.focal_length * y + P.H / 2.0));
    lift_pinhole(P.cam[0], (double)cp.x, (double)cp.y, x, y);
    S.p2[tid] = make_float2((float)(P.focal_length * x + P.W / 2.0),
                            (float)(P.focal_length * y + P.H / 2.0));
  }
  __syncthreads();
  int iters = 0;
  __shared__ int s_iters;
  fundamental_mask(S, n, P.f_threshold, 0.99, 1000, &s_iters);
  __syncthreads();
  iters = s_iters;
  const int keep = tid < n ? S.mask[tid] : 0;
  // order-preserving compaction (reduceVector)
  const int lane = lane_id(), warp = tid >> 5;
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
    st->stat_ransac_iters = iters;
  }
}

__global__ void __launch_bounds__(kRansacThreads)
k_ransac_stage(const float2* __restrict__ p1, const float2* __restrict__ p2, int n, double thresh,
               uint8_t* __restrict__ mask, int* __restrict__ iters) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  RansacShared& S = *reinterpret_cast<RansacShared*>(s_raw);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    S.p1[i] = p1[i];
    S.p2[i] = p2[i];
  }
  __syncthreads();
  __shared__ int s_iters;
  if (threadIdx.x == 0) s_iters = 0;
  const int ok = fundamental_mask(S, n, thresh, 0.99, 1000, &s_iters);
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) mask[i] = ok ? S.mask[i] : 0;
  if (threadIdx.x == 0) *iters = s_iters;
}

static int ransac_configure() {
  static int done = 0;
  cudaError_t e = cudaFuncSetAttribute(k_ransac_tracks, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(RansacShared));
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(k_ransac_stage, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(RansacShared));
  done = e == cudaSuccess;
  return done ? 0 : -1;
}

void launch_ransac(const TrackParams& P, const TrackBuffers& B, cudaStream_t s,
                   int64_t* launches) {
  ransac_configure();
  k_ransac_tracks<<<1, kRansacThreads, sizeof(RansacShared), s>>>(P, B);
  ++*launches;
}

void launch_ransac_stage(const float2* p1, const float2* p2, int n, double thresh, uint8_t* mask,
                         int* iters, cudaStream_t s, int64_t* launches) {
  ransac_configure();
  k_ransac_stage<<<1, kRansacThreads, sizeof(RansacShared), s>>>(p1, p2, n, thresh, mask, iters);
  ++*launches;
}

}  // namespace esvio
