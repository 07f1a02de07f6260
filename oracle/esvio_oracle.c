/*
 * esvio_oracle.c -- CPU oracle (test infrastructure, NOT product code).
 * See esvio_oracle.h for the reference file:line map and parity status.
 *
 * Build: gcc -O3 -ffp-contract=off -fPIC -shared (the reference builds with
 * -O3 and no -march flag, feature_tracker/CMakeLists.txt:4-6, so no FMA
 * contraction happens there either).
 */
#include "esvio_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define ORA_API __attribute__((visibility("default")))

static double now_sec(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ======================================================================== */
/* SAE                                                                      */
/* ======================================================================== */

ORA_API ora_sae *ora_sae_create(int W, int H) {
  ora_sae *s = (ora_sae *)calloc(1, sizeof(ora_sae));
  s->W = W;
  s->H = H;
  for (int k = 0; k < 2; ++k) {
    s->sae[k] = (double *)calloc((size_t)W * H, sizeof(double));
    s->latest[k] = (double *)calloc((size_t)W * H, sizeof(double));
  }
  return s;
}

ORA_API void ora_sae_destroy(ora_sae *s) {
  if (!s) return;
  for (int k = 0; k < 2; ++k) {
    free(s->sae[k]);
    free(s->latest[k]);
  }
  free(s);
}

ORA_API void ora_sae_reset(ora_sae *s) {
  size_t n = (size_t)s->W * s->H * sizeof(double);
  for (int k = 0; k < 2; ++k) {
    memset(s->sae[k], 0, n);
    memset(s->latest[k], 0, n);
  }
}

/* event_detector.cc:149-166: an event is ACCEPTED into sae[pol] when it is more
 * than filter_threshold after the previous same-polarity event at the pixel, or
 * when the opposite polarity fired after that previous event; latest[pol] is
 * always advanced. */
ORA_API void ora_sae_update(ora_sae *s, const uint16_t *x, const uint16_t *y, const double *t,
                            const uint8_t *p, size_t n, double filter_threshold) {
  const int W = s->W;
  for (size_t i = 0; i < n; ++i) {
    const size_t px = (size_t)x[i] + (size_t)y[i] * W;
    const int pol = p[i] ? 1 : 0;
    const double prev_same = s->latest[pol][px];
    const double prev_opp = s->latest[1 - pol][px];
    const double et = t[i];
    if (et > prev_same + filter_threshold || prev_opp > prev_same) s->sae[pol][px] = et;
    s->latest[pol][px] = et;
  }
}

/* ======================================================================== */
/* Motion-compensated SAE update (event_detector.cc:102-147,168-210,547-591)   */
/* ======================================================================== */
/* The reference does this arithmetic with Eigen fixed-size float types (Matrix3f / Vector3f,
 * unsupported/MatrixFunctions exp(), inverse(), PartialPivLU).  Eigen is not in
 * /root/reference (eigen_catkin downloads >= 3.3.4) nor in this container, so the evaluation
 * order below restates Eigen 3.3's fixed-size kernels as published: coefficient-based 3x3
 * products summed as p0 + (p1 + p2) (redux_novec_unroller), cofactor inverse with one
 * 1/det, matrix exponential = Pade 3/5/7 by L1 norm with scaling and squaring, solved by
 * partial-pivot LU with column-oriented triangular solves (reciprocal of the diagonal).
 * PARITY UNPINNED against a real Eigen build; the CUDA path is bit-exact against this. */
static float dot3f(float a0, float b0, float a1, float b1, float a2, float b2) {
  return a0 * b0 + (a1 * b1 + a2 * b2);
}
static void mat3_mul_f(const float *A, const float *B, float *C) { /* row-major, C != A,B */
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      C[i * 3 + j] = dot3f(A[i * 3], B[j], A[i * 3 + 1], B[3 + j], A[i * 3 + 2], B[6 + j]);
}
static void mat3_vec_f(const float *A, const float *v, float *o) {
  for (int i = 0; i < 3; ++i) o[i] = dot3f(A[i * 3], v[0], A[i * 3 + 1], v[1], A[i * 3 + 2], v[2]);
}
static float cof3f(const float *m, int i, int j) { /* Eigen cofactor_3x3<i,j> */
  const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
  return m[i1 * 3 + j1] * m[i2 * 3 + j2] - m[i1 * 3 + j2] * m[i2 * 3 + j1];
}
static void mat3_inv_f(const float *m, float *r) { /* compute_inverse<Matrix3f> */
  const float c0 = cof3f(m, 0, 0), c1 = cof3f(m, 1, 0), c2 = cof3f(m, 2, 0);
  const float det = dot3f(c0, m[0], c1, m[3], c2, m[6]);
  const float invdet = 1.0f / det;
  r[0] = c0 * invdet;
  r[1] = c1 * invdet;
  r[2] = c2 * invdet;
  r[3] = cof3f(m, 0, 1) * invdet;
  r[4] = cof3f(m, 1, 1) * invdet;
  r[5] = cof3f(m, 2, 1) * invdet;
  r[6] = cof3f(m, 0, 2) * invdet;
  r[7] = cof3f(m, 1, 2) * invdet;
  r[8] = cof3f(m, 2, 2) * invdet;
}

/* Matrix3f::exp() (unsupported/Eigen/src/MatrixFunctions/MatrixExponential.h, float) */
ORA_API void ora_mat3_exp_f(const float *A_in, float *R) {
  float A[9], A2[9], A4[9], A6[9], tmp[9], U[9], V[9];
  memcpy(A, A_in, sizeof(A));
  float l1 = 0.f;
  for (int j = 0; j < 3; ++j) {
    const float cs = fabsf(A[j]) + (fabsf(A[3 + j]) + fabsf(A[6 + j]));
    if (cs > l1) l1 = cs;
  }
  int squarings = 0;
  if (l1 < 4.258730016922831e-001f) {
    const float b[] = {120.f, 60.f, 12.f, 1.f};
    mat3_mul_f(A, A, A2);
    for (int i = 0; i < 9; ++i) {
      const float id = (i % 4 == 0) ? 1.f : 0.f;
      tmp[i] = b[3] * A2[i] + b[1] * id;
      V[i] = b[2] * A2[i] + b[0] * id;
    }
    mat3_mul_f(A, tmp, U);
  } else if (l1 < 1.880152677804762e+000f) {
    const float b[] = {30240.f, 15120.f, 3360.f, 420.f, 30.f, 1.f};
    mat3_mul_f(A, A, A2);
    mat3_mul_f(A2, A2, A4);
    for (int i = 0; i < 9; ++i) {
      const float id = (i % 4 == 0) ? 1.f : 0.f;
      tmp[i] = b[5] * A4[i] + b[3] * A2[i] + b[1] * id;
      V[i] = b[4] * A4[i] + b[2] * A2[i] + b[0] * id;
    }
    mat3_mul_f(A, tmp, U);
  } else {
    const float maxnorm = 3.925724783138660f;
    frexpf(l1 / maxnorm, &squarings);
    if (squarings < 0) squarings = 0;
    for (int i = 0; i < 9; ++i) A[i] = ldexpf(A[i], -squarings);
    const float b[] = {17297280.f, 8648640.f, 1995840.f, 277200.f, 25200.f, 1512.f, 56.f, 1.f};
    mat3_mul_f(A, A, A2);
    mat3_mul_f(A2, A2, A4);
    mat3_mul_f(A4, A2, A6);
    for (int i = 0; i < 9; ++i) {
      const float id = (i % 4 == 0) ? 1.f : 0.f;
      tmp[i] = b[7] * A6[i] + b[5] * A4[i] + b[3] * A2[i] + b[1] * id;
      V[i] = b[6] * A6[i] + b[4] * A4[i] + b[2] * A2[i] + b[0] * id;
    }
    mat3_mul_f(A, tmp, U);
  }
  float lu[9], x[9];
  for (int i = 0; i < 9; ++i) {
    x[i] = U[i] + V[i];    /* numer */
    lu[i] = -U[i] + V[i];  /* denom */
  }
  /* PartialPivLU (unblocked) */
  int piv[3];
  for (int k = 0; k < 3; ++k) {
    int best = k;
    float score = fabsf(lu[k * 3 + k]);
    for (int i = k + 1; i < 3; ++i)
      if (fabsf(lu[i * 3 + k]) > score) {
        score = fabsf(lu[i * 3 + k]);
        best = i;
      }
    piv[k] = best;
    if (score != 0.f) {
      if (best != k)
        for (int j = 0; j < 3; ++j) {
          const float t = lu[k * 3 + j];
          lu[k * 3 + j] = lu[best * 3 + j];
          lu[best * 3 + j] = t;
        }
      for (int i = k + 1; i < 3; ++i) lu[i * 3 + k] /= lu[k * 3 + k];
    }
    for (int i = k + 1; i < 3; ++i)
      for (int j = k + 1; j < 3; ++j) lu[i * 3 + j] -= lu[i * 3 + k] * lu[k * 3 + j];
  }
  /* solve: P * numer, unit-lower then upper (triangular_solve_matrix, column-major lhs) */
  for (int k = 0; k < 3; ++k)
    if (piv[k] != k)
      for (int j = 0; j < 3; ++j) {
        const float t = x[k * 3 + j];
        x[k * 3 + j] = x[piv[k] * 3 + j];
        x[piv[k] * 3 + j] = t;
      }
  for (int k = 0; k < 3; ++k)
    for (int j = 0; j < 3; ++j) {
      const float b = x[k * 3 + j];
      for (int i = k + 1; i < 3; ++i) x[i * 3 + j] -= b * lu[i * 3 + k];
    }
  for (int k = 2; k >= 0; --k) {
    const float a = 1.0f / lu[k * 3 + k];
    for (int j = 0; j < 3; ++j) {
      const float b = (x[k * 3 + j] *= a);
      for (int i = 0; i < k; ++i) x[i * 3 + j] -= b * lu[i * 3 + k];
    }
  }
  for (int s = 0; s < squarings; ++s) {
    float t[9];
    mat3_mul_f(x, x, t);
    memcpy(x, t, sizeof(t));
  }
  memcpy(R, x, sizeof(x));
}

/* EventDetector::motioncorrection (event_detector.cc:547-591): pixel (ex, ey) of an event
 * `dt` seconds after the window's first event, warped back to the first event's time. */
ORA_API void ora_motion_correct(const ora_motion *m, int W, int H, double ex, double ey, double dt,
                                int *ox, int *oy) {
  *ox = (int)ex;
  *oy = (int)ey;
  const int border = 6;
  if (!(ex > border && ex <= (W - border) && ey > border && ey <= (H - border))) return;
  const float fdt = (float)dt; /* Vector3f * double: the scalar is converted to float */
  const float rv[3] = {m->omega[0] * fdt, m->omega[1] * fdt, m->omega[2] * fdt};
  const float skew[9] = {0.f, -rv[2], rv[1], rv[2], 0.f, -rv[0], -rv[1], rv[0], 0.f};
  float R[9], Rt[9], K[9] = {m->K[0], 0.f, m->K[2], 0.f, m->K[1], m->K[3], 0.f, 0.f, 1.f};
  float Kinv[9], KR[9], rotK[9];
  ora_mat3_exp_f(skew, R);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Rt[i * 3 + j] = R[j * 3 + i];
  mat3_inv_f(K, Kinv);
  mat3_mul_f(K, Rt, KR);
  mat3_mul_f(KR, Kinv, rotK);
  const float c = (float)(0.5 * dt);
  const float tv[3] = {(float)m->state_v[0], (float)m->state_v[1], (float)m->state_v[2]};
  float tr[3], w[3], nrotK[9], transK[3], ev[3] = {(float)ex, (float)ey, 1.f}, o[3];
  for (int i = 0; i < 3; ++i) tr[i] = c * (tv[i] + m->v_pre[i]);
  mat3_vec_f(Kinv, tr, w);
  for (int i = 0; i < 9; ++i) nrotK[i] = -rotK[i];
  mat3_vec_f(nrotK, w, transK);
  mat3_vec_f(rotK, ev, o);
  for (int i = 0; i < 3; ++i) o[i] = o[i] + transK[i];
  o[0] = o[0] / o[2];
  o[1] = o[1] / o[2];
  const int x = (int)floorf(o[0]), y = (int)floorf(o[1]);
  if (x > 0 && x < W - 1 && y > 0 && y < H - 1) {
    *ox = x;
    *oy = y;
  }
}

/* |accel| > a_motion_compensation_threshold (event_detector.cc:125, event_detector.h:51) */
ORA_API int ora_motion_active(const ora_motion *m) {
  const double a0 = m->accel[0], a1 = m->accel[1], a2 = m->accel[2];
  return sqrt(a0 * a0 + a1 * a1 + a2 * a2) > 5.0;
}

/* trackEvent(..., measurements) HOT LOOP A (feature_tracker.cpp:628-642) for one camera:
 * t0 = time of the first LEFT event, t1 = left header stamp. */
ORA_API void ora_sae_update_mc(ora_sae *s, const uint16_t *x, const uint16_t *y, const double *t,
                               const uint8_t *p, size_t n, double filter_threshold,
                               const ora_motion *m, double t0) {
  const int W = s->W, H = s->H;
  const double dtw = m->t1 - t0;
  const int active = ora_motion_active(m);
  for (size_t i = 0; i < n; ++i) {
    int ex = x[i], ey = y[i];
    const double et = t[i];
    if (dtw > 0 && (et - t0) / dtw < 1 && active) ora_motion_correct(m, W, H, ex, ey, et - t0, &ex, &ey);
    const size_t px = (size_t)ex + (size_t)ey * W;
    const int pol = p[i] ? 1 : 0;
    const double prev_same = s->latest[pol][px];
    const double prev_opp = s->latest[1 - pol][px];
    if (et > prev_same + filter_threshold || prev_opp > prev_same) s->sae[pol][px] = et;
    s->latest[pol][px] = et;
  }
}

static inline uint8_t sat_u8_from_double(double v) {
  long r = lrint(v); /* round-half-even, like cv::saturate_cast<uchar>(double) */
  if (r < 0) r = 0;
  if (r > 255) r = 255;
  return (uint8_t)r;
}

/* event_detector.cc:230-267.  The MatExpr 255*(m+1)/2 folds to convertTo(alpha=127.5,
 * beta=127.5); an untouched pixel is 127.5 -> 128 under round-half-even. */
ORA_API void ora_time_surface(const ora_sae *s, double t_ref, double decay_ms, int ignore_polarity,
                              uint8_t *out) {
  const double decay_sec = decay_ms / 1000.0;
  const size_t N = (size_t)s->W * s->H;
  const double *pos = s->sae[1], *neg = s->sae[0];
  for (size_t i = 0; i < N; ++i) {
    const int pos_newer = pos[i] > neg[i];
    const double stamp = pos_newer ? pos[i] : neg[i];
    double v = 0.0;
    if (stamp > 0) {
      const double dt = t_ref - stamp;
      v = exp(-dt / decay_sec);
      if (!ignore_polarity) v *= pos_newer ? 1.0 : -1.0;
    }
    const double scaled = ignore_polarity ? v * 255.0 : v * 127.5 + 127.5;
    out[i] = sat_u8_from_double(scaled);
  }
}

/* ======================================================================== */
/* Arc* (event_detector.cc:308-544)                                         */
/* ======================================================================== */

static const int8_t kRing3[16][2] = {{0, 3},  {1, 3},   {2, 2},   {3, 1},  {3, 0},  {3, -1},
                                     {2, -2}, {1, -3},  {0, -3},  {-1, -3}, {-2, -2}, {-3, -1},
                                     {-3, 0}, {-3, 1},  {-2, 2},  {-1, 3}};
static const int8_t kRing4[20][2] = {{0, 4},   {1, 4},   {2, 3},   {3, 2},  {4, 1},
                                     {4, 0},   {4, -1},  {3, -2},  {2, -3}, {1, -4},
                                     {0, -4},  {-1, -4}, {-2, -3}, {-3, -2}, {-4, -1},
                                     {-4, 0},  {-4, 1},  {-3, 2},  {-2, 3}, {-1, 4}};

typedef struct {
  int idx;
  double value, minimum;
} arc_arm;

static inline void arm_step(arc_arm *a, const double *ring, int n, int dir) {
  a->idx = (a->idx + dir + n) % n;
  a->value = ring[a->idx];
  if (a->value < a->minimum) a->minimum = a->value;
}

/* One ring of the Arc* test: grow an arc from the newest ring element towards
 * whichever neighbour is newer; returns 1 when the "newest segment" length is
 * in [.., hi] or in [n-hi, n-lo]. */
static int arc_ring_valid(const double *ring, int n, int lo, int hi) {
  int newest = 0;
  for (int i = 1; i < n; ++i)
    if (ring[i] > ring[newest]) newest = i;
  double seg_min = ring[newest];
  arc_arm cw = {(newest + 1) % n, 0, 0}, ccw = {(newest - 1 + n) % n, 0, 0};
  cw.value = cw.minimum = ring[cw.idx];
  ccw.value = ccw.minimum = ring[ccw.idx];
  int it = 1;
  for (; it < lo; ++it) {
    arc_arm *a = (cw.value > ccw.value) ? &cw : &ccw;
    if (a->minimum < seg_min) seg_min = a->minimum;
    arm_step(a, ring, n, a == &cw ? +1 : -1);
  }
  int seg_len = lo;
  for (; it < n; ++it) {
    arc_arm *a = (cw.value > ccw.value) ? &cw : &ccw;
    if (a->value >= seg_min) {
      seg_len = it + 1;
      if (a->minimum < seg_min) seg_min = a->minimum;
    }
    arm_step(a, ring, n, a == &cw ? +1 : -1);
  }
  return (seg_len <= hi) || (seg_len >= n - hi && seg_len <= n - lo);
}

ORA_API int ora_is_corner(const ora_sae *s, double t, int x, int y, int p, double filter_threshold,
                          int min_dist) {
  const int W = s->W, H = s->H;
  const int pol = p ? 1 : 0;
  const size_t px = (size_t)x + (size_t)y * W;
  const double last_same = s->latest[pol][px];
  const double last_opp = s->latest[1 - pol][px];
  if (t > last_same + filter_threshold || last_opp > last_same) return 0;
  const int border = min_dist + 1;
  if (x < border || x >= W - border || y < border || y >= H - border) return 0;
  const double *S = s->sae[pol];
  double ring[20];
  for (int i = 0; i < 16; ++i) ring[i] = S[(x + kRing3[i][0]) + (size_t)(y + kRing3[i][1]) * W];
  if (!arc_ring_valid(ring, 16, 4, 6)) return 0;
  for (int i = 0; i < 20; ++i) ring[i] = S[(x + kRing4[i][0]) + (size_t)(y + kRing4[i][1]) * W];
  return arc_ring_valid(ring, 20, 5, 8);
}

ORA_API void ora_corner_flags(const ora_sae *s, const uint16_t *x, const uint16_t *y,
                              const double *t, const uint8_t *p, size_t n,
                              double filter_threshold, int min_dist, uint8_t *flags) {
  for (size_t i = 0; i < n; ++i)
    flags[i] = (uint8_t)ora_is_corner(s, t[i], x[i], y[i], p[i], filter_threshold, min_dist);
}

/* ======================================================================== */
/* Filled circle raster (OpenCV drawing.cpp, Circle(..., fill=1))            */
/* ======================================================================== */

ORA_API void ora_disc_half_widths(int r, int *hw) {
  for (int k = 0; k <= r; ++k) hw[k] = -1;
  int err = 0, dx = r, dy = 0, plus = 1, minus = (r << 1) - 1;
  while (dx >= dy) {
    if (dx > hw[dy]) hw[dy] = dx; /* rows cy +- dy span cx +- dx */
    if (dy > hw[dx]) hw[dx] = dy; /* rows cy +- dx span cx +- dy */
    dy++;
    err += plus;
    plus += 2;
    int m = (err <= 0) - 1;
    err -= minus & m;
    dx += m;
    minus -= m & 2;
  }
}

ORA_API void ora_fill_disc_u8(uint8_t *mask, int W, int H, int cx, int cy, int r, uint8_t value) {
  int hw[r + 1];
  ora_disc_half_widths(r, hw);
  for (int k = -r; k <= r; ++k) {
    const int yy = cy + k;
    if (yy < 0 || yy >= H) continue;
    const int h = hw[k < 0 ? -k : k];
    if (h < 0) continue;
    int x0 = cx - h, x1 = cx + h;
    if (x0 < 0) x0 = 0;
    if (x1 > W - 1) x1 = W - 1;
    for (int xx = x0; xx <= x1; ++xx) mask[(size_t)yy * W + xx] = value;
  }
}

static inline int cv_round_f(float v) { return (int)lrintf(v); }

/* ---- std::sort as libstdc++ implements it (bits/stl_algo.h: __sort -> __introsort_loop +
 * __final_insertion_sort), on a permutation `o` of indices with the comparator of
 * feature_tracker.cpp:100-103,132-135: comp(a, b) = key[a] > key[b].  std::sort leaves the order
 * of equal keys unspecified; the reference is built with GCC (ROS), so "what the reference
 * does" on ties is what this algorithm does: median-of-three pivot moved to the front,
 * unguarded Hoare partition, recursion on the right part, ranges of <= 16 left to one final
 * insertion sort, heap sort when the depth budget 2*floor(log2 n) runs out.  Pinned against the
 * real std::sort by tests/test_oracle_ref_tracker.py (the reference's own Event_setMask and a
 * direct randomized comparison). ---- */
#define SORT_CMP(a, b) (key[(a)] > key[(b)])
static void ss_adjust_heap(const int *key, int *o, int hole, int len, int value) {
  const int top = hole;
  int child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (SORT_CMP(o[child], o[child - 1])) child--;
    o[hole] = o[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    o[hole] = o[child - 1];
    hole = child - 1;
  }
  int parent = (hole - 1) / 2; /* __push_heap */
  while (hole > top && SORT_CMP(o[parent], value)) {
    o[hole] = o[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  o[hole] = value;
}
static void ss_heap_sort(const int *key, int *o, int len) { /* __partial_sort(first, last, last) */
  if (len >= 2)
    for (int parent = (len - 2) / 2;; --parent) { /* __make_heap */
      ss_adjust_heap(key, o, parent, len, o[parent]);
      if (parent == 0) break;
    }
  for (int last = len; last > 1;) { /* __sort_heap: __pop_heap(first, last - 1, last - 1) */
    --last;
    const int value = o[last];
    o[last] = o[0];
    ss_adjust_heap(key, o, 0, last, value);
  }
}
static void ss_introsort_loop(const int *key, int *o, int first, int last, int depth) {
  while (last - first > 16) {
    if (depth == 0) {
      ss_heap_sort(key, o + first, last - first);
      return;
    }
    --depth;
    /* __unguarded_partition_pivot: __move_median_to_first(first, first + 1, mid, last - 1) */
    const int mid = first + (last - first) / 2, a = first + 1, b = mid, c = last - 1;
    int m;
    if (SORT_CMP(o[a], o[b])) m = SORT_CMP(o[b], o[c]) ? b : (SORT_CMP(o[a], o[c]) ? c : a);
    else m = SORT_CMP(o[a], o[c]) ? a : (SORT_CMP(o[b], o[c]) ? c : b);
    int tmp = o[first];
    o[first] = o[m];
    o[m] = tmp;
    /* __unguarded_partition(first + 1, last, pivot = first) */
    int i = first + 1, j = last;
    for (;;) {
      while (SORT_CMP(o[i], o[first])) ++i;
      --j;
      while (SORT_CMP(o[first], o[j])) --j;
      if (!(i < j)) break;
      tmp = o[i];
      o[i] = o[j];
      o[j] = tmp;
      ++i;
    }
    ss_introsort_loop(key, o, i, last, depth);
    last = i;
  }
}
/* order[k] = index of the element std::sort puts at position k; depth_limit < 0: the library's
 * own budget 2 * floor(log2 n) (tests force small budgets to reach the heap-sort branch) */
ORA_API void ora_std_sort_order(const int *key, int n, int depth_limit, int *order) {
  for (int i = 0; i < n; ++i) order[i] = i;
  if (n <= 1) return;
  if (depth_limit < 0) {
    int lg = 0;
    while ((n >> (lg + 1)) > 0) ++lg;
    depth_limit = 2 * lg;
  }
  ss_introsort_loop(key, order, 0, n, depth_limit);
  /* __final_insertion_sort: guarded insertion over the first 16, unguarded over the rest */
  const int head = n > 16 ? 16 : n;
  for (int i = 1; i < head; ++i) {
    const int v = order[i];
    if (SORT_CMP(v, order[0])) {
      memmove(order + 1, order, sizeof(int) * i);
      order[0] = v;
    } else {
      int j = i;
      while (SORT_CMP(v, order[j - 1])) {
        order[j] = order[j - 1];
        --j;
      }
      order[j] = v;
    }
  }
  for (int i = head; i < n; ++i) {
    const int v = order[i];
    int j = i;
    while (SORT_CMP(v, order[j - 1])) {
      order[j] = order[j - 1];
      --j;
    }
    order[j] = v;
  }
}
#undef SORT_CMP

/* feature_tracker.cpp:123-151.  Points are visited in the order std::sort (libstdc++) leaves
 * them in: track_cnt descending, ties as ora_std_sort_order documents; the CUDA path
 * reproduces the same permutation. */
ORA_API int ora_set_mask(int W, int H, int min_dist, int n, float *pts, int *ids, int *track_cnt,
                         uint8_t *mask) {
  memset(mask, 0, (size_t)W * H);
  if (n <= 0) return 0;
  int *order = (int *)malloc(sizeof(int) * n);
  ora_std_sort_order(track_cnt, n, -1, order);
  float *np = (float *)malloc(sizeof(float) * 2 * n);
  int *ni = (int *)malloc(sizeof(int) * n), *nc = (int *)malloc(sizeof(int) * n);
  int m = 0;
  for (int k = 0; k < n; ++k) {
    const int i = order[k];
    const int cx = cv_round_f(pts[2 * i]), cy = cv_round_f(pts[2 * i + 1]);
    if (cx < 0 || cx >= W || cy < 0 || cy >= H) continue; /* cannot happen after the border test */
    if (mask[(size_t)cy * W + cx] == 0) {
      np[2 * m] = pts[2 * i];
      np[2 * m + 1] = pts[2 * i + 1];
      ni[m] = ids[i];
      nc[m] = track_cnt[i];
      ++m;
      ora_fill_disc_u8(mask, W, H, cx, cy, min_dist, 255);
    }
  }
  memcpy(pts, np, sizeof(float) * 2 * m);
  memcpy(ids, ni, sizeof(int) * m);
  memcpy(track_cnt, nc, sizeof(int) * m);
  free(order);
  free(np);
  free(ni);
  free(nc);
  return m;
}

/* feature_tracker.cpp:13-38: first-come in stream order until the quota is filled */
ORA_API int ora_features_to_track(const ora_sae *left, const uint16_t *x, const uint16_t *y,
                                  const double *t, const uint8_t *p, size_t n, int max_corners,
                                  int min_dist, const uint8_t *mask, const uint8_t *ts,
                                  double ts_lk_threshold, double filter_threshold, float *out_pts,
                                  uint8_t *mask_out) {
  const int W = left->W, H = left->H;
  uint8_t *m = (uint8_t *)malloc((size_t)W * H);
  memcpy(m, mask, (size_t)W * H);
  int found = 0;
  if (max_corners > 0) {
    for (size_t i = 0; i < n && found < max_corners; ++i) {
      const size_t px = (size_t)y[i] * W + x[i];
      if (m[px] == 255) continue;
      if ((double)ts[px] == ts_lk_threshold) continue;
      if (!ora_is_corner(left, t[i], x[i], y[i], p[i], filter_threshold, min_dist)) continue;
      out_pts[2 * found] = (float)x[i];
      out_pts[2 * found + 1] = (float)y[i];
      ++found;
      ora_fill_disc_u8(m, W, H, x[i], y[i], min_dist, 255);
    }
  }
  if (mask_out) memcpy(mask_out, m, (size_t)W * H);
  free(m);
  return found;
}

/* ======================================================================== */
/* Pyramid (OpenCV pyrDown, 8U, BORDER_REFLECT_101) and Scharr derivative    */
/* ======================================================================== */

static inline int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) {
    if (i < 0) i = -i;
    else i = 2 * n - 2 - i;
  }
  return i;
}

ORA_API int ora_pyramid_sizes(int W, int H, int max_level, int win, int *w_out, int *h_out) {
  int w = W, h = H, levels = 0;
  for (int l = 0; l <= max_level; ++l) {
    w_out[l] = w;
    h_out[l] = h;
    levels = l + 1;
    w = (w + 1) / 2;
    h = (h + 1) / 2;
    if (w <= win || h <= win) break; /* buildOpticalFlowPyramid stops here */
  }
  return levels;
}

ORA_API void ora_pyr_down(const uint8_t *src, int sw, int sh, uint8_t *dst, int dw, int dh) {
  int *rows = (int *)malloc(sizeof(int) * 5 * dw);
  for (int y = 0; y < dh; ++y) {
    for (int k = 0; k < 5; ++k) {
      const uint8_t *s = src + (size_t)reflect101(2 * y - 2 + k, sh) * sw;
      int *r = rows + k * dw;
      for (int x = 0; x < dw; ++x) {
        const int c = 2 * x;
        r[x] = s[reflect101(c, sw)] * 6 + (s[reflect101(c - 1, sw)] + s[reflect101(c + 1, sw)]) * 4 +
               s[reflect101(c - 2, sw)] + s[reflect101(c + 2, sw)];
      }
    }
    for (int x = 0; x < dw; ++x) {
      const int v = rows[2 * dw + x] * 6 + (rows[dw + x] + rows[3 * dw + x]) * 4 + rows[x] +
                    rows[4 * dw + x];
      dst[(size_t)y * dw + x] = (uint8_t)((v + 128) >> 8);
    }
  }
  free(rows);
}

/* lkpyramid.cpp calcSharrDeriv: 3-10-3 smoothing x [-1 0 1], reflect-101 borders */
ORA_API void ora_scharr_deriv(const uint8_t *src, int w, int h, int16_t *dst) {
  int *sm = (int *)malloc(sizeof(int) * (w + 2) * 2);
  int *t0 = sm + 1, *t1 = sm + (w + 2) + 1;
  for (int y = 0; y < h; ++y) {
    const uint8_t *up = src + (size_t)reflect101(y - 1, h) * w;
    const uint8_t *mid = src + (size_t)y * w;
    const uint8_t *dn = src + (size_t)reflect101(y + 1, h) * w;
    for (int x = 0; x < w; ++x) {
      t0[x] = (up[x] + dn[x]) * 3 + mid[x] * 10;
      t1[x] = dn[x] - up[x];
    }
    const int xl = w > 1 ? 1 : 0, xr = w > 1 ? w - 2 : 0;
    t0[-1] = t0[xl];
    t0[w] = t0[xr];
    t1[-1] = t1[xl];
    t1[w] = t1[xr];
    int16_t *d = dst + (size_t)y * w * 2;
    for (int x = 0; x < w; ++x) {
      d[2 * x] = (int16_t)(t0[x + 1] - t0[x - 1]);
      d[2 * x + 1] = (int16_t)((t1[x + 1] + t1[x - 1]) * 3 + t1[x] * 10);
    }
  }
  free(sm);
}

/* ======================================================================== */
/* Pyramidal LK (OpenCV lkpyramid.cpp, scalar path)                          */
/* ======================================================================== */

typedef struct {
  int w, h;
  uint8_t *img;   /* padded by win on every side, REFLECT_101 */
  int16_t *deriv; /* padded by win on every side, zeros; 2 channels */
} lk_level;

static void pad_reflect_u8(const uint8_t *src, int w, int h, int pad, uint8_t *dst) {
  const int pw = w + 2 * pad;
  for (int y = -pad; y < h + pad; ++y) {
    const uint8_t *s = src + (size_t)reflect101(y, h) * w;
    uint8_t *d = dst + (size_t)(y + pad) * pw;
    for (int x = -pad; x < w + pad; ++x) d[x + pad] = s[reflect101(x, w)];
  }
}

static int build_levels(const uint8_t *img, int W, int H, int win, int max_level, int with_deriv,
                        lk_level *lv) {
  int ws[16], hs[16];
  const int n = ora_pyramid_sizes(W, H, max_level, win, ws, hs);
  uint8_t *cur = (uint8_t *)malloc((size_t)W * H);
  memcpy(cur, img, (size_t)W * H);
  for (int l = 0; l < n; ++l) {
    const int w = ws[l], h = hs[l], pw = w + 2 * win, ph = h + 2 * win;
    lv[l].w = w;
    lv[l].h = h;
    lv[l].img = (uint8_t *)malloc((size_t)pw * ph);
    pad_reflect_u8(cur, w, h, win, lv[l].img);
    lv[l].deriv = NULL;
    if (with_deriv) {
      int16_t *d = (int16_t *)malloc(sizeof(int16_t) * 2 * (size_t)w * h);
      ora_scharr_deriv(cur, w, h, d);
      lv[l].deriv = (int16_t *)calloc((size_t)pw * ph * 2, sizeof(int16_t));
      for (int y = 0; y < h; ++y)
        memcpy(lv[l].deriv + ((size_t)(y + win) * pw + win) * 2, d + (size_t)y * w * 2,
               sizeof(int16_t) * 2 * w);
      free(d);
    }
    if (l + 1 < n) {
      uint8_t *nxt = (uint8_t *)malloc((size_t)ws[l + 1] * hs[l + 1]);
      ora_pyr_down(cur, w, h, nxt, ws[l + 1], hs[l + 1]);
      free(cur);
      cur = nxt;
    }
  }
  free(cur);
  return n;
}

static void free_levels(lk_level *lv, int n) {
  for (int l = 0; l < n; ++l) {
    free(lv[l].img);
    free(lv[l].deriv);
  }
}

#define LK_W_BITS 14
static inline int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

ORA_API void ora_calc_optical_flow_pyr_lk(const uint8_t *prev, const uint8_t *next, int W, int H,
                                          const float *prev_pts, float *next_pts, int n,
                                          uint8_t *status, int win, int max_level, int max_count,
                                          double epsilon, int use_initial_flow,
                                          double min_eig_threshold) {
  if (n <= 0) return;
  lk_level LI[16], LJ[16];
  const int nl_i = build_levels(prev, W, H, win, max_level, 1, LI);
  const int nl_j = build_levels(next, W, H, win, max_level, 0, LJ);
  const int top = (nl_i < nl_j ? nl_i : nl_j) - 1;
  if (max_count < 0) max_count = 0;
  if (max_count > 100) max_count = 100;
  if (epsilon < 0) epsilon = 0;
  if (epsilon > 10) epsilon = 10;
  const double eps2 = epsilon * epsilon;
  const float half = (win - 1) * 0.5f;
  const float flt_scale = 1.f / (1 << 20);
  int16_t *Iw = (int16_t *)malloc(sizeof(int16_t) * win * win * 3);
  int16_t *dIw = Iw + win * win;

  for (int i = 0; i < n; ++i) status[i] = 1;
  if (!use_initial_flow)
    for (int i = 0; i < 2 * n; ++i) next_pts[i] = 0.f;

  for (int level = top; level >= 0; --level) {
    const lk_level *I = &LI[level], *J = &LJ[level];
    const int pw = I->w + 2 * win; /* same size for I and J */
    const float sc = (float)(1. / (1 << level));
    for (int k = 0; k < n; ++k) {
      float ppx = prev_pts[2 * k] * sc, ppy = prev_pts[2 * k + 1] * sc;
      float npx, npy;
      if (level == top) {
        if (use_initial_flow) {
          npx = next_pts[2 * k] * sc;
          npy = next_pts[2 * k + 1] * sc;
        } else {
          npx = ppx;
          npy = ppy;
        }
      } else {
        npx = next_pts[2 * k] * 2.f;
        npy = next_pts[2 * k + 1] * 2.f;
      }
      next_pts[2 * k] = npx;
      next_pts[2 * k + 1] = npy;

      ppx -= half;
      ppy -= half;
      const int ipx = (int)floorf(ppx), ipy = (int)floorf(ppy);
      if (ipx < -win || ipx >= I->w || ipy < -win || ipy >= I->h) {
        if (level == 0) status[k] = 0;
        continue;
      }
      float a = ppx - ipx, b = ppy - ipy;
      int iw00 = cv_round_f((1.f - a) * (1.f - b) * (1 << LK_W_BITS));
      int iw01 = cv_round_f(a * (1.f - b) * (1 << LK_W_BITS));
      int iw10 = cv_round_f((1.f - a) * b * (1 << LK_W_BITS));
      int iw11 = (1 << LK_W_BITS) - iw00 - iw01 - iw10;
      float sA11 = 0, sA12 = 0, sA22 = 0;
      for (int y = 0; y < win; ++y) {
        const uint8_t *src = I->img + (size_t)(y + ipy + win) * pw + (ipx + win);
        const int16_t *ds = I->deriv + ((size_t)(y + ipy + win) * pw + (ipx + win)) * 2;
        for (int x = 0; x < win; ++x) {
          const int iv = descale(src[x] * iw00 + src[x + 1] * iw01 + src[x + pw] * iw10 +
                                     src[x + pw + 1] * iw11,
                                 LK_W_BITS - 5);
          const int ix = descale(ds[2 * x] * iw00 + ds[2 * x + 2] * iw01 + ds[2 * (x + pw)] * iw10 +
                                     ds[2 * (x + pw) + 2] * iw11,
                                 LK_W_BITS);
          const int iy = descale(ds[2 * x + 1] * iw00 + ds[2 * x + 3] * iw01 +
                                     ds[2 * (x + pw) + 1] * iw10 + ds[2 * (x + pw) + 3] * iw11,
                                 LK_W_BITS);
          Iw[y * win + x] = (int16_t)iv;
          dIw[2 * (y * win + x)] = (int16_t)ix;
          dIw[2 * (y * win + x) + 1] = (int16_t)iy;
          sA11 += (float)(ix * ix);
          sA12 += (float)(ix * iy);
          sA22 += (float)(iy * iy);
        }
      }
      const float A11 = sA11 * flt_scale, A12 = sA12 * flt_scale, A22 = sA22 * flt_scale;
      float D = A11 * A22 - A12 * A12;
      const float min_eig =
          (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (2 * win * win);
      if (min_eig < min_eig_threshold || D < FLT_EPSILON) {
        if (level == 0) status[k] = 0;
        continue;
      }
      D = 1.f / D;
      npx -= half;
      npy -= half;
      float pdx = 0, pdy = 0;
      for (int j = 0; j < max_count; ++j) {
        const int inx = (int)floorf(npx), iny = (int)floorf(npy);
        if (inx < -win || inx >= J->w || iny < -win || iny >= J->h) {
          if (level == 0) status[k] = 0;
          break;
        }
        a = npx - inx;
        b = npy - iny;
        iw00 = cv_round_f((1.f - a) * (1.f - b) * (1 << LK_W_BITS));
        iw01 = cv_round_f(a * (1.f - b) * (1 << LK_W_BITS));
        iw10 = cv_round_f((1.f - a) * b * (1 << LK_W_BITS));
        iw11 = (1 << LK_W_BITS) - iw00 - iw01 - iw10;
        float sb1 = 0, sb2 = 0;
        for (int y = 0; y < win; ++y) {
          const uint8_t *jp = J->img + (size_t)(y + iny + win) * pw + (inx + win);
          for (int x = 0; x < win; ++x) {
            const int diff = descale(jp[x] * iw00 + jp[x + 1] * iw01 + jp[x + pw] * iw10 +
                                         jp[x + pw + 1] * iw11,
                                     LK_W_BITS - 5) -
                             Iw[y * win + x];
            sb1 += (float)(diff * dIw[2 * (y * win + x)]);
            sb2 += (float)(diff * dIw[2 * (y * win + x) + 1]);
          }
        }
        const float b1 = sb1 * flt_scale, b2 = sb2 * flt_scale;
        const float dx = (float)((A12 * b2 - A22 * b1) * D);
        const float dy = (float)((A12 * b1 - A11 * b2) * D);
        npx += dx;
        npy += dy;
        next_pts[2 * k] = npx + half;
        next_pts[2 * k + 1] = npy + half;
        if ((double)dx * dx + (double)dy * dy <= eps2) break;
        if (j > 0 && fabsf(dx + pdx) < 0.01 && fabsf(dy + pdy) < 0.01) {
          next_pts[2 * k] -= dx * 0.5f;
          next_pts[2 * k + 1] -= dy * 0.5f;
          break;
        }
        pdx = dx;
        pdy = dy;
      }
      /* the reference passes an err vector, so OpenCV re-checks the final window */
      if (status[k] && level == 0) {
        const float fx = next_pts[2 * k] - half, fy = next_pts[2 * k + 1] - half;
        const int inx = (int)floorf(fx), iny = (int)floorf(fy);
        if (inx < -win || inx >= J->w || iny < -win || iny >= J->h) status[k] = 0;
      }
    }
  }
  free(Iw);
  free_levels(LI, nl_i);
  free_levels(LJ, nl_j);
}

/* ======================================================================== */
/* camodocal pinhole (PinholeCamera.cc:450-510, 646-662)                    */
/* ======================================================================== */

ORA_API void ora_lift_projective(const ora_pinhole *c, double u, double v, double *ox, double *oy) {
  const double inv_fx = 1.0 / c->fx, inv_fy = 1.0 / c->fy;
  const double off_x = -c->cx / c->fx, off_y = -c->cy / c->fy;
  const double xd = inv_fx * u + off_x, yd = inv_fy * v + off_y;
  const int no_dist = (c->k1 == 0.0 && c->k2 == 0.0 && c->p1 == 0.0 && c->p2 == 0.0);
  double xu = xd, yu = yd;
  if (!no_dist) {
    for (int it = 0; it < 8; ++it) {
      const double xx = xu * xu, yy = yu * yu, xy = xu * yu;
      const double r2 = xx + yy;
      const double rad = c->k1 * r2 + c->k2 * r2 * r2;
      const double ddx = xu * rad + 2.0 * c->p1 * xy + c->p2 * (r2 + 2.0 * xx);
      const double ddy = yu * rad + 2.0 * c->p2 * xy + c->p1 * (r2 + 2.0 * yy);
      xu = xd - ddx;
      yu = yd - ddy;
    }
  }
  *ox = xu;
  *oy = yu;
}

/* ======================================================================== */
/* F-matrix RANSAC / LMedS (OpenCV fundam.cpp + ptsetreg.cpp)                */
/* ======================================================================== */

ORA_API int ora_solve_cubic(const double *c, double *roots) {
  double a0 = c[0], a1 = c[1], a2 = c[2], a3 = c[3];
  double x0 = 0, x1 = 0, x2 = 0;
  int n = 0;
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
      x1 = t0 * cos(t1 + (2. * M_PI / 3)) - t2;
      x2 = t0 * cos(t1 + (4. * M_PI / 3)) - t2;
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

/* cyclic Jacobi eigen-decomposition of a symmetric 9x9 matrix; V columns = eigenvectors */
static void jacobi_eig9(double A[9][9], double V[9][9], double w[9]) {
  for (int i = 0; i < 9; ++i)
    for (int j = 0; j < 9; ++j) V[i][j] = (i == j);
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0;
    for (int i = 0; i < 9; ++i)
      for (int j = i + 1; j < 9; ++j) off += A[i][j] * A[i][j];
    if (off < 1e-300) break;
    for (int p = 0; p < 8; ++p)
      for (int q = p + 1; q < 9; ++q) {
        if (fabs(A[p][q]) < 1e-300) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
        for (int k = 0; k < 9; ++k) {
          const double akp = A[k][p], akq = A[k][q];
          A[k][p] = cs * akp - sn * akq;
          A[k][q] = sn * akp + cs * akq;
        }
        for (int k = 0; k < 9; ++k) {
          const double apk = A[p][k], aqk = A[q][k];
          A[p][k] = cs * apk - sn * aqk;
          A[q][k] = sn * apk + cs * aqk;
        }
        for (int k = 0; k < 9; ++k) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = cs * vkp - sn * vkq;
          V[k][q] = sn * vkp + cs * vkq;
        }
      }
  }
  for (int i = 0; i < 9; ++i) w[i] = A[i][i];
}

/* fundam.cpp run7Point: normalised 7-point algorithm, up to 3 solutions.
 * The null-space basis comes from the eigenvectors of A^T A with the two
 * smallest eigenvalues (OpenCV takes the last two rows of V^T of an SVD; any
 * basis of the same 2-D null space yields the same set of F matrices). */
ORA_API int ora_run_7point(const float *m1, const float *m2, double *Fout) {
  double c1x = 0, c1y = 0, c2x = 0, c2y = 0;
  for (int i = 0; i < 7; ++i) {
    c1x += m1[2 * i];
    c1y += m1[2 * i + 1];
    c2x += m2[2 * i];
    c2y += m2[2 * i + 1];
  }
  const double t = 1. / 7;
  c1x *= t;
  c1y *= t;
  c2x *= t;
  c2y *= t;
  double s1 = 0, s2 = 0;
  for (int i = 0; i < 7; ++i) {
    const double ax = m1[2 * i] - c1x, ay = m1[2 * i + 1] - c1y;
    const double bx = m2[2 * i] - c2x, by = m2[2 * i + 1] - c2y;
    s1 += sqrt(ax * ax + ay * ay);
    s2 += sqrt(bx * bx + by * by);
  }
  s1 *= t;
  s2 *= t;
  if (s1 < FLT_EPSILON || s2 < FLT_EPSILON) return 0;
  s1 = sqrt(2.) / s1;
  s2 = sqrt(2.) / s2;

  double AtA[9][9], V[9][9], w[9];
  memset(AtA, 0, sizeof(AtA));
  for (int i = 0; i < 7; ++i) {
    const double x0 = (m1[2 * i] - c1x) * s1, y0 = (m1[2 * i + 1] - c1y) * s1;
    const double x1 = (m2[2 * i] - c2x) * s2, y1 = (m2[2 * i + 1] - c2y) * s2;
    const double r[9] = {x1 * x0, x1 * y0, x1, y1 * x0, y1 * y0, y1, x0, y0, 1};
    for (int a = 0; a < 9; ++a)
      for (int b = 0; b < 9; ++b) AtA[a][b] += r[a] * r[b];
  }
  jacobi_eig9(AtA, V, w);
  int i0 = 0, i1 = 1;
  if (w[i1] < w[i0]) {
    int tmp = i0;
    i0 = i1;
    i1 = tmp;
  }
  for (int i = 2; i < 9; ++i) {
    if (w[i] < w[i0]) {
      i1 = i0;
      i0 = i;
    } else if (w[i] < w[i1])
      i1 = i;
  }
  double f1[9], f2[9];
  for (int i = 0; i < 9; ++i) {
    f1[i] = V[i][i1];
    f2[i] = V[i][i0];
  }
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
  const int n = ora_solve_cubic(c, r);
  if (n < 1 || n > 3) return n;
  const double T1[9] = {s1, 0, -s1 * c1x, 0, s1, -s1 * c1y, 0, 0, 1};
  const double T2[9] = {s2, 0, -s2 * c2x, 0, s2, -s2 * c2y, 0, 0, 1};
  for (int k = 0; k < n; ++k) {
    double *F = Fout + 9 * k;
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
    /* F = T2^T * G * T1 */
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

typedef struct {
  uint64_t state;
} cv_rng;
static inline unsigned rng_next(cv_rng *r) {
  r->state = (uint64_t)(unsigned)r->state * 4164903690U + (unsigned)(r->state >> 32);
  return (unsigned)r->state;
}
static inline int rng_uniform(cv_rng *r, int a, int b) {
  return a == b ? a : (int)(rng_next(r) % (unsigned)(b - a) + a);
}

static int collinear_with_last(const float *m, int count) {
  const int i = count - 1;
  for (int j = 0; j < i; ++j) {
    const double dx1 = m[2 * j] - m[2 * i], dy1 = m[2 * j + 1] - m[2 * i + 1];
    for (int k = 0; k < j; ++k) {
      const double dx2 = m[2 * k] - m[2 * i], dy2 = m[2 * k + 1] - m[2 * i + 1];
      if (fabs(dx2 * dy1 - dy2 * dx1) <=
          FLT_EPSILON * (fabs(dx1) + fabs(dy1) + fabs(dx2) + fabs(dy2)))
        return 1;
    }
  }
  return 0;
}

static int get_subset7(const float *p1, const float *p2, int count, cv_rng *rng, int max_attempts,
                       float *s1, float *s2) {
  int idx[7];
  for (int iters = 0; iters < max_attempts; ++iters) {
    int i;
    for (i = 0; i < 7; ++i) {
      int cand;
      for (;;) {
        cand = rng_uniform(rng, 0, count);
        int dup = 0;
        for (int q = 0; q < i; ++q) dup |= (idx[q] == cand);
        if (!dup) break;
      }
      idx[i] = cand;
      s1[2 * i] = p1[2 * cand];
      s1[2 * i + 1] = p1[2 * cand + 1];
      s2[2 * i] = p2[2 * cand];
      s2[2 * i + 1] = p2[2 * cand + 1];
    }
    if (!collinear_with_last(s1, i) && !collinear_with_last(s2, i)) return 1;
  }
  return 0;
}

static void fm_errors(const float *m1, const float *m2, int n, const double *F, float *err) {
  for (int i = 0; i < n; ++i) {
    const double x1 = m1[2 * i], y1 = m1[2 * i + 1], x2 = m2[2 * i], y2 = m2[2 * i + 1];
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
    err[i] = (float)(e1 > e2 ? e1 : e2);
  }
}

static int ransac_update_iters(double p, double ep, int model_points, int max_iters) {
  if (p < 0) p = 0;
  if (p > 1) p = 1;
  if (ep < 0) ep = 0;
  if (ep > 1) ep = 1;
  double num = 1. - p;
  if (num < DBL_MIN) num = DBL_MIN;
  double denom = 1. - pow(1. - ep, model_points);
  if (denom < DBL_MIN) return 0;
  num = log(num);
  denom = log(denom);
  return (denom >= 0 || -num >= max_iters * (-denom)) ? max_iters : (int)lrint(num / denom);
}

static int cmp_float(const void *a, const void *b) {
  const float x = *(const float *)a, y = *(const float *)b;
  return (x > y) - (x < y);
}

ORA_API int ora_find_fundamental_mask(const float *p1, const float *p2, int n, double thresh,
                                      double confidence, int max_iters, uint8_t *mask) {
  const int MP = 7;
  if (n < 7) return 0;
  double F[27];
  if (n == 7) {
    const int k = ora_run_7point(p1, p2, F);
    memset(mask, 1, n);
    return k > 0;
  }
  if (thresh <= 0) thresh = 3;
  if (confidence < DBL_EPSILON || confidence > 1 - DBL_EPSILON) confidence = 0.99;
  cv_rng rng = {(uint64_t)-1};
  float s1[14], s2[14];
  float *err = (float *)malloc(sizeof(float) * n * 2);
  float *tmp = err + n;
  uint8_t *cur = (uint8_t *)malloc(n);
  int result = 0;
  if (n >= 15) { /* RANSACPointSetRegistrator::run */
    int niters = max_iters > 1 ? max_iters : 1, best = 0;
    const float t2 = (float)(thresh * thresh);
    for (int iter = 0; iter < niters; ++iter) {
      if (!get_subset7(p1, p2, n, &rng, 10000, s1, s2)) {
        if (iter == 0) {
          free(err);
          free(cur);
          return 0;
        }
        break;
      }
      const int nm = ora_run_7point(s1, s2, F);
      if (nm <= 0) continue;
      for (int m = 0; m < nm; ++m) {
        fm_errors(p1, p2, n, F + 9 * m, err);
        int good = 0;
        for (int i = 0; i < n; ++i) {
          cur[i] = err[i] <= t2;
          good += cur[i];
        }
        if (good > (best > MP - 1 ? best : MP - 1)) {
          memcpy(mask, cur, n);
          best = good;
          niters = ransac_update_iters(confidence, (double)(n - good) / n, MP, niters);
        }
      }
    }
    result = best > 0;
    if (!result) memset(mask, 0, n);
  } else { /* LMeDSPointSetRegistrator::run */
    double min_median = DBL_MAX, bestF[9];
    int niters = ransac_update_iters(confidence, 0.45, MP, max_iters);
    if (niters < 3) niters = 3;
    for (int iter = 0; iter < niters; ++iter) {
      if (!get_subset7(p1, p2, n, &rng, 1000, s1, s2)) {
        if (iter == 0) {
          free(err);
          free(cur);
          return 0;
        }
        break;
      }
      const int nm = ora_run_7point(s1, s2, F);
      if (nm <= 0) continue;
      for (int m = 0; m < nm; ++m) {
        fm_errors(p1, p2, n, F + 9 * m, err);
        memcpy(tmp, err, sizeof(float) * n);
        qsort(tmp, n, sizeof(float), cmp_float);
        const double median = tmp[n / 2];
        if (median < min_median) {
          min_median = median;
          memcpy(bestF, F + 9 * m, sizeof(bestF));
        }
      }
    }
    if (min_median < DBL_MAX) {
      double sigma = 2.5 * 1.4826 * (1 + 5. / (n - MP)) * sqrt(min_median);
      if (sigma < 0.001) sigma = 0.001;
      fm_errors(p1, p2, n, bestF, err);
      const float t2 = (float)(sigma * sigma);
      int good = 0;
      for (int i = 0; i < n; ++i) {
        mask[i] = err[i] <= t2;
        good += mask[i];
      }
      result = good >= MP;
    } else
      memset(mask, 0, n);
  }
  free(err);
  free(cur);
  return result;
}

/* ======================================================================== */
/* Optional image conditioning of the time surface (OpenCV imgproc, restated)   */
/* ======================================================================== */

/* cv::medianBlur(src, dst, ksize) for CV_8U (event_detector.cc:262-264): the exact median of
 * the ksize x ksize window, BORDER_REPLICATE (every OpenCV code path -- sorting network for 3/5,
 * histogram for larger -- computes this same value). */
ORA_API void ora_median_blur_u8(const uint8_t *src, int W, int H, int ksize, uint8_t *dst) {
  const int r = ksize / 2, half = (ksize * ksize) / 2;
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      int hist[256];
      memset(hist, 0, sizeof(hist));
      for (int dy = -r; dy <= r; ++dy) {
        int yy = y + dy;
        yy = yy < 0 ? 0 : (yy >= H ? H - 1 : yy);
        for (int dx = -r; dx <= r; ++dx) {
          int xx = x + dx;
          xx = xx < 0 ? 0 : (xx >= W ? W - 1 : xx);
          hist[src[(size_t)yy * W + xx]]++;
        }
      }
      int acc = 0, v = 0;
      for (; v < 256; ++v) {
        acc += hist[v];
        if (acc > half) break;
      }
      dst[(size_t)y * W + x] = (uint8_t)v;
    }
}

static int reflect101_i(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
  return i;
}

static uint8_t sat_u8_f(float v) { /* saturate_cast<uchar>(float): cvRound then clamp */
  long r = lrintf(v);
  return (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
}

/* cv::createCLAHE(clip_limit, Size(tiles, tiles))->apply(src, dst) for CV_8U
 * (feature_tracker.cpp:377-379 uses the defaults 40.0, 8x8); follows modules/imgproc/src/
 * clahe.cpp: pad right/bottom with BORDER_REFLECT_101 when either dimension is not a multiple
 * of the grid (a dimension that IS a multiple is then padded by a whole grid step, as OpenCV
 * does), per-tile clipped histogram -> LUT, bilinear blend of the four nearest tile LUTs. */
ORA_API void ora_clahe_u8(const uint8_t *src, int W, int H, double clip_limit_, int tiles,
                          uint8_t *dst) {
  const int tx_n = tiles, ty_n = tiles;
  int EW = W, EH = H;
  if (W % tx_n != 0 || H % ty_n != 0) {
    EW = W + (tx_n - W % tx_n);
    EH = H + (ty_n - H % ty_n);
  }
  const int tw = EW / tx_n, th = EH / ty_n;
  const int area = tw * th;
  const float lut_scale = (float)255 / area;
  int clip = 0;
  if (clip_limit_ > 0.0) {
    clip = (int)(clip_limit_ * area / 256);
    if (clip < 1) clip = 1;
  }
  uint8_t *lut = (uint8_t *)malloc((size_t)tx_n * ty_n * 256);
  for (int ty = 0; ty < ty_n; ++ty)
    for (int tx = 0; tx < tx_n; ++tx) {
      int hist[256];
      memset(hist, 0, sizeof(hist));
      for (int y = ty * th; y < (ty + 1) * th; ++y) {
        const int sy = reflect101_i(y, H);
        for (int x = tx * tw; x < (tx + 1) * tw; ++x) hist[src[(size_t)sy * W + reflect101_i(x, W)]]++;
      }
      if (clip > 0) {
        int clipped = 0;
        for (int i = 0; i < 256; ++i)
          if (hist[i] > clip) {
            clipped += hist[i] - clip;
            hist[i] = clip;
          }
        const int batch = clipped / 256;
        int residual = clipped - batch * 256;
        for (int i = 0; i < 256; ++i) hist[i] += batch;
        if (residual != 0) {
          int step = 256 / residual;
          if (step < 1) step = 1;
          for (int i = 0; i < 256 && residual > 0; i += step, --residual) hist[i]++;
        }
      }
      uint8_t *l = lut + ((size_t)ty * tx_n + tx) * 256;
      int sum = 0;
      for (int i = 0; i < 256; ++i) {
        sum += hist[i];
        l[i] = sat_u8_f((float)sum * lut_scale);
      }
    }
  const float inv_tw = 1.0f / tw, inv_th = 1.0f / th;
  for (int y = 0; y < H; ++y) {
    const float tyf = y * inv_th - 0.5f;
    int ty1 = (int)floorf(tyf), ty2 = ty1 + 1;
    const float ya = tyf - ty1, ya1 = 1.0f - ya;
    if (ty1 < 0) ty1 = 0;
    if (ty2 > ty_n - 1) ty2 = ty_n - 1;
    for (int x = 0; x < W; ++x) {
      const float txf = x * inv_tw - 0.5f;
      int tx1 = (int)floorf(txf), tx2 = tx1 + 1;
      const float xa = txf - tx1, xa1 = 1.0f - xa;
      if (tx1 < 0) tx1 = 0;
      if (tx2 > tx_n - 1) tx2 = tx_n - 1;
      const int v = src[(size_t)y * W + x];
      const uint8_t *p1 = lut + ((size_t)ty1 * tx_n) * 256, *p2 = lut + ((size_t)ty2 * tx_n) * 256;
      const float res = (p1[tx1 * 256 + v] * xa1 + p1[tx2 * 256 + v] * xa) * ya1 +
                        (p2[tx1 * 256 + v] * xa1 + p2[tx2 * 256 + v] * xa) * ya;
      dst[(size_t)y * W + x] = sat_u8_f(res);
    }
  }
  free(lut);
}

/* cv::normalize(src, dst, 0, 255, NORM_MINMAX) for CV_8U (feature_tracker.cpp:380-381):
 * scale = 255 / (max - min) (0 when max == min), shift = -min * scale in double, then
 * convertTo's float path: saturate_cast<uchar>(v * (float)scale + (float)shift), evaluated
 * with one rounding (fused) as OpenCV's AVX2 build does. */
ORA_API void ora_normalize_minmax_u8(const uint8_t *src, size_t n, uint8_t *dst) {
  int mn = 255, mx = 0;
  for (size_t i = 0; i < n; ++i) {
    if (src[i] < mn) mn = src[i];
    if (src[i] > mx) mx = src[i];
  }
  const double scale = 255.0 * ((double)(mx - mn) > 2.220446049250313e-16 ? 1.0 / (mx - mn) : 0.0);
  const double shift = 0.0 - mn * scale;
  const float a = (float)scale, b = (float)shift;
  for (size_t i = 0; i < n; ++i) dst[i] = sat_u8_f(fmaf((float)src[i], a, b));
}

/* the EQUALIZE branch of trackEvent (feature_tracker.cpp:375-382) */
ORA_API void ora_equalize_u8(const uint8_t *src, int W, int H, uint8_t *dst) {
  uint8_t *tmp = (uint8_t *)malloc((size_t)W * H);
  ora_clahe_u8(src, W, H, 40.0, 8, tmp);
  ora_normalize_minmax_u8(tmp, (size_t)W * H, dst);
  free(tmp);
}

/* ======================================================================== */
/* cv::goodFeaturesToTrack as trackImage calls it (feature_tracker.cpp:228):   */
/* minimum-eigenvalue corners, blockSize 3, Sobel aperture 3, quality 0.01     */
/* ======================================================================== */
/* OpenCV modules/imgproc/src/corner.cpp (cornerEigenValsVecs + calcMinEigenVal) and
 * featureselect.cpp -- sources not under /root/reference; restated from the published
 * algorithm and pinned bit for bit against cv2 4.13.0 (AVX2/FMA3 dispatch) in
 * tests/test_oracle_golden.py.  The float operation order below is the one that build
 * executes:
 *   Dx = Sobel(1,0): row pass [-1 0 1] exact, column pass fma(S0+S2, k, S1*2k), k = 1/3060;
 *   Dy = Sobel(0,1): row pass k*A, fma(2k,B,.), fma(k,C,.) in the 32-pixel vector body and plain
 *        multiply/add in the row tail (x >= 32*(W/32)), column pass S2 - S0;
 *   cov = (Dx*Dx, Dx*Dy, Dy*Dy) in f32, 3x3 unnormalised box sums in f64 (running column sums);
 *   minEig = (a + c) - sqrt((a - c)^2 + b*b) with a = cov0/2, c = cov2/2, no fma.
 * All borders are BORDER_REFLECT_101. */
static inline float sobel_src(const uint8_t *img, int W, int H, int y, int x) {
  return (float)img[(size_t)reflect101(y, H) * W + reflect101(x, W)];
}

ORA_API void ora_corner_min_eigen_val_u8(const uint8_t *img, int W, int H, float *eig) {
  const size_t N = (size_t)W * H;
  const float k1 = (float)(1.0 / 3060.0), k0 = (float)(2.0 * (1.0 / 3060.0));
  float *dx = (float *)malloc(sizeof(float) * N * 5), *dy = dx + N;
  float *cxx = dy + N, *cxy = cxx + N, *cyy = cxy + N;
  const int nv = (W / 32) * 32;
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      float d[3], s[3];
      for (int r = 0; r < 3; ++r) {
        const float a = sobel_src(img, W, H, y - 1 + r, x - 1), b = sobel_src(img, W, H, y - 1 + r, x),
                    cc = sobel_src(img, W, H, y - 1 + r, x + 1);
        d[r] = cc - a;
        if (x < nv) s[r] = fmaf(cc, k1, fmaf(b, k0, a * k1));
        else s[r] = (a * k1 + b * k0) + cc * k1;
      }
      const float gx = fmaf(d[0] + d[2], k1, d[1] * k0), gy = s[2] - s[0];
      const size_t i = (size_t)y * W + x;
      dx[i] = gx;
      dy[i] = gy;
      cxx[i] = gx * gx;
      cxy[i] = gx * gy;
      cyy[i] = gy * gy;
    }
  /* boxFilter(3x3, normalize=false) on CV_32F: row sums (S0 + S1) + S2 in f64 (RowSum, ksize 3),
   * then ONE running column sum per pixel column in f64 down the whole image (ColumnSum:
   * s = SUM + row[y+1]; out = (float)s; SUM = s - row[y-1]) -- its rounding history is part of
   * the result, so it is restated literally. */
  double *sum = (double *)calloc((size_t)W * 3, sizeof(double));
  double *ring = (double *)malloc(sizeof(double) * (size_t)W * 3 * 3); /* row sums of y-1, y, y+1 */
  const float *plane[3] = {cxx, cxy, cyy};
#define ROWSUM(dst, yy)                                                                   \
  do {                                                                                    \
    const size_t row_ = (size_t)reflect101((yy), H) * W;                                  \
    for (int ch_ = 0; ch_ < 3; ++ch_)                                                     \
      for (int x_ = 0; x_ < W; ++x_)                                                      \
        (dst)[(size_t)ch_ * W + x_] = ((double)plane[ch_][row_ + reflect101(x_ - 1, W)] +  \
                                       (double)plane[ch_][row_ + x_]) +                   \
                                      (double)plane[ch_][row_ + reflect101(x_ + 1, W)];   \
  } while (0)
  double *r0 = ring, *r1 = ring + (size_t)W * 3, *r2 = ring + (size_t)W * 6;
  ROWSUM(r0, -1);
  ROWSUM(r1, 0);
  for (size_t i = 0; i < (size_t)W * 3; ++i) sum[i] = (0.0 + r0[i]) + r1[i];
  for (int y = 0; y < H; ++y) {
    ROWSUM(r2, y + 1);
    for (int x = 0; x < W; ++x) {
      float cv[3];
      for (int ch = 0; ch < 3; ++ch) {
        const size_t i = (size_t)ch * W + x;
        const double s0 = sum[i] + r2[i];
        cv[ch] = (float)s0;
        sum[i] = s0 - r0[i];
      }
      const float a = cv[0] * 0.5f, b = cv[1], c = cv[2] * 0.5f;
      const float tt = a - c;
      eig[(size_t)y * W + x] = (a + c) - sqrtf(tt * tt + b * b);
    }
    double *tmp = r0;
    r0 = r1, r1 = r2, r2 = tmp;
  }
#undef ROWSUM
  free(sum);
  free(ring);
  free(dx);
}

typedef struct gf_cand {
  float v;
  int idx;
} gf_cand;
/* featureselect.cpp greaterThanPtr: value descending, then address (= pixel index) descending */
static int gf_cmp(const void *pa, const void *pb) {
  const gf_cand *a = (const gf_cand *)pa, *b = (const gf_cand *)pb;
  if (a->v > b->v) return -1;
  if (a->v < b->v) return 1;
  return a->idx > b->idx ? -1 : (a->idx < b->idx ? 1 : 0);
}

/* mask: NULL or W*H bytes, non-zero = allowed.  Writes <= max_corners (x, y) pairs in the
 * order cv2 returns them; returns their number. */
ORA_API int ora_good_features_to_track(const uint8_t *img, int W, int H, const uint8_t *mask,
                                       int max_corners, double quality, double min_distance,
                                       float *out_xy) {
  if (W < 3 || H < 3) return 0;
  const size_t N = (size_t)W * H;
  float *eig = (float *)malloc(sizeof(float) * N);
  ora_corner_min_eigen_val_u8(img, W, H, eig);
  /* minMaxLoc under the mask, threshold(THRESH_TOZERO) on the whole image */
  double max_val = 0;
  int any = 0;
  for (size_t i = 0; i < N; ++i)
    if (!mask || mask[i]) {
      if (!any || eig[i] > max_val) max_val = eig[i];
      any = 1;
    }
  if (!any) max_val = 0; /* minMaxLoc leaves maxVal 0 under an all-zero mask */
  const float thr = (float)(max_val * quality);
  for (size_t i = 0; i < N; ++i)
    if (!(eig[i] > thr)) eig[i] = 0.f;
  /* local maxima of the 3x3 dilation (pixels outside the image do not take part) */
  gf_cand *cand = (gf_cand *)malloc(sizeof(gf_cand) * N);
  size_t nc = 0;
  for (int y = 1; y < H - 1; ++y)
    for (int x = 1; x < W - 1; ++x) {
      const size_t i = (size_t)y * W + x;
      const float v = eig[i];
      if (v == 0.f || (mask && !mask[i])) continue;
      float m = v;
      for (int r = -1; r <= 1; ++r)
        for (int q = -1; q <= 1; ++q) {
          const float w = eig[i + (ptrdiff_t)r * W + q];
          if (w > m) m = w;
        }
      if (v == m) cand[nc].v = v, cand[nc].idx = (int)i, ++nc;
    }
  qsort(cand, nc, sizeof(gf_cand), gf_cmp);
  /* greedy minimum-distance filter in that order (the grid of the original only speeds it up) */
  int n = 0;
  const float md2 = (float)(min_distance * min_distance);
  for (size_t k = 0; k < nc && (max_corners <= 0 || n < max_corners); ++k) {
    const float x = (float)(cand[k].idx % W), y = (float)(cand[k].idx / W);
    int good = 1;
    if (min_distance >= 1)
      for (int j = 0; j < n && good; ++j) {
        const float dx = x - out_xy[2 * j], dy = y - out_xy[2 * j + 1];
        if (dx * dx + dy * dy < md2) good = 0;
      }
    if (good) out_xy[2 * n] = x, out_xy[2 * n + 1] = y, ++n;
  }
  free(cand);
  free(eig);
  return n;
}

/* ======================================================================== */
/* Whole-window tracker: FeatureTracker::trackEvent (feature_tracker.cpp:340-603) */
/* ======================================================================== */

struct ora_tracker {
  ora_config cfg;
  ora_sae *sae[2];
  uint8_t *ts[2];   /* current time surfaces (pre-equalisation) */
  uint8_t *img[2];  /* images fed to LK (== ts unless equalize) */
  uint8_t *prev_img;
  int have_prev_img;
  int n_prev;
  float *prev_pts;
  int *ids, *track_cnt;
  int n_prev_un;
  int *prev_un_ids;
  float *prev_un;
  int n_prev_un_r;
  int *prev_un_r_ids;
  float *prev_un_r;
  double prev_time;
  int next_id;
  uint8_t *mask;
  ora_lk_fn lk;
  ora_fmat_fn fm;
  ora_equalize_fn eq;
  int ransac_disabled;
  double timers[6];
};

/* FeatureTracker::n_id (feature_tracker.cpp:9): the id the next new corner gets */
ORA_API int ora_tracker_next_id(const ora_tracker *t) { return t->next_id; }

ORA_API ora_tracker *ora_tracker_create(const ora_config *cfg) {
  ora_tracker *t = (ora_tracker *)calloc(1, sizeof(ora_tracker));
  t->cfg = *cfg;
  const size_t N = (size_t)cfg->width * cfg->height;
  const int M = cfg->max_cnt > 0 ? cfg->max_cnt : 1;
  for (int c = 0; c < 2; ++c) {
    t->sae[c] = ora_sae_create(cfg->width, cfg->height);
    t->ts[c] = (uint8_t *)malloc(N);
    t->img[c] = (uint8_t *)malloc(N);
  }
  t->prev_img = (uint8_t *)malloc(N);
  t->mask = (uint8_t *)malloc(N);
  t->prev_pts = (float *)malloc(sizeof(float) * 2 * M);
  t->ids = (int *)malloc(sizeof(int) * M);
  t->track_cnt = (int *)malloc(sizeof(int) * M);
  t->prev_un_ids = (int *)malloc(sizeof(int) * M);
  t->prev_un = (float *)malloc(sizeof(float) * 2 * M);
  t->prev_un_r_ids = (int *)malloc(sizeof(int) * M);
  t->prev_un_r = (float *)malloc(sizeof(float) * 2 * M);
  return t;
}

ORA_API void ora_tracker_destroy(ora_tracker *t) {
  if (!t) return;
  for (int c = 0; c < 2; ++c) {
    ora_sae_destroy(t->sae[c]);
    free(t->ts[c]);
    free(t->img[c]);
  }
  free(t->prev_img);
  free(t->mask);
  free(t->prev_pts);
  free(t->ids);
  free(t->track_cnt);
  free(t->prev_un_ids);
  free(t->prev_un);
  free(t->prev_un_r_ids);
  free(t->prev_un_r);
  free(t);
}

ORA_API void ora_tracker_set_hooks(ora_tracker *t, ora_lk_fn lk, ora_fmat_fn fm,
                                   ora_equalize_fn eq) {
  t->lk = lk;
  t->fm = fm;
  t->eq = eq;
}
ORA_API void ora_tracker_disable_ransac(ora_tracker *t, int d) { t->ransac_disabled = d; }
ORA_API const ora_sae *ora_tracker_sae(const ora_tracker *t, int cam) { return t->sae[cam]; }
ORA_API const uint8_t *ora_tracker_time_surface(const ora_tracker *t, int cam) {
  return t->ts[cam];
}
ORA_API const uint8_t *ora_tracker_lk_image(const ora_tracker *t, int cam) { return t->img[cam]; }
ORA_API void ora_tracker_timers(const ora_tracker *t, double *o) {
  memcpy(o, t->timers, sizeof(t->timers));
}

static void run_lk(ora_tracker *t, const uint8_t *a, const uint8_t *b, const float *pp, float *np,
                   int n, uint8_t *st, int max_level, int init) {
  if (t->lk) t->lk(a, b, t->cfg.width, t->cfg.height, pp, np, n, st, max_level, init);
  else
    ora_calc_optical_flow_pyr_lk(a, b, t->cfg.width, t->cfg.height, pp, np, n, st, 21, max_level,
                                 30, 0.01, init, 1e-4);
}

/* feature_tracker.cpp:48-54 */
static int in_border(const ora_config *c, float x, float y) {
  const int ix = cv_round_f(x), iy = cv_round_f(y);
  return 1 <= ix && ix < c->width - 1 && 1 <= iy && iy < c->height - 1;
}

/* feature_tracker.cpp:1314-1319 */
static double pt_dist(float ax, float ay, float bx, float by) {
  const double dx = ax - bx, dy = ay - by;
  return sqrt(dx * dx + dy * dy);
}

/* feature_tracker.cpp:1004-1045 with the id->point map as two flat arrays */
static void velocity(const int *ids, const float *un, int n, const int *pids, const float *pun,
                     int np, double dt, float *vx, float *vy) {
  for (int i = 0; i < n; ++i) {
    vx[i] = vy[i] = 0.f;
    if (np == 0 || ids[i] == -1) continue;
    for (int j = 0; j < np; ++j)
      if (pids[j] == ids[i]) {
        const double ax = (un[2 * i] - pun[2 * j]) / dt;
        const double ay = (un[2 * i + 1] - pun[2 * j + 1]) / dt;
        vx[i] = (float)ax;
        vy[i] = (float)ay;
        break;
      }
  }
}

static int track_tail(ora_tracker *t, double cur_time, int pub, int image_mode, int have_right,
                      const uint16_t *lx, const uint16_t *ly, const double *lt,
                      const uint8_t *lp, size_t nl, ora_tracks *out);

/* FeatureTracker::trackImage (feature_tracker.cpp:164-338): cfg.max_cnt / cfg.min_dist play
 * MAX_CNT_IMG / MIN_DIST_IMG, cfg.width / height play COL / ROW.  right == NULL: img_right.empty() */
ORA_API int ora_tracker_track_image(ora_tracker *t, double cur_time, const uint8_t *left,
                                    const uint8_t *right, int pub, ora_tracks *out) {
  const size_t N = (size_t)t->cfg.width * t->cfg.height;
  memcpy(t->img[0], left, N);
  memcpy(t->ts[0], left, N);
  if (right) memcpy(t->img[1], right, N), memcpy(t->ts[1], right, N);
  if (!t->have_prev_img) { /* :172-174 */
    memcpy(t->prev_img, t->img[0], N);
    t->have_prev_img = 1;
  }
  return track_tail(t, cur_time, pub, 1, right != NULL, NULL, NULL, NULL, NULL, 0, out);
}

ORA_API int ora_tracker_track(ora_tracker *t, double cur_time, const uint16_t *lx,
                              const uint16_t *ly, const double *lt, const uint8_t *lp, size_t nl,
                              const uint16_t *rx, const uint16_t *ry, const double *rt,
                              const uint8_t *rp, size_t nr, int pub, ora_tracks *out) {
  return ora_tracker_track_mc(t, cur_time, lx, ly, lt, lp, nl, rx, ry, rt, rp, nr, pub, NULL, out);
}

/* mc != NULL: FeatureTracker::trackEvent(..., measurements) (feature_tracker.cpp:605-877) */
ORA_API int ora_tracker_track_mc(ora_tracker *t, double cur_time, const uint16_t *lx,
                                 const uint16_t *ly, const double *lt, const uint8_t *lp,
                                 size_t nl, const uint16_t *rx, const uint16_t *ry,
                                 const double *rt, const uint8_t *rp, size_t nr, int pub,
                                 const ora_motion *mc, ora_tracks *out) {
  const ora_config *c = &t->cfg;
  const int W = c->width, H = c->height, M = c->max_cnt;
  const size_t N = (size_t)W * H;
  double t0 = now_sec(), t1;

  /* HOT LOOP A (feature_tracker.cpp:356-362) */
  if (mc && nl > 0) {
    ora_sae_update_mc(t->sae[0], lx, ly, lt, lp, nl, c->feature_filter_threshold, mc, lt[0]);
    ora_sae_update_mc(t->sae[1], rx, ry, rt, rp, nr, c->feature_filter_threshold, mc, lt[0]);
  } else {
    ora_sae_update(t->sae[0], lx, ly, lt, lp, nl, c->feature_filter_threshold);
    ora_sae_update(t->sae[1], rx, ry, rt, rp, nr, c->feature_filter_threshold);
  }
  t1 = now_sec();
  t->timers[0] += t1 - t0;
  t0 = t1;
  /* HOT LOOP B (:367-368) */
  for (int cam = 0; cam < 2; ++cam) {
    ora_time_surface(t->sae[cam], cur_time, c->decay_ms, c->ignore_polarity, t->ts[cam]);
    if (c->median_blur_kernel_size > 0) { /* event_detector.cc:262-264 */
      ora_median_blur_u8(t->ts[cam], W, H, 2 * c->median_blur_kernel_size + 1, t->img[cam]);
      memcpy(t->ts[cam], t->img[cam], N);
    }
    if (c->equalize) { /* feature_tracker.cpp:375-382 */
      if (t->eq) t->eq(t->ts[cam], W, H, t->img[cam]);
      else ora_equalize_u8(t->ts[cam], W, H, t->img[cam]);
    } else
      memcpy(t->img[cam], t->ts[cam], N);
  }
  if (!t->have_prev_img) {
    memcpy(t->prev_img, t->img[0], N);
    t->have_prev_img = 1;
  }
  t1 = now_sec();
  t->timers[1] += t1 - t0;
  return track_tail(t, cur_time, pub, 0, 1, lx, ly, lt, lp, nl, out);
}

/* Everything of trackEvent behind the images (feature_tracker.cpp:405-603), and -- with
 * image_mode -- of trackImage (:178-338): there the backward temporal LK is a full 4-level
 * call without initial flow (:190), there is no F-RANSAC, and new points come from
 * goodFeaturesToTrack under the Image_setMask mask (:213-237). */
static int track_tail(ora_tracker *t, double cur_time, int pub, int image_mode, int have_right,
                      const uint16_t *lx, const uint16_t *ly, const double *lt,
                      const uint8_t *lp, size_t nl, ora_tracks *out) {
  const ora_config *c = &t->cfg;
  const int W = c->width, H = c->height, M = c->max_cnt;
  const size_t N = (size_t)W * H;
  double t0 = now_sec(), t1;

  float *cur = (float *)malloc(sizeof(float) * 2 * (M + 1));
  float *rev = (float *)malloc(sizeof(float) * 2 * (M + 1));
  uint8_t *st = (uint8_t *)malloc(M + 1), *st2 = (uint8_t *)malloc(M + 1);
  int n = 0;
  out->n_prev = t->n_prev;

  /* temporal tracking (:405-437) */
  if (t->n_prev > 0) {
    n = t->n_prev;
    run_lk(t, t->prev_img, t->img[0], t->prev_pts, cur, n, st, 3, 0);
    if (c->flow_back) {
      memcpy(rev, t->prev_pts, sizeof(float) * 2 * n);
      if (image_mode) run_lk(t, t->img[0], t->prev_img, cur, rev, n, st2, 3, 0);
      else run_lk(t, t->img[0], t->prev_img, cur, rev, n, st2, 1, 1);
      for (int i = 0; i < n; ++i)
        st[i] = st[i] && st2[i] &&
                pt_dist(t->prev_pts[2 * i], t->prev_pts[2 * i + 1], rev[2 * i], rev[2 * i + 1]) <= 0.5;
    }
    for (int i = 0; i < n; ++i)
      if (st[i] && !in_border(c, cur[2 * i], cur[2 * i + 1])) st[i] = 0;
    int m = 0;
    for (int i = 0; i < n; ++i)
      if (st[i]) {
        t->prev_pts[2 * m] = t->prev_pts[2 * i];
        t->prev_pts[2 * m + 1] = t->prev_pts[2 * i + 1];
        cur[2 * m] = cur[2 * i];
        cur[2 * m + 1] = cur[2 * i + 1];
        t->ids[m] = t->ids[i];
        t->track_cnt[m] = t->track_cnt[i];
        ++m;
      }
    n = m;
  }
  for (int i = 0; i < n; ++i) t->track_cnt[i]++;
  out->n_after_temporal = n;
  t1 = now_sec();
  t->timers[2] += t1 - t0;
  t0 = t1;

  out->n_after_ransac = n;
  out->n_after_mask = n;
  out->n_new = 0;
  if (pub) {
    /* rejectWithF_event (:910-947) */
    if (n >= 8 && !t->ransac_disabled && !image_mode) {
      float *a = (float *)malloc(sizeof(float) * 4 * n), *b = a + 2 * n;
      for (int i = 0; i < n; ++i) {
        double x, y;
        ora_lift_projective(&c->cam[0], t->prev_pts[2 * i], t->prev_pts[2 * i + 1], &x, &y);
        a[2 * i] = (float)(c->focal_length * x + W / 2.0);
        a[2 * i + 1] = (float)(c->focal_length * y + H / 2.0);
        ora_lift_projective(&c->cam[0], cur[2 * i], cur[2 * i + 1], &x, &y);
        b[2 * i] = (float)(c->focal_length * x + W / 2.0);
        b[2 * i + 1] = (float)(c->focal_length * y + H / 2.0);
      }
      memset(st, 0, n);
      if (t->fm) t->fm(a, b, n, c->f_threshold, st);
      else
        ora_find_fundamental_mask(a, b, n, c->f_threshold, 0.99, 1000, st);
      free(a);
      int m = 0;
      for (int i = 0; i < n; ++i)
        if (st[i]) {
          cur[2 * m] = cur[2 * i];
          cur[2 * m + 1] = cur[2 * i + 1];
          t->ids[m] = t->ids[i];
          t->track_cnt[m] = t->track_cnt[i];
          ++m;
        }
      n = m;
    }
    out->n_after_ransac = n;
    n = ora_set_mask(W, H, c->min_dist, n, cur, t->ids, t->track_cnt, t->mask);
    out->n_after_mask = n;
    const int want = M - n;
    if (want > 0) {
      float *np = (float *)malloc(sizeof(float) * 2 * want);
      int k;
      if (image_mode) { /* goodFeaturesToTrack wants 255 = allowed; ora_set_mask marks blocked */
        uint8_t *allow = (uint8_t *)malloc(N);
        for (size_t i = 0; i < N; ++i) allow[i] = t->mask[i] ? 0 : 255;
        k = ora_good_features_to_track(t->img[0], W, H, allow, want, 0.01, (double)c->min_dist, np);
        free(allow);
      } else
        k = ora_features_to_track(t->sae[0], lx, ly, lt, lp, nl, want, c->min_dist, t->mask,
                                  t->ts[0], c->ts_lk_threshold, c->feature_filter_threshold, np,
                                  NULL);
      for (int i = 0; i < k; ++i) {
        cur[2 * n] = np[2 * i];
        cur[2 * n + 1] = np[2 * i + 1];
        t->ids[n] = t->next_id++;
        t->track_cnt[n] = 1;
        ++n;
      }
      out->n_new = k;
      free(np);
    }
  }
  t1 = now_sec();
  t->timers[3] += t1 - t0;
  t0 = t1;

  /* undistort + velocity, left (:470-473) */
  const double dt = cur_time - t->prev_time;
  float *un = (float *)malloc(sizeof(float) * 2 * (M + 1));
  for (int i = 0; i < n; ++i) {
    double x, y;
    ora_lift_projective(&c->cam[0], cur[2 * i], cur[2 * i + 1], &x, &y);
    un[2 * i] = (float)x;
    un[2 * i + 1] = (float)y;
  }
  velocity(t->ids, un, n, t->prev_un_ids, t->prev_un, t->n_prev_un, dt, out->vx, out->vy);
  out->n_left = n;
  for (int i = 0; i < n; ++i) {
    out->id[i] = t->ids[i];
    out->track_cnt[i] = t->track_cnt[i];
    out->u[i] = cur[2 * i];
    out->v[i] = cur[2 * i + 1];
    out->un_x[i] = un[2 * i];
    out->un_y[i] = un[2 * i + 1];
  }
  t1 = now_sec();
  t->timers[5] += t1 - t0;
  t0 = t1;

  /* stereo matching (:475-575) */
  int nrgt = 0;
  float *unr = (float *)malloc(sizeof(float) * 2 * (M + 1));
  int *idr = (int *)malloc(sizeof(int) * (M + 1));
  if (n > 0 && have_right) {
    float *rpts = (float *)malloc(sizeof(float) * 2 * n);
    run_lk(t, t->img[0], t->img[1], cur, rpts, n, st, 3, 0);
    if (c->flow_back) {
      run_lk(t, t->img[1], t->img[0], rpts, rev, n, st2, 3, 0);
      for (int i = 0; i < n; ++i)
        st[i] = st[i] && st2[i] && in_border(c, rpts[2 * i], rpts[2 * i + 1]) &&
                pt_dist(cur[2 * i], cur[2 * i + 1], rev[2 * i], rev[2 * i + 1]) <= 0.5;
    }
    for (int i = 0; i < n; ++i)
      if (st[i]) {
        out->ru[nrgt] = rpts[2 * i];
        out->rv[nrgt] = rpts[2 * i + 1];
        idr[nrgt] = t->ids[i];
        ++nrgt;
      }
    free(rpts);
    t1 = now_sec();
    t->timers[4] += t1 - t0;
    t0 = t1;
    for (int i = 0; i < nrgt; ++i) {
      double x, y;
      ora_lift_projective(&c->cam[1], out->ru[i], out->rv[i], &x, &y);
      unr[2 * i] = (float)x;
      unr[2 * i + 1] = (float)y;
      out->run_x[i] = (float)x;
      out->run_y[i] = (float)y;
      out->id_right[i] = idr[i];
    }
    velocity(idr, unr, nrgt, t->prev_un_r_ids, t->prev_un_r, t->n_prev_un_r, dt, out->rvx,
             out->rvy);
  }
  out->n_right = nrgt;
  /* prev_un_right_pts_map = cur_un_right_pts_map (:574) -- empty when no left points; the
   * whole block is skipped when trackImage gets no right image (:245, :322), which leaves the
   * map as it was */
  if (have_right) {
    t->n_prev_un_r = nrgt;
    memcpy(t->prev_un_r_ids, idr, sizeof(int) * nrgt);
    memcpy(t->prev_un_r, unr, sizeof(float) * 2 * nrgt);
  }

  /* state roll (:585-590) */
  memcpy(t->prev_img, t->img[0], N);
  t->n_prev = n;
  memcpy(t->prev_pts, cur, sizeof(float) * 2 * n);
  t->n_prev_un = n;
  memcpy(t->prev_un_ids, t->ids, sizeof(int) * n);
  memcpy(t->prev_un, un, sizeof(float) * 2 * n);
  t->prev_time = cur_time;
  t1 = now_sec();
  t->timers[5] += t1 - t0;

  free(cur);
  free(rev);
  free(st);
  free(st2);
  free(un);
  free(unr);
  free(idr);
  return 0;
}
