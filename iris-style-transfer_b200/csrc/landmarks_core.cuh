// Core of the eye-landmark path (models/gaze_estimators/gaze_estimators.py:55-178): the pieces of
// OpenCV's findContours(RETR_EXTERNAL, CHAIN_APPROX_SIMPLE) / contourArea / fitEllipse that landmarks.cu runs per
// (frame, class) CTA, written as host+device functions over BIT PLANES so that tests/test_landmarks_core_host.py can
// compile this very header with g++ and pin it against cv2 on the CPU (the kernels add only the block-level glue).
//
// Planes: one bit per pixel, rows of `Ww` 32-bit words, with a one-pixel zero frame like OpenCV's copyMakeBorder: pixel
// (x, y) of the frame is bit (x + 1) of row (y + 1); "padded" coordinates below include that offset.
//   F  foreground (the pixel belongs to the class)
//   M  marked: a border trace visited the pixel (OpenCV's value 2 or -126 instead of 1)
//   N  the mark is NEGATIVE (-126): the pixel's right-hand neighbour was examined as background during a trace
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define ISX_HD __host__ __device__ __forceinline__
#else
#define ISX_HD inline
#endif

namespace isx_lm {

ISX_HD int lm_clz(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return __clz(static_cast<int>(v));
#else
  return v ? __builtin_clz(v) : 32;
#endif
}
ISX_HD int lm_ffs(uint32_t v) {   // index of the lowest set bit, v != 0
#if defined(__CUDA_ARCH__)
  return __ffs(static_cast<int>(v)) - 1;
#else
  return __builtin_ctz(v);
#endif
}
ISX_HD int lm_bit(const uint32_t* P, int Ww, int x, int y) { return (P[y * Ww + (x >> 5)] >> (x & 31)) & 1u; }
ISX_HD void lm_set(uint32_t* P, int Ww, int x, int y) { P[y * Ww + (x >> 5)] |= 1u << (x & 31); }

// Raster scan of ONE padded row (cvFindNextContour, mode RETR_EXTERNAL): the first pixel with x > x_after that starts an
// outer border -- value 1 (foreground, unmarked), left neighbour 0 -- and is not rejected: OpenCV skips the start when the
// last marked pixel met on this row (`lnbd`) carries a POSITIVE mark (the scan is inside a traced border).  -1: none.
// x_after > 0 must be a MARKED pixel (the start of the trace that has just run): every candidate right of it finds a mark in
// its own word or in a later one, so the words left of x_after need not be read.
ISX_HD int lm_row_first_start(const uint32_t* F, const uint32_t* M, const uint32_t* N, int Ww, int y, int x_after) {
  const uint32_t* f = F + y * Ww;
  const uint32_t* m = M + y * Ww;
  const uint32_t* n = N + y * Ww;
  int inside = 0;          // 1: the last marked pixel so far is positive
  uint32_t prev_top = 0;   // foreground bit of the pixel left of the current word
  for (int w = x_after > 0 ? (x_after >> 5) : 0; w < Ww; ++w) {
    const uint32_t fw = f[w];
    if (fw == 0) { prev_top = 0; continue; }   // no foreground: no marks either
    const uint32_t mw = m[w], nw = n[w];
    uint32_t cand = fw & ~mw & ~((fw << 1) | prev_top);
    const int lo = x_after - w * 32;           // candidates need bit index > lo
    if (lo >= 31) cand = 0;
    else if (lo >= 0) cand &= ~((2u << lo) - 1u);
    while (cand) {
      const int k = lm_ffs(cand);
      const uint32_t left = mw & ((1u << k) - 1u);   // marked pixels of this word left of the candidate
      int in = inside;
      if (left) in = ((nw >> (31 - lm_clz(left))) & 1u) ? 0 : 1;
      if (!in) return w * 32 + k;
      cand &= cand - 1;
    }
    if (mw) inside = ((nw >> (31 - lm_clz(mw))) & 1u) ? 0 : 1;
    prev_top = fw >> 31;
  }
  return -1;
}

// ---- the same row scan split by WORD, for a warp that gives every lane one word (landmarks.cu lm_row_first_start_warp) ----
// The serial scan carries two things from word to word: the foreground bit left of the word (`prev_top`) and whether the last
// marked pixel so far is positive (`inside`).  Per lane: the candidate mask of its word, whether its word's last mark is
// positive, and -- given the carried `inside` -- its first accepted candidate.  The warp combines them with one shuffle and two
// ballots; tests/landmarks_core_host.cpp combines them with loops and checks the result against lm_row_first_start.
ISX_HD uint32_t lm_word_candidates(uint32_t fw, uint32_t mw, uint32_t prev_top, int x_after, int w) {
  uint32_t cand = fw & ~mw & ~((fw << 1) | prev_top);
  const int lo = x_after - w * 32;   // candidates need bit index > lo
  if (lo >= 31) cand = 0u;
  else if (lo >= 0) cand &= ~((2u << lo) - 1u);
  return cand;
}
ISX_HD bool lm_word_last_mark_positive(uint32_t mw, uint32_t nw) { return mw != 0u && !((nw >> (31 - lm_clz(mw))) & 1u); }
ISX_HD int lm_word_first_accepted(uint32_t cand, uint32_t mw, uint32_t nw, int inside_in) {   // bit index, -1: none
  while (cand) {
    const int k = lm_ffs(cand);
    const uint32_t left = mw & ((1u << k) - 1u);
    const int in = left ? (((nw >> (31 - lm_clz(left))) & 1u) ? 0 : 1) : inside_in;
    if (!in) return k;
    cand &= cand - 1u;
  }
  return -1;
}

struct LmTrace {
  int n;              // points emitted (CHAIN_APPROX_SIMPLE)
  int y_max;          // last padded row the border touches (the first is the start's row: a start is the border's raster-first pixel)
  long long cross;    // sum over the closed polygon of prev.x * p.y - prev.y * p.x   (contourArea = |cross| / 2)
};

// chain code -> step (0 = east, then counter-clockwise on the screen: NE, N, NW, W, SW, S, SE), OpenCV's icvCodeDeltas
ISX_HD int lm_dx(int s) { return (s == 0 || s == 1 || s == 7) ? 1 : (s >= 3 && s <= 5) ? -1 : 0; }
ISX_HD int lm_dy(int s) { return (s >= 1 && s <= 3) ? -1 : (s >= 5 && s <= 7) ? 1 : 0; }

// the eight neighbours of padded (x, y) at once: bit s of the result = foreground bit of the neighbour in chain direction s.
// Three independent pairs of word loads instead of up to eight dependent single-bit probes per border step.
ISX_HD uint32_t lm_neighbours(const uint32_t* F, int Ww, int x, int y) {
  const int xl = x - 1, w = xl >> 5, sh = xl & 31, up = sh > 29 ? 1 : 0;   // bits x-1 .. x+1 straddle two words when sh > 29
  const uint32_t* r = F + (y - 1) * Ww + w;
  const uint32_t top = static_cast<uint32_t>(((static_cast<uint64_t>(r[up]) << 32 | r[0]) >> sh) & 7u);
  const uint32_t mid = static_cast<uint32_t>(((static_cast<uint64_t>(r[Ww + up]) << 32 | r[Ww]) >> sh) & 7u);
  const uint32_t bot = static_cast<uint32_t>(((static_cast<uint64_t>(r[2 * Ww + up]) << 32 | r[2 * Ww]) >> sh) & 7u);
  return (mid >> 2) | ((top >> 2) << 1) | (((top >> 1) & 1u) << 2) | ((top & 1u) << 3) | ((mid & 1u) << 4) | ((bot & 1u) << 5) |
         (((bot >> 1) & 1u) << 6) | ((bot >> 2) << 7);
}

// icvFetchContour for the outer border that starts at padded (x0, y0): marks M / N, emits the points where the chain
// direction changes as (x | y << 16) in FRAME coordinates into pts[0..cap) (points beyond cap are counted, not stored).
ISX_HD LmTrace lm_trace(const uint32_t* F, uint32_t* M, uint32_t* N, int Ww, int x0, int y0, uint32_t* pts, int cap) {
  LmTrace r;
  r.n = 0;
  r.cross = 0;
  r.y_max = y0;
  int s = 4, s_end = 4;
  uint32_t nb = lm_neighbours(F, Ww, x0, y0);
  if (nb != 0) {      // OpenCV probes 3, 2, 1, 0, 7, 6, 5, 4 and stops at the first foreground neighbour (4 = west is background)
    do {
      s = (s - 1) & 7;
    } while (!((nb >> s) & 1u) && s != s_end);
  }
  if (s == s_end) {   // single-pixel domain
    lm_set(M, Ww, x0, y0);
    lm_set(N, Ww, x0, y0);
    if (cap > 0) pts[0] = static_cast<uint32_t>(x0 - 1) | (static_cast<uint32_t>(y0 - 1) << 16);
    r.n = 1;
    return r;
  }
  const int x1 = x0 + lm_dx(s), y1 = y0 + lm_dy(s);
  int x3 = x0, y3 = y0;
  int prev_s = s ^ 4;
  int fx = 0, fy = 0, lx = 0, ly = 0;   // first / last emitted point
  for (;;) {
    s_end = s;
    // OpenCV examines s_end + 1, s_end + 2, ... (at most up to 15) until a nonzero neighbour: the lowest set bit of the
    // direction byte, doubled to 16 bits and shifted so that bit 0 is direction s_end + 1
    const uint32_t ahead = ((nb | (nb << 8)) & 0xFFFFu) >> (s_end + 1);
    s = ahead ? s_end + 1 + lm_ffs(ahead) : 15;
    const int x4 = x3 + lm_dx(s & 7), y4 = y3 + lm_dy(s & 7);
    s &= 7;
    lm_set(M, Ww, x3, y3);   // 1 -> 2; an earlier negative mark stays
    if (static_cast<unsigned>(s - 1) < static_cast<unsigned>(s_end)) lm_set(N, Ww, x3, y3);   // the east neighbour was examined as background: -126
    if (s != prev_s) {
      const int px = x3 - 1, py = y3 - 1;
      if (r.n < cap) pts[r.n] = static_cast<uint32_t>(px) | (static_cast<uint32_t>(py) << 16);
      if (r.n == 0) { fx = px; fy = py; }
      else r.cross += static_cast<long long>(lx) * py - static_cast<long long>(ly) * px;
      lx = px; ly = py;
      ++r.n;
      prev_s = s;
    }
    if (x4 == x0 && y4 == y0 && x3 == x1 && y3 == y1) break;
    x3 = x4; y3 = y4;
    if (y3 > r.y_max) r.y_max = y3;
    nb = lm_neighbours(F, Ww, x3, y3);
    s = (s + 4) & 7;
  }
  r.cross += static_cast<long long>(lx) * fy - static_cast<long long>(ly) * fx;
  return r;
}

// ---- fitEllipse (shapedescr.cpp fitEllipseNoDirect) as sums over the points ----
// OpenCV solves both least-squares systems with an SVD of the n x 5 (n x 3) matrix in double.  Here the systems are reduced
// to their normal equations, which squares the condition number -- harmless for a pupil or an iris (cond ~ 1e2), fatal in
// plain double for a near-degenerate speck (six points on almost a line pair: cond 3e6, cond^2 7e12).  So the products are
// accumulated and the 5 x 5 / 3 x 3 systems solved in DOUBLE-DOUBLE arithmetic (two_sum / fma two_prod, ~31 digits): the
// result is the exact least-squares solution of OpenCV's double matrix to far below float32 rounding whenever OpenCV itself
// does not declare the system rank deficient.
struct dd {
  double hi, lo;
};
ISX_HD dd dd_make(double a) { dd r; r.hi = a; r.lo = 0.0; return r; }
ISX_HD dd dd_quick(double a, double b) {   // |a| >= |b|
  dd r;
#if defined(__CUDA_ARCH__)
  r.hi = __dadd_rn(a, b);
  r.lo = __dsub_rn(b, __dsub_rn(r.hi, a));
#else
  r.hi = a + b;
  r.lo = b - (r.hi - a);
#endif
  return r;
}
ISX_HD dd dd_two_sum(double a, double b) {
  dd r;
#if defined(__CUDA_ARCH__)
  r.hi = __dadd_rn(a, b);
  const double bb = __dsub_rn(r.hi, a);
  r.lo = __dadd_rn(__dsub_rn(a, __dsub_rn(r.hi, bb)), __dsub_rn(b, bb));
#else
  r.hi = a + b;
  const double bb = r.hi - a;
  r.lo = (a - (r.hi - bb)) + (b - bb);
#endif
  return r;
}
ISX_HD dd dd_two_prod(double a, double b) {
  dd r;
#if defined(__CUDA_ARCH__)
  r.hi = __dmul_rn(a, b);   // never contracted: with nvcc's default fmad the product would fuse into the two_sum that consumes
                            // it (hi = fma(a, b, c) instead of fl(c + fl(ab))) and the error terms would no longer be exact
#else
  r.hi = a * b;
#endif
  r.lo = fma(a, b, -r.hi);
  return r;
}
ISX_HD dd dd_add(dd a, dd b) {
  dd s = dd_two_sum(a.hi, b.hi);
  const dd t = dd_two_sum(a.lo, b.lo);
  s.lo += t.hi;
  s = dd_quick(s.hi, s.lo);
  s.lo += t.lo;
  return dd_quick(s.hi, s.lo);
}
ISX_HD dd dd_neg(dd a) { dd r; r.hi = -a.hi; r.lo = -a.lo; return r; }
ISX_HD dd dd_sub(dd a, dd b) { return dd_add(a, dd_neg(b)); }
ISX_HD dd dd_mul(dd a, dd b) {
  dd p = dd_two_prod(a.hi, b.hi);
  p.lo += a.hi * b.lo + a.lo * b.hi;
  return dd_quick(p.hi, p.lo);
}
ISX_HD dd dd_div(dd a, dd b) {
  const double q1 = a.hi / b.hi;
  dd r = dd_sub(a, dd_mul(b, dd_make(q1)));
  const double q2 = r.hi / b.hi;
  r = dd_sub(r, dd_mul(b, dd_make(q2)));
  const double q3 = r.hi / b.hi;
  return dd_add(dd_quick(q1, q2), dd_make(q3));
}

// first system: rows a = (-px^2, -py^2, -px py, px, py), right-hand side 10000; px, py = (p - c) * scale in double
enum { kLmSums1 = 20, kLmSums2 = 9 };
ISX_HD void lm_acc1(double px, double py, dd* acc) {   // acc[0..14] upper triangle of sum a a^T (row-major), acc[15..19] sum 10000 a
  const double a[5] = {-px * px, -py * py, -px * py, px, py};   // OpenCV's matrix entries, rounded to double like there
  int k = 0;
  for (int i = 0; i < 5; ++i)
    for (int j = i; j < 5; ++j) { acc[k] = dd_add(acc[k], dd_two_prod(a[i], a[j])); ++k; }
  for (int i = 0; i < 5; ++i) acc[15 + i] = dd_add(acc[15 + i], dd_two_prod(10000.0, a[i]));
}
// second system (centre fixed): rows a = ((px-rx)^2, (py-ry)^2, (px-rx)(py-ry)), right-hand side 1
ISX_HD void lm_acc2(double px, double py, double rx, double ry, dd* acc) {
  const double u = px - rx, v = py - ry;
  const double a[3] = {u * u, v * v, u * v};
  int k = 0;
  for (int i = 0; i < 3; ++i)
    for (int j = i; j < 3; ++j) { acc[k] = dd_add(acc[k], dd_two_prod(a[i], a[j])); ++k; }
  for (int i = 0; i < 3; ++i) acc[6 + i] = dd_add(acc[6 + i], dd_make(a[i]));
}

// symmetric n x n system (upper triangle in `tri`, row-major) by Gaussian elimination with partial pivoting in double-double;
// false = a pivot is exactly zero.  *min_pivot = smallest |pivot| relative to the largest diagonal entry: a healthy system
// keeps it above ~1e-14 (cond^2 of a near-degenerate speck), an exactly rank-deficient one -- e.g. a speck whose points are
// symmetric about the fitted centre, so that two columns of the refit coincide -- leaves double-double noise (< 1e-28).
// There OpenCV's answer is the minimum-norm solution its SVD back-substitution picks after dropping the zero singular value;
// the caller flags such systems (kLmPivotFloor) instead of imitating that.  *det_abs = |product of the pivots| = det of the matrix.
#define kLmPivotFloor 1e-22
template <int NN>
ISX_HD bool lm_solve_sym(const dd* tri, const dd* rhs, double* x, double* min_pivot, double* det_abs = nullptr) {
  dd A[NN][NN + 1];
  dd xs[NN];
  int k = 0;
  for (int i = 0; i < NN; ++i)
    for (int j = i; j < NN; ++j) { A[i][j] = tri[k]; A[j][i] = tri[k]; ++k; }
  for (int i = 0; i < NN; ++i) A[i][NN] = rhs[i];
  double dmax = 0.0, det = 1.0;
  for (int i = 0; i < NN; ++i) dmax = fmax(dmax, fabs(A[i][i].hi));
  *min_pivot = 1.0;
  for (int c = 0; c < NN; ++c) {
    int p = c;
    for (int r = c + 1; r < NN; ++r)
      if (fabs(A[r][c].hi) > fabs(A[p][c].hi)) p = r;
    if (A[p][c].hi == 0.0) { *min_pivot = 0.0; if (det_abs) *det_abs = 0.0; return false; }
    *min_pivot = fmin(*min_pivot, fabs(A[p][c].hi) / dmax);
    det *= fabs(A[p][c].hi);
    if (p != c)
      for (int j = 0; j <= NN; ++j) { const dd t = A[c][j]; A[c][j] = A[p][j]; A[p][j] = t; }
    for (int r = c + 1; r < NN; ++r) {
      const dd f = dd_div(A[r][c], A[c][c]);
      for (int j = c; j <= NN; ++j) A[r][j] = dd_sub(A[r][j], dd_mul(f, A[c][j]));
    }
  }
  for (int i = NN - 1; i >= 0; --i) {
    dd v = A[i][NN];
    for (int j = i + 1; j < NN; ++j) v = dd_sub(v, dd_mul(A[i][j], xs[j]));
    xs[i] = dd_div(v, A[i][i]);
    x[i] = xs[i].hi;
  }
  if (det_abs) *det_abs = det;
  return true;
}

// OpenCV's rank test can only fire when lambda_min / lambda_max < FLT_EPSILON^2 = 1.4e-14.  A rigorous lower bound that costs
// nothing: the other four eigenvalues multiply to at most (trace / 4)^4 (AM-GM) and lambda_max <= trace, so
// lambda_min / lambda_max >= det / ((trace / 4)^4 * trace) with det = the product of the pivots of the elimination that has
// just run.  A pupil or an iris is at ~1e-6: the Jacobi sweeps below are needed for near-degenerate specks only.
ISX_HD bool lm_rank_surely_full(const dd* tri, double det_abs) {
  double tr = 0.0;
  int k = 0;
  for (int i = 0; i < 5; ++i) { tr += tri[k].hi; k += 5 - i; }
  const double q = tr / 4.0;
  return det_abs > 1e-12 * (q * q * q * q * tr);   // 70 x the threshold: rounding of det and trace cannot matter
}

// eigenvalue range of the 5 x 5 normal matrix (cyclic Jacobi): OpenCV's rank test sigma_max * FLT_EPSILON > sigma_min on
// the singular values of the system matrix = lambda_max * FLT_EPSILON^2 > lambda_min here
ISX_HD bool lm_rank_deficient(const dd* tri) {
  double A[5][5];
  int k = 0;
  for (int i = 0; i < 5; ++i)
    for (int j = i; j < 5; ++j) { A[i][j] = tri[k].hi; A[j][i] = tri[k].hi; ++k; }
  for (int sweep = 0; sweep < 12; ++sweep) {
    double off = 0;
    for (int i = 0; i < 5; ++i)
      for (int j = i + 1; j < 5; ++j) off += A[i][j] * A[i][j];
    if (off == 0.0) break;
    for (int p = 0; p < 4; ++p)
      for (int q = p + 1; q < 5; ++q) {
        if (A[p][q] == 0.0) continue;
        const double th = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        const double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
        for (int r = 0; r < 5; ++r) {   // columns p, q
          const double arp = A[r][p], arq = A[r][q];
          A[r][p] = c * arp - sn * arq;
          A[r][q] = sn * arp + c * arq;
        }
        for (int r = 0; r < 5; ++r) {   // rows p, q
          const double apr = A[p][r], aqr = A[q][r];
          A[p][r] = c * apr - sn * aqr;
          A[q][r] = sn * apr + c * aqr;
        }
      }
  }
  double lo = A[0][0], hi = A[0][0];
  for (int i = 1; i < 5; ++i) { lo = fmin(lo, A[i][i]); hi = fmax(hi, A[i][i]); }
  const double eps = 1.1920928955078125e-07;
  return hi * eps * eps > lo;
}

// centre of the conic of the first fit: [2A C; C 2B] r = (D, E)   (OpenCV solves it with DECOMP_SVD; it is 2 x 2)
ISX_HD bool lm_centre(const double* gfp, double* rx, double* ry) {
  const double det = 4.0 * gfp[0] * gfp[1] - gfp[2] * gfp[2];
  if (det == 0.0) return false;
  *rx = (2.0 * gfp[1] * gfp[3] - gfp[2] * gfp[4]) / det;
  *ry = (2.0 * gfp[0] * gfp[4] - gfp[2] * gfp[3]) / det;
  return true;
}

// angle and axes from the refit (g = A, B, C), then the RotatedRect OpenCV returns: out = (cx, cy, width, height, angle)
ISX_HD void lm_box(const double* g, double rx, double ry, double scale, float cx, float cy, float* out) {
  const double min_eps = 1e-8;
  const double ang = -0.5 * atan2(g[2], g[1] - g[0]);
  double t;
  if (fabs(g[2]) > min_eps) t = g[2] / sin(-2.0 * ang);
  else t = g[1] - g[0];
  double r2 = fabs(g[0] + g[1] - t);
  if (r2 > min_eps) r2 = sqrt(2.0 / r2);
  double r3 = fabs(g[0] + g[1] + t);
  if (r3 > min_eps) r3 = sqrt(2.0 / r3);
  out[0] = static_cast<float>(rx / scale) + cx;
  out[1] = static_cast<float>(ry / scale) + cy;
  float w = static_cast<float>(r2 * 2 / scale), h = static_cast<float>(r3 * 2 / scale);
  float a = 0.f;   // RotatedRect's default: OpenCV assigns the angle only when it swaps the axes (always, for a real ellipse)
  if (w > h) {
    const float tmp = w; w = h; h = tmp;
    a = static_cast<float>(90 + ang * 180 / 3.1415926535897932384626433832795);
  }
  if (a < -180) a += 360;
  if (a > 360) a -= 360;
  out[2] = w; out[3] = h; out[4] = a;
}

// the 19 landmarks of extract_eye_landmarks (gaze_estimators.py:139-177) from the per-class results; Python computes the
// derived ones in double (numpy int64 / Python float arithmetic) and stores float32
ISX_HD void lm_assemble(const float* pupil, int has_pupil, const float* iris, int has_iris, const int* sclera_bbox /* xmin, xmax,
                        ymin, ymax */, int has_sclera, double epsilon, float* out) {
  for (int i = 0; i < 19; ++i) out[i] = 0.f;
  if (has_pupil) for (int i = 0; i < 5; ++i) out[i] = pupil[i];
  if (has_iris) for (int i = 0; i < 5; ++i) out[5 + i] = iris[i];
  if (has_sclera) {
    const int left = sclera_bbox[0], right = sclera_bbox[1], bottom = sclera_bbox[2], top = sclera_bbox[3];
    const int ew = right - left, eh = top - bottom;
    out[10] = static_cast<float>(left); out[11] = static_cast<float>(right);
    out[12] = static_cast<float>(bottom); out[13] = static_cast<float>(top);
    out[14] = static_cast<float>(ew); out[15] = static_cast<float>(eh);
    out[16] = static_cast<float>(static_cast<double>(eh) / (static_cast<double>(ew) + epsilon));
    if (has_pupil) {
      out[17] = static_cast<float>((static_cast<double>(pupil[0]) - static_cast<double>(left + right) / 2.0) / (static_cast<double>(ew) + epsilon));
      out[18] = static_cast<float>((static_cast<double>(pupil[1]) - static_cast<double>(bottom + top) / 2.0) / (static_cast<double>(eh) + epsilon));
    }
  }
}

}  // namespace isx_lm
