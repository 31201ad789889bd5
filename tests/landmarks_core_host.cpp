// CPU harness around iris-style-transfer_b200/csrc/landmarks_core.cuh (TEST INFRASTRUCTURE): the same host+device
// functions the CUDA kernels of landmarks.cu call, driven serially -- rows scanned one after the other, the point sums
// accumulated in a loop -- so that tests/test_landmarks_core_host.py can pin them against cv2 without a GPU.
// Built by the test with:  g++ -O2 -shared -fPIC -I iris-style-transfer_b200/csrc tests/landmarks_core_host.cpp
#include <float.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "landmarks_core.cuh"

using namespace isx_lm;

// mask uint8 [H,W] (nonzero = foreground) -> out[5] = cx, cy, width, height, angle; info[4] = points of the chosen contour,
// contours found, flags (1 = rank deficient or five points, 2 = more than cap points, 16 = the determinant bound
// could not rule out rank deficiency), has_result.  pts_out (optional):
// the chosen contour's points as (x | y << 16).
extern "C" int lm_host_ellipse_features(const unsigned char* mask, int H, int W, int cap, float* out, int* info, unsigned* pts_out) {
  const int Ww = (W + 2 + 31) / 32;
  std::vector<uint32_t> F((H + 2) * Ww, 0u), M((H + 2) * Ww, 0u), N((H + 2) * Ww, 0u);
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x)
      if (mask[y * W + x]) lm_set(F.data(), Ww, x + 1, y + 1);
  std::vector<uint32_t> cur(cap), best(cap);
  long long best_area = -1;
  int best_n = 0, ncont = 0;
  // the kernel's schedule: cand[y] = first accepted start of row y, filled for all rows, then per contour only the rows
  // the trace touched (start row .. y_max) are recomputed -- a trace marks nothing else, and rows above the start are done
  std::vector<int> cand(H + 2, -1);
  for (int y = 1; y <= H; ++y) cand[y] = lm_row_first_start(F.data(), M.data(), N.data(), Ww, y, 0);
  int sy = 1;
  for (;;) {
    int fy = -1;
    for (int y = sy; y <= H; ++y)
      if (cand[y] >= 0) { fy = y; break; }
    if (fy < 0) break;
    const int fx = cand[fy];
    const LmTrace t = lm_trace(F.data(), M.data(), N.data(), Ww, fx, fy, cur.data(), cap);
    ++ncont;
    const long long a = t.cross < 0 ? -t.cross : t.cross;
    if (a >= best_area) { best_area = a; best_n = t.n; cur.swap(best); }   // cv2 lists the LAST discovered contour first
    for (int y = fy; y <= t.y_max; ++y) cand[y] = lm_row_first_start(F.data(), M.data(), N.data(), Ww, y, y == fy ? fx : 0);
    sy = fy;
  }
  info[0] = best_n; info[1] = ncont; info[2] = 0; info[3] = 0;
  for (int i = 0; i < 5; ++i) out[i] = 0.f;
  if (best_n > cap) { info[2] |= 2; return 0; }
  if (pts_out) memcpy(pts_out, best.data(), sizeof(uint32_t) * best_n);
  if (best_n < 5) return 0;
  long long sxi = 0, syi = 0;
  for (int i = 0; i < best_n; ++i) { sxi += best[i] & 0xFFFF; syi += best[i] >> 16; }
  const float cx = static_cast<float>(sxi) / static_cast<float>(best_n), cy = static_cast<float>(syi) / static_cast<float>(best_n);
  double s = 0;
  for (int i = 0; i < best_n; ++i) {
    const float dx = static_cast<float>(best[i] & 0xFFFF) - cx, dy = static_cast<float>(best[i] >> 16) - cy;
    s += fabsf(dx) + fabsf(dy);
  }
  const double scale = 100.0 / (s > FLT_EPSILON ? s : FLT_EPSILON);
  dd a1[kLmSums1];
  for (int i = 0; i < kLmSums1; ++i) a1[i] = dd_make(0.0);
  for (int i = 0; i < best_n; ++i) {
    const float dx = static_cast<float>(best[i] & 0xFFFF) - cx, dy = static_cast<float>(best[i] >> 16) - cy;
    lm_acc1(dx * scale, dy * scale, a1);
  }
  double gfp[5], rx = 0, ry = 0;
  double piv = 1.0, det = 0.0;
  bool ok = lm_solve_sym<5>(a1, a1 + 15, gfp, &piv, &det) && lm_centre(gfp, &rx, &ry);
  const bool quick = lm_rank_surely_full(a1, det);
  if (quick && lm_rank_deficient(a1)) return 99;   // the bound must never contradict the eigenvalues
  info[2] |= quick ? 0 : 16;                       // diagnostic for the tests (bit 4): the Jacobi sweeps were needed
  if (best_n == 5 || piv < kLmPivotFloor || (!quick && lm_rank_deficient(a1))) info[2] |= 1;
  dd a2[kLmSums2];
  for (int i = 0; i < kLmSums2; ++i) a2[i] = dd_make(0.0);
  double g[3];
  if (ok) {
    for (int i = 0; i < best_n; ++i) {
      const float dx = static_cast<float>(best[i] & 0xFFFF) - cx, dy = static_cast<float>(best[i] >> 16) - cy;
      lm_acc2(dx * scale, dy * scale, rx, ry, a2);
    }
    ok = lm_solve_sym<3>(a2, a2 + 6, g, &piv);
    if (piv < kLmPivotFloor) info[2] |= 1;
  }
  if (!ok) { info[2] |= 1; info[3] = 1; return 0; }
  lm_box(g, rx, ry, scale, cx, cy, out);
  info[3] = 1;
  return 0;
}

// The warp formulation of the row scan (landmarks.cu lm_row_first_start_warp) with the shuffle and the ballots written as
// loops over 32 "lanes", against the serial scan, on arbitrary planes: F / M / N rows of Ww <= 32 words given by the caller
// (marks only where F is set).  Returns the number of (row, x_after) pairs on which the two disagree.
extern "C" int lm_host_row_scan_mismatches(const uint32_t* F, const uint32_t* M, const uint32_t* N, int rows, int Ww) {
  int bad = 0;
  for (int y = 0; y < rows; ++y) {
    // x_after = 0 (a fresh row) and every marked pixel of the row (the precondition of a rescan: x_after is a marked pixel)
    std::vector<int> afters(1, 0);
    for (int x = 1; x < Ww * 32; ++x)
      if (lm_bit(M, Ww, x, y)) afters.push_back(x);
    for (int x_after : afters) {
      const int want = lm_row_first_start(F, M, N, Ww, y, x_after);
      uint32_t fw[32], mw[32], nw[32];
      bool has[32], pos[32];
      int xs[32];
      for (int lane = 0; lane < 32; ++lane) {
        fw[lane] = lane < Ww ? F[y * Ww + lane] : 0u;
        mw[lane] = fw[lane] ? M[y * Ww + lane] : 0u;
        nw[lane] = fw[lane] ? N[y * Ww + lane] : 0u;
        has[lane] = mw[lane] != 0u;
        pos[lane] = lm_word_last_mark_positive(mw[lane], nw[lane]);
      }
      for (int lane = 0; lane < 32; ++lane) {
        const uint32_t prev_top = lane == 0 ? 0u : fw[lane - 1] >> 31;               // __shfl_up_sync
        const uint32_t cand = lm_word_candidates(fw[lane], mw[lane], prev_top, x_after, lane);
        int inside_in = 0;                                                           // the two ballots
        for (int l = lane - 1; l >= 0; --l)
          if (has[l]) { inside_in = pos[l] ? 1 : 0; break; }
        const int k = lm_word_first_accepted(cand, mw[lane], nw[lane], inside_in);
        xs[lane] = k >= 0 ? lane * 32 + k : -1;
      }
      int got = -1;
      for (int lane = 0; lane < 32 && got < 0; ++lane) got = xs[lane];
      if (got != want) ++bad;
    }
  }
  return bad;
}

extern "C" void lm_host_assemble(const float* pupil, int has_pupil, const float* iris, int has_iris, const int* sclera_bbox,
                                 int has_sclera, double epsilon, float* out19) {
  lm_assemble(pupil, has_pupil, iris, has_iris, sclera_bbox, has_sclera, epsilon, out19);
}
