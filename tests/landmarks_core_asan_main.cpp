// AddressSanitizer / UBSan driver for iris-style-transfer_b200/csrc/landmarks_core.cuh (TEST INFRASTRUCTURE): random masks of
// awkward shapes (widths around the 32-bit word boundaries, foreground on the frame edge) through the same host+device
// functions the CUDA kernels call, with the planes in exactly-sized heap buffers so that any out-of-bounds word access of
// the scan / neighbour / trace code aborts.  compute-sanitizer is not available on the GPU pool; this is its stand-in for
// the shared-memory index arithmetic.  Built and run by tests/test_landmarks_core_host.py::test_core_under_asan.
#include <stdio.h>
#include <stdlib.h>

#include <vector>

extern "C" int lm_host_ellipse_features(const unsigned char* mask, int H, int W, int cap, float* out, int* info, unsigned* pts_out);

int main() {
  unsigned long long s = 88172645463325252ull;
  auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
  const int widths[] = {1, 2, 5, 29, 30, 31, 32, 33, 34, 61, 62, 63, 64, 65, 66, 95, 96, 97, 126, 127, 128, 129, 640};
  long total = 0;
  for (int it = 0; it < 600; ++it) {
    const int W = widths[rnd() % (sizeof(widths) / sizeof(widths[0]))], H = 1 + static_cast<int>(rnd() % 40);
    std::vector<unsigned char> m(static_cast<size_t>(H) * W);
    const unsigned dens = 1 + rnd() % 9;
    for (auto& v : m) v = (rnd() % 10) < dens;
    if (it % 3 == 0) for (int x = 0; x < W; ++x) { m[x] = 1; m[static_cast<size_t>(H - 1) * W + x] = 1; }
    if (it % 3 == 1) for (int y = 0; y < H; ++y) { m[static_cast<size_t>(y) * W] = 1; m[static_cast<size_t>(y) * W + W - 1] = 1; }
    float out[5];
    int info[4];
    const int cap = it % 5 == 0 ? 8 : 4096;
    std::vector<unsigned> pts(cap);
    if (lm_host_ellipse_features(m.data(), H, W, cap, out, info, pts.data()) != 0) { printf("rank bound contradiction\n"); return 1; }
    total += info[0] + info[1];
  }
  printf("asan driver ok %ld\n", total);
  return 0;
}
