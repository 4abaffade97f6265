// Host build of combat_b200/csrc/dct_butterfly.cuh (the register network dct32.cu runs per image row / column), so
// that tests/test_dct_butterfly_cpu.py can hold it to scipy's orthonormal DCT-II / DCT-III without a GPU.
// Built by the test with g++; not part of the product library.
#include <cmath>
#define DCT_FN static inline
#define DCT_CONST static constexpr
#include "../combat_b200/csrc/dct_butterfly.cuh"

template <int N>
static void run(int inverse, const float* in, float* out, long rows) {
  for (long r = 0; r < rows; ++r) {
    float a[N], b[N];
    for (int i = 0; i < N; ++i) a[i] = in[r * N + i];
    if (inverse) Dct<N, N>::inv(a, b); else Dct<N, N>::fwd(a, b);
    for (int i = 0; i < N; ++i) out[r * N + i] = b[i];
  }
}

extern "C" int dct_rows(int n, int inverse, const float* in, float* out, long rows) {
  if (n == 32) run<32>(inverse, in, out, rows);
  else if (n == 64) run<64>(inverse, in, out, rows);
  else return -1;
  return 0;
}
