// Nesterov SGD over flat parameter buffers, and the master-weight -> compute-layout preparation.
// Replaces torch.optim.SGD(momentum=.9, weight_decay=5e-4, nesterov=True).step()
// (train_generator.py:123-126,212,255).
#include "common.cuh"

__global__ void __launch_bounds__(256) sgd_nesterov_k(float* __restrict__ p, const float* __restrict__ g,
                                                      float* __restrict__ buf, long long n4, long long n,
                                                      const float* __restrict__ lr_dev, float mu, float wd, int first) {
  const float lr = lr_dev[0];
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) {
    float4 pv = ((float4*)p)[i], gv = ((const float4*)g)[i], bv;
    if (!first) bv = ((float4*)buf)[i];
#define UPD(f)                                       \
  {                                                  \
    float gg = fmaf(wd, pv.f, gv.f);                 \
    float b = first ? gg : fmaf(mu, bv.f, gg);       \
    bv.f = b;                                        \
    pv.f = pv.f - lr * fmaf(mu, b, gg);              \
  }
    UPD(x) UPD(y) UPD(z) UPD(w)
#undef UPD
    ((float4*)p)[i] = pv;
    ((float4*)buf)[i] = bv;
  }
  // tail
  for (long long j = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
    float gg = fmaf(wd, p[j], g[j]);
    float b = first ? gg : fmaf(mu, buf[j], gg);
    buf[j] = b;
    p[j] = p[j] - lr * fmaf(mu, b, gg);
  }
}

extern "C" int combat_sgd_nesterov(float* p, const float* g, float* buf, long long n, const float* lr_dev, float momentum,
                                   float wd, int first_step, void* stream) {
  COMBAT_ARG(p && g && buf && lr_dev, 0);
  if (n <= 0) return 0;
  COMBAT_ARG((((uintptr_t)p | (uintptr_t)g | (uintptr_t)buf) & 15) == 0, 1);
  long long n4 = n / 4;
  int grid = (int)((n4 + 255) / 256);
  if (grid > 148 * 16) grid = 148 * 16;
  if (grid < 1) grid = 1;
  sgd_nesterov_k<<<grid, 256, 0, (cudaStream_t)stream>>>(p, g, buf, n4, n, lr_dev, momentum, wd, first_step);
  COMBAT_RETURN_LAUNCH("sgd_nesterov");
}

// grid.y = descriptor, grid.x strides over the layer's elements (dst-major so that writes coalesce)
template <typename T>
__global__ void __launch_bounds__(256) prep_weights_k(const float* __restrict__ params, T* __restrict__ wbuf,
                                                      const combat_wprep_desc* __restrict__ table) {
  const combat_wprep_desc d = table[blockIdx.y];
  const int KK = d.KH * d.KW;
  const long long total = (long long)d.Cout * d.Cin * KK;
  const float* src = params + d.src_off;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    // master and fwd layout are both channels-last [co][kk][ci]: the forward copy is a pure cast
    int ci = (int)(e % d.Cin);
    long long r = e / d.Cin;
    int kk = (int)(r % KK);
    int co = (int)(r / KK);
    float v = src[e];
    wbuf[d.fwd_off + e] = from_f<T>(v);
    if (d.dgrad_off >= 0) {
      // dgrad layout [ci][KK-1-kk][co]
      wbuf[d.dgrad_off + ((long long)ci * KK + (KK - 1 - kk)) * d.Cout + co] = from_f<T>(v);
    }
  }
}

extern "C" int combat_prep_weights(const float* params, void* wbuf, int dtype, const combat_wprep_desc* table_dev,
                                   int n_desc, long long max_elems, void* stream) {
  COMBAT_ARG(params && wbuf && table_dev, 0);
  if (n_desc <= 0) return 0;
  int gx = (int)((max_elems + 255) / 256);
  if (gx > 64) gx = 64;
  if (gx < 1) gx = 1;
  dim3 grid(gx, n_desc);
  DISPATCH_DTYPE(dtype, prep_weights_k<T><<<grid, 256, 0, (cudaStream_t)stream>>>(params, (T*)wbuf, table_dev);)
  COMBAT_RETURN_LAUNCH("prep_weights");
}
