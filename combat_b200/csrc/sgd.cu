// Nesterov SGD over flat parameter buffers, and the master-weight -> compute-layout preparation.
// Replaces torch.optim.SGD(momentum=.9, weight_decay=5e-4, nesterov=True).step()
// (train_generator.py:123-126,212,255).
#include "common.cuh"

__global__ void __launch_bounds__(256) sgd_nesterov_k(float* __restrict__ p, const float* __restrict__ g,
                                                      float* __restrict__ buf, long long n4, long long n,
                                                      const float* __restrict__ lr_dev, float mu, float wd, int first) {
  pdl_entry();
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
  pdl_launch(sgd_nesterov_k, grid, 256, 0, (cudaStream_t)stream, p, g, buf, n4, n, lr_dev, momentum, wd, first_step);
  COMBAT_RETURN_LAUNCH("sgd_nesterov");
}

// grid.y = descriptor, grid.x strides over 32x32 (co, ci) tiles of every filter tap.  The forward copy is a cast of the
// channels-last master; the input-gradient copy is its (co <-> ci) transpose with the taps reversed, staged through a
// padded shared-memory tile so that both the reads (along ci) and the writes (along co) are coalesced.
template <typename T>
__global__ void __launch_bounds__(256) prep_weights_k(const float* __restrict__ params, T* __restrict__ wbuf,
                                                      const combat_wprep_desc* __restrict__ table) {
  pdl_entry();
  __shared__ float tile[32][33];
  const combat_wprep_desc d = table[blockIdx.y];
  const int KK = d.KH * d.KW;
  const int tco = (d.Cout + 31) / 32, tci = (d.Cin + 31) / 32;
  const int ntiles = KK * tco * tci;
  const float* src = params + d.src_off;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int ci0 = (t % tci) * 32;
    const int r = t / tci;
    const int co0 = (r % tco) * 32;
    const int kk = r / tco;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + ty + 8 * j, ci = ci0 + tx;
      float v = 0.f;
      if (co < d.Cout && ci < d.Cin) {
        const long long e = ((long long)co * KK + kk) * d.Cin + ci;
        v = src[e];
        wbuf[d.fwd_off + e] = from_f<T>(v);
      }
      tile[ty + 8 * j][tx] = v;
    }
    __syncthreads();
    if (d.dgrad_off >= 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ci = ci0 + ty + 8 * j, co = co0 + tx;
        if (co < d.Cout && ci < d.Cin)  // dgrad layout [ci][KK-1-kk][co]
          wbuf[d.dgrad_off + ((long long)ci * KK + (KK - 1 - kk)) * d.Cout + co] = from_f<T>(tile[tx][ty + 8 * j]);
      }
    }
    __syncthreads();
  }
}

extern "C" int combat_prep_weights(const float* params, void* wbuf, int dtype, const combat_wprep_desc* table_dev,
                                   int n_desc, long long max_elems, void* stream) {
  COMBAT_ARG(params && wbuf && table_dev, 0);
  if (n_desc <= 0) return 0;
  int gx = (int)((max_elems + 1023) / 1024);  // one 32x32 tile per block iteration
  if (gx > 592) gx = 592;  // the big layers need ~2300 tiles: latency-bound below ~4 blocks per SM per descriptor
  if (gx < 1) gx = 1;
  dim3 grid(gx, n_desc);
  DISPATCH_DTYPE(dtype, pdl_launch(prep_weights_k<T>, grid, 256, 0, (cudaStream_t)stream, params, (T*)wbuf, table_dev);)
  COMBAT_RETURN_LAUNCH("prep_weights");
}
