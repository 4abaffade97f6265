// Training-only kernels of the frequency detector (defenses/frequency_based/train.py:178-221, model.py:8-52 in train
// mode): ELU backward, 2x2 max-pool backward, dropout as a host-drawn keep mask, Adadelta over the flat parameter buffer.
// All four are HBM-bound element-wise passes (8-wide where the channel count allows, grid-stride, non-persistent grid).
// Forward ELU is fused into the convolution epilogues (act = 2); BatchNorm statistics / apply / backward are the kernels
// of norm.cu.  NOT part of the alternated step; written at the close of round 1 and not yet run on a GPU.
#include "common.cuh"

static inline int det_grid(long long n) {
  long long g = (n + 255) / 256;
  if (g > 148 * 32) g = 148 * 32;
  if (g < 1) g = 1;
  return (int)g;
}

// ---------------------------------------------------------------------------------- ELU backward (alpha = 1)
// a = elu(z) is what the forward kept: d elu / dz = 1 for z > 0, exp(z) = a + 1 otherwise (torch: elu_backward on the result)
template <typename T>
__global__ void __launch_bounds__(256) elu_bwd_k(const T* __restrict__ da, const T* __restrict__ a, T* __restrict__ dz, long long n) {
  pdl_entry();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = to_f<T>(a[i]);
    dz[i] = from_f<T>(to_f<T>(da[i]) * (v > 0.f ? 1.f : v + 1.f));
  }
}

extern "C" int combat_elu_bwd(const void* da, const void* a, void* dz, int dtype, long long n, void* stream) {
  COMBAT_ARG(da && a && dz, 0);
  if (n <= 0) return 0;
  DISPATCH_DTYPE(dtype, pdl_launch(elu_bwd_k<T>, det_grid(n), 256, 0, (cudaStream_t)stream, (const T*)da, (const T*)a, (T*)dz, n);)
  COMBAT_RETURN_LAUNCH("elu_bwd");
}

// ---------------------------------------------------------------------------------- MaxPool2d(2) backward, NHWC
// One thread per (n, oh, ow, c): the gradient goes to the FIRST window position in row-major order that holds the
// maximum (torch's max_pool2d keeps the first index: it replaces the running maximum only on `val > max`), zeros to the
// other three.  Every input element belongs to exactly one window, so dx is written completely, without atomics.
template <typename T>
__global__ void __launch_bounds__(256) maxpool2_bwd_k(const T* __restrict__ dy, const T* __restrict__ x, T* __restrict__ dx,
                                                      long long total, int H, int W, int C) {
  pdl_entry();
  const int Ho = H / 2, Wo = W / 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long p = i / C;
    const int ow = (int)(p % Wo);
    p /= Wo;
    const int oh = (int)(p % Ho);
    const long long n = p / Ho;
    const long long b = ((n * H + 2 * oh) * W + 2 * ow) * C + c;
    const long long off[4] = {0, C, (long long)W * C, (long long)W * C + C};
    int arg = 0;
    float m = to_f<T>(x[b]);
#pragma unroll
    for (int k = 1; k < 4; ++k) {
      const float v = to_f<T>(x[b + off[k]]);
      if (v > m) { m = v; arg = k; }
    }
    const T g = dy[i];
#pragma unroll
    for (int k = 0; k < 4; ++k) dx[b + off[k]] = k == arg ? g : from_f<T>(0.f);
  }
}

extern "C" int combat_maxpool2_bwd(const void* dy, const void* x, void* dx, int dtype, int N, int H, int W, int C, void* stream) {
  COMBAT_ARG(dy && x && dx, 0);
  COMBAT_ARG(N >= 0 && H > 0 && W > 0 && C > 0 && (H % 2) == 0 && (W % 2) == 0, 4);
  const long long total = (long long)N * (H / 2) * (W / 2) * C;
  if (total <= 0) return 0;
  DISPATCH_DTYPE(dtype, pdl_launch(maxpool2_bwd_k<T>, det_grid(total), 256, 0, (cudaStream_t)stream, (const T*)dy, (const T*)x, (T*)dx, total, H, W, C);)
  COMBAT_RETURN_LAUNCH("maxpool2_bwd");
}

// ---------------------------------------------------------------------------------- dropout with a given keep mask
// y = x * keep * scale (scale = 1 / (1 - p)); the same map is its own backward.  The mask is drawn on the HOST from the
// torch CPU generator (so that the reference's / oracle's stream can be reproduced) and uploaded as one byte per element.
template <typename T>
__global__ void __launch_bounds__(256) mask_scale_k(const T* __restrict__ x, const unsigned char* __restrict__ keep, T* __restrict__ y,
                                                    long long n, float scale) {
  pdl_entry();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = from_f<T>(keep[i] ? to_f<T>(x[i]) * scale : 0.f);
}

extern "C" int combat_mask_scale(const void* x, const unsigned char* keep, void* y, int dtype, long long n, float scale, void* stream) {
  COMBAT_ARG(x && keep && y, 0);
  if (n <= 0) return 0;
  DISPATCH_DTYPE(dtype, pdl_launch(mask_scale_k<T>, det_grid(n), 256, 0, (cudaStream_t)stream, (const T*)x, keep, (T*)y, n, scale);)
  COMBAT_RETURN_LAUNCH("mask_scale");
}

// ---------------------------------------------------------------------------------- Adadelta (train.py:152)
// torch.optim.Adadelta(lr, rho = 0.9, eps = 1e-6, weight_decay) over the flat buffers, 28 B / parameter:
//   g' = g + wd p;  v = rho v + (1 - rho) g'^2;  d = sqrt(u + eps) / sqrt(v + eps) * g';  u = rho u + (1 - rho) d^2;  p -= lr d
// lr is read from device memory (a captured graph follows the host's value, as in sgd_nesterov_k).
__global__ void __launch_bounds__(256) adadelta_k(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ v,
                                                  float* __restrict__ u, long long n, const float* __restrict__ lr_dev, float rho,
                                                  float eps, float wd) {
  pdl_entry();
  const float lr = lr_dev[0];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float pi = p[i];
    const float gi = fmaf(wd, pi, g[i]);
    const float vi = fmaf(1.f - rho, gi * gi, rho * v[i]);
    const float d = sqrtf(u[i] + eps) / sqrtf(vi + eps) * gi;
    v[i] = vi;
    u[i] = fmaf(1.f - rho, d * d, rho * u[i]);
    p[i] = fmaf(-lr, d, pi);
  }
}

extern "C" int combat_adadelta(float* p, const float* g, float* square_avg, float* acc_delta, long long n, const float* lr_dev,
                               float rho, float eps, float wd, void* stream) {
  COMBAT_ARG(p && g && square_avg && acc_delta && lr_dev, 0);
  if (n <= 0) return 0;
  pdl_launch(adadelta_k, det_grid(n), 256, 0, (cudaStream_t)stream, p, g, square_avg, acc_delta, n, lr_dev, rho, eps, wd);
  COMBAT_RETURN_LAUNCH("adadelta");
}
