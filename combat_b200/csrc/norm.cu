// Normalisation / activation / resampling / pooling kernels on NHWC activations (float32 or bf16 storage,
// float32 arithmetic and statistics).  Replaces, for the hot path of the reference:
//   nn.BatchNorm2d + F.relu          classifier_models/preact_resnet.py:20,22,32,35 ; resnet.py:18-35,75,90
//   nn.InstanceNorm2d + LeakyReLU    networks/models.py:273-340
//   nn.Upsample(bilinear, x2)        networks/models.py:274
//   AvgPool2d(4) + Linear            classifier_models/preact_resnet.py:99-101 ; resnet.py:95-97
//   MaxPool2d(2), ELU+BN(eval)       defenses/frequency_based/model.py:13-44
#include <cooperative_groups.h>

#include "common.cuh"

// ---- two adjacent channels per thread
template <typename T>
__device__ __forceinline__ float2 ld2(const T* p);
template <>
__device__ __forceinline__ float2 ld2<float>(const float* p) { return *(const float2*)p; }
template <>
__device__ __forceinline__ float2 ld2<bf16>(const bf16* p) { return __bfloat1622float2(*(const __nv_bfloat162*)p); }
template <typename T>
__device__ __forceinline__ void st2(T* p, float2 v);
template <>
__device__ __forceinline__ void st2<float>(float* p, float2 v) { *(float2*)p = v; }
template <>
__device__ __forceinline__ void st2<bf16>(bf16* p, float2 v) { *(__nv_bfloat162*)p = __float22bfloat162_rn(v); }

// 8 consecutive channels per thread (16 bytes of bf16, 32 of float32): the elementwise kernels below are HBM-bound and
// spend their issue slots on index arithmetic when they move 4 bytes per thread
template <typename T>
__device__ __forceinline__ void ld8(const T* p, float* v);
template <>
__device__ __forceinline__ void ld8<float>(const float* p, float* v) {
  const float4 a = ((const float4*)p)[0], b = ((const float4*)p)[1];
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void ld8<bf16>(const bf16* p, float* v) {
  const uint4 u = *(const uint4*)p;
  const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(*(const __nv_bfloat162*)&w4[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}
template <typename T>
__device__ __forceinline__ void st8(T* p, const float* v);
template <>
__device__ __forceinline__ void st8<float>(float* p, const float* v) {
  ((float4*)p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  ((float4*)p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
template <>
__device__ __forceinline__ void st8<bf16>(bf16* p, const float* v) {
  __nv_bfloat162 h[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *(uint4*)p = *(uint4*)h;
}

// Column-reduction geometry: blockDim = (32, 8); a thread owns channels 2*(blockIdx.y*32+tx)+{0,1};
// row lanes ty stride over the block's row range.
#define CR_TX 32
#define CR_TY 8
#define CR_CPB 64  // channels per block

// reduce (a,b) float2 pairs across threadIdx.y; result valid on ty == 0
__device__ __forceinline__ void block_reduce_y(float2& a, float2& b) {
  __shared__ float2 sa[CR_TY][CR_TX], sb[CR_TY][CR_TX];
  sa[threadIdx.y][threadIdx.x] = a;
  sb[threadIdx.y][threadIdx.x] = b;
  __syncthreads();
  if (threadIdx.y == 0) {
    for (int y = 1; y < CR_TY; ++y) {
      float2 u = sa[y][threadIdx.x], v = sb[y][threadIdx.x];
      a.x += u.x; a.y += u.y; b.x += v.x; b.y += v.y;
    }
  }
  __syncthreads();
}

// ------------------------------------------------------------------ BatchNorm
template <typename T>
__global__ void __launch_bounds__(256) bn_stats_k(const T* __restrict__ x, long long R, int C, long long rows_per_block,
                                                  float* __restrict__ partial) {
  pdl_entry();
  const int c = 2 * (blockIdx.y * CR_TX + threadIdx.x);
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > R) r1 = R;
  float2 s = {0.f, 0.f}, ss = {0.f, 0.f};
  if (c < C) {
    for (long long r = r0 + threadIdx.y; r < r1; r += CR_TY) {
      float2 v = ld2<T>(x + r * C + c);
      s.x += v.x; s.y += v.y;
      ss.x = fmaf(v.x, v.x, ss.x); ss.y = fmaf(v.y, v.y, ss.y);
    }
  }
  block_reduce_y(s, ss);
  if (threadIdx.y == 0 && c < C) {
    float* p0 = partial + ((long long)blockIdx.x * 2) * C + c;
    p0[0] = s.x; p0[1] = s.y;
    p0[C] = ss.x; p0[C + 1] = ss.y;
  }
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-level sum of two doubles over FIN_THREADS threads (4 warps); result valid on thread 0
#define FIN_THREADS 128
__device__ __forceinline__ void fin_block_sum(double& a, double& b) {
  __shared__ double sa[FIN_THREADS / 32], sb[FIN_THREADS / 32];
  a = warp_sum_d(a);
  b = warp_sum_d(b);
  if ((threadIdx.x & 31) == 0) { sa[threadIdx.x >> 5] = a; sb[threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    a = sa[0]; b = sb[0];
#pragma unroll
    for (int w = 1; w < FIN_THREADS / 32; ++w) { a += sa[w]; b += sb[w]; }
  }
}

// one BLOCK of 4 warps per channel: threads stride over the partial blocks (double accumulation; up to ~1200 blocks when the
// producer conv or an 8-CTA-per-SM reduction wrote them)
__global__ void bn_finalize_k(const float* __restrict__ partial, int nblk, long long R, int C,
                              const float* __restrict__ gamma, const float* __restrict__ beta, float* running_mean,
                              float* running_var, float momentum, float eps, float* __restrict__ scale,
                              float* __restrict__ shift, float* save_mean, float* save_invstd) {
  pdl_entry();
  const int c = blockIdx.x;
  const int lane = threadIdx.x;  // thread 0 finishes the channel
  float mean = 0.f, invstd = 0.f;
  if (nblk > 0) {
    double s = 0.0, ss = 0.0;
    for (int b = threadIdx.x; b < nblk; b += FIN_THREADS) {
      s += (double)partial[((long long)b * 2) * C + c];
      ss += (double)partial[((long long)b * 2 + 1) * C + c];
    }
    fin_block_sum(s, ss);
    if (lane != 0) return;
    double m = s / (double)R;
    double var = ss / (double)R - m * m;
    if (var < 0.0) var = 0.0;
    mean = (float)m;
    invstd = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean && lane == 0) {
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      double unb = R > 1 ? var * (double)R / (double)(R - 1) : var;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
    }
  } else {
    if (lane != 0) return;
    mean = running_mean[c];
    invstd = 1.0f / sqrtf(running_var[c] + eps);
  }
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  const float sc = g * invstd;
  scale[c] = sc;
  shift[c] = b - mean * sc;
  if (save_mean) { save_mean[c] = mean; save_invstd[c] = invstd; }
}

template <typename TX, typename T>
__global__ void __launch_bounds__(256) affine_act_k(const TX* __restrict__ x, const T* __restrict__ res, T* __restrict__ y,
                                                    long long n2, int C, const float* __restrict__ scale,
                                                    const float* __restrict__ shift, int relu) {
  pdl_entry();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) {
    const long long e = i * 2;
    const int c = (int)(e % C);
    float2 v = ld2<TX>(x + e);
    v.x = fmaf(v.x, scale[c], shift[c]);
    v.y = fmaf(v.y, scale[c + 1], shift[c + 1]);
    if (res) {
      float2 r = ld2<T>(res + e);
      v.x += r.x; v.y += r.y;
    }
    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); }
    st2<T>(y + e, v);
  }
}

template <typename TX, typename T>
__global__ void __launch_bounds__(256) affine_act_v8_k(const TX* __restrict__ x, const T* __restrict__ res, T* __restrict__ y,
                                                       long long n8, int C, const float* __restrict__ scale,
                                                       const float* __restrict__ shift, int relu) {
  pdl_entry();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const long long e = i * 8;
    const int c = (int)(e % C);
    float v[8], sc[8], sh[8];
    ld8<TX>(x + e, v);
    ld8<float>(scale + c, sc);
    ld8<float>(shift + c, sh);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], sc[j], sh[j]);
    if (res) {
      float r[8];
      ld8<T>(res + e, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += r[j];
    }
    if (relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    st8<T>(y + e, v);
  }
}

template <typename TX, typename T>
__global__ void __launch_bounds__(256) bn_bwd_reduce_k(const T* __restrict__ dy, const TX* __restrict__ x,
                                                       const T* __restrict__ y, long long R, int C,
                                                       long long rows_per_block, const float* __restrict__ mean,
                                                       const float* __restrict__ invstd, float* __restrict__ partial,
                                                       int relu) {
  pdl_entry();
  const int c = 2 * (blockIdx.y * CR_TX + threadIdx.x);
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > R) r1 = R;
  float2 s = {0.f, 0.f}, sx = {0.f, 0.f};
  if (c < C) {
    const float m0 = mean[c], m1 = mean[c + 1], i0 = invstd[c], i1 = invstd[c + 1];
    for (long long r = r0 + threadIdx.y; r < r1; r += CR_TY) {
      float2 g = ld2<T>(dy + r * C + c);
      if (relu) {
        float2 yy = ld2<T>(y + r * C + c);
        if (!(yy.x > 0.f)) g.x = 0.f;
        if (!(yy.y > 0.f)) g.y = 0.f;
      }
      float2 v = ld2<TX>(x + r * C + c);
      s.x += g.x; s.y += g.y;
      sx.x = fmaf(g.x, (v.x - m0) * i0, sx.x);
      sx.y = fmaf(g.y, (v.y - m1) * i1, sx.y);
    }
  }
  block_reduce_y(s, sx);
  if (threadIdx.y == 0 && c < C) {
    float* p0 = partial + ((long long)blockIdx.x * 2) * C + c;
    p0[0] = s.x; p0[1] = s.y;
    p0[C] = sx.x; p0[C + 1] = sx.y;
  }
}

__global__ void bn_bwd_finalize_k(const float* __restrict__ partial, int nblk, int C, float* __restrict__ dgamma,
                                  float* __restrict__ dbeta) {
  pdl_entry();
  const int c = blockIdx.x;  // one block of 4 warps per channel
  double s = 0.0, sx = 0.0;
  for (int b = threadIdx.x; b < nblk; b += FIN_THREADS) {
    s += (double)partial[((long long)b * 2) * C + c];
    sx += (double)partial[((long long)b * 2 + 1) * C + c];
  }
  fin_block_sum(s, sx);
  if (threadIdx.x == 0) {
    dbeta[c] = (float)s;
    dgamma[c] = (float)sx;
  }
}

template <typename TX, typename T>
__global__ void __launch_bounds__(256) bn_bwd_apply_k(const T* __restrict__ dy, const TX* __restrict__ x,
                                                      const T* __restrict__ y, const T* __restrict__ dadd,
                                                      T* __restrict__ dx, T* __restrict__ dres, long long n2, long long R,
                                                      int C, const float* __restrict__ gamma,
                                                      const float* __restrict__ mean, const float* __restrict__ invstd,
                                                      const float* __restrict__ dgamma, const float* __restrict__ dbeta,
                                                      const float* __restrict__ eval_scale, int relu) {
  pdl_entry();
  const float invR = 1.f / (float)R;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) {
    const long long e = i * 2;
    const int c = (int)(e % C);
    float2 g = ld2<T>(dy + e);
    if (relu) {
      float2 yy = ld2<T>(y + e);
      if (!(yy.x > 0.f)) g.x = 0.f;
      if (!(yy.y > 0.f)) g.y = 0.f;
    }
    if (dres) st2<T>(dres + e, g);
    float2 o;
    if (eval_scale) {
      o.x = g.x * eval_scale[c];
      o.y = g.y * eval_scale[c + 1];
    } else {
      float2 v = ld2<TX>(x + e);
      const float g0 = gamma ? gamma[c] : 1.f, g1 = gamma ? gamma[c + 1] : 1.f;
      float xh0 = (v.x - mean[c]) * invstd[c], xh1 = (v.y - mean[c + 1]) * invstd[c + 1];
      o.x = g0 * invstd[c] * (g.x - dbeta[c] * invR - xh0 * dgamma[c] * invR);
      o.y = g1 * invstd[c + 1] * (g.y - dbeta[c + 1] * invR - xh1 * dgamma[c + 1] * invR);
    }
    if (dadd) {
      float2 a = ld2<T>(dadd + e);
      o.x += a.x; o.y += a.y;
    }
    st2<T>(dx + e, o);
  }
}

template <typename TX, typename T>
__global__ void __launch_bounds__(256) bn_bwd_apply_v8_k(const T* __restrict__ dy, const TX* __restrict__ x,
                                                         const T* __restrict__ y, const T* __restrict__ dadd,
                                                         T* __restrict__ dx, T* __restrict__ dres, long long n8, long long R,
                                                         int C, const float* __restrict__ gamma,
                                                         const float* __restrict__ mean, const float* __restrict__ invstd,
                                                         const float* __restrict__ dgamma, const float* __restrict__ dbeta,
                                                         const float* __restrict__ eval_scale, int relu) {
  pdl_entry();
  const float invR = 1.f / (float)R;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const long long e = i * 8;
    const int c = (int)(e % C);
    float g[8], o[8];
    ld8<T>(dy + e, g);
    if (relu) {
      float yy[8];
      ld8<T>(y + e, yy);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (!(yy[j] > 0.f)) g[j] = 0.f;
    }
    if (dres) st8<T>(dres + e, g);
    if (eval_scale) {
      float es[8];
      ld8<float>(eval_scale + c, es);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = g[j] * es[j];
    } else {
      float v[8], mu[8], is[8], dg[8], db[8], gm[8];
      ld8<TX>(x + e, v);
      ld8<float>(mean + c, mu);
      ld8<float>(invstd + c, is);
      ld8<float>(dgamma + c, dg);
      ld8<float>(dbeta + c, db);
      if (gamma) {
        ld8<float>(gamma + c, gm);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) gm[j] = 1.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (v[j] - mu[j]) * is[j];
        o[j] = gm[j] * is[j] * (g[j] - db[j] * invR - xh * dg[j] * invR);
      }
    }
    if (dadd) {
      float a[8];
      ld8<T>(dadd + e, a);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] += a[j];
    }
    st8<T>(dx + e, o);
  }
}

// TX = dtype of the pre-normalisation tensor (conv output), T = dtype of activations / gradients
#define DISPATCH_2(xdt, dt, ...)                 \
  if ((xdt) == COMBAT_F32 && (dt) == COMBAT_F32) { typedef float TX; typedef float T; __VA_ARGS__ } \
  else if ((xdt) == COMBAT_F32) { typedef float TX; typedef bf16 T; __VA_ARGS__ }                   \
  else if ((dt) == COMBAT_F32) { typedef bf16 TX; typedef float T; __VA_ARGS__ }                    \
  else { typedef bf16 TX; typedef bf16 T; __VA_ARGS__ }

static inline int ew_grid(long long n) {
  long long g = (n + 255) / 256;
  if (g > 148 * 32) g = 148 * 32;
  if (g < 1) g = 1;
  return (int)g;
}

static inline int cr_plan(long long R, int max_blocks, long long* rows_per_block) {
  int nblk = (int)((R + 255) / 256);  // >= 32 rows per row-lane
  if (nblk > max_blocks) nblk = max_blocks;
  if (nblk < 1) nblk = 1;
  long long rpb = (R + nblk - 1) / nblk;
  nblk = (int)((R + rpb - 1) / rpb);
  *rows_per_block = rpb;
  return nblk;
}

extern "C" int combat_bn_stats(const void* x, int dtype, long long R, int C, float* partial, int max_blocks,
                               int* nblk_out_host, void* stream) {
  COMBAT_ARG(x && partial && nblk_out_host, 0);
  COMBAT_ARG(R > 0 && C > 0 && (C % 2) == 0 && max_blocks > 0, 2);
  long long rpb;
  int nblk = cr_plan(R, max_blocks, &rpb);
  *nblk_out_host = nblk;
  dim3 grid(nblk, cdiv(C, CR_CPB)), block(CR_TX, CR_TY);
  DISPATCH_DTYPE(dtype, pdl_launch(bn_stats_k<T>, grid, block, 0, (cudaStream_t)stream, (const T*)x, R, C, rpb, partial);)
  COMBAT_RETURN_LAUNCH("bn_stats");
}

extern "C" int combat_bn_finalize(const float* partial, int nblk, long long R, int C, const float* gamma, const float* beta,
                                  float* running_mean, float* running_var, float momentum, float eps, float* scale,
                                  float* shift, float* save_mean, float* save_invstd, void* stream) {
  COMBAT_ARG(scale && shift, 10);
  COMBAT_ARG(nblk > 0 ? partial != nullptr : (running_mean && running_var), 0);
  pdl_launch(bn_finalize_k, C, FIN_THREADS, 0, (cudaStream_t)stream, partial, nblk, R, C, gamma, beta, running_mean, running_var,
                                                                momentum, eps, scale, shift, save_mean, save_invstd);
  COMBAT_RETURN_LAUNCH("bn_finalize");
}

// ---- train-mode BatchNorm backward in ONE cooperative launch: column reduction -> grid sync -> per-channel finalize -> grid
// sync -> elementwise apply.  Replaces three launches (bn_bwd_reduce / _finalize / _apply) per BatchNorm of the C-step
// backward; the second pass over dy / x / y is served from the 126 MB L2 for every layer but the first stage's.
// Geometry: 256 threads viewed as (32, 8) in the reduction and flat in the apply phase; the grid is exactly the number of
// co-resident CTAs (cooperative launch), every phase strides over its own virtual blocks.
template <typename TX, typename T>
__global__ void __launch_bounds__(256) bn_bwd_fused_k(const T* __restrict__ dy, const TX* __restrict__ x, const T* __restrict__ y,
                                                      const T* __restrict__ dadd, T* __restrict__ dx, T* __restrict__ dres,
                                                      long long R, int C, long long rows_per_block, int nblk,
                                                      const float* __restrict__ gamma, const float* __restrict__ mean,
                                                      const float* __restrict__ invstd, float* __restrict__ partial,
                                                      float* __restrict__ dgamma, float* __restrict__ dbeta, int relu) {
  pdl_entry();
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  __shared__ float2 sa[CR_TY][CR_TX], sb[CR_TY][CR_TX];
  // ---- phase 1: partial sums of dyh and dyh * xhat per (row block, channel)
  const int ngroups = (C + CR_CPB - 1) / CR_CPB;
  for (int vb = blockIdx.x; vb < nblk * ngroups; vb += gridDim.x) {
    const int bx = vb % nblk, by = vb / nblk;
    const int c = 2 * (by * CR_TX + tx);
    const long long r0 = (long long)bx * rows_per_block;
    long long r1 = r0 + rows_per_block;
    if (r1 > R) r1 = R;
    float2 s1 = {0.f, 0.f}, s2 = {0.f, 0.f};
    if (c < C) {
      const float m0 = mean[c], m1 = mean[c + 1], i0 = invstd[c], i1 = invstd[c + 1];
      for (long long r = r0 + ty; r < r1; r += CR_TY) {
        float2 g = ld2<T>(dy + r * C + c);
        if (relu) {
          const float2 yy = ld2<T>(y + r * C + c);
          if (!(yy.x > 0.f)) g.x = 0.f;
          if (!(yy.y > 0.f)) g.y = 0.f;
        }
        const float2 v = ld2<TX>(x + r * C + c);
        s1.x += g.x; s1.y += g.y;
        s2.x = fmaf(g.x, (v.x - m0) * i0, s2.x);
        s2.y = fmaf(g.y, (v.y - m1) * i1, s2.y);
      }
    }
    sa[ty][tx] = s1;
    sb[ty][tx] = s2;
    __syncthreads();
    if (ty == 0 && c < C) {
      for (int yy = 1; yy < CR_TY; ++yy) {
        const float2 u = sa[yy][tx], v = sb[yy][tx];
        s1.x += u.x; s1.y += u.y; s2.x += v.x; s2.y += v.y;
      }
      float* p0 = partial + ((long long)bx * 2) * C + c;
      p0[0] = s1.x; p0[1] = s1.y;
      p0[C] = s2.x; p0[C + 1] = s2.y;
    }
    __syncthreads();
  }
  grid.sync();
  // ---- phase 2: one CTA per channel sums the partial blocks in double precision
  __shared__ double da[8], db_[8];
  for (int c = blockIdx.x; c < C; c += gridDim.x) {
    double a = 0.0, b = 0.0;
    for (int k = threadIdx.x; k < nblk; k += 256) {
      a += (double)partial[((long long)k * 2) * C + c];
      b += (double)partial[((long long)k * 2 + 1) * C + c];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (tx == 0) { da[ty] = a; db_[ty] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < 8; ++w) { a += da[w]; b += db_[w]; }
      dbeta[c] = (float)a;
      dgamma[c] = (float)b;
    }
    __syncthreads();
  }
  grid.sync();
  // ---- phase 3: dx = gamma * invstd * (dyh - dbeta / R - xhat * dgamma / R) (+ dadd), 8 channels per thread
  const float invR = 1.f / (float)R;
  const long long n8 = R * C / 8;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n8; i += (long long)gridDim.x * 256) {
    const long long e = i * 8;
    const int c = (int)(e % C);
    float g[8], o[8], v[8], mu[8], is[8], dg[8], dbv[8], gm[8];
    ld8<T>(dy + e, g);
    if (relu) {
      float yy[8];
      ld8<T>(y + e, yy);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (!(yy[j] > 0.f)) g[j] = 0.f;
    }
    if (dres) st8<T>(dres + e, g);
    ld8<TX>(x + e, v);
    ld8<float>(mean + c, mu);
    ld8<float>(invstd + c, is);
    ld8<float>(dgamma + c, dg);
    ld8<float>(dbeta + c, dbv);
    if (gamma) {
      ld8<float>(gamma + c, gm);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) gm[j] = 1.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = (v[j] - mu[j]) * is[j];
      o[j] = gm[j] * is[j] * (g[j] - dbv[j] * invR - xh * dg[j] * invR);
    }
    if (dadd) {
      float a[8];
      ld8<T>(dadd + e, a);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] += a[j];
    }
    st8<T>(dx + e, o);
  }
}


// eval-mode scale/shift of EVERY BatchNorm of a network in one launch: table4[i] = (gamma, beta, mean, var) indices of
// flattened channel i into the flat parameter / running-stat buffers
__global__ void bn_eval_affine_k(const float* __restrict__ params, const float* __restrict__ bufs, const int4* __restrict__ table,
                                 int n, float eps, float* __restrict__ scale, float* __restrict__ shift) {
  pdl_entry();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int4 t = table[i];
  const float invstd = 1.0f / sqrtf(bufs[t.w] + eps);
  const float sc = params[t.x] * invstd;
  scale[i] = sc;
  shift[i] = params[t.y] - bufs[t.z] * sc;
}

extern "C" int combat_bn_eval_affine(const float* params, const float* bufs, const int* table4, int n, float eps, float* scale,
                                     float* shift, void* stream) {
  COMBAT_ARG(params && bufs && table4 && scale && shift, 0);
  if (n <= 0) return 0;
  pdl_launch(bn_eval_affine_k, cdiv(n, 256), 256, 0, (cudaStream_t)stream, params, bufs, (const int4*)table4, n, eps, scale, shift);
  COMBAT_RETURN_LAUNCH("bn_eval_affine");
}

extern "C" int combat_affine_act(const void* x, int x_dtype, const void* residual, void* y, int dtype, long long R, int C,
                                 const float* scale, const float* shift, int relu, void* stream) {
  COMBAT_ARG(x && y && scale && shift, 0);
  COMBAT_ARG((C % 2) == 0, 5);
  long long n2 = R * C / 2;
  if (n2 <= 0) return 0;
  if ((C % 8) == 0) {
    const long long n8 = R * C / 8;
    DISPATCH_2(x_dtype, dtype, pdl_launch(affine_act_v8_k<TX, T>, ew_grid(n8), 256, 0, (cudaStream_t)stream, 
                                   (const TX*)x, (const T*)residual, (T*)y, n8, C, scale, shift, relu);)
    COMBAT_RETURN_LAUNCH("affine_act");
  }
  DISPATCH_2(x_dtype, dtype, pdl_launch(affine_act_k<TX, T>, ew_grid(n2), 256, 0, (cudaStream_t)stream, 
                                 (const TX*)x, (const T*)residual, (T*)y, n2, C, scale, shift, relu);)
  COMBAT_RETURN_LAUNCH("affine_act");
}

extern "C" int combat_bn_bwd_reduce(const void* dy, const void* x, int x_dtype, const void* y, int dtype, long long R, int C,
                                    const float* mean, const float* invstd, float* partial, int max_blocks,
                                    int* nblk_out_host, int relu, void* stream) {
  COMBAT_ARG(dy && x && partial && mean && invstd && nblk_out_host, 0);
  COMBAT_ARG(!relu || y, 2);
  long long rpb;
  int nblk = cr_plan(R, max_blocks, &rpb);
  *nblk_out_host = nblk;
  dim3 grid(nblk, cdiv(C, CR_CPB)), block(CR_TX, CR_TY);
  DISPATCH_2(x_dtype, dtype, pdl_launch(bn_bwd_reduce_k<TX, T>, grid, block, 0, (cudaStream_t)stream, 
                                 (const T*)dy, (const TX*)x, (const T*)y, R, C, rpb, mean, invstd, partial, relu);)
  COMBAT_RETURN_LAUNCH("bn_bwd_reduce");
}

extern "C" int combat_bn_bwd_finalize(const float* partial, int nblk, int C, float* dgamma, float* dbeta, void* stream) {
  COMBAT_ARG(partial && dgamma && dbeta && nblk > 0, 0);
  pdl_launch(bn_bwd_finalize_k, C, FIN_THREADS, 0, (cudaStream_t)stream, partial, nblk, C, dgamma, dbeta);
  COMBAT_RETURN_LAUNCH("bn_bwd_finalize");
}

extern "C" int combat_bn_bwd_apply(const void* dy, const void* x, int x_dtype, const void* y, const void* dadd, void* dx,
                                   void* dres, int dtype, long long R, int C, const float* gamma, const float* mean,
                                   const float* invstd, const float* dgamma, const float* dbeta, const float* eval_scale,
                                   int relu, void* stream) {
  COMBAT_ARG(dy && dx, 0);
  COMBAT_ARG(eval_scale || (x && mean && invstd && dgamma && dbeta), 1);
  COMBAT_ARG(!relu || y, 2);
  long long n2 = R * C / 2;
  if (n2 <= 0) return 0;
  if ((C % 8) == 0) {
    const long long n8 = R * C / 8;
    DISPATCH_2(x_dtype, dtype, pdl_launch(bn_bwd_apply_v8_k<TX, T>, ew_grid(n8), 256, 0, (cudaStream_t)stream, 
                                   (const T*)dy, (const TX*)x, (const T*)y, (const T*)dadd, (T*)dx, (T*)dres, n8, R, C, gamma,
                                   mean, invstd, dgamma, dbeta, eval_scale, relu);)
    COMBAT_RETURN_LAUNCH("bn_bwd_apply");
  }
  DISPATCH_2(x_dtype, dtype, pdl_launch(bn_bwd_apply_k<TX, T>, ew_grid(n2), 256, 0, (cudaStream_t)stream, 
                                 (const T*)dy, (const TX*)x, (const T*)y, (const T*)dadd, (T*)dx, (T*)dres, n2, R, C, gamma,
                                 mean, invstd, dgamma, dbeta, eval_scale, relu);)
  COMBAT_RETURN_LAUNCH("bn_bwd_apply");
}

extern "C" int combat_bn_bwd_fused(const void* dy, const void* x, int x_dtype, const void* y, const void* dadd, void* dx,
                                   void* dres, int dtype, long long R, int C, const float* gamma, const float* mean,
                                   const float* invstd, float* partial, int max_blocks, float* dgamma, float* dbeta, int relu,
                                   void* stream) {
  COMBAT_ARG(dy && x && dx && partial && mean && invstd && dgamma && dbeta, 0);
  COMBAT_ARG(!relu || y, 2);
  COMBAT_ARG((C % 8) == 0 && R > 0, 9);
  long long rpb;
  int nblk = cr_plan(R, max_blocks, &rpb);
  int relu_ = relu;
  void* args[] = {(void*)&dy, (void*)&x, (void*)&y, (void*)&dadd, (void*)&dx, (void*)&dres, (void*)&R, (void*)&C, (void*)&rpb,
                  (void*)&nblk, (void*)&gamma, (void*)&mean, (void*)&invstd, (void*)&partial, (void*)&dgamma, (void*)&dbeta,
                  (void*)&relu_};
  cudaError_t e = cudaErrorInvalidValue;
  DISPATCH_2(x_dtype, dtype, {
    static int resident = 0;   // co-resident CTAs of this instantiation (cooperative launches must fit in one wave)
    if (!resident) resident = resident_ctas(bn_bwd_fused_k<TX, T>, 256);
    if (resident <= 0) resident = 148;
    e = cudaLaunchCooperativeKernel((const void*)bn_bwd_fused_k<TX, T>, dim3(resident), dim3(256), args, 0, (cudaStream_t)stream);
  })
  COMBAT_COUNT_LAUNCH();
  if (e != cudaSuccess) {
    combat_set_err("bn_bwd_fused", e);
    return -(int)e;
  }
  return 0;
}

// ------------------------------------------------------------------ InstanceNorm + LeakyReLU (+ skip)
// grid = (N, C/64), block (32, 8).  Three cached passes: mean, centred variance, apply.
template <typename TX, typename T>
__global__ void __launch_bounds__(256) instnorm_fwd_k(const TX* __restrict__ x, const T* __restrict__ skip, T* __restrict__ y,
                                                      int HW, int C, float eps, float slope, int act,
                                                      float* __restrict__ save_mean, float* __restrict__ save_invstd) {
  pdl_entry();
  const int n = blockIdx.x;
  const int c = 2 * (blockIdx.y * CR_TX + threadIdx.x);
  const bool ok = c < C;
  const long long base = (long long)n * HW * C + c;
  float2 s = {0.f, 0.f}, dummy = {0.f, 0.f};
  if (ok)
    for (int r = threadIdx.y; r < HW; r += CR_TY) {
      float2 v = ld2<TX>(x + base + (long long)r * C);
      s.x += v.x; s.y += v.y;
    }
  block_reduce_y(s, dummy);
  __shared__ float2 bc[CR_TX], bi[CR_TX];
  if (threadIdx.y == 0) bc[threadIdx.x] = make_float2(s.x / (float)HW, s.y / (float)HW);
  __syncthreads();
  const float2 mean = bc[threadIdx.x];
  float2 q = {0.f, 0.f};
  if (ok)
    for (int r = threadIdx.y; r < HW; r += CR_TY) {
      float2 v = ld2<TX>(x + base + (long long)r * C);
      float a = v.x - mean.x, b = v.y - mean.y;
      q.x = fmaf(a, a, q.x); q.y = fmaf(b, b, q.y);
    }
  block_reduce_y(q, dummy);
  if (threadIdx.y == 0) {
    float2 iv = make_float2(1.0f / sqrtf(q.x / (float)HW + eps), 1.0f / sqrtf(q.y / (float)HW + eps));
    bi[threadIdx.x] = iv;
    if (ok) {
      save_mean[(long long)n * C + c] = mean.x; save_mean[(long long)n * C + c + 1] = mean.y;
      save_invstd[(long long)n * C + c] = iv.x; save_invstd[(long long)n * C + c + 1] = iv.y;
    }
  }
  __syncthreads();
  const float2 inv = bi[threadIdx.x];
  if (ok)
    for (int r = threadIdx.y; r < HW; r += CR_TY) {
      const long long o = base + (long long)r * C;
      float2 v = ld2<TX>(x + o);
      v.x = (v.x - mean.x) * inv.x;
      v.y = (v.y - mean.y) * inv.y;
      if (act) {
        v.x = v.x > 0.f ? v.x : v.x * slope;
        v.y = v.y > 0.f ? v.y : v.y * slope;
      }
      if (skip) {
        float2 k = ld2<T>(skip + o);
        v.x += k.x; v.y += k.y;
      }
      st2<T>(y + o, v);
    }
}

template <typename TX, typename T>
__global__ void __launch_bounds__(256) instnorm_bwd_k(const T* __restrict__ dy1, const T* __restrict__ dy2,
                                                      const TX* __restrict__ x, T* __restrict__ dx, int HW, int C,
                                                      float slope, int act, const float* __restrict__ mean_,
                                                      const float* __restrict__ invstd_) {
  pdl_entry();
  const int n = blockIdx.x;
  const int c = 2 * (blockIdx.y * CR_TX + threadIdx.x);
  const bool ok = c < C;
  const long long base = (long long)n * HW * C + c;
  float2 mean = {0.f, 0.f}, inv = {0.f, 0.f};
  if (ok) {
    mean = make_float2(mean_[(long long)n * C + c], mean_[(long long)n * C + c + 1]);
    inv = make_float2(invstd_[(long long)n * C + c], invstd_[(long long)n * C + c + 1]);
  }
  float2 s = {0.f, 0.f}, sx = {0.f, 0.f};
  if (ok)
    for (int r = threadIdx.y; r < HW; r += CR_TY) {
      const long long o = base + (long long)r * C;
      float2 g = ld2<T>(dy1 + o);
      if (dy2) { float2 g2 = ld2<T>(dy2 + o); g.x += g2.x; g.y += g2.y; }
      float2 v = ld2<TX>(x + o);
      float xh0 = (v.x - mean.x) * inv.x, xh1 = (v.y - mean.y) * inv.y;
      if (act) {
        if (!(xh0 > 0.f)) g.x *= slope;
        if (!(xh1 > 0.f)) g.y *= slope;
      }
      s.x += g.x; s.y += g.y;
      sx.x = fmaf(g.x, xh0, sx.x); sx.y = fmaf(g.y, xh1, sx.y);
    }
  block_reduce_y(s, sx);
  __shared__ float2 b0[CR_TX], b1[CR_TX];
  if (threadIdx.y == 0) {
    b0[threadIdx.x] = make_float2(s.x / (float)HW, s.y / (float)HW);
    b1[threadIdx.x] = make_float2(sx.x / (float)HW, sx.y / (float)HW);
  }
  __syncthreads();
  const float2 mg = b0[threadIdx.x], mgx = b1[threadIdx.x];
  if (ok)
    for (int r = threadIdx.y; r < HW; r += CR_TY) {
      const long long o = base + (long long)r * C;
      float2 g = ld2<T>(dy1 + o);
      if (dy2) { float2 g2 = ld2<T>(dy2 + o); g.x += g2.x; g.y += g2.y; }
      float2 v = ld2<TX>(x + o);
      float xh0 = (v.x - mean.x) * inv.x, xh1 = (v.y - mean.y) * inv.y;
      if (act) {
        if (!(xh0 > 0.f)) g.x *= slope;
        if (!(xh1 > 0.f)) g.y *= slope;
      }
      float2 d;
      d.x = inv.x * (g.x - mg.x - xh0 * mgx.x);
      d.y = inv.y * (g.y - mg.y - xh1 * mgx.y);
      st2<T>(dx + o, d);
    }
}

// ---- large planes (ImageNet-10 shape: 112 x 112 = 12,544 pixels per (sample, channel), batch 64 -> only N * C/64 = 64 CTAs above, each
// streaming 3.2 MB three times: 732 us per launch, 12 % of that step).  The plane is split over K CTAs: the first kernel leaves
// per-chunk (mean, centred sum of squares) partials, the second merges them (Chan's parallel-variance formula: the same centred
// two-pass statistic up to rounding) and applies the normalisation to its chunk.  grid = (N, C/64, K).
template <typename TX>
__global__ void __launch_bounds__(256) instnorm_part_k(const TX* __restrict__ x, int HW, int C, int K, float* __restrict__ part) {
  pdl_entry();
  const int n = blockIdx.x, k = blockIdx.z;
  const int c = 2 * (blockIdx.y * CR_TX + threadIdx.x);
  const bool ok = c < C;
  const int rows = (HW + K - 1) / K, r0 = k * rows, r1 = min(HW, r0 + rows);
  const long long base = (long long)n * HW * C + c;
  float2 s = {0.f, 0.f}, dummy = {0.f, 0.f};
  if (ok)
    for (int r = r0 + threadIdx.y; r < r1; r += CR_TY) {
      float2 v = ld2<TX>(x + base + (long long)r * C);
      s.x += v.x; s.y += v.y;
    }
  block_reduce_y(s, dummy);
  __shared__ float2 bc[CR_TX];
  const float cnt = (float)max(r1 - r0, 1);
  if (threadIdx.y == 0) bc[threadIdx.x] = make_float2(s.x / cnt, s.y / cnt);
  __syncthreads();
  const float2 mean = bc[threadIdx.x];
  float2 q = {0.f, 0.f};
  if (ok)
    for (int r = r0 + threadIdx.y; r < r1; r += CR_TY) {
      float2 v = ld2<TX>(x + base + (long long)r * C);
      float a = v.x - mean.x, b = v.y - mean.y;
      q.x = fmaf(a, a, q.x); q.y = fmaf(b, b, q.y);
    }
  block_reduce_y(q, dummy);
  if (threadIdx.y == 0 && ok) {
    float* p0 = part + (((long long)n * K + k) * 2) * C + c;   // [N][K][2][C]: chunk mean | chunk centred sum of squares
    p0[0] = mean.x; p0[1] = mean.y;
    p0[C] = q.x; p0[C + 1] = q.y;
  }
}

template <typename TX, typename T>
__global__ void __launch_bounds__(256) instnorm_apply_k(const TX* __restrict__ x, const T* __restrict__ skip, T* __restrict__ y,
                                                        int HW, int C, int K, float eps, float slope, int act,
                                                        const float* __restrict__ part, float* __restrict__ save_mean,
                                                        float* __restrict__ save_invstd) {
  pdl_entry();
  const int n = blockIdx.x, k = blockIdx.z;
  const int c = 2 * (blockIdx.y * CR_TX + threadIdx.x);
  const bool ok = c < C;
  const int rows = (HW + K - 1) / K, r0 = k * rows, r1 = min(HW, r0 + rows);
  const long long base = (long long)n * HW * C + c;
  __shared__ float2 bc[CR_TX], bi[CR_TX];
  if (threadIdx.y == 0) {
    float2 mean = {0.f, 0.f}, m2 = {0.f, 0.f};
    if (ok) {
      const float* p0 = part + ((long long)n * K * 2) * C + c;
      for (int j = 0; j < K; ++j) {   // overall mean: chunk means weighted by their row counts
        const float cj = (float)max(min(HW, (j + 1) * rows) - j * rows, 0);
        mean.x += cj * p0[(long long)j * 2 * C]; mean.y += cj * p0[(long long)j * 2 * C + 1];
      }
      mean.x /= (float)HW; mean.y /= (float)HW;
      for (int j = 0; j < K; ++j) {   // M2 = sum_j (M2_j + n_j (mean_j - mean)^2)
        const float cj = (float)max(min(HW, (j + 1) * rows) - j * rows, 0);
        const float dx = p0[(long long)j * 2 * C] - mean.x, dy = p0[(long long)j * 2 * C + 1] - mean.y;
        m2.x += p0[(long long)j * 2 * C + C] + cj * dx * dx; m2.y += p0[(long long)j * 2 * C + C + 1] + cj * dy * dy;
      }
    }
    const float2 iv = make_float2(1.0f / sqrtf(m2.x / (float)HW + eps), 1.0f / sqrtf(m2.y / (float)HW + eps));
    bc[threadIdx.x] = mean;
    bi[threadIdx.x] = iv;
    if (ok && k == 0) {
      save_mean[(long long)n * C + c] = mean.x; save_mean[(long long)n * C + c + 1] = mean.y;
      save_invstd[(long long)n * C + c] = iv.x; save_invstd[(long long)n * C + c + 1] = iv.y;
    }
  }
  __syncthreads();
  const float2 mean = bc[threadIdx.x], inv = bi[threadIdx.x];
  if (ok)
    for (int r = r0 + threadIdx.y; r < r1; r += CR_TY) {
      const long long o = base + (long long)r * C;
      float2 v = ld2<TX>(x + o);
      v.x = (v.x - mean.x) * inv.x;
      v.y = (v.y - mean.y) * inv.y;
      if (act) {
        v.x = v.x > 0.f ? v.x : v.x * slope;
        v.y = v.y > 0.f ? v.y : v.y * slope;
      }
      if (skip) {
        float2 kk = ld2<T>(skip + o);
        v.x += kk.x; v.y += kk.y;
      }
      st2<T>(y + o, v);
    }
}

// backward, split the same way: chunk sums of g and g * xhat, then merge + apply
template <typename TX, typename T>
__global__ void __launch_bounds__(256) instnorm_bwd_part_k(const T* __restrict__ dy1, const T* __restrict__ dy2,
                                                           const TX* __restrict__ x, int HW, int C, int K, float slope, int act,
                                                           const float* __restrict__ mean_, const float* __restrict__ invstd_,
                                                           float* __restrict__ part) {
  pdl_entry();
  const int n = blockIdx.x, k = blockIdx.z;
  const int c = 2 * (blockIdx.y * CR_TX + threadIdx.x);
  const bool ok = c < C;
  const int rows = (HW + K - 1) / K, r0 = k * rows, r1 = min(HW, r0 + rows);
  const long long base = (long long)n * HW * C + c;
  float2 mean = {0.f, 0.f}, inv = {0.f, 0.f};
  if (ok) {
    mean = make_float2(mean_[(long long)n * C + c], mean_[(long long)n * C + c + 1]);
    inv = make_float2(invstd_[(long long)n * C + c], invstd_[(long long)n * C + c + 1]);
  }
  float2 s = {0.f, 0.f}, sx = {0.f, 0.f};
  if (ok)
    for (int r = r0 + threadIdx.y; r < r1; r += CR_TY) {
      const long long o = base + (long long)r * C;
      float2 g = ld2<T>(dy1 + o);
      if (dy2) { float2 g2 = ld2<T>(dy2 + o); g.x += g2.x; g.y += g2.y; }
      float2 v = ld2<TX>(x + o);
      float xh0 = (v.x - mean.x) * inv.x, xh1 = (v.y - mean.y) * inv.y;
      if (act) {
        if (!(xh0 > 0.f)) g.x *= slope;
        if (!(xh1 > 0.f)) g.y *= slope;
      }
      s.x += g.x; s.y += g.y;
      sx.x = fmaf(g.x, xh0, sx.x); sx.y = fmaf(g.y, xh1, sx.y);
    }
  block_reduce_y(s, sx);
  if (threadIdx.y == 0 && ok) {
    float* p0 = part + (((long long)n * K + k) * 2) * C + c;
    p0[0] = s.x; p0[1] = s.y;
    p0[C] = sx.x; p0[C + 1] = sx.y;
  }
}

template <typename TX, typename T>
__global__ void __launch_bounds__(256) instnorm_bwd_apply_k(const T* __restrict__ dy1, const T* __restrict__ dy2,
                                                            const TX* __restrict__ x, T* __restrict__ dx, int HW, int C, int K,
                                                            float slope, int act, const float* __restrict__ mean_,
                                                            const float* __restrict__ invstd_, const float* __restrict__ part) {
  pdl_entry();
  const int n = blockIdx.x, k = blockIdx.z;
  const int c = 2 * (blockIdx.y * CR_TX + threadIdx.x);
  const bool ok = c < C;
  const int rows = (HW + K - 1) / K, r0 = k * rows, r1 = min(HW, r0 + rows);
  const long long base = (long long)n * HW * C + c;
  float2 mean = {0.f, 0.f}, inv = {0.f, 0.f}, mg = {0.f, 0.f}, mgx = {0.f, 0.f};
  if (ok) {
    mean = make_float2(mean_[(long long)n * C + c], mean_[(long long)n * C + c + 1]);
    inv = make_float2(invstd_[(long long)n * C + c], invstd_[(long long)n * C + c + 1]);
    const float* p0 = part + ((long long)n * K * 2) * C + c;
    for (int j = 0; j < K; ++j) {
      mg.x += p0[(long long)j * 2 * C]; mg.y += p0[(long long)j * 2 * C + 1];
      mgx.x += p0[(long long)j * 2 * C + C]; mgx.y += p0[(long long)j * 2 * C + C + 1];
    }
    mg.x /= (float)HW; mg.y /= (float)HW; mgx.x /= (float)HW; mgx.y /= (float)HW;
  }
  if (ok)
    for (int r = r0 + threadIdx.y; r < r1; r += CR_TY) {
      const long long o = base + (long long)r * C;
      float2 g = ld2<T>(dy1 + o);
      if (dy2) { float2 g2 = ld2<T>(dy2 + o); g.x += g2.x; g.y += g2.y; }
      float2 v = ld2<TX>(x + o);
      float xh0 = (v.x - mean.x) * inv.x, xh1 = (v.y - mean.y) * inv.y;
      if (act) {
        if (!(xh0 > 0.f)) g.x *= slope;
        if (!(xh1 > 0.f)) g.y *= slope;
      }
      float2 d;
      d.x = inv.x * (g.x - mg.x - xh0 * mgx.x);
      d.y = inv.y * (g.y - mg.y - xh1 * mgx.y);
      st2<T>(dx + o, d);
    }
}

// number of plane chunks: 1 (the single-kernel path) unless the plane is large and the grid would leave most SMs idle
static int instnorm_splits(int N, int HW, int C) {
  const long long ctas = (long long)N * cdiv(C, CR_CPB);
  if (HW < 4096 || ctas >= 296) return 1;
  int k = (int)((592 + ctas - 1) / ctas);   // ~4 CTAs per SM
  if (k > 32) k = 32;
  while (k > 1 && HW / k < 256) --k;
  return k;
}
extern "C" int combat_instnorm_splits(int N, int HW, int C) { return instnorm_splits(N, HW, C); }

extern "C" int combat_instnorm_fwd(const void* x, int x_dtype, const void* skip, void* y, int dtype, int N, int HW, int C, float eps,
                                   float slope, int act, float* save_mean, float* save_invstd, void* stream) {
  COMBAT_ARG(x && y && save_mean && save_invstd, 0);
  COMBAT_ARG(N > 0 && HW > 0 && C > 0 && (C % 2) == 0, 4);
  dim3 grid(N, cdiv(C, CR_CPB)), block(CR_TX, CR_TY);
  DISPATCH_2(x_dtype, dtype, pdl_launch(instnorm_fwd_k<TX, T>, grid, block, 0, (cudaStream_t)stream, 
                                 (const TX*)x, (const T*)skip, (T*)y, HW, C, eps, slope, act, save_mean, save_invstd);)
  COMBAT_RETURN_LAUNCH("instnorm_fwd");
}

extern "C" int combat_instnorm_fwd_split(const void* x, int x_dtype, const void* skip, void* y, int dtype, int N, int HW, int C,
                                         float eps, float slope, int act, float* save_mean, float* save_invstd, float* part, int K,
                                         void* stream) {
  COMBAT_ARG(x && y && save_mean && save_invstd && part, 0);
  COMBAT_ARG(N > 0 && HW > 0 && C > 0 && (C % 2) == 0 && K >= 1 && K <= 32, 4);
  dim3 grid(N, cdiv(C, CR_CPB), K), block(CR_TX, CR_TY);
  DISPATCH_2(x_dtype, dtype, pdl_launch(instnorm_part_k<TX>, grid, block, 0, (cudaStream_t)stream, (const TX*)x, HW, C, K, part);)
  COMBAT_CHECK_LAUNCH("instnorm_part");
  DISPATCH_2(x_dtype, dtype, pdl_launch(instnorm_apply_k<TX, T>, grid, block, 0, (cudaStream_t)stream, (const TX*)x, (const T*)skip,
                                        (T*)y, HW, C, K, eps, slope, act, (const float*)part, save_mean, save_invstd);)
  COMBAT_RETURN_LAUNCH("instnorm_apply");
}

extern "C" int combat_instnorm_bwd(const void* dy1, const void* dy2, const void* x, int x_dtype, void* dx, int dtype, int N, int HW,
                                   int C, float slope, int act, const float* mean, const float* invstd, void* stream) {
  COMBAT_ARG(dy1 && x && dx && mean && invstd, 0);
  COMBAT_ARG(N > 0 && HW > 0 && C > 0 && (C % 2) == 0, 5);
  dim3 grid(N, cdiv(C, CR_CPB)), block(CR_TX, CR_TY);
  DISPATCH_2(x_dtype, dtype, pdl_launch(instnorm_bwd_k<TX, T>, grid, block, 0, (cudaStream_t)stream, 
                                 (const T*)dy1, (const T*)dy2, (const TX*)x, (T*)dx, HW, C, slope, act, mean, invstd);)
  COMBAT_RETURN_LAUNCH("instnorm_bwd");
}

extern "C" int combat_instnorm_bwd_split(const void* dy1, const void* dy2, const void* x, int x_dtype, void* dx, int dtype, int N,
                                         int HW, int C, float slope, int act, const float* mean, const float* invstd, float* part,
                                         int K, void* stream) {
  COMBAT_ARG(dy1 && x && dx && mean && invstd && part, 0);
  COMBAT_ARG(N > 0 && HW > 0 && C > 0 && (C % 2) == 0 && K >= 1 && K <= 32, 5);
  dim3 grid(N, cdiv(C, CR_CPB), K), block(CR_TX, CR_TY);
  DISPATCH_2(x_dtype, dtype, pdl_launch(instnorm_bwd_part_k<TX, T>, grid, block, 0, (cudaStream_t)stream, (const T*)dy1, (const T*)dy2,
                                        (const TX*)x, HW, C, K, slope, act, mean, invstd, part);)
  COMBAT_CHECK_LAUNCH("instnorm_bwd_part");
  DISPATCH_2(x_dtype, dtype, pdl_launch(instnorm_bwd_apply_k<TX, T>, grid, block, 0, (cudaStream_t)stream, (const T*)dy1,
                                        (const T*)dy2, (const TX*)x, (T*)dx, HW, C, K, slope, act, mean, invstd, (const float*)part);)
  COMBAT_RETURN_LAUNCH("instnorm_bwd_apply");
}

// ------------------------------------------------------------------ bilinear x2 upsample (+ LeakyReLU)
__device__ __forceinline__ void up_src(int o, int n_in, int& i0, int& i1, float& l0, float& l1) {
  // PyTorch area_pixel_compute_source_index, align_corners=False, scale 0.5
  float r = 0.5f * ((float)o + 0.5f) - 0.5f;
  if (r < 0.f) r = 0.f;
  i0 = (int)r;
  i1 = i0 + (i0 < n_in - 1 ? 1 : 0);
  l1 = r - (float)i0;
  l0 = 1.f - l1;
}

template <typename T>
__global__ void __launch_bounds__(256) upsample2x_act_k(const T* __restrict__ x, T* __restrict__ y, long long total2, int H,
                                                        int W, int C, float slope) {
  pdl_entry();
  const int C2 = C / 2, Ho = 2 * H, Wo = 2 * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total2; i += (long long)gridDim.x * blockDim.x) {
    const int c = 2 * (int)(i % C2);
    long long p = i / C2;
    const int ow = (int)(p % Wo);
    p /= Wo;
    const int oh = (int)(p % Ho);
    const long long n = p / Ho;
    int h0, h1, w0, w1;
    float lh0, lh1, lw0, lw1;
    up_src(oh, H, h0, h1, lh0, lh1);
    up_src(ow, W, w0, w1, lw0, lw1);
    const T* xb = x + n * H * W * C + c;
    float2 a = ld2<T>(xb + ((long long)h0 * W + w0) * C), b = ld2<T>(xb + ((long long)h0 * W + w1) * C);
    float2 cc = ld2<T>(xb + ((long long)h1 * W + w0) * C), d = ld2<T>(xb + ((long long)h1 * W + w1) * C);
    float2 v;
    v.x = lh0 * (lw0 * a.x + lw1 * b.x) + lh1 * (lw0 * cc.x + lw1 * d.x);
    v.y = lh0 * (lw0 * a.y + lw1 * b.y) + lh1 * (lw0 * cc.y + lw1 * d.y);
    v.x = v.x > 0.f ? v.x : v.x * slope;
    v.y = v.y > 0.f ? v.y : v.y * slope;
    st2<T>(y + i * 2, v);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) upsample2x_act_v8_k(const T* __restrict__ x, T* __restrict__ y, long long total8, int H,
                                                           int W, int C, float slope) {
  pdl_entry();
  const int C8 = C / 8, Ho = 2 * H, Wo = 2 * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (long long)gridDim.x * blockDim.x) {
    const int c = 8 * (int)(i % C8);
    long long p = i / C8;
    const int ow = (int)(p % Wo);
    p /= Wo;
    const int oh = (int)(p % Ho);
    const long long n = p / Ho;
    int h0, h1, w0, w1;
    float lh0, lh1, lw0, lw1;
    up_src(oh, H, h0, h1, lh0, lh1);
    up_src(ow, W, w0, w1, lw0, lw1);
    const T* xb = x + n * H * W * C + c;
    float a[8], b[8], cc[8], d[8], v[8];
    ld8<T>(xb + ((long long)h0 * W + w0) * C, a);
    ld8<T>(xb + ((long long)h0 * W + w1) * C, b);
    ld8<T>(xb + ((long long)h1 * W + w0) * C, cc);
    ld8<T>(xb + ((long long)h1 * W + w1) * C, d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float t = lh0 * (lw0 * a[j] + lw1 * b[j]) + lh1 * (lw0 * cc[j] + lw1 * d[j]);
      v[j] = t > 0.f ? t : t * slope;
    }
    st8<T>(y + i * 8, v);
  }
}

// weight with which input index `i` enters output index `o` (0 if not referenced)
__device__ __forceinline__ float up_w(int o, int i, int n_in) {
  int i0, i1;
  float l0, l1;
  up_src(o, n_in, i0, i1, l0, l1);
  float w = 0.f;
  if (i0 == i) w += l0;
  if (i1 == i) w += l1;
  return w;
}

template <typename T>
__global__ void __launch_bounds__(256) upsample2x_act_bwd_k(const T* __restrict__ dy, const T* __restrict__ y,
                                                            T* __restrict__ dx, long long total2, int H, int W, int C,
                                                            float slope) {
  pdl_entry();
  const int C2 = C / 2, Ho = 2 * H, Wo = 2 * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total2; i += (long long)gridDim.x * blockDim.x) {
    const int c = 2 * (int)(i % C2);
    long long p = i / C2;
    const int w = (int)(p % W);
    p /= W;
    const int h = (int)(p % H);
    const long long n = p / H;
    float2 acc = {0.f, 0.f};
    for (int oh = 2 * h - 1; oh <= 2 * h + 2; ++oh) {
      if (oh < 0 || oh >= Ho) continue;
      const float wh = up_w(oh, h, H);
      if (wh == 0.f) continue;
      for (int ow = 2 * w - 1; ow <= 2 * w + 2; ++ow) {
        if (ow < 0 || ow >= Wo) continue;
        const float ww = up_w(ow, w, W);
        if (ww == 0.f) continue;
        const long long o = ((n * Ho + oh) * Wo + ow) * C + c;
        float2 g = ld2<T>(dy + o);
        if (slope != 1.f) {
          float2 yy = ld2<T>(y + o);
          if (!(yy.x > 0.f)) g.x *= slope;
          if (!(yy.y > 0.f)) g.y *= slope;
        }
        acc.x = fmaf(wh * ww, g.x, acc.x);
        acc.y = fmaf(wh * ww, g.y, acc.y);
      }
    }
    st2<T>(dx + i * 2, acc);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) upsample2x_act_bwd_v8_k(const T* __restrict__ dy, const T* __restrict__ y,
                                                               T* __restrict__ dx, long long total8, int H, int W, int C,
                                                               float slope) {
  pdl_entry();
  const int C8 = C / 8, Ho = 2 * H, Wo = 2 * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (long long)gridDim.x * blockDim.x) {
    const int c = 8 * (int)(i % C8);
    long long p = i / C8;
    const int w = (int)(p % W);
    p /= W;
    const int h = (int)(p % H);
    const long long n = p / H;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int oh = 2 * h - 1; oh <= 2 * h + 2; ++oh) {
      if (oh < 0 || oh >= Ho) continue;
      const float wh = up_w(oh, h, H);
      if (wh == 0.f) continue;
      for (int ow = 2 * w - 1; ow <= 2 * w + 2; ++ow) {
        if (ow < 0 || ow >= Wo) continue;
        const float ww = up_w(ow, w, W);
        if (ww == 0.f) continue;
        const long long o = ((n * Ho + oh) * Wo + ow) * C + c;
        float g[8];
        ld8<T>(dy + o, g);
        if (slope != 1.f) {
          float yy[8];
          ld8<T>(y + o, yy);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (!(yy[j] > 0.f)) g[j] *= slope;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(wh * ww, g[j], acc[j]);
      }
    }
    st8<T>(dx + i * 8, acc);
  }
}

extern "C" int combat_upsample2x_act(const void* x, void* y, int dtype, int N, int H, int W, int C, float slope,
                                     void* stream) {
  COMBAT_ARG(x && y && (C % 2) == 0, 0);
  long long total2 = (long long)N * 4 * H * W * C / 2;
  if (total2 <= 0) return 0;
  if ((C % 8) == 0) {
    const long long total8 = total2 / 4;
    DISPATCH_DTYPE(dtype, pdl_launch(upsample2x_act_v8_k<T>, ew_grid(total8), 256, 0, (cudaStream_t)stream, (const T*)x, (T*)y, total8, H,
                                                                                                   W, C, slope);)
    COMBAT_RETURN_LAUNCH("upsample2x_act");
  }
  DISPATCH_DTYPE(dtype, pdl_launch(upsample2x_act_k<T>, ew_grid(total2), 256, 0, (cudaStream_t)stream, (const T*)x, (T*)y, total2, H, W,
                                                                                              C, slope);)
  COMBAT_RETURN_LAUNCH("upsample2x_act");
}

extern "C" int combat_upsample2x_act_bwd(const void* dy, const void* y, void* dx, int dtype, int N, int H, int W, int C,
                                         float slope, void* stream) {
  COMBAT_ARG(dy && dx && (C % 2) == 0, 0);
  COMBAT_ARG(slope == 1.f || y, 1);
  long long total2 = (long long)N * H * W * C / 2;
  if (total2 <= 0) return 0;
  if ((C % 8) == 0) {
    const long long total8 = total2 / 4;
    DISPATCH_DTYPE(dtype, pdl_launch(upsample2x_act_bwd_v8_k<T>, ew_grid(total8), 256, 0, (cudaStream_t)stream, 
                              (const T*)dy, (const T*)y, (T*)dx, total8, H, W, C, slope);)
    COMBAT_RETURN_LAUNCH("upsample2x_act_bwd");
  }
  DISPATCH_DTYPE(dtype, pdl_launch(upsample2x_act_bwd_k<T>, ew_grid(total2), 256, 0, (cudaStream_t)stream, 
                            (const T*)dy, (const T*)y, (T*)dx, total2, H, W, C, slope);)
  COMBAT_RETURN_LAUNCH("upsample2x_act_bwd");
}

// ------------------------------------------------------------------ elementwise
template <typename T>
__global__ void __launch_bounds__(256) leaky_relu_k(const T* __restrict__ x, T* __restrict__ y, long long n, float slope) {
  pdl_entry();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = to_f<T>(x[i]);
    y[i] = from_f<T>(v > 0.f ? v : v * slope);
  }
}
template <typename T>
__global__ void __launch_bounds__(256) leaky_relu_bwd_k(const T* __restrict__ dy, const T* __restrict__ x, T* __restrict__ dx,
                                                        long long n, float slope) {
  pdl_entry();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float g = to_f<T>(dy[i]);
    dx[i] = from_f<T>(to_f<T>(x[i]) > 0.f ? g : g * slope);
  }
}
__global__ void __launch_bounds__(256) tanh_bwd_k(const float* __restrict__ dy, const float* __restrict__ y,
                                                  float* __restrict__ dz, long long n) {
  pdl_entry();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float t = y[i];
    dz[i] = dy[i] * (1.f - t * t);
  }
}

extern "C" int combat_leaky_relu(const void* x, void* y, int dtype, long long n, float slope, void* stream) {
  COMBAT_ARG(x && y, 0);
  if (n <= 0) return 0;
  DISPATCH_DTYPE(dtype, pdl_launch(leaky_relu_k<T>, ew_grid(n), 256, 0, (cudaStream_t)stream, (const T*)x, (T*)y, n, slope);)
  COMBAT_RETURN_LAUNCH("leaky_relu");
}
extern "C" int combat_leaky_relu_bwd(const void* dy, const void* x, void* dx, int dtype, long long n, float slope,
                                     void* stream) {
  COMBAT_ARG(dy && x && dx, 0);
  if (n <= 0) return 0;
  DISPATCH_DTYPE(dtype,
                 pdl_launch(leaky_relu_bwd_k<T>, ew_grid(n), 256, 0, (cudaStream_t)stream, (const T*)dy, (const T*)x, (T*)dx, n, slope);)
  COMBAT_RETURN_LAUNCH("leaky_relu_bwd");
}
extern "C" int combat_tanh_bwd(const float* dy, const float* y, float* dz, long long n, void* stream) {
  COMBAT_ARG(dy && y && dz, 0);
  if (n <= 0) return 0;
  pdl_launch(tanh_bwd_k, ew_grid(n), 256, 0, (cudaStream_t)stream, dy, y, dz, n);
  COMBAT_RETURN_LAUNCH("tanh_bwd");
}

// column sums (conv bias gradient): one block column per 64 channels, all rows; accumulate flag
template <typename T>
__global__ void __launch_bounds__(256) colsum_k(const T* __restrict__ x, long long R, int C, long long rows_per_block,
                                                float* __restrict__ out) {
  pdl_entry();
  const int c = 2 * (blockIdx.y * CR_TX + threadIdx.x);
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > R) r1 = R;
  float2 s = {0.f, 0.f}, dummy = {0.f, 0.f};
  if (c < C)
    for (long long r = r0 + threadIdx.y; r < r1; r += CR_TY) {
      float2 v = ld2<T>(x + r * C + c);
      s.x += v.x; s.y += v.y;
    }
  block_reduce_y(s, dummy);
  if (threadIdx.y == 0 && c < C) {
    atomicAdd(out + c, s.x);
    atomicAdd(out + c + 1, s.y);
  }
}
extern "C" int combat_colsum(const void* x, int dtype, long long R, int C, float* out, void* stream) {
  COMBAT_ARG(x && out && (C % 2) == 0, 0);
  long long rpb;
  int nblk = cr_plan(R, 128, &rpb);
  dim3 grid(nblk, cdiv(C, CR_CPB)), block(CR_TX, CR_TY);
  DISPATCH_DTYPE(dtype, pdl_launch(colsum_k<T>, grid, block, 0, (cudaStream_t)stream, (const T*)x, R, C, rpb, out);)
  COMBAT_RETURN_LAUNCH("colsum");
}

// ------------------------------------------------------------------ avg_pool(P) + flatten + linear
// one CTA per sample.  feature index f = c*(ph*pw) + hp*pw + wp (NCHW flatten order of the reference)
template <typename T>
__global__ void __launch_bounds__(256) pool_linear_fwd_k(const T* __restrict__ x, int Hf, int Wf, int C, int P,
                                                         const float* __restrict__ Wt, const float* __restrict__ bias,
                                                         int ncls, float* __restrict__ pooled, float* __restrict__ logits) {
  pdl_entry();
  extern __shared__ float feat[];  // F
  const int b = blockIdx.x;
  const int ph = Hf / P, pw = Wf / P, F = C * ph * pw;
  const float inv = 1.f / (float)(P * P);
  const T* xb = x + (long long)b * Hf * Wf * C;
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    // iterate f in an order where consecutive threads take consecutive channels
    const int c = f % C, cell = f / C;
    const int hp = cell / pw, wp = cell % pw;
    float s = 0.f;
    for (int i = 0; i < P; ++i)
      for (int j = 0; j < P; ++j) s += to_f<T>(xb[((long long)(hp * P + i) * Wf + (wp * P + j)) * C + c]);
    const int fi = c * (ph * pw) + cell;
    float v = s * inv;
    feat[fi] = v;
    pooled[(long long)b * F + fi] = v;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int k = warp; k < ncls; k += nwarp) {
    float s = 0.f;
    for (int f = lane; f < F; f += 32) s = fmaf(Wt[(long long)k * F + f], feat[f], s);
    s = warp_sum(s);
    if (lane == 0) logits[(long long)b * ncls + k] = s + bias[k];
  }
}

template <typename T>
__global__ void __launch_bounds__(256) pool_linear_bwd_dx_k(const float* __restrict__ dlogits, const float* __restrict__ Wt,
                                                            int Hf, int Wf, int C, int P, int ncls, T* __restrict__ dx) {
  pdl_entry();
  extern __shared__ float dfeat[];  // F
  __shared__ float dl[64];
  const int b = blockIdx.x;
  const int ph = Hf / P, pw = Wf / P, F = C * ph * pw;
  if ((int)threadIdx.x < ncls) dl[threadIdx.x] = dlogits[(long long)b * ncls + threadIdx.x];
  __syncthreads();
  const float inv = 1.f / (float)(P * P);
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < ncls; ++k) s = fmaf(dl[k], Wt[(long long)k * F + f], s);
    dfeat[f] = s * inv;
  }
  __syncthreads();
  T* xb = dx + (long long)b * Hf * Wf * C;
  const int tot = Hf * Wf * C;
  if ((C & 7) == 0) {  // 8 channels (one 16-byte bf16 store) per thread
    const int C8 = C >> 3, pp = ph * pw;
    for (int e8 = threadIdx.x; e8 < tot / 8; e8 += blockDim.x) {
      const int c = (e8 % C8) * 8, pix = e8 / C8;
      const int h = pix / Wf, w = pix % Wf;
      const int hp = h / P, wp = w / P;
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (hp < ph && wp < pw) ? dfeat[(c + j) * pp + hp * pw + wp] : 0.f;
      st8<T>(xb + (long long)e8 * 8, v);
    }
    return;
  }
  for (int e = threadIdx.x; e < tot; e += blockDim.x) {
    const int c = e % C, pix = e / C;
    const int h = pix / Wf, w = pix % Wf;
    const int hp = h / P, wp = w / P;
    float v = (hp < ph && wp < pw) ? dfeat[c * (ph * pw) + hp * pw + wp] : 0.f;
    xb[e] = from_f<T>(v);
  }
}

// dW[k][f] = sum_b dlogits[b][k] * pooled[b][f]; db[k] = sum_b dlogits[b][k]
// grid (F / 32, ncls), block (32 features, 8 batch slices): fixed-order shared-memory reduction over the slices
__global__ void __launch_bounds__(256) linear_wgrad_k(const float* __restrict__ dlogits, const float* __restrict__ pooled,
                                                      int B, int F, int ncls, float* __restrict__ dW, float* __restrict__ db) {
  pdl_entry();
  __shared__ float red[8][33];
  const int f = blockIdx.x * 32 + threadIdx.x;
  const int k = blockIdx.y;
  float s = 0.f, sb = 0.f;
  for (int b = threadIdx.y; b < B; b += 8) {
    const float g = dlogits[(long long)b * ncls + k];
    if (f < F) s = fmaf(g, pooled[(long long)b * F + f], s);
    sb += g;
  }
  red[threadIdx.y][threadIdx.x] = s;
  if (threadIdx.x == 0) red[threadIdx.y][32] = sb;
  __syncthreads();
  if (threadIdx.y == 0) {
    float t = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) t += red[y][threadIdx.x];
    if (f < F) dW[(long long)k * F + f] = t;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      float tb = 0.f;
#pragma unroll
      for (int y = 0; y < 8; ++y) tb += red[y][32];
      db[k] = tb;
    }
  }
}

extern "C" int combat_pool_linear_fwd(const void* x, int dtype, int B, int Hf, int Wf, int C, int P, const float* W,
                                      const float* b, int ncls, float* pooled, float* logits, void* stream) {
  COMBAT_ARG(x && W && b && pooled && logits, 0);
  COMBAT_ARG(P > 0 && Hf >= P && Wf >= P, 6);
  int F = C * (Hf / P) * (Wf / P);
  size_t smem = (size_t)F * sizeof(float);
  COMBAT_ARG(smem <= 200 * 1024, 5);
  DISPATCH_DTYPE(dtype, {
    cudaFuncSetAttribute(pool_linear_fwd_k<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    pdl_launch(pool_linear_fwd_k<T>, B, 256, smem, (cudaStream_t)stream, (const T*)x, Hf, Wf, C, P, W, b, ncls, pooled, logits);
  })
  COMBAT_RETURN_LAUNCH("pool_linear_fwd");
}

extern "C" int combat_pool_linear_bwd(const float* dlogits, const float* pooled, const float* W, int B, int Hf, int Wf,
                                      int C, int P, int ncls, void* dx, int dtype, float* dW, float* db, void* stream) {
  COMBAT_ARG(dlogits && W, 0);
  COMBAT_ARG(ncls <= 64, 8);
  int F = C * (Hf / P) * (Wf / P);
  size_t smem = (size_t)F * sizeof(float);
  if (dx) {
    DISPATCH_DTYPE(dtype, {
      cudaFuncSetAttribute(pool_linear_bwd_dx_k<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      pdl_launch(pool_linear_bwd_dx_k<T>, B, 256, smem, (cudaStream_t)stream, dlogits, W, Hf, Wf, C, P, ncls, (T*)dx);
    })
    COMBAT_CHECK_LAUNCH("pool_linear_bwd_dx");
  }
  if (dW) {
    COMBAT_ARG(pooled && db, 1);
    dim3 grid(cdiv(F, 32), ncls), block(32, 8);
    pdl_launch(linear_wgrad_k, grid, block, 0, (cudaStream_t)stream, dlogits, pooled, B, F, ncls, dW, db);
    COMBAT_CHECK_LAUNCH("linear_wgrad");
  }
  return 0;
}

// ------------------------------------------------------------------ maxpool 2x2, layout conversion, one-hot planes
template <typename T>
__global__ void __launch_bounds__(256) maxpool2_k(const T* __restrict__ x, T* __restrict__ y, long long total, int H, int W,
                                                  int C) {
  pdl_entry();
  const int Ho = H / 2, Wo = W / 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long p = i / C;
    const int ow = (int)(p % Wo);
    p /= Wo;
    const int oh = (int)(p % Ho);
    const long long n = p / Ho;
    const T* xb = x + ((n * H + 2 * oh) * W + 2 * ow) * C + c;
    float m = fmaxf(fmaxf(to_f<T>(xb[0]), to_f<T>(xb[C])), fmaxf(to_f<T>(xb[(long long)W * C]), to_f<T>(xb[(long long)W * C + C])));
    y[i] = from_f<T>(m);
  }
}
template <typename T>
__global__ void __launch_bounds__(256) maxpool2_v8_k(const T* __restrict__ x, T* __restrict__ y, long long total8, int H, int W,
                                                     int C) {
  pdl_entry();
  const int Ho = H / 2, Wo = W / 2, C8 = C / 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (long long)gridDim.x * blockDim.x) {
    const int c = 8 * (int)(i % C8);
    long long p = i / C8;
    const int ow = (int)(p % Wo);
    p /= Wo;
    const int oh = (int)(p % Ho);
    const long long n = p / Ho;
    const T* xb = x + ((n * H + 2 * oh) * W + 2 * ow) * C + c;
    float a[8], b[8], cc[8], d[8], m[8];
    ld8<T>(xb, a);
    ld8<T>(xb + C, b);
    ld8<T>(xb + (long long)W * C, cc);
    ld8<T>(xb + (long long)W * C + C, d);
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = fmaxf(fmaxf(a[j], b[j]), fmaxf(cc[j], d[j]));
    st8<T>(y + i * 8, m);
  }
}
extern "C" int combat_maxpool2(const void* x, void* y, int dtype, int N, int H, int W, int C, void* stream) {
  COMBAT_ARG(x && y, 0);
  long long total = (long long)N * (H / 2) * (W / 2) * C;
  if (total <= 0) return 0;
  if ((C % 8) == 0) {
    DISPATCH_DTYPE(dtype, pdl_launch(maxpool2_v8_k<T>, ew_grid(total / 8), 256, 0, (cudaStream_t)stream, (const T*)x, (T*)y, total / 8, H, W, C);)
    COMBAT_RETURN_LAUNCH("maxpool2");
  }
  DISPATCH_DTYPE(dtype, pdl_launch(maxpool2_k<T>, ew_grid(total), 256, 0, (cudaStream_t)stream, (const T*)x, (T*)y, total, H, W, C);)
  COMBAT_RETURN_LAUNCH("maxpool2");
}

template <typename T>
__global__ void __launch_bounds__(256) nchw_to_nhwc_k(const float* __restrict__ x, T* __restrict__ y, long long total, int C,
                                                      int HW) {
  pdl_entry();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long p = i / C;
    const int hw = (int)(p % HW);
    const long long n = p / HW;
    y[i] = from_f<T>(x[(n * C + c) * HW + hw]);
  }
}
template <typename T>
__global__ void __launch_bounds__(256) nhwc_to_nchw_k(const T* __restrict__ x, float* __restrict__ y, long long total, int C,
                                                      int HW) {
  pdl_entry();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int hw = (int)(i % HW);
    long long p = i / HW;
    const int c = (int)(p % C);
    const long long n = p / C;
    y[i] = to_f<T>(x[(n * HW + hw) * C + c]);
  }
}
extern "C" int combat_nchw_to_nhwc(const float* x, void* y, int dtype, int N, int C, int H, int W, void* stream) {
  COMBAT_ARG(x && y, 0);
  long long total = (long long)N * C * H * W;
  if (total <= 0) return 0;
  DISPATCH_DTYPE(dtype, pdl_launch(nchw_to_nhwc_k<T>, ew_grid(total), 256, 0, (cudaStream_t)stream, x, (T*)y, total, C, H * W);)
  COMBAT_RETURN_LAUNCH("nchw_to_nhwc");
}
extern "C" int combat_nhwc_to_nchw(const void* x, int dtype, float* y, int N, int C, int H, int W, void* stream) {
  COMBAT_ARG(x && y, 0);
  long long total = (long long)N * C * H * W;
  if (total <= 0) return 0;
  DISPATCH_DTYPE(dtype, pdl_launch(nhwc_to_nchw_k<T>, ew_grid(total), 256, 0, (cudaStream_t)stream, (const T*)x, y, total, C, H * W);)
  COMBAT_RETURN_LAUNCH("nhwc_to_nchw");
}

// writes one-hot(label) planes into channels [c_off, c_off+ncls) of an NHWC tensor with Ctot channels
template <typename T>
__global__ void __launch_bounds__(256) onehot_planes_k(T* __restrict__ y, const long long* __restrict__ labels, long long total,
                                                       int HW, int Ctot, int c_off, int ncls) {
  pdl_entry();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % ncls);
    const long long pix = i / ncls;
    const long long n = pix / HW;
    y[pix * Ctot + c_off + k] = from_f<T>(labels[n] == k ? 1.f : 0.f);
  }
}
extern "C" int combat_onehot_planes(void* y, int dtype, const long long* labels, int N, int HW, int Ctot, int c_off,
                                    int ncls, void* stream) {
  COMBAT_ARG(y && labels, 0);
  long long total = (long long)N * HW * ncls;
  if (total <= 0) return 0;
  DISPATCH_DTYPE(dtype, pdl_launch(onehot_planes_k<T>, ew_grid(total), 256, 0, (cudaStream_t)stream, (T*)y, labels, total, HW, Ctot,
                                                                                            c_off, ncls);)
  COMBAT_RETURN_LAUNCH("onehot_planes");
}

// dst[pix, c_off + c] = leaky_relu(src[pix, c]) for c < Csrc: writes an activated copy into a channel slice of a wider
// NHWC tensor (CUnetGeneratorv1 concatenation, networks/models.py:524-531)
template <typename T>
__global__ void __launch_bounds__(256) lrelu_into_slice_k(const T* __restrict__ src, T* __restrict__ dst, long long total,
                                                          int Csrc, int Cdst, int c_off, float slope) {
  pdl_entry();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Csrc);
    const long long pix = i / Csrc;
    float v = to_f<T>(src[i]);
    dst[pix * Cdst + c_off + c] = from_f<T>(v > 0.f ? v : v * slope);
  }
}
extern "C" int combat_lrelu_into_slice(const void* src, void* dst, int dtype, long long npix, int Csrc, int Cdst, int c_off,
                                       float slope, void* stream) {
  COMBAT_ARG(src && dst && c_off + Csrc <= Cdst, 0);
  long long total = npix * Csrc;
  if (total <= 0) return 0;
  DISPATCH_DTYPE(dtype, pdl_launch(lrelu_into_slice_k<T>, ew_grid(total), 256, 0, (cudaStream_t)stream, (const T*)src, (T*)dst, total,
                                                                                               Csrc, Cdst, c_off, slope);)
  COMBAT_RETURN_LAUNCH("lrelu_into_slice");
}
