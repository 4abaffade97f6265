// Direct convolutions for the image-boundary layers that are not tensor-core shaped (SURVEY.md section 7 "hard
// parts"): 3 input channels (K = 27: classifier conv1, generator conv0_0, FrequencyModel conv1, and the input
// gradient of upconv0_0) and 3 output channels (generator upconv0_0 and the input gradient of the classifiers'
// conv1).  They are memory-bound: the 3-channel side is NCHW float32 (the reference's image tensors), the wide
// side is NHWC in the activation dtype.  Replaces the corresponding aten::conv2d / convolution_backward calls.
#include "common.cuh"

template <typename T>
__device__ __forceinline__ void store8(T* p, const float* v);
template <>
__device__ __forceinline__ void store8<float>(float* p, const float* v) {
  ((float4*)p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  ((float4*)p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
template <>
__device__ __forceinline__ void store8<bf16>(bf16* p, const float* v) {
  __nv_bfloat162 h[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *(uint4*)p = *(uint4*)h;
}

// ---------------------------------------------------------------------------------- 3 -> Co (3x3, pad 1, stride 1|2)
// x: NCHW float32 [N,3,H,W]; w: [Co][9][3] (TW); out: NHWC [N,Ho,Wo,Co] (TO).  8 channels per thread, Co/8 threads per
// pixel, weights staged in shared memory as [27][Co] float32.
template <typename TW, typename TO>
__global__ void __launch_bounds__(256) conv_cin3_k(const float* __restrict__ x, const TW* __restrict__ w,
                                                   const float* __restrict__ bias, TO* __restrict__ out, int N, int H, int W,
                                                   int Ho, int Wo, int Co, int stride, int act,
                                                   const float* __restrict__ post_scale, const float* __restrict__ post_shift) {
  extern __shared__ float ws[];  // [27][Co]
  for (int e = threadIdx.x; e < 27 * Co; e += blockDim.x) {
    const int co = e / 27, k = e % 27;
    ws[k * Co + co] = to_f<TW>(w[e]);
  }
  __syncthreads();
  const int tpp = Co >> 3;                 // threads per pixel
  const int ppb = blockDim.x / tpp;        // pixels per block iteration
  const int cg = (threadIdx.x % tpp) * 8;  // first channel of this thread
  const int pl = threadIdx.x / tpp;
  const long long M = (long long)N * Ho * Wo;
  const long long HW = (long long)H * W;
  for (long long m = (long long)blockIdx.x * ppb + pl; m < M; m += (long long)gridDim.x * ppb) {
    const int ow = (int)(m % Wo);
    const long long q = m / Wo;
    const int oh = (int)(q % Ho);
    const long long n = q / Ho;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = bias ? bias[cg + j] : 0.f;
    const float* xb = x + n * 3 * HW;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int ih = oh * stride - 1 + kh;
      if (ih < 0 || ih >= H) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int iw = ow * stride - 1 + kw;
        if (iw < 0 || iw >= W) continue;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) {
          const float v = xb[ci * HW + (long long)ih * W + iw];
          const float* wp = ws + ((kh * 3 + kw) * 3 + ci) * Co + cg;
          const float4 w0 = *(const float4*)wp, w1 = *(const float4*)(wp + 4);
          acc[0] = fmaf(v, w0.x, acc[0]); acc[1] = fmaf(v, w0.y, acc[1]);
          acc[2] = fmaf(v, w0.z, acc[2]); acc[3] = fmaf(v, w0.w, acc[3]);
          acc[4] = fmaf(v, w1.x, acc[4]); acc[5] = fmaf(v, w1.y, acc[5]);
          acc[6] = fmaf(v, w1.z, acc[6]); acc[7] = fmaf(v, w1.w, acc[7]);
        }
      }
    }
    if (act == 2) {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = acc[j] > 0.f ? acc[j] : expm1f(acc[j]);
    }
    if (post_scale) {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(acc[j], post_scale[cg + j], post_shift[cg + j]);
    }
    store8<TO>(out + m * Co + cg, acc);
  }
}

// ---------------------------------------------------------------------------------- 64 -> 3 (3x3, pad 1, stride 1)
// in: NHWC [N,H,W,64] (TI); w: [3][9][64] (TW); out: NCHW float32 [N,3,H,W] (+bias, optional tanh).
// One warp per pixel, 2 channels per lane, weights in registers, 3 warp reductions per pixel.
template <typename TI, typename TW>
__global__ void __launch_bounds__(256) conv_cout3_k(const TI* __restrict__ in, const TW* __restrict__ w,
                                                    const float* __restrict__ bias, float* __restrict__ out, int N, int H, int W,
                                                    int act) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  float wr[3][9][2];
#pragma unroll
  for (int co = 0; co < 3; ++co)
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      wr[co][t][0] = to_f<TW>(w[(co * 9 + t) * 64 + 2 * lane]);
      wr[co][t][1] = to_f<TW>(w[(co * 9 + t) * 64 + 2 * lane + 1]);
    }
  const float b0 = bias ? bias[0] : 0.f, b1 = bias ? bias[1] : 0.f, b2 = bias ? bias[2] : 0.f;
  const long long M = (long long)N * H * W, HW = (long long)H * W;
  for (long long m = warp; m < M; m += nwarps) {
    const int ow = (int)(m % W);
    const long long q = m / W;
    const int oh = (int)(q % H);
    const long long n = q / H;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int ih = oh - 1 + kh;
      if (ih < 0 || ih >= H) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int iw = ow - 1 + kw;
        if (iw < 0 || iw >= W) continue;
        const TI* p = in + ((n * H + ih) * W + iw) * 64 + 2 * lane;
        const float v0 = to_f<TI>(p[0]), v1 = to_f<TI>(p[1]);
        const int t = kh * 3 + kw;
        a0 = fmaf(v0, wr[0][t][0], fmaf(v1, wr[0][t][1], a0));
        a1 = fmaf(v0, wr[1][t][0], fmaf(v1, wr[1][t][1], a1));
        a2 = fmaf(v0, wr[2][t][0], fmaf(v1, wr[2][t][1], a2));
      }
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
    if (lane < 3) {
      float v = lane == 0 ? a0 + b0 : (lane == 1 ? a1 + b1 : a2 + b2);
      if (act == 1) v = tanhf(v);
      out[(n * 3 + lane) * HW + (long long)oh * W + ow] = v;
    }
  }
}

// ---------------------------------------------------------------------------------- wgrad, 3 input channels
// dw[co][9][3] (+)= sum_pix dy[pix][co] * x[pix @ tap][ci];  db[co] += sum_pix dy[pix][co].
// x NCHW float32, dy NHWC [N,Ho,Wo,Co] (T).  blockDim = (Co, 256/Co): a thread owns one co and strides over pixels.
template <typename T>
__global__ void __launch_bounds__(256) wgrad_cin3_k(const float* __restrict__ x, const T* __restrict__ dy,
                                                    float* __restrict__ dw, float* __restrict__ db, int N, int H, int W, int Ho,
                                                    int Wo, int Co, int stride) {
  const int co = threadIdx.x;
  float acc[28];
#pragma unroll
  for (int j = 0; j < 28; ++j) acc[j] = 0.f;
  const long long M = (long long)N * Ho * Wo, HW = (long long)H * W;
  for (long long m = (long long)blockIdx.x * blockDim.y + threadIdx.y; m < M; m += (long long)gridDim.x * blockDim.y) {
    const int ow = (int)(m % Wo);
    const long long q = m / Wo;
    const int oh = (int)(q % Ho);
    const long long n = q / Ho;
    const float g = to_f<T>(dy[m * Co + co]);
    acc[27] += g;
    const float* xb = x + n * 3 * HW;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int ih = oh * stride - 1 + kh;
      if (ih < 0 || ih >= H) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int iw = ow * stride - 1 + kw;
        if (iw < 0 || iw >= W) continue;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) acc[(kh * 3 + kw) * 3 + ci] = fmaf(g, xb[ci * HW + (long long)ih * W + iw], acc[(kh * 3 + kw) * 3 + ci]);
      }
    }
  }
  // reduce over threadIdx.y through shared memory, then one atomic per (co, k) per block
  extern __shared__ float red[];  // [blockDim.y][Co][28]
  float* mine = red + ((size_t)threadIdx.y * Co + co) * 28;
#pragma unroll
  for (int j = 0; j < 28; ++j) mine[j] = acc[j];
  __syncthreads();
  if (threadIdx.y == 0) {
    for (int y = 1; y < (int)blockDim.y; ++y) {
      const float* o = red + ((size_t)y * Co + co) * 28;
#pragma unroll
      for (int j = 0; j < 28; ++j) acc[j] += o[j];
    }
#pragma unroll
    for (int j = 0; j < 27; ++j) atomicAdd(dw + (long long)co * 27 + j, acc[j]);
    if (db) atomicAdd(db + co, acc[27]);
  }
}

// ---------------------------------------------------------------------------------- wgrad, 3 output channels
// dw[co<3][9][64] += sum_pix dz[n,co,oh,ow] * a[pix @ tap][ci];  db[co] += sum dz.   a NHWC [N,H,W,64] (T), dz NCHW float32.
template <typename T>
__global__ void __launch_bounds__(256) wgrad_cout3_k(const T* __restrict__ a, const float* __restrict__ dz,
                                                     float* __restrict__ dw, float* __restrict__ db, int N, int H, int W) {
  const int ci = threadIdx.x;  // 64
  float acc[27];
#pragma unroll
  for (int j = 0; j < 27; ++j) acc[j] = 0.f;
  float bsum = 0.f;
  const long long M = (long long)N * H * W, HW = (long long)H * W;
  for (long long m = (long long)blockIdx.x * blockDim.y + threadIdx.y; m < M; m += (long long)gridDim.x * blockDim.y) {
    const int ow = (int)(m % W);
    const long long q = m / W;
    const int oh = (int)(q % H);
    const long long n = q / H;
    const float* zp = dz + n * 3 * HW + (long long)oh * W + ow;
    const float g0 = zp[0], g1 = zp[HW], g2 = zp[2 * HW];
    if (ci < 3) bsum += ci == 0 ? g0 : (ci == 1 ? g1 : g2);
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int ih = oh - 1 + kh;
      if (ih < 0 || ih >= H) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int iw = ow - 1 + kw;
        if (iw < 0 || iw >= W) continue;
        const float v = to_f<T>(a[((n * H + ih) * W + iw) * 64 + ci]);
        const int t = kh * 3 + kw;
        acc[t] = fmaf(g0, v, acc[t]);
        acc[9 + t] = fmaf(g1, v, acc[9 + t]);
        acc[18 + t] = fmaf(g2, v, acc[18 + t]);
      }
    }
  }
  extern __shared__ float red[];  // [blockDim.y][64][28]
  float* mine = red + ((size_t)threadIdx.y * 64 + ci) * 28;
#pragma unroll
  for (int j = 0; j < 27; ++j) mine[j] = acc[j];
  mine[27] = bsum;
  __syncthreads();
  if (threadIdx.y == 0) {
    for (int y = 1; y < (int)blockDim.y; ++y) {
      const float* o = red + ((size_t)y * 64 + ci) * 28;
#pragma unroll
      for (int j = 0; j < 27; ++j) acc[j] += o[j];
      bsum += o[27];
    }
#pragma unroll
    for (int co = 0; co < 3; ++co)
#pragma unroll
      for (int t = 0; t < 9; ++t) atomicAdd(dw + ((long long)co * 9 + t) * 64 + ci, acc[co * 9 + t]);
    if (db && ci < 3) atomicAdd(db + ci, bsum);
  }
}

static int grid_for(long long work_items, int per_block) {
  long long g = (work_items + per_block - 1) / per_block;
  const long long cap = 148 * 8;
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

extern "C" int combat_conv_cin3(const float* x, const void* w, int w_dtype, const float* bias, void* out, int out_dtype, int N,
                                int H, int W, int Co, int stride, int act, const float* post_scale, const float* post_shift,
                                void* stream) {
  COMBAT_ARG(x && w && out, 0);
  COMBAT_ARG(Co % 8 == 0 && Co <= 256 && 256 % (Co / 8) == 0 && (stride == 1 || stride == 2), 9);
  const int Ho = (H + 2 - 3) / stride + 1, Wo = (W + 2 - 3) / stride + 1;
  const long long M = (long long)N * Ho * Wo;
  if (M <= 0) return 0;
  const int ppb = 256 / (Co / 8);
  const int grid = grid_for(M, ppb * 4);
  const size_t smem = (size_t)27 * Co * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
#define LCI(TW, TO) conv_cin3_k<TW, TO><<<grid, 256, smem, st>>>(x, (const TW*)w, bias, (TO*)out, N, H, W, Ho, Wo, Co, stride, act, post_scale, post_shift)
  if (w_dtype == COMBAT_F32) { if (out_dtype == COMBAT_F32) LCI(float, float); else LCI(float, bf16); }
  else { if (out_dtype == COMBAT_F32) LCI(bf16, float); else LCI(bf16, bf16); }
#undef LCI
  COMBAT_RETURN_LAUNCH("conv_cin3");
}

extern "C" int combat_conv_cout3(const void* in, int in_dtype, const void* w, int w_dtype, const float* bias, float* out, int N,
                                 int H, int W, int Ci, int act, void* stream) {
  COMBAT_ARG(in && w && out, 0);
  COMBAT_ARG(Ci == 64, 9);
  const long long M = (long long)N * H * W;
  if (M <= 0) return 0;
  const int grid = grid_for(M, 8 * 16);
  cudaStream_t st = (cudaStream_t)stream;
#define LCO(TI, TW) conv_cout3_k<TI, TW><<<grid, 256, 0, st>>>((const TI*)in, (const TW*)w, bias, out, N, H, W, act)
  if (in_dtype == COMBAT_F32) { if (w_dtype == COMBAT_F32) LCO(float, float); else LCO(float, bf16); }
  else { if (w_dtype == COMBAT_F32) LCO(bf16, float); else LCO(bf16, bf16); }
#undef LCO
  COMBAT_RETURN_LAUNCH("conv_cout3");
}

extern "C" int combat_wgrad_cin3(const float* x, const void* dy, int dy_dtype, float* dw, float* db, int N, int H, int W, int Co,
                                 int stride, void* stream) {
  COMBAT_ARG(x && dy && dw, 0);
  COMBAT_ARG(Co >= 32 && Co <= 256 && 256 % Co == 0 && (stride == 1 || stride == 2), 8);
  const int Ho = (H + 2 - 3) / stride + 1, Wo = (W + 2 - 3) / stride + 1;
  const long long M = (long long)N * Ho * Wo;
  if (M <= 0) return 0;
  dim3 block(Co, 256 / Co);
  const int grid = grid_for(M, block.y * 64);
  const size_t smem = (size_t)256 * 28 * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  if (dy_dtype == COMBAT_F32)
    wgrad_cin3_k<float><<<grid, block, smem, st>>>(x, (const float*)dy, dw, db, N, H, W, Ho, Wo, Co, stride);
  else
    wgrad_cin3_k<bf16><<<grid, block, smem, st>>>(x, (const bf16*)dy, dw, db, N, H, W, Ho, Wo, Co, stride);
  COMBAT_RETURN_LAUNCH("wgrad_cin3");
}

extern "C" int combat_wgrad_cout3(const void* a, int a_dtype, const float* dz, float* dw, float* db, int N, int H, int W, int Ci,
                                  void* stream) {
  COMBAT_ARG(a && dz && dw, 0);
  COMBAT_ARG(Ci == 64, 9);
  const long long M = (long long)N * H * W;
  if (M <= 0) return 0;
  dim3 block(64, 4);
  const int grid = grid_for(M, 4 * 64);
  const size_t smem = (size_t)256 * 28 * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  if (a_dtype == COMBAT_F32)
    wgrad_cout3_k<float><<<grid, block, smem, st>>>((const float*)a, dz, dw, db, N, H, W);
  else
    wgrad_cout3_k<bf16><<<grid, block, smem, st>>>((const bf16*)a, dz, dw, db, N, H, W);
  COMBAT_RETURN_LAUNCH("wgrad_cout3");
}
