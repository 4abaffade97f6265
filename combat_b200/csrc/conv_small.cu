// Direct convolutions for the image-boundary layers that are not tensor-core shaped (SURVEY.md section 7 "hard
// parts"): 3 input channels (K = 27: classifier conv1, generator conv0_0, FrequencyModel conv1, and the input
// gradient of upconv0_0) and 3 output channels (generator upconv0_0 and the input gradient of the classifiers'
// conv1).  They are memory-bound: the 3-channel side is NCHW float32 (the reference's image tensors), the wide
// side is NHWC in the activation dtype.  Replaces the corresponding aten::conv2d / convolution_backward calls.
#include <stdlib.h>
#include <string.h>

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
__global__ void __launch_bounds__(256) conv_cin3_generic_k(const float* __restrict__ x, const TW* __restrict__ w,
                                                   const float* __restrict__ bias, TO* __restrict__ out, int N, int H, int W,
                                                   int Ho, int Wo, int Co, int stride, int act,
                                                   const float* __restrict__ post_scale, const float* __restrict__ post_shift) {
  pdl_entry();
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
__global__ void __launch_bounds__(256) conv_cout3_generic_k(const TI* __restrict__ in, const TW* __restrict__ w,
                                                    const float* __restrict__ bias, float* __restrict__ out, int N, int H, int W,
                                                    int act) {
  pdl_entry();
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

// =====================================================================================================================
// Register-tiled kernels (the paths the step uses).  All three are fp32-FMA bound at ~0.9 GFMA per 512-image batch
// (K = 27 or N = 3 cannot feed a 128-wide tensor tile); the tiling keeps shared-memory and global load instructions
// at <= 1 per 4..8 FMAs so the FMA pipe, not the load/store unit, is the limiter.
// =====================================================================================================================

// ---------------------------------------------------------------------------------- 3 -> Co, PX pixels x 8 channels per thread
// x: NCHW float32 [N,3,H,W]; w: [Co][9][3] (TW); out: NHWC [N,Ho,Wo,Co] (TO); Wo % PX == 0.
// A weight float read from shared memory costs one LSU cycle per warp however it is vectorised or broadcast, an FMA
// warp-instruction a quarter of an SM cycle: PX = 8 pixels per weight keeps the FMA pipe, not the LSU, the limiter.
template <typename TW, typename TO, int S, int PX>
__global__ void __launch_bounds__(256) conv_cin3_k(const float* __restrict__ x, const TW* __restrict__ w,
                                                   const float* __restrict__ bias, TO* __restrict__ out, int N, int H, int W,
                                                   int Ho, int Wo, int Co, int act, const float* __restrict__ post_scale,
                                                   const float* __restrict__ post_shift, bf16* __restrict__ out2,
                                                   const float* __restrict__ scale2, const float* __restrict__ shift2) {
  pdl_entry();
  extern __shared__ float ws[];  // [27][Co]
  for (int e = threadIdx.x; e < 27 * Co; e += blockDim.x) {
    const int co = e / 27, k = e % 27;
    ws[k * Co + co] = to_f<TW>(w[e]);
  }
  __syncthreads();
  constexpr int NV = (PX - 1) * S + 3;     // input columns touched by PX adjacent output pixels
  const int tpq = Co >> 3;                 // threads per pixel group
  const int qpb = blockDim.x / tpq;        // groups per block iteration
  const int cg = (threadIdx.x % tpq) * 8;
  const int ql = threadIdx.x / tpq;
  const int Wq = Wo / PX;
  const long long Q = (long long)N * Ho * Wq;
  const long long HW = (long long)H * W;
  for (long long q = (long long)blockIdx.x * qpb + ql; q < Q; q += (long long)gridDim.x * qpb) {
    const int ow0 = (int)(q % Wq) * PX;
    const long long r = q / Wq;
    const int oh = (int)(r % Ho);
    const long long n = r / Ho;
    float acc[PX][8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float b = bias ? bias[cg + c] : 0.f;
#pragma unroll
      for (int j = 0; j < PX; ++j) acc[j][c] = b;
    }
    const float* xb = x + n * 3 * HW;
    const int iw0 = ow0 * S - 1;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int ih = oh * S - 1 + kh;
      if (ih < 0 || ih >= H) continue;
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) {
        const float* xr = xb + ci * HW + (long long)ih * W;
        float v[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int iw = iw0 + i;
          v[i] = (iw >= 0 && iw < W) ? __ldg(xr + iw) : 0.f;
        }
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const float* wp = ws + ((kh * 3 + kw) * 3 + ci) * Co + cg;
          const float4 w0 = *(const float4*)wp, w1 = *(const float4*)(wp + 4);
#pragma unroll
          for (int j = 0; j < PX; ++j) {
            const float a = v[j * S + kw];
            acc[j][0] = fmaf(a, w0.x, acc[j][0]); acc[j][1] = fmaf(a, w0.y, acc[j][1]);
            acc[j][2] = fmaf(a, w0.z, acc[j][2]); acc[j][3] = fmaf(a, w0.w, acc[j][3]);
            acc[j][4] = fmaf(a, w1.x, acc[j][4]); acc[j][5] = fmaf(a, w1.y, acc[j][5]);
            acc[j][6] = fmaf(a, w1.z, acc[j][6]); acc[j][7] = fmaf(a, w1.w, acc[j][7]);
          }
        }
      }
    }
    const long long m0 = (r * Wo + ow0);
#pragma unroll
    for (int j = 0; j < PX; ++j) {
      if (act == 2) {
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[j][c] = acc[j][c] > 0.f ? acc[j][c] : expm1f(acc[j][c]);
      }
      if (post_scale) {
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[j][c] = fmaf(acc[j][c], post_scale[cg + c], post_shift[cg + c]);
      }
      store8<TO>(out + (m0 + j) * Co + cg, acc[j]);
      if (out2) {  // eval-mode BatchNorm + ReLU of the first PreAct block, fused (preact_resnet.py:29-31)
        float u[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) u[c] = fmaxf(fmaf(acc[j][c], scale2[cg + c], shift2[cg + c]), 0.f);
        store8<bf16>(out2 + (m0 + j) * Co + cg, u);
      }
    }
  }
}

template <typename T>
__device__ __forceinline__ void load8(const T* p, float* v);
template <>
__device__ __forceinline__ void load8<float>(const float* p, float* v) {
  const float4 a = ((const float4*)p)[0], b = ((const float4*)p)[1];
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<bf16>(const bf16* p, float* v) {
  const uint4 u = *(const uint4*)p;
  const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(*(const __nv_bfloat162*)&w4[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}

// ---------------------------------------------------------------------------------- 64 -> 3, PX pixels per thread
// in: NHWC [N,H,W,64] (TI); w: [3][9][64] (TW); out: NCHW float32 [N,3,H,W]; W % PX == 0.
// Weights sit in shared memory as [9][64] float4 (co0, co1, co2, 0): every weight read is a warp-wide broadcast and
// feeds 3 * PX FMAs.  PX = 8 keeps the kernel FMA-bound (see conv_cin3_k); PX = 2 serves narrow / small problems.
template <typename TI, typename TW, int PX>
__global__ void __launch_bounds__(128) conv_cout3_k(const TI* __restrict__ in, const TW* __restrict__ w,
                                                    const float* __restrict__ bias, float* __restrict__ out, int N, int H, int W,
                                                    int act) {
  pdl_entry();
  __shared__ float4 wsm[9 * 64];
  for (int e = threadIdx.x; e < 9 * 64; e += blockDim.x)
    wsm[e] = make_float4(to_f<TW>(w[e]), to_f<TW>(w[576 + e]), to_f<TW>(w[1152 + e]), 0.f);
  __syncthreads();
  const int Wp = W / PX;
  const long long P = (long long)N * H * Wp, HW = (long long)H * W;
  const float b0 = bias ? bias[0] : 0.f, b1 = bias ? bias[1] : 0.f, b2 = bias ? bias[2] : 0.f;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
    const int ow0 = (int)(p % Wp) * PX;
    const long long r = p / Wp;
    const int oh = (int)(r % H);
    const long long n = r / H;
    float a[PX][3];
#pragma unroll
    for (int j = 0; j < PX; ++j) { a[j][0] = b0; a[j][1] = b1; a[j][2] = b2; }
#pragma unroll 1
    for (int kh = 0; kh < 3; ++kh) {
      const int ih = oh - 1 + kh;
      if (ih < 0 || ih >= H) continue;
      const TI* rowp = in + ((n * H + ih) * W) * 64;
#pragma unroll 1
      for (int ch = 0; ch < 8; ++ch) {
        float v[PX + 2][8];
#pragma unroll
        for (int c = 0; c < PX + 2; ++c) {
          const int iw = ow0 - 1 + c;
          if (iw >= 0 && iw < W) {
            load8<TI>(rowp + (long long)iw * 64 + ch * 8, v[c]);
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[c][i] = 0.f;
          }
        }
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const float4* wp = wsm + (kh * 3 + kw) * 64 + ch * 8;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 ww = wp[i];
#pragma unroll
            for (int j = 0; j < PX; ++j) {
              a[j][0] = fmaf(v[j + kw][i], ww.x, a[j][0]);
              a[j][1] = fmaf(v[j + kw][i], ww.y, a[j][1]);
              a[j][2] = fmaf(v[j + kw][i], ww.z, a[j][2]);
            }
          }
        }
      }
    }
    float* op = out + n * 3 * HW + (long long)oh * W + ow0;
#pragma unroll
    for (int co = 0; co < 3; ++co) {
      float y[PX];
#pragma unroll
      for (int j = 0; j < PX; ++j) y[j] = act == 1 ? tanhf(a[j][co]) : a[j][co];
#pragma unroll
      for (int j = 0; j < PX; j += 2) *(float2*)(op + co * HW + j) = make_float2(y[j], y[j + 1]);
    }
  }
}

// ---------------------------------------------------------------------------------- weight gradient of both boundary convs
// G[c][t][j] = sum_m wide[m][c] * narrow[n, j, oh*S - 1 + kh, ow*S - 1 + kw]      (t = kh*3 + kw, j < 3, c < 64)
//   MODE 0 (3 -> 64 conv, wgrad_cin3): wide = dy NHWC, narrow = x NCHW;  dw[(c*9 + t)*3 + j] += G;  db[c] += sum_m wide[m][c]
//   MODE 1 (64 -> 3 conv, wgrad_cout3, S = 1): wide = layer input a NHWC, narrow = dz NCHW;
//                                              dw[(j*9 + 8 - t)*64 + c] += G;  db[j] += sum dz[j]
// Per tile of 64 wide pixels the block builds the im2col patch of the narrow tensor in shared memory as
// [64 pixels][4 slices][8] (7 patch columns + 1 pad per slice, 27 = 4*7 - 1 used).  A warp streams over pixels; lane
// (cg, ks) owns an 8-channel x 7-column register tile: per pixel one 16-byte global load of 8 wide channels, two
// broadcast 16-byte shared loads of the slice, 56 FMAs -- FMA-bound (a shared-memory float costs an LSU cycle per warp).
#define WG_TP 64
template <typename T>
__device__ __forceinline__ void load8w(const T* p, float* v);
template <>
__device__ __forceinline__ void load8w<float>(const float* p, float* v) { load8<float>(p, v); }
template <>
__device__ __forceinline__ void load8w<bf16>(const bf16* p, float* v) { load8<bf16>(p, v); }

template <typename T, int MODE>
__global__ void __launch_bounds__(256) wgrad3_k(const float* __restrict__ narrow, const T* __restrict__ wide,
                                                float* __restrict__ dw, float* __restrict__ db, int N, int Hn, int Wn, int Hw,
                                                int Ww, int S) {
  pdl_entry();
  __shared__ __align__(16) float patch[WG_TP * 32];
  __shared__ __align__(16) float red[4 * 32 * 64];  // tree reduction over the 8 warps: 4 x (32 lanes x 64 values)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cg = lane & 7, ks = lane >> 3;  // channels 8*cg .. 8*cg+7, patch columns 7*ks .. 7*ks+6
  float acc[8][7];
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int k = 0; k < 7; ++k) acc[c][k] = 0.f;
  float wsum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // MODE 0 bias gradient (lanes with ks == 0)
  float nsum[3] = {0.f, 0.f, 0.f};                           // MODE 1 bias gradient: centre patch column j, summed at build time
  const long long M = (long long)N * Hw * Ww, HWn = (long long)Hn * Wn;
  const long long tiles = (M + WG_TP - 1) / WG_TP;
  // patch builder: thread -> (pixel p = tid / 4, slice q = tid % 4): 7 columns k = 7q .. 7q+6, one pixel decode per tile
  const int bp = tid >> 2, bq = tid & 3;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long m0 = tile * WG_TP;
    __syncthreads();  // previous tile's readers are done
    {
      const long long m = m0 + bp;
      float pv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) pv[i] = 0.f;
      if (m < M) {
        const int ow = (int)(m % Ww);
        const long long r = m / Ww;
        const int oh = (int)(r % Hw);
        const long long n = r / Hw;
        const float* nb = narrow + n * 3 * HWn;
#pragma unroll
        for (int i = 0; i < 7; ++i) {
          const int k = bq * 7 + i;  // < 28; k == 27 is the unused tail
          const int t = k / 3, j = k - t * 3;
          const int ih = oh * S - 1 + t / 3, iw = ow * S - 1 + t % 3;
          if (k < 27 && ih >= 0 && ih < Hn && iw >= 0 && iw < Wn) pv[i] = __ldg(nb + j * HWn + (long long)ih * Wn + iw);
          if (MODE == 1 && t == 4) nsum[j] += pv[i];  // centre tap: every narrow pixel exactly once over all tiles
        }
      }
      float4* dst = (float4*)(patch + bp * 32 + bq * 8);
      dst[0] = make_float4(pv[0], pv[1], pv[2], pv[3]);
      dst[1] = make_float4(pv[4], pv[5], pv[6], 0.f);
    }
    __syncthreads();
    for (int p = warp; p < WG_TP; p += 8) {
      const long long m = m0 + p;
      if (m >= M) break;
      float g[8];
      load8w<T>(wide + m * 64 + cg * 8, g);
      const float4* pp = (const float4*)(patch + p * 32 + ks * 8);
      const float4 q0 = pp[0], q1 = pp[1];
      const float v[7] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z};
#pragma unroll
      for (int c = 0; c < 8; ++c)
#pragma unroll
        for (int k = 0; k < 7; ++k) acc[c][k] = fmaf(g[c], v[k], acc[c][k]);
      if (MODE == 0 && ks == 0) {
#pragma unroll
        for (int c = 0; c < 8; ++c) wsum[c] += g[c];
      }
    }
  }
  // tree reduction over the 8 warps (each holds a full partial [64 channels][28 columns]), then atomics by warp 0
  for (int half = 4; half >= 1; half >>= 1) {
    __syncthreads();
    if (warp >= half && warp < 2 * half) {
      float* o = red + ((warp - half) * 32 + lane) * 64;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
#pragma unroll
        for (int k = 0; k < 7; ++k) o[c * 8 + k] = acc[c][k];
        o[c * 8 + 7] = wsum[c];
      }
    }
    __syncthreads();
    if (warp < half) {
      const float* o = red + (warp * 32 + lane) * 64;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
#pragma unroll
        for (int k = 0; k < 7; ++k) acc[c][k] += o[c * 8 + k];
        wsum[c] += o[c * 8 + 7];
      }
    }
  }
  if (warp == 0) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int ch = cg * 8 + c;
#pragma unroll
      for (int i = 0; i < 7; ++i) {
        const int k = ks * 7 + i;
        if (k < 27) {
          const int t = k / 3, j = k - t * 3;
          if (MODE == 0) atomicAdd(dw + (long long)ch * 27 + k, acc[c][i]);
          else atomicAdd(dw + ((long long)j * 9 + (8 - t)) * 64 + ch, acc[c][i]);
        }
      }
      if (MODE == 0 && db && ks == 0) atomicAdd(db + ch, wsum[c]);
    }
  }
  if (MODE == 1 && db) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float v = warp_sum(nsum[j]);
      if (lane == 0) atomicAdd(db + j, v);
    }
  }
}

static int grid_for(long long work_items, int per_block) {
  long long g = (work_items + per_block - 1) / per_block;
  const long long cap = 148 * 8;
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

// ---------------------------------------------------------------------------------- im2col of a 3-channel image for tcgen05
// A[pix][64] (bf16, NHWC with 64 "channels") = [ hi(patch[0..26]), 0 x5, lo(patch[0..26]), 0 x5 ] with patch[(kh*3+kw)*3+ci] =
// x[n, ci, oh*S-1+kh, ow*S-1+kw] (zero outside), hi = bf16(v), lo = bf16(v - hi).  A 3x3 conv over 3 input channels then
// is a 1x1 tensor-core conv over A with the filter duplicated into both halves: K = 27 would waste a 64-deep MMA chunk
// anyway, so the second half carries the rounding residual and the image enters with ~16 mantissa bits at no extra cost.
__global__ void __launch_bounds__(256) im2col3_k(const float* __restrict__ x, bf16* __restrict__ A, int N, int H, int W, int Ho,
                                                 int Wo, int S) {
  pdl_entry();
  const long long total = (long long)N * Ho * Wo * 8;  // one 16-byte chunk (8 columns) per thread
  const long long HW = (long long)H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i & 7);
    const long long m = i >> 3;
    const int ow = (int)(m % Wo);
    const long long r = m / Wo;
    const int oh = (int)(r % Ho);
    const long long n = r / Ho;
    const bool lo = j >= 4;
    const int kk0 = (j & 3) * 8;
    const float* xb = x + n * 3 * HW;
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int kk = kk0 + q;
      float t = 0.f;
      if (kk < 27) {
        const int tap = kk / 3, ci = kk - tap * 3;
        const int ih = oh * S - 1 + tap / 3, iw = ow * S - 1 + tap % 3;
        if (ih >= 0 && ih < H && iw >= 0 && iw < W) t = __ldg(xb + ci * HW + (long long)ih * W + iw);
        if (lo) t = t - __bfloat162float(__float2bfloat16_rn(t));
      }
      v[q] = t;
    }
    store8<bf16>(A + m * 64 + j * 8, v);
  }
}

// Folds the weight gradient computed over the im2col3 operand, dW'[row][64] = [hi-part k<27 | 0 | lo-part | 0], into the
// filter gradient (ACCUMULATES):
//   mode 0 (3 -> rows conv)   : dw[row][k]                  += dW'[row][k] + dW'[row][32+k]
//   mode 1 (rows=64 -> 3 conv): dw[(j*9 + 8 - t)*64 + row]  += dW'[row][k] + dW'[row][32+k],  k = t*3 + j   (wgrad3_k MODE 1)
//                               db[j] += colsum[12+j] + colsum[44+j]  (centre tap of the operand == the 3-channel tensor)
__global__ void fold_w64_k(const float* __restrict__ dw64, float* __restrict__ dw, int rows, int mode,
                           const float* __restrict__ colsum, float* __restrict__ db) {
  pdl_entry();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows * 27) {
    const int row = i / 27, k = i - row * 27;
    const float v = dw64[row * 64 + k] + dw64[row * 64 + 32 + k];
    if (mode == 0) {
      dw[i] += v;
    } else {
      const int t = k / 3, j = k - t * 3;
      dw[(j * 9 + 8 - t) * rows + row] += v;
    }
  }
  if (mode == 1 && db && colsum && i < 3) db[i] += colsum[12 + i] + colsum[44 + i];
}

extern "C" int combat_fold_w64(const float* dw64, float* dw, int rows, int mode, const float* colsum, float* db, void* stream) {
  COMBAT_ARG(dw64 && dw && rows > 0 && (mode == 0 || mode == 1), 0);
  pdl_launch(fold_w64_k, cdiv(rows * 27, 256), 256, 0, (cudaStream_t)stream, dw64, dw, rows, mode, colsum, db);
  COMBAT_RETURN_LAUNCH("fold_w64");
}

extern "C" int combat_im2col3(const float* x, void* A, int N, int H, int W, int stride, void* stream) {
  COMBAT_ARG(x && A, 0);
  COMBAT_ARG(stride == 1 || stride == 2, 5);
  const int Ho = (H + 2 - 3) / stride + 1, Wo = (W + 2 - 3) / stride + 1;
  const long long total = (long long)N * Ho * Wo * 8;
  if (total <= 0) return 0;
  pdl_launch(im2col3_k, grid_for(total, 256), 256, 0, (cudaStream_t)stream, x, (bf16*)A, N, H, W, Ho, Wo, stride);
  COMBAT_RETURN_LAUNCH("im2col3");
}


extern "C" int combat_conv_cin3(const float* x, const void* w, int w_dtype, const float* bias, void* out, int out_dtype, int N,
                                int H, int W, int Co, int stride, int act, const float* post_scale, const float* post_shift,
                                void* out2, const float* scale2, const float* shift2, void* stream) {
  COMBAT_ARG(x && w && out, 0);
  COMBAT_ARG(!out2 || (scale2 && shift2), 15);
  COMBAT_ARG(Co % 8 == 0 && Co <= 256 && 256 % (Co / 8) == 0 && (stride == 1 || stride == 2), 9);
  const int Ho = (H + 2 - 3) / stride + 1, Wo = (W + 2 - 3) / stride + 1;
  const long long M = (long long)N * Ho * Wo;
  if (M <= 0) return 0;
  const size_t smem = (size_t)27 * Co * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  if (Wo % 4 == 0) {
    const int qpb = 256 / (Co / 8);
    // measured on B200 (scripts/bench_small.py): 8 pixels per thread halves the shared-memory weight reads per FMA but
    // its 190 registers leave one CTA per SM and the kernel latency-bound -- no gain over 4, which stays the default
    const int px = (Wo % 8 == 0 && getenv("COMBAT_CIN3_PX8")) ? 8 : 4;
    const int grid = grid_for(M / px, qpb);
#define LCI2(TW, TO, SS, PP)                                                                                      \
  pdl_launch(conv_cin3_k<TW, TO, SS, PP>, grid, 256, smem, st, x, (const TW*)w, bias, (TO*)out, N, H, W, Ho, Wo, Co, act, \
                                                       post_scale, post_shift, (bf16*)out2, scale2, shift2)
#define LCI(TW, TO)                                          \
  {                                                          \
    if (stride == 1) { if (px == 8) LCI2(TW, TO, 1, 8); else LCI2(TW, TO, 1, 4); } \
    else { if (px == 8) LCI2(TW, TO, 2, 8); else LCI2(TW, TO, 2, 4); }             \
  }
    if (w_dtype == COMBAT_F32) { if (out_dtype == COMBAT_F32) LCI(float, float) else LCI(float, bf16) }
    else { if (out_dtype == COMBAT_F32) LCI(bf16, float) else LCI(bf16, bf16) }
#undef LCI
#undef LCI2
    COMBAT_RETURN_LAUNCH("conv_cin3");
  }
  COMBAT_ARG(!out2, 15);  // the fused second output needs Wo % 4 == 0
  const int ppb = 256 / (Co / 8);
  const int grid = grid_for(M, ppb * 4);
#define LCI(TW, TO) pdl_launch(conv_cin3_generic_k<TW, TO>, grid, 256, smem, st, x, (const TW*)w, bias, (TO*)out, N, H, W, Ho, Wo, Co, stride, act, post_scale, post_shift)
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
  cudaStream_t st = (cudaStream_t)stream;
  if (W % 2 == 0) {
    const int px = (W % 8 == 0 && M / 8 >= 148 * 128) ? 8 : 2;  // 8 pixels per thread once that still fills the GPU
    const int grid = grid_for(M / px, 128);
#define LCO(TI, TW)                                                                                                     \
  {                                                                                                                     \
    if (px == 8) pdl_launch(conv_cout3_k<TI, TW, 8>, grid, 128, 0, st, (const TI*)in, (const TW*)w, bias, out, N, H, W, act);   \
    else pdl_launch(conv_cout3_k<TI, TW, 2>, grid, 128, 0, st, (const TI*)in, (const TW*)w, bias, out, N, H, W, act);           \
  }
    if (in_dtype == COMBAT_F32) { if (w_dtype == COMBAT_F32) LCO(float, float) else LCO(float, bf16) }
    else { if (w_dtype == COMBAT_F32) LCO(bf16, float) else LCO(bf16, bf16) }
#undef LCO
    COMBAT_RETURN_LAUNCH("conv_cout3");
  }
  const int grid = grid_for(M, 8 * 16);
#define LCO(TI, TW) pdl_launch(conv_cout3_generic_k<TI, TW>, grid, 256, 0, st, (const TI*)in, (const TW*)w, bias, out, N, H, W, act)
  if (in_dtype == COMBAT_F32) { if (w_dtype == COMBAT_F32) LCO(float, float); else LCO(float, bf16); }
  else { if (w_dtype == COMBAT_F32) LCO(bf16, float); else LCO(bf16, bf16); }
#undef LCO
  COMBAT_RETURN_LAUNCH("conv_cout3");
}

static int wgrad_grid(long long M) {
  const long long tiles = (M + WG_TP - 1) / WG_TP;
  const long long cap = 148 * 2;
  return (int)(tiles < cap ? (tiles < 1 ? 1 : tiles) : cap);
}

extern "C" int combat_wgrad_cin3(const float* x, const void* dy, int dy_dtype, float* dw, float* db, int N, int H, int W, int Co,
                                 int stride, void* stream) {
  COMBAT_ARG(x && dy && dw, 0);
  COMBAT_ARG(stride == 1 || stride == 2, 9);
  if (Co != 64) {  // other widths: generic strided kernel (not on the measured path)
    combat_conv_desc d;
    memset(&d, 0, sizeof(d));
    const int Ho_ = (H + 2 - 3) / stride + 1, Wo_ = (W + 2 - 3) / stride + 1;
    d.in = x; d.bias = db;
    d.N = N; d.Hi = H; d.Wi = W; d.Ci = 3; d.Ho = Ho_; d.Wo = Wo_; d.Co = Co; d.KH = d.KW = 3; d.stride = stride; d.pad = 1; d.up = 1;
    d.in_sn = 3LL * H * W; d.in_sc = (long long)H * W; d.in_sh = W; d.in_sw = 1;
    d.out_sn = (long long)Ho_ * Wo_ * Co; d.out_sh = (long long)Wo_ * Co; d.out_sw = Co; d.out_sc = 1;
    d.in_dtype = COMBAT_F32; d.w_dtype = COMBAT_F32; d.out_dtype = dy_dtype;
    return combat_conv_wgrad_simt(&d, dy, dy_dtype, dw, stream);
  }
  const int Ho = (H + 2 - 3) / stride + 1, Wo = (W + 2 - 3) / stride + 1;
  const long long M = (long long)N * Ho * Wo;
  if (M <= 0) return 0;
  const int grid = wgrad_grid(M);
  cudaStream_t st = (cudaStream_t)stream;
  if (dy_dtype == COMBAT_F32)
    pdl_launch(wgrad3_k<float, 0>, grid, 256, 0, st, x, (const float*)dy, dw, db, N, H, W, Ho, Wo, stride);
  else
    pdl_launch(wgrad3_k<bf16, 0>, grid, 256, 0, st, x, (const bf16*)dy, dw, db, N, H, W, Ho, Wo, stride);
  COMBAT_RETURN_LAUNCH("wgrad_cin3");
}

extern "C" int combat_wgrad_cout3(const void* a, int a_dtype, const float* dz, float* dw, float* db, int N, int H, int W, int Ci,
                                  void* stream) {
  COMBAT_ARG(a && dz && dw, 0);
  COMBAT_ARG(Ci == 64, 9);
  const long long M = (long long)N * H * W;
  if (M <= 0) return 0;
  const int grid = wgrad_grid(M);
  cudaStream_t st = (cudaStream_t)stream;
  if (a_dtype == COMBAT_F32)
    pdl_launch(wgrad3_k<float, 1>, grid, 256, 0, st, dz, (const float*)a, dw, db, N, H, W, H, W, 1);
  else
    pdl_launch(wgrad3_k<bf16, 1>, grid, 256, 0, st, dz, (const bf16*)a, dw, db, N, H, W, H, W, 1);
  COMBAT_RETURN_LAUNCH("wgrad_cout3");
}
