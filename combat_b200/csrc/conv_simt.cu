// Generic CUDA-core implicit-GEMM convolution (float32 accumulate), any strides/layouts/dtypes.
// Role: (1) the float32 parity path for every nn.Conv2d of the hot path, (2) the production kernel for the
// layers that are not tensor-core shaped (K = 27 first convs, Cout = 3 last conv), (3) FrequencyModel.
// Replaces aten::conv2d / convolution_backward of the reference.
#include "common.cuh"

#define BM 64
#define BN 64
#define BK 16

struct ConvK {
  combat_conv_desc d;
  int M, K;
};

__device__ __forceinline__ float act_apply(float v, int act) {
  if (act == 1) return tanhf(v);
  if (act == 2) return v > 0.f ? v : expm1f(v);
  return v;
}

// gather one A element: pixel (n,oh,ow), reduction index k=(tap,ci)
__device__ __forceinline__ float gather_a(const combat_conv_desc& d, int n, int oh, int ow, int k) {
  const int ci = k % d.Ci, tap = k / d.Ci;
  const int kh = tap / d.KW, kw = tap % d.KW;
  int ih = oh * d.stride - d.pad + kh, iw = ow * d.stride - d.pad + kw;
  if (ih < 0 || iw < 0) return 0.f;
  if (d.up > 1) {
    if ((ih % d.up) | (iw % d.up)) return 0.f;
    ih /= d.up;
    iw /= d.up;
  }
  if (ih >= d.Hi || iw >= d.Wi) return 0.f;
  return ld_any(d.in, n * d.in_sn + ih * d.in_sh + iw * d.in_sw + ci * d.in_sc, d.in_dtype);
}

__global__ void __launch_bounds__(256) conv_simt_k(const ConvK p) {
  pdl_entry();
  const combat_conv_desc& d = p.d;
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int t = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int tx = t & 15, ty = t >> 4;  // compute mapping: 4 rows (ty*4..) x 4 cols (tx*4..)
  const int lk = t & 15, lr = t >> 4;  // load mapping
  // pixel decode for the 4 A rows this thread loads
  int pn[4], poh[4], pow_[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int m = m0 + lr + 16 * j;
    if (m < p.M) {
      pow_[j] = m % d.Wo;
      int q = m / d.Wo;
      poh[j] = q % d.Ho;
      pn[j] = q / d.Ho;
    } else {
      pn[j] = -1; poh[j] = 0; pow_[j] = 0;
    }
  }
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < p.K; k0 += BK) {
    const int k = k0 + lk;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float a = 0.f;
      if (k < p.K && pn[j] >= 0) a = gather_a(d, pn[j], poh[j], pow_[j], k);
      As[lk][lr + 16 * j] = a;
      const int co = n0 + lr + 16 * j;
      float b = 0.f;
      if (k < p.K && co < d.Co) b = ld_any(d.w, (long long)co * p.K + k, d.w_dtype);
      Bs[lk][lr + 16 * j] = b;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 av = *(const float4*)&As[kk][ty * 4];
      const float4 bv = *(const float4*)&Bs[kk][tx * 4];
      const float a[4] = {av.x, av.y, av.z, av.w}, b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
    const int ow = m % d.Wo;
    const int q = m / d.Wo;
    const int oh = q % d.Ho, n = q / d.Ho;
    const long long ob = n * d.out_sn + oh * d.out_sh + ow * d.out_sw;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = n0 + tx * 4 + j;
      if (co >= d.Co) continue;
      float v = acc[i][j];
      if (d.bias) v += d.bias[co];
      v = act_apply(v, d.act);
      if (d.post_scale) v = fmaf(v, d.post_scale[co], d.post_shift[co]);
      const long long o = ob + co * d.out_sc;
      if (d.residual) v += ld_any(d.residual, o, d.out_dtype);
      st_any(d.out, o, d.out_dtype, v);
    }
  }
}

// wgrad: dW[co][k] += sum_p dy[p][co] * A[p][k], p split over blockIdx.z
__global__ void __launch_bounds__(256) conv_wgrad_simt_k(const ConvK p, const void* __restrict__ dy, int dy_dtype,
                                                         float* __restrict__ dw, int p_per_z) {
  pdl_entry();
  const combat_conv_desc& d = p.d;
  __shared__ float Ds[BK][BM + 4];  // [pp][co]
  __shared__ float As[BK][BN + 4];  // [pp][k]
  const int t = threadIdx.x;
  const int co0 = blockIdx.x * BM, kb0 = blockIdx.y * BN;
  const int tx = t & 15, ty = t >> 4;
  const int lc = t & 63, lp = t >> 6;  // load mapping: column (co or k) x 4 row lanes
  const int pz0 = blockIdx.z * p_per_z;
  int pz1 = pz0 + p_per_z;
  if (pz1 > p.M) pz1 = p.M;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float bsum = 0.f;  // bias gradient: column sum of dy, taken by the k-tile 0 blocks
  float* db = (blockIdx.y == 0) ? (float*)d.bias : nullptr;
  for (int pb = pz0; pb < pz1; pb += BK) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int pp = lp + 4 * j;
      const int pix = pb + pp;
      float dv = 0.f, av = 0.f;
      if (pix < pz1) {
        const int ow = pix % d.Wo;
        const int q = pix / d.Wo;
        const int oh = q % d.Ho, n = q / d.Ho;
        const int co = co0 + lc;
        if (co < d.Co) dv = ld_any(dy, n * d.out_sn + oh * d.out_sh + ow * d.out_sw + co * d.out_sc, dy_dtype);
        const int k = kb0 + lc;
        if (k < p.K) av = gather_a(d, n, oh, ow, k);
      }
      Ds[pp][lc] = dv;
      As[pp][lc] = av;
    }
    __syncthreads();
    if (db && t < BM) {
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) bsum += Ds[kk][t];
    }
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a4 = *(const float4*)&Ds[kk][ty * 4];
      const float4 b4 = *(const float4*)&As[kk][tx * 4];
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  if (db && t < BM && co0 + t < d.Co) atomicAdd(db + co0 + t, bsum);
  // gradient buffer is channels-last like the master weights: dw[co][tap][ci] == dw[co * K + k]
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    if (co >= d.Co) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = kb0 + tx * 4 + j;
      if (k >= p.K) continue;
      atomicAdd(dw + (long long)co * p.K + k, acc[i][j]);
    }
  }
}

static int check_desc(const combat_conv_desc* d) {
  COMBAT_ARG(d && d->in, 0);
  COMBAT_ARG(d->N > 0 && d->Ci > 0 && d->Co > 0 && d->KH > 0 && d->KW > 0 && d->stride > 0 && d->up > 0, 0);
  COMBAT_ARG((long long)d->N * d->Ho * d->Wo < (1ll << 31), 0);
  return 0;
}

extern "C" int combat_conv_simt(const combat_conv_desc* d, void* stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  COMBAT_ARG(d->out, 0);
  ConvK p;
  p.d = *d;
  p.M = d->N * d->Ho * d->Wo;
  p.K = d->KH * d->KW * d->Ci;
  dim3 grid(cdiv(p.M, BM), cdiv(d->Co, BN));
  pdl_launch(conv_simt_k, grid, 256, 0, (cudaStream_t)stream, p);
  COMBAT_RETURN_LAUNCH("conv_simt");
}

extern "C" int combat_conv_wgrad_simt(const combat_conv_desc* d, const void* dy, int dy_dtype, float* dw_ohwi,
                                      void* stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  COMBAT_ARG(dy && dw_ohwi, 1);
  COMBAT_ARG(d->up == 1, 0);
  ConvK p;
  p.d = *d;
  p.M = d->N * d->Ho * d->Wo;
  p.K = d->KH * d->KW * d->Ci;
  const int tiles = cdiv(d->Co, BM) * cdiv(p.K, BN);
  int z = cdiv(148 * 4, tiles);
  const int zmax = cdiv(p.M, 4 * BK);
  if (z > zmax) z = zmax;
  if (z < 1) z = 1;
  int p_per_z = cdiv(p.M, z);
  p_per_z = cdiv(p_per_z, BK) * BK;
  z = cdiv(p.M, p_per_z);
  dim3 grid(cdiv(d->Co, BM), cdiv(p.K, BN), z);
  pdl_launch(conv_wgrad_simt_k, grid, 256, 0, (cudaStream_t)stream, p, dy, dy_dtype, dw_ohwi, p_per_z);
  COMBAT_RETURN_LAUNCH("conv_wgrad_simt");
}
