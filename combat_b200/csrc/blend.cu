// Fused poisoned-batch builder: gather rows, blend, clamp, 3x3 Gaussian blur (reflect), scatter into
// the [bd | rest-of-target | non-target] order, plus its backward and the fused MSE partial sums.
// Replaces train_generator.py:188-195,223-226,234 and torchvision GaussianBlur.
#include "common.cuh"

__device__ __forceinline__ int reflect1(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }
__device__ __forceinline__ float clamp1(float v) { return fminf(fmaxf(v, -1.f), 1.f); }

// one CTA per (row, channel) plane; the clamped plane is staged in shared memory when it fits
__global__ void __launch_bounds__(256) poison_blend_fwd_k(const float* __restrict__ x, const float* __restrict__ noise,
                                                          const int* __restrict__ perm, const int* __restrict__ nperm,
                                                          int num_bd, float rate, float k0, float k1,
                                                          float* __restrict__ out, float* __restrict__ sq_partial, int C,
                                                          int H, int W, int use_smem, const float* __restrict__ taps_dev,
                                                          const int* __restrict__ num_bd_dev) {
  extern __shared__ float tile[];  // H*W clamped values (use_smem)
  if (taps_dev) { k0 = taps_dev[0]; k1 = taps_dev[1]; }
  if (num_bd_dev) num_bd = num_bd_dev[0];
  const int plane = blockIdx.x;
  const int r = plane / C, c = plane % C;
  const int src = perm ? perm[r] : r;
  const int HW = H * W;
  const float* xp = x + ((long long)src * C + c) * HW;
  float* op = out + ((long long)r * C + c) * HW;
  float sq = 0.f;
  if (r >= num_bd) {  // pass-through row of the concatenation
    if ((HW & 3) == 0) {
      for (int i = threadIdx.x; i < HW / 4; i += blockDim.x) ((float4*)op)[i] = ((const float4*)xp)[i];
    } else {
      for (int i = threadIdx.x; i < HW; i += blockDim.x) op[i] = xp[i];
    }
  } else {
    const int nsrc = nperm ? nperm[r] : src;
    const float* np_ = noise + ((long long)nsrc * C + c) * HW;
    if (use_smem) {
      if ((HW & 3) == 0) {
        for (int i = threadIdx.x; i < HW / 4; i += blockDim.x) {
          float4 a = ((const float4*)xp)[i], b = ((const float4*)np_)[i];
          float4 v;
          v.x = clamp1(a.x + b.x * rate);
          v.y = clamp1(a.y + b.y * rate);
          v.z = clamp1(a.z + b.z * rate);
          v.w = clamp1(a.w + b.w * rate);
          ((float4*)tile)[i] = v;
        }
      } else {
        for (int i = threadIdx.x; i < HW; i += blockDim.x) tile[i] = clamp1(xp[i] + np_[i] * rate);
      }
      __syncthreads();
    }
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
      const int h = i / W, w = i % W;
      const int hm = reflect1(h - 1, H), hp = reflect1(h + 1, H);
      const int wm = reflect1(w - 1, W), wp = reflect1(w + 1, W);
      float v[9];
      const int hh[3] = {hm, h, hp}, ww[3] = {wm, w, wp};
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) {
          int j = hh[a] * W + ww[b];
          v[a * 3 + b] = use_smem ? tile[j] : clamp1(xp[j] + np_[j] * rate);
        }
      // depthwise conv with kernel2d = k1d (x) k1d, accumulated in the row-major tap order of the reference conv
      const float kk[3] = {k1, k0, k1};
      float acc = 0.f;
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) acc = fmaf(kk[a] * kk[b], v[a * 3 + b], acc);
      op[i] = acc;
      float d = acc - xp[i];
      sq = fmaf(d, d, sq);
    }
  }
  if (sq_partial) {
    __shared__ float red[8];
    sq = warp_sum(sq);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sq;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
      sq_partial[plane] = s;
    }
  }
}

// adjoint coefficient of the 1-D reflect-padded 3-tap filter: d out[p] / d v[q]
__device__ __forceinline__ float adj_coef(int p, int q, int n, float k0, float k1) {
  if (p == q) return k0;
  float c = k1;
  if ((p == 0 && q == 1) || (p == n - 1 && q == n - 2)) c += k1;  // the reflected tap lands on q as well
  return c;
}

__global__ void __launch_bounds__(256) poison_blend_bwd_k(const float* __restrict__ x, const float* __restrict__ noise,
                                                          const float* __restrict__ x_bd, const float* __restrict__ g1,
                                                          const float* __restrict__ g2, float mse_scale, float rate,
                                                          float k0, float k1, float* __restrict__ dnoise, int HW, int H,
                                                          int W, const float* __restrict__ taps_dev) {
  extern __shared__ float gt[];  // total upstream gradient of the plane
  if (taps_dev) { k0 = taps_dev[0]; k1 = taps_dev[1]; }
  const long long base = (long long)blockIdx.x * HW;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) {
    float g = g1[base + i];
    if (g2) g += g2[base + i];
    g = fmaf(mse_scale, x_bd[base + i] - x[base + i], g);
    gt[i] = g;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < HW; i += blockDim.x) {
    const int h = i / W, w = i % W;
    float acc = 0.f;
#pragma unroll
    for (int dh = -1; dh <= 1; ++dh) {
      int ph = h + dh;
      if (ph < 0 || ph >= H) continue;
      float ch = adj_coef(ph, h, H, k0, k1);
#pragma unroll
      for (int dw = -1; dw <= 1; ++dw) {
        int pw = w + dw;
        if (pw < 0 || pw >= W) continue;
        acc = fmaf(ch * adj_coef(pw, w, W, k0, k1), gt[ph * W + pw], acc);
      }
    }
    float pre = x[base + i] + noise[base + i] * rate;
    dnoise[base + i] = (pre >= -1.f && pre <= 1.f) ? acc * rate : 0.f;
  }
}

extern "C" int combat_poison_blend_fwd(const float* x, const float* noise, const int* perm, const int* nperm, int rows,
                                       int num_bd, float noise_rate, float k0, float k1, float* out, float* sq_partial,
                                       int C, int H, int W, const float* taps_dev, const int* num_bd_dev,
                                       void* stream) {
  COMBAT_ARG(x && out, 0);
  COMBAT_ARG((num_bd == 0 && !num_bd_dev) || noise, 1);
  COMBAT_ARG(H >= 2 && W >= 2 && C > 0 && num_bd >= 0 && num_bd <= rows, 5);
  if (rows <= 0) return 0;
  size_t smem = (size_t)H * W * sizeof(float);
  int use_smem = smem <= 64 * 1024;
  if (use_smem) cudaFuncSetAttribute(poison_blend_fwd_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  poison_blend_fwd_k<<<rows * C, 256, use_smem ? smem : 0, (cudaStream_t)stream>>>(x, noise, perm, nperm, num_bd, noise_rate,
                                                                                  k0, k1, out, sq_partial, C, H, W, use_smem,
                                                                                  taps_dev, num_bd_dev);
  COMBAT_RETURN_LAUNCH("poison_blend_fwd");
}

extern "C" int combat_poison_blend_bwd(const float* x, const float* noise, const float* x_bd, const float* g1,
                                       const float* g2, float mse_scale, float noise_rate, float k0, float k1,
                                       float* dnoise, int rows, int C, int H, int W, const float* taps_dev,
                                       void* stream) {
  COMBAT_ARG(x && noise && x_bd && g1 && dnoise, 0);
  COMBAT_ARG(H >= 3 && W >= 3, 12);
  if (rows <= 0) return 0;
  size_t smem = (size_t)H * W * sizeof(float);
  COMBAT_ARG(smem <= 200 * 1024, 13);
  cudaFuncSetAttribute(poison_blend_bwd_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  poison_blend_bwd_k<<<rows * C, 256, smem, (cudaStream_t)stream>>>(x, noise, x_bd, g1, g2, mse_scale, noise_rate, k0, k1,
                                                                    dnoise, H * W, H, W, taps_dev);
  COMBAT_RETURN_LAUNCH("poison_blend_bwd");
}
