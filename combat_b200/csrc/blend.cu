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
                                                          const int* __restrict__ num_bd_dev,
                                                          const float* __restrict__ taps_rows) {
  pdl_entry();
  extern __shared__ float tile[];  // H*W clamped values (use_smem)
  if (taps_dev) { k0 = taps_dev[0]; k1 = taps_dev[1]; }
  if (num_bd_dev) num_bd = num_bd_dev[0];
  const int plane = blockIdx.x;
  const int r = plane / C, c = plane % C;
  if (taps_rows) { k0 = taps_rows[2 * r]; k1 = taps_rows[2 * r + 1]; }
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

// Vectorised forward for W % 4 == 0 and W <= 128 (every shape on the path): a thread owns 4 horizontally adjacent pixels;
// the rows above / below are re-read from global memory (L1 hits: they are other threads' centre rows), the left / right
// neighbours come from the adjacent lanes by warp shuffle (a row is W/4 <= 32 consecutive lanes) -- no shared-memory
// tile, no barrier before the stencil.  Same arithmetic (and accumulation order) per pixel as the scalar kernel.
__device__ __forceinline__ float4 blend_clamp4(const float4 a, const float4 b, float rate) {
  return make_float4(clamp1(a.x + b.x * rate), clamp1(a.y + b.y * rate), clamp1(a.z + b.z * rate), clamp1(a.w + b.w * rate));
}

__global__ void __launch_bounds__(256) poison_blend_fwd_v4_k(const float* __restrict__ x, const float* __restrict__ noise,
                                                             const int* __restrict__ perm, const int* __restrict__ nperm,
                                                             int num_bd, float rate, float k0, float k1,
                                                             float* __restrict__ out, float* __restrict__ sq_partial, int C,
                                                             int H, int W, const float* __restrict__ taps_dev,
                                                             const int* __restrict__ num_bd_dev,
                                                             const float* __restrict__ taps_rows) {
  pdl_entry();
  if (taps_dev) { k0 = taps_dev[0]; k1 = taps_dev[1]; }
  if (num_bd_dev) num_bd = num_bd_dev[0];
  const int plane = blockIdx.x;
  const int r = plane / C, c = plane % C;
  if (taps_rows) { k0 = taps_rows[2 * r]; k1 = taps_rows[2 * r + 1]; }
  const int src = perm ? perm[r] : r;
  const int HW4 = (H * W) >> 2, W4 = W >> 2;
  const float4* xp = (const float4*)(x + ((long long)src * C + c) * H * W);
  float4* op = (float4*)(out + ((long long)r * C + c) * H * W);
  float sq = 0.f;
  if (r >= num_bd) {  // pass-through row of the concatenation
    for (int i = threadIdx.x; i < HW4; i += blockDim.x) op[i] = xp[i];
  } else {
    const int nsrc = nperm ? nperm[r] : src;
    const float4* np_ = (const float4*)(noise + ((long long)nsrc * C + c) * H * W);
    const float kk[3] = {k1, k0, k1};
    const int lane = threadIdx.x & 31;
    // every lane of a warp takes part in the shuffles: iterate whole warps (HW4 is a multiple of W4, W4 divides 32)
    for (int i0 = threadIdx.x - lane; i0 < HW4; i0 += blockDim.x) {
      const int i = i0 + lane;
      const bool act = i < HW4;
      const int ii = act ? i : HW4 - 1;
      const int h = ii / W4, q = ii - h * W4;
      const int hh[3] = {reflect1(h - 1, H), h, reflect1(h + 1, H)};
      const float4 xc = xp[ii];
      float v[3][6];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const int j = hh[a] * W4 + q;
        const float4 m = a == 1 ? blend_clamp4(xc, np_[j], rate) : blend_clamp4(xp[j], np_[j], rate);
        const float lft = __shfl_up_sync(0xffffffffu, m.w, 1), rgt = __shfl_down_sync(0xffffffffu, m.x, 1);
        v[a][0] = q == 0 ? m.y : lft;          // reflect: column -1 -> column 1
        v[a][1] = m.x; v[a][2] = m.y; v[a][3] = m.z; v[a][4] = m.w;
        v[a][5] = q == W4 - 1 ? m.z : rgt;     // reflect: column W -> column W-2
      }
      if (!act) continue;
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        // depthwise conv with kernel2d = k1d (x) k1d, accumulated in the row-major tap order of the reference conv
        float acc = 0.f;
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int b = 0; b < 3; ++b) acc = fmaf(kk[a] * kk[b], v[a][j + b], acc);
        o[j] = acc;
      }
      op[i] = make_float4(o[0], o[1], o[2], o[3]);
      float d = o[0] - xc.x; sq = fmaf(d, d, sq);
      d = o[1] - xc.y; sq = fmaf(d, d, sq);
      d = o[2] - xc.z; sq = fmaf(d, d, sq);
      d = o[3] - xc.w; sq = fmaf(d, d, sq);
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

// 4 x 4 pixel block per thread (H % 4 == 0, W % 4 == 0, W/4 * H/4 threads per plane, a multiple of 32 dividing 256): six
// staged rows feed four output rows, i.e. 12 independent 16-byte loads in flight per thread and a 1.5x (L1-served) re-read
// instead of 3x; 256 / (W/4 * H/4) planes per CTA.
__global__ void __launch_bounds__(256) poison_blend_fwd_b16_k(const float* __restrict__ x, const float* __restrict__ noise,
                                                              const int* __restrict__ perm, const int* __restrict__ nperm,
                                                              int num_bd, float rate, float k0, float k1,
                                                              float* __restrict__ out, float* __restrict__ sq_partial, int C,
                                                              int H, int W, int planes, const float* __restrict__ taps_dev,
                                                              const int* __restrict__ num_bd_dev,
                                                              const float* __restrict__ taps_rows) {
  pdl_entry();
  if (taps_dev) { k0 = taps_dev[0]; k1 = taps_dev[1]; }
  if (num_bd_dev) num_bd = num_bd_dev[0];
  const int W4 = W >> 2, H4 = H >> 2, TPP = W4 * H4;  // threads per plane
  const int ppc = blockDim.x / TPP;                   // planes per CTA
  const int pl = threadIdx.x / TPP, t = threadIdx.x - pl * TPP;
  const int plane = blockIdx.x * ppc + pl;
  const bool pvalid = plane < planes;
  const int pc = pvalid ? plane : planes - 1;
  const int r = pc / C, c = pc % C;
  if (taps_rows) { k0 = taps_rows[2 * r]; k1 = taps_rows[2 * r + 1]; }
  const int src = perm ? perm[r] : r;
  const float4* xp = (const float4*)(x + ((long long)src * C + c) * H * W);
  float4* op = (float4*)(out + ((long long)r * C + c) * H * W);
  const int hb = (t / W4) * 4, q = t % W4;
  float sq = 0.f;
  if (r >= num_bd) {  // pass-through row of the concatenation (whole warps share a plane: no divergence around shuffles)
    if (pvalid) {
#pragma unroll
      for (int j = 0; j < 4; ++j) op[(hb + j) * W4 + q] = xp[(hb + j) * W4 + q];
    }
  } else {
    const int nsrc = nperm ? nperm[r] : src;
    const float4* np_ = (const float4*)(noise + ((long long)nsrc * C + c) * H * W);
    float4 xv[6], nv[6];
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      const int j = reflect1(hb - 1 + a, H) * W4 + q;
      xv[a] = xp[j];
      nv[a] = np_[j];
    }
    float v[6][6];
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      const float4 m = blend_clamp4(xv[a], nv[a], rate);
      const float lft = __shfl_up_sync(0xffffffffu, m.w, 1), rgt = __shfl_down_sync(0xffffffffu, m.x, 1);
      v[a][0] = q == 0 ? m.y : lft;
      v[a][1] = m.x; v[a][2] = m.y; v[a][3] = m.z; v[a][4] = m.w;
      v[a][5] = q == W4 - 1 ? m.z : rgt;
    }
    const float kk[3] = {k1, k0, k1};
    if (pvalid) {
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float acc = 0.f;  // row-major tap order of the reference's depthwise conv
#pragma unroll
          for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) acc = fmaf(kk[a] * kk[b], v[rr + a][j + b], acc);
          o[j] = acc;
        }
        op[(hb + rr) * W4 + q] = make_float4(o[0], o[1], o[2], o[3]);
        const float4 xc = xv[rr + 1];
        float d = o[0] - xc.x; sq = fmaf(d, d, sq);
        d = o[1] - xc.y; sq = fmaf(d, d, sq);
        d = o[2] - xc.z; sq = fmaf(d, d, sq);
        d = o[3] - xc.w; sq = fmaf(d, d, sq);
      }
    }
  }
  if (sq_partial) {
    __shared__ float red[8];
    sq = warp_sum(sq);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sq;
    __syncthreads();
    if (t == 0 && pvalid) {
      const int w0 = threadIdx.x >> 5, nw = TPP >> 5;
      float s = 0.f;
      for (int i = 0; i < nw; ++i) s += red[w0 + i];
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
                                                          int W, const float* __restrict__ taps_dev, int C,
                                                          const float* __restrict__ taps_rows) {
  pdl_entry();
  extern __shared__ float gt[];  // total upstream gradient of the plane
  if (taps_dev) { k0 = taps_dev[0]; k1 = taps_dev[1]; }
  if (taps_rows) { k0 = taps_rows[2 * (blockIdx.x / C)]; k1 = taps_rows[2 * (blockIdx.x / C) + 1]; }
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

// Vectorised backward for W % 4 == 0: same per-pixel arithmetic, 4 pixels per thread.
__global__ void __launch_bounds__(256) poison_blend_bwd_v4_k(const float* __restrict__ x, const float* __restrict__ noise,
                                                             const float* __restrict__ x_bd, const float* __restrict__ g1,
                                                             const float* __restrict__ g2, float mse_scale, float rate,
                                                             float k0, float k1, float* __restrict__ dnoise, int H, int W,
                                                             const float* __restrict__ taps_dev, int C,
                                                             const float* __restrict__ taps_rows) {
  pdl_entry();
  extern __shared__ float gt[];  // total upstream gradient of the plane
  if (taps_dev) { k0 = taps_dev[0]; k1 = taps_dev[1]; }
  if (taps_rows) { k0 = taps_rows[2 * (blockIdx.x / C)]; k1 = taps_rows[2 * (blockIdx.x / C) + 1]; }
  const int HW4 = (H * W) >> 2, W4 = W >> 2;
  const long long base4 = (long long)blockIdx.x * HW4;
  const float4* x4 = (const float4*)x + base4;
  const float4* n4 = (const float4*)noise + base4;
  const float4* b4 = (const float4*)x_bd + base4;
  const float4* g14 = (const float4*)g1 + base4;
  const float4* g24 = g2 ? (const float4*)g2 + base4 : nullptr;
  float4* d4 = (float4*)dnoise + base4;
  float4 xkeep = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = threadIdx.x; i < HW4; i += blockDim.x) {
    float4 g = g14[i];
    if (g24) { const float4 t = g24[i]; g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w; }
    const float4 xb = b4[i], xx = x4[i];
    xkeep = xx;
    g.x = fmaf(mse_scale, xb.x - xx.x, g.x); g.y = fmaf(mse_scale, xb.y - xx.y, g.y);
    g.z = fmaf(mse_scale, xb.z - xx.z, g.z); g.w = fmaf(mse_scale, xb.w - xx.w, g.w);
    ((float4*)gt)[i] = g;
  }
  __syncthreads();
  const bool one_pass = HW4 <= (int)blockDim.x;
  for (int i = threadIdx.x; i < HW4; i += blockDim.x) {
    const int h = i / W4, w0 = (i - h * W4) * 4;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int dh = -1; dh <= 1; ++dh) {
      const int ph = h + dh;
      if (ph < 0 || ph >= H) continue;
      const float ch = adj_coef(ph, h, H, k0, k1);
      const float* row = gt + ph * W;
      const float4 m = *(const float4*)(row + w0);
      const float gv[6] = {w0 > 0 ? row[w0 - 1] : 0.f, m.x, m.y, m.z, m.w, w0 + 4 < W ? row[w0 + 4] : 0.f};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int w = w0 + j;
#pragma unroll
        for (int dw = -1; dw <= 1; ++dw) {
          const int pw = w + dw;
          if (pw < 0 || pw >= W) continue;
          acc[j] = fmaf(ch * adj_coef(pw, w, W, k0, k1), gv[j + 1 + dw], acc[j]);
        }
      }
    }
    const float4 xx = one_pass ? xkeep : x4[i], nn = n4[i];
    float4 o;
    float pre = xx.x + nn.x * rate; o.x = (pre >= -1.f && pre <= 1.f) ? acc[0] * rate : 0.f;
    pre = xx.y + nn.y * rate; o.y = (pre >= -1.f && pre <= 1.f) ? acc[1] * rate : 0.f;
    pre = xx.z + nn.z * rate; o.z = (pre >= -1.f && pre <= 1.f) ? acc[2] * rate : 0.f;
    pre = xx.w + nn.w * rate; o.w = (pre >= -1.f && pre <= 1.f) ? acc[3] * rate : 0.f;
    d4[i] = o;
  }
}

extern "C" int combat_poison_blend_fwd(const float* x, const float* noise, const int* perm, const int* nperm, int rows,
                                       int num_bd, float noise_rate, float k0, float k1, float* out, float* sq_partial,
                                       int C, int H, int W, const float* taps_dev, const int* num_bd_dev,
                                       const float* taps_rows, void* stream) {
  COMBAT_ARG(x && out, 0);
  COMBAT_ARG((num_bd == 0 && !num_bd_dev) || noise, 1);
  COMBAT_ARG(H >= 2 && W >= 2 && C > 0 && num_bd >= 0 && num_bd <= rows, 5);
  if (rows <= 0) return 0;
  size_t smem = (size_t)H * W * sizeof(float);
  int use_smem = smem <= 64 * 1024;
  {
    const int tpp = (W / 4) * (H / 4);
    if ((W & 3) == 0 && (H & 3) == 0 && tpp % 32 == 0 && tpp <= 256 && 256 % tpp == 0 && 32 % (W / 4) == 0) {
      const int planes = rows * C, ppc = 256 / tpp;
      pdl_launch(poison_blend_fwd_b16_k, (planes + ppc - 1) / ppc, 256, 0, (cudaStream_t)stream, 
          x, noise, perm, nperm, num_bd, noise_rate, k0, k1, out, sq_partial, C, H, W, planes, taps_dev, num_bd_dev, taps_rows);
      COMBAT_RETURN_LAUNCH("poison_blend_fwd");
    }
  }
  if ((W & 3) == 0 && W <= 128 && 32 % (W / 4) == 0) {  // a row = W/4 consecutive lanes of one warp
    pdl_launch(poison_blend_fwd_v4_k, rows * C, 256, 0, (cudaStream_t)stream, x, noise, perm, nperm, num_bd, noise_rate, k0, k1, out,
                                                                         sq_partial, C, H, W, taps_dev, num_bd_dev, taps_rows);
    COMBAT_RETURN_LAUNCH("poison_blend_fwd");
  }
  if (use_smem) cudaFuncSetAttribute(poison_blend_fwd_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  pdl_launch(poison_blend_fwd_k, rows * C, 256, use_smem ? smem : 0, (cudaStream_t)stream, x, noise, perm, nperm, num_bd, noise_rate,
                                                                                  k0, k1, out, sq_partial, C, H, W, use_smem,
                                                                                  taps_dev, num_bd_dev, taps_rows);
  COMBAT_RETURN_LAUNCH("poison_blend_fwd");
}

extern "C" int combat_poison_blend_bwd(const float* x, const float* noise, const float* x_bd, const float* g1,
                                       const float* g2, float mse_scale, float noise_rate, float k0, float k1,
                                       float* dnoise, int rows, int C, int H, int W, const float* taps_dev,
                                       const float* taps_rows, void* stream) {
  COMBAT_ARG(x && noise && x_bd && g1 && dnoise, 0);
  COMBAT_ARG(H >= 3 && W >= 3, 12);
  if (rows <= 0) return 0;
  size_t smem = (size_t)H * W * sizeof(float);
  COMBAT_ARG(smem <= 200 * 1024, 13);
  if ((W & 3) == 0) {
    cudaFuncSetAttribute(poison_blend_bwd_v4_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    pdl_launch(poison_blend_bwd_v4_k, rows * C, 256, smem, (cudaStream_t)stream, x, noise, x_bd, g1, g2, mse_scale, noise_rate, k0, k1,
                                                                         dnoise, H, W, taps_dev, C, taps_rows);
    COMBAT_RETURN_LAUNCH("poison_blend_bwd");
  }
  cudaFuncSetAttribute(poison_blend_bwd_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  pdl_launch(poison_blend_bwd_k, rows * C, 256, smem, (cudaStream_t)stream, x, noise, x_bd, g1, g2, mse_scale, noise_rate, k0, k1,
                                                                    dnoise, H * W, H, W, taps_dev, C, taps_rows);
  COMBAT_RETURN_LAUNCH("poison_blend_bwd");
}

// ---------------------------------------------------------------------------------------------- "Grad L2 Loss" (logged only)
// train_generator.py:235-243: with e = inputs - inputs_bd zero-padded by F.pad(., (1, 1, 2, 1)),
//   loss_grad_l2 = mean((e_ext[:, :, 1:] - e_ext[:, :, :-1])^2) + mean((e_ext[..., 1:] - e_ext[..., :-1])^2)
// (MSE is taken between the difference images of inputs and inputs_bd; differences are linear, so it is the squared
// difference image of e).  One CTA per plane writes the two partial sums; combat_grad_l2 then reduces them with the two
// element counts B*C*(H+2)*(W+2) and B*C*(H+3)*(W+1).  The zero rows / columns of the padding only contribute through the
// border terms e[0,c]^2, e[H-1,c]^2, e[r,0]^2, e[r,W-1]^2.
__global__ void __launch_bounds__(256) grad_l2_partial_k(const float* __restrict__ x, const float* __restrict__ x_bd,
                                                         float* __restrict__ partial, int H, int W) {
  pdl_entry();
  const long long base = (long long)blockIdx.x * H * W;
  float sv = 0.f, sh = 0.f;
  for (int i = threadIdx.x; i < H * W; i += blockDim.x) {
    const int r = i / W, c = i - r * W;
    const float e = x[base + i] - x_bd[base + i];
    if (r == 0) sv = fmaf(e, e, sv);
    if (r == H - 1) sv = fmaf(e, e, sv);
    else { const float d = (x[base + i + W] - x_bd[base + i + W]) - e; sv = fmaf(d, d, sv); }
    if (c == 0) sh = fmaf(e, e, sh);
    if (c == W - 1) sh = fmaf(e, e, sh);
    else { const float d = (x[base + i + 1] - x_bd[base + i + 1]) - e; sh = fmaf(d, d, sh); }
  }
  __shared__ float red[2][8];
  sv = warp_sum(sv);
  sh = warp_sum(sh);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = sv; red[1][threadIdx.x >> 5] = sh; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < 8; ++w) { a += red[0][w]; b += red[1][w]; }
    partial[2 * blockIdx.x] = a;
    partial[2 * blockIdx.x + 1] = b;
  }
}

__global__ void __launch_bounds__(256) grad_l2_final_k(const float* __restrict__ partial, int planes, double inv_v, double inv_h,
                                                       float* __restrict__ out) {
  pdl_entry();
  double av = 0.0, ah = 0.0;
  for (int i = threadIdx.x; i < planes; i += blockDim.x) { av += (double)partial[2 * i]; ah += (double)partial[2 * i + 1]; }
  __shared__ double red[2][256];
  red[0][threadIdx.x] = av;
  red[1][threadIdx.x] = ah;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) { red[0][threadIdx.x] += red[0][threadIdx.x + s]; red[1][threadIdx.x] += red[1][threadIdx.x + s]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (float)(red[0][0] * inv_v + red[1][0] * inv_h);
}

extern "C" int combat_grad_l2(const float* x, const float* x_bd, float* partial, float* out, int rows, int C, int H, int W,
                              void* stream) {
  COMBAT_ARG(x && x_bd && partial && out, 0);
  COMBAT_ARG(rows > 0 && C > 0 && H >= 2 && W >= 2, 4);
  const int planes = rows * C;
  pdl_launch(grad_l2_partial_k, planes, 256, 0, (cudaStream_t)stream, x, x_bd, partial, H, W);
  COMBAT_CHECK_LAUNCH("grad_l2_partial");
  const double inv_v = 1.0 / ((double)planes * (H + 2) * (W + 2)), inv_h = 1.0 / ((double)planes * (H + 3) * (W + 1));
  pdl_launch(grad_l2_final_k, 1, 256, 0, (cudaStream_t)stream, partial, planes, inv_v, inv_h, out);
  COMBAT_RETURN_LAUNCH("grad_l2");
}

// ---------------------------------------------------------------------------------------------- total-variation loss (imperceptible variant)
// train_generator_imperceptible.py:228: loss_tv = kornia.losses.total_variation(inputs_bd).mean(), kornia 0.6.6 (requirements.txt:12):
//   total_variation(img) = sum_{c,h,w} |img[c, h+1, w] - img[c, h, w]| + sum_{c,h,w} |img[c, h, w+1] - img[c, h, w]|   per image.
// One CTA per (image, channel) plane: partial sum of the absolute differences, and -- when `grad` is given -- the gradient
//   d loss / d x = weight * ( sign(x - up) - sign(down - x) + sign(x - left) - sign(right - x) )      (sign(0) = 0, torch's abs backward)
// ADDED to grad (the gradient already arriving at inputs_bd from the classifier legs), weight = tv_weight / batch.
__device__ __forceinline__ float sgnf(float v) { return (float)((v > 0.f) - (v < 0.f)); }

__global__ void __launch_bounds__(256) tv_loss_k(const float* __restrict__ x, float* __restrict__ grad, float weight,
                                                 float* __restrict__ partial, int H, int W) {
  pdl_entry();
  const long long base = (long long)blockIdx.x * H * W;
  float s = 0.f;
  for (int i = threadIdx.x; i < H * W; i += blockDim.x) {
    const int r = i / W, c = i - r * W;
    const float v = x[base + i];
    float g = 0.f;
    if (r + 1 < H) { const float d = x[base + i + W] - v; s += fabsf(d); g -= sgnf(d); }
    if (c + 1 < W) { const float d = x[base + i + 1] - v; s += fabsf(d); g -= sgnf(d); }
    if (r > 0) g += sgnf(v - x[base + i - W]);
    if (c > 0) g += sgnf(v - x[base + i - 1]);
    if (grad) grad[base + i] = fmaf(weight, g, grad[base + i]);
  }
  __shared__ float red[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f;
    for (int w = 0; w < 8; ++w) a += red[w];
    partial[blockIdx.x] = a;
  }
}

__global__ void __launch_bounds__(256) tv_final_k(const float* __restrict__ partial, int planes, double scale, float* __restrict__ out) {
  pdl_entry();
  double a = 0.0;
  for (int i = threadIdx.x; i < planes; i += blockDim.x) a += (double)partial[i];
  __shared__ double red[256];
  red[threadIdx.x] = a;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (float)(red[0] * scale);
}

extern "C" int combat_tv_loss(const float* x, float* grad, float grad_weight, float* partial, float* out, int rows, int C, int H, int W,
                              void* stream) {
  COMBAT_ARG(x && partial && out, 0);
  COMBAT_ARG(rows > 0 && C > 0 && H >= 1 && W >= 1, 5);
  const int planes = rows * C;
  pdl_launch(tv_loss_k, planes, 256, 0, (cudaStream_t)stream, x, grad, grad_weight, partial, H, W);
  COMBAT_CHECK_LAUNCH("tv_loss");
  pdl_launch(tv_final_k, 1, 256, 0, (cudaStream_t)stream, partial, planes, 1.0 / (double)rows, out);
  COMBAT_RETURN_LAUNCH("tv_loss");
}
