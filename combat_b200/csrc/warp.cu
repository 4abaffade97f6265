// WaNet-style warp trigger of train_generator_wanet.py (:151-158 / :196-203 of the reference) as ONE kernel and its backward:
//   flow       = netG(inputs)                 the GridGenerator's tanh output, [N, 2, S, S]                (networks/models.py:383-385)
//   noise_grid = bicubic_upsample(flow, H x H, align_corners=True).permute(0, 2, 3, 1)                      (:152-154)
//   grid       = clamp(identity_grid * (1 - grid_rescale) + noise_grid * grid_rescale, -1, 1)               (:155-156)
//   inputs_bd  = grid_sample(inputs, grid, bilinear, zeros padding, align_corners=True)                     (:157)
// plus the batch assembly of the C-step (:159: row i of the output is the warped image of sample perm[i] for i < num_bd and a
// plain copy of sample perm[i] otherwise), the sum of squares of noise_grid (loss_l2 = MSE(noise_grid, 0), :212) and the
// logged-only finite-difference term (:213-222).
//
// Bicubic upsampling is linear and separable: noise(c, h, w) = sum_{py, px} Wt[py][h] * Wt[px][w] * flow(c, py, px), with
// Wt[p][o] = sum of the four cubic-convolution coefficients (A = -0.75) of output index o whose clamped source index is p.
// S is tiny (--s 2: a 2 x 2 control grid), so the table lives in shared memory and a pixel costs S*S multiply-adds per
// channel; the backward is the same table transposed (per-thread register accumulators, one block reduction, no atomics).
// Images are NCHW float32 (the reference's tensors), square (identity_grid is built from input_height alone, :560-562).
// HBM-bound: read + write of one image per row; the gather goes through L1 (a CIFAR image is 12 KB).
#include "common.cuh"

namespace {

constexpr int WARP_MAX_S = 4;     // 2 * S * S <= 32 accumulators per thread; the reference default is S = 2
constexpr int WARP_MAX_HW = 256;  // table rows in shared memory

struct WarpShared {
  float flow2[2 * 32];                  // fast kernels: the flow of the current image (+ the next one when pipelined)
  float flow[2 * WARP_MAX_S * WARP_MAX_S];
  float wt[WARP_MAX_S * WARP_MAX_HW];   // wt[p * H + o]
  float red[32 * 8];
};

__device__ __forceinline__ void cubic_coeffs(float t, float c[4]) {
  const float A = -0.75f;
  const float x0 = t + 1.f, x3 = 2.f - t, u = 1.f - t;
  c[0] = ((A * x0 - 5.f * A) * x0 + 8.f * A) * x0 - 4.f * A;
  c[1] = ((A + 2.f) * t - (A + 3.f)) * t * t + 1.f;
  c[2] = ((A + 2.f) * u - (A + 3.f)) * u * u + 1.f;
  c[3] = ((A * x3 - 5.f * A) * x3 + 8.f * A) * x3 - 4.f * A;
}

// the sample's flow and the separable bicubic table of an S -> H upsampling with align_corners=True
__device__ __forceinline__ void warp_setup(WarpShared& sh, const float* __restrict__ z, int S, int H) {
  const int nf = 2 * S * S;
  if ((int)threadIdx.x < nf) sh.flow[threadIdx.x] = z[threadIdx.x];
  const float scale = H > 1 ? (float)(S - 1) / (float)(H - 1) : 0.f;
  for (int o = threadIdx.x; o < H; o += blockDim.x) {
    const float real = scale * (float)o;
    const float fl = floorf(real);
    const int in = (int)fl;
    float c[4];
    cubic_coeffs(real - fl, c);
    float acc[WARP_MAX_S];
#pragma unroll
    for (int p = 0; p < WARP_MAX_S; ++p) acc[p] = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int p = min(max(in - 1 + i, 0), S - 1);
#pragma unroll
      for (int q = 0; q < WARP_MAX_S; ++q)
        if (q == p) acc[q] += c[i];
    }
#pragma unroll
    for (int p = 0; p < WARP_MAX_S; ++p)
      if (p < S) sh.wt[p * H + o] = acc[p];
  }
  __syncthreads();
}

__device__ __forceinline__ void noise_at(const WarpShared& sh, int S, int H, int h, int w, float& nx, float& ny) {
  nx = 0.f;
  ny = 0.f;
  const int SS = S * S;
  for (int py = 0; py < S; ++py) {
    const float wy = sh.wt[py * H + h];
    for (int px = 0; px < S; ++px) {
      const float k = wy * sh.wt[px * H + w];
      nx = fmaf(k, sh.flow[py * S + px], nx);
      ny = fmaf(k, sh.flow[SS + py * S + px], ny);
    }
  }
}

struct Bilin {
  int x0, y0;
  float wx, wy;   // weight of the x0+1 / y0+1 taps
};

__device__ __forceinline__ Bilin unnormalise(float gx, float gy, int H, int W) {
  const float ix = (gx + 1.f) * 0.5f * (float)(W - 1);
  const float iy = (gy + 1.f) * 0.5f * (float)(H - 1);
  const float fx = floorf(ix), fy = floorf(iy);
  Bilin b;
  b.x0 = (int)fx;
  b.y0 = (int)fy;
  b.wx = ix - fx;
  b.wy = iy - fy;
  return b;
}

__device__ __forceinline__ float tap(const float* __restrict__ pl, int x, int y, int H, int W) {
  return (x >= 0 && x < W && y >= 0 && y < H) ? __ldg(pl + y * W + x) : 0.f;
}

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  if (warp == 0) {
    t = lane < (int)(blockDim.x >> 5) ? red[lane] : 0.f;
    t = warp_sum(t);
  }
  return t;   // valid in warp 0
}

__global__ void __launch_bounds__(256) wanet_warp_fwd_k(const float* __restrict__ x, const float* __restrict__ z,
                                                        const float* __restrict__ ident, const int* __restrict__ perm,
                                                        int num_bd, const int* __restrict__ num_bd_dev, float rescale,
                                                        float* __restrict__ out, float* __restrict__ noise_grid,
                                                        float* __restrict__ sq_partial, float* __restrict__ gl_partial, int C,
                                                        int H, int S) {
  pdl_entry();
  __shared__ WarpShared sh;
  const int row = blockIdx.x;
  const int W = H, HW = H * H;
  const int src = perm ? perm[row] : row;
  const int nbd = num_bd_dev ? *num_bd_dev : num_bd;
  const float* xs = x + (long long)src * C * HW;
  float* dst = out + (long long)row * C * HW;
  if (row >= nbd) {  // pass-through row of the C-step batch (bit-exact copy)
    for (int i = threadIdx.x; i < C * HW; i += blockDim.x) dst[i] = __ldg(xs + i);
    return;
  }
  warp_setup(sh, z + (long long)src * 2 * S * S, S, H);
  float sq = 0.f, gl1 = 0.f, gl2 = 0.f;
  for (int p = threadIdx.x; p < HW; p += blockDim.x) {
    const int h = p / W, w = p - h * W;
    float nx, ny;
    noise_at(sh, S, H, h, w, nx, ny);
    const float gx = fminf(fmaxf(__ldg(ident + w) * (1.f - rescale) + nx * rescale, -1.f), 1.f);
    const float gy = fminf(fmaxf(__ldg(ident + h) * (1.f - rescale) + ny * rescale, -1.f), 1.f);
    const Bilin b = unnormalise(gx, gy, H, W);
    const float w00 = (1.f - b.wx) * (1.f - b.wy), w01 = b.wx * (1.f - b.wy), w10 = (1.f - b.wx) * b.wy, w11 = b.wx * b.wy;
    for (int c = 0; c < C; ++c) {
      const float* pl = xs + c * HW;
      dst[c * HW + p] = tap(pl, b.x0, b.y0, H, W) * w00 + tap(pl, b.x0 + 1, b.y0, H, W) * w01 +
                        tap(pl, b.x0, b.y0 + 1, H, W) * w10 + tap(pl, b.x0 + 1, b.y0 + 1, H, W) * w11;
    }
    if (noise_grid) *(float2*)(noise_grid + ((long long)row * HW + p) * 2) = make_float2(nx, ny);
    sq += nx * nx + ny * ny;
    if (gl_partial) {
      // :213-222 on F.pad(noise_grid, (1, 1, 2, 1)): differences along the padded W axis ([0, 0, v_0 .. v_{W-1}, 0]) and along
      // the padded channel axis ([0, nx, ny, 0]); the zero columns contribute nothing
      float px_ = 0.f, py_ = 0.f;
      if (w > 0) noise_at(sh, S, H, h, w - 1, px_, py_);
      gl1 += (nx - px_) * (nx - px_) + (ny - py_) * (ny - py_);
      if (w == W - 1) gl1 += nx * nx + ny * ny;
      gl2 += nx * nx + (ny - nx) * (ny - nx) + ny * ny;
    }
  }
  if (sq_partial) {
    const float t = block_sum(sq, sh.red);
    if (threadIdx.x == 0) sq_partial[row] = t;
  }
  if (gl_partial) {
    const float t1 = block_sum(gl1, sh.red);
    const float t2 = block_sum(gl2, sh.red);
    if (threadIdx.x == 0) gl_partial[row] = t1 / (float)(H * (W + 2) * 4) + t2 / (float)(H * (W + 3) * 3);
  }
}

// dz[n] = d loss / d flow[n] for  loss = <g1 + g2, inputs_bd> + (l2_scale / 2) * |noise_grid|^2
__global__ void __launch_bounds__(256) wanet_warp_bwd_k(const float* __restrict__ x, const float* __restrict__ z,
                                                        const float* __restrict__ ident, const float* __restrict__ g1,
                                                        const float* __restrict__ g2, float rescale, float l2_scale,
                                                        float* __restrict__ dz, int C, int H, int S) {
  pdl_entry();
  __shared__ WarpShared sh;
  const int n = blockIdx.x;
  const int W = H, HW = H * H, SS = S * S;
  const float* xs = x + (long long)n * C * HW;
  warp_setup(sh, z + (long long)n * 2 * SS, S, H);
  float acc[2 * WARP_MAX_S * WARP_MAX_S];
#pragma unroll
  for (int k = 0; k < 2 * WARP_MAX_S * WARP_MAX_S; ++k) acc[k] = 0.f;
  for (int p = threadIdx.x; p < HW; p += blockDim.x) {
    const int h = p / W, w = p - h * W;
    float nx, ny;
    noise_at(sh, S, H, h, w, nx, ny);
    const float rx = __ldg(ident + w) * (1.f - rescale) + nx * rescale;
    const float ry = __ldg(ident + h) * (1.f - rescale) + ny * rescale;
    const Bilin b = unnormalise(fminf(fmaxf(rx, -1.f), 1.f), fminf(fmaxf(ry, -1.f), 1.f), H, W);
    float dix = 0.f, diy = 0.f;
    for (int c = 0; c < C; ++c) {
      const long long gi = ((long long)n * C + c) * HW + p;
      float g = __ldg(g1 + gi);
      if (g2) g += __ldg(g2 + gi);
      const float* pl = xs + c * HW;
      const float v00 = tap(pl, b.x0, b.y0, H, W), v01 = tap(pl, b.x0 + 1, b.y0, H, W);
      const float v10 = tap(pl, b.x0, b.y0 + 1, H, W), v11 = tap(pl, b.x0 + 1, b.y0 + 1, H, W);
      dix = fmaf(g, (v01 - v00) * (1.f - b.wy) + (v11 - v10) * b.wy, dix);
      diy = fmaf(g, (v10 - v00) * (1.f - b.wx) + (v11 - v01) * b.wx, diy);
    }
    // clamp passes the gradient where -1 <= raw <= 1 (torch.clamp backward), then the blend with the identity grid
    const float dnx = ((rx >= -1.f && rx <= 1.f) ? dix * 0.5f * (float)(W - 1) * rescale : 0.f) + l2_scale * nx;
    const float dny = ((ry >= -1.f && ry <= 1.f) ? diy * 0.5f * (float)(H - 1) * rescale : 0.f) + l2_scale * ny;
#pragma unroll
    for (int py = 0; py < WARP_MAX_S; ++py) {
      if (py < S) {
        const float wy = sh.wt[py * H + h];
#pragma unroll
        for (int px = 0; px < WARP_MAX_S; ++px) {
          if (px < S) {
            const float k = wy * sh.wt[px * H + w];
            acc[py * WARP_MAX_S + px] = fmaf(k, dnx, acc[py * WARP_MAX_S + px]);
            acc[WARP_MAX_S * WARP_MAX_S + py * WARP_MAX_S + px] = fmaf(k, dny, acc[WARP_MAX_S * WARP_MAX_S + py * WARP_MAX_S + px]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 2; ++c)
#pragma unroll
    for (int py = 0; py < WARP_MAX_S; ++py)
#pragma unroll
      for (int px = 0; px < WARP_MAX_S; ++px) {
        if (py < S && px < S) {   // uniform across the block
          const float t = block_sum(acc[c * WARP_MAX_S * WARP_MAX_S + py * WARP_MAX_S + px], sh.red);
          if (threadIdx.x == 0) dz[(long long)n * 2 * SS + c * SS + py * S + px] = t;
        }
      }
}

// ------------------------------------------------------------------------------------------------------------------------------
// Fast path (W % 4 == 0, the shapes of the reference: 32 / 64 / 224): PERSISTENT CTAs (a few per SM) that build the bicubic table
// once and loop over images; a thread owns four consecutive pixels of a row -- 16-byte loads of the gradients, 16-byte stores of
// the warped image and the noise grid; S is a template parameter (compile-time S*S loops); one warp-shuffle + one shared-memory
// pass reduces all per-image sums at once.  First version (one CTA per image, scalar pixels, a block reduction per sum):
// 1057 us forward / 2180 us backward for 65,536 CIFAR images (23 % / 11 % of the HBM bound).
template <int S>
__device__ __forceinline__ void build_table(WarpShared& sh, int H) {
  const float scale = H > 1 ? (float)(S - 1) / (float)(H - 1) : 0.f;
  for (int o = threadIdx.x; o < H; o += blockDim.x) {
    const float real = scale * (float)o;
    const float fl = floorf(real);
    const int in = (int)fl;
    float c[4];
    cubic_coeffs(real - fl, c);
    float acc[S];
#pragma unroll
    for (int p = 0; p < S; ++p) acc[p] = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int p = min(max(in - 1 + i, 0), S - 1);
#pragma unroll
      for (int q = 0; q < S; ++q)
        if (q == p) acc[q] += c[i];
    }
#pragma unroll
    for (int p = 0; p < S; ++p) sh.wt[p * H + o] = acc[p];
  }
}

template <int S>
__device__ __forceinline__ void noise_one(const WarpShared& sh, const float* flow, int H, int h, int w, float& nx, float& ny) {
  nx = ny = 0.f;
#pragma unroll
  for (int py = 0; py < S; ++py) {
    const float wy = sh.wt[py * H + h];
#pragma unroll
    for (int px = 0; px < S; ++px) {
      const float k = wy * sh.wt[px * H + w];
      nx = fmaf(k, flow[py * S + px], nx);
      ny = fmaf(k, flow[S * S + py * S + px], ny);
    }
  }
}

template <int S>
__device__ __forceinline__ void noise4(const WarpShared& sh, const float* flow, int H, int h, int w0, float nx[4], float ny[4]) {
  float wy[S];
#pragma unroll
  for (int py = 0; py < S; ++py) wy[py] = sh.wt[py * H + h];
#pragma unroll
  for (int j = 0; j < 4; ++j) nx[j] = ny[j] = 0.f;
#pragma unroll
  for (int px = 0; px < S; ++px) {
    const float4 wx = *(const float4*)(sh.wt + px * H + w0);
    const float wxv[4] = {wx.x, wx.y, wx.z, wx.w};
#pragma unroll
    for (int py = 0; py < S; ++py) {
      const float fx = flow[py * S + px], fy = flow[S * S + py * S + px];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float k = wy[py] * wxv[j];
        nx[j] = fmaf(k, fx, nx[j]);
        ny[j] = fmaf(k, fy, ny[j]);
      }
    }
  }
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// The clamped grid keeps ix in [0, W-1] and iy in [0, H-1]: the top-left tap is always inside the image and a right / bottom tap
// is outside only when ix == W-1 / iy == H-1 exactly, where its bilinear weight is exactly 0.  So the four taps are read from
// offsets folded back into the image (off, off + dx, off + dy, off + dx + dy with dx in {0, 1}, dy in {0, W}) without a predicate
// per load; the backward, which needs the VALUES of the outside taps to be zero (zeros padding), selects on two flags per pixel.
struct Tap {
  int off, dx, dy;
  float wx, wy;
};
__device__ __forceinline__ Tap make_tap(float gx, float gy, int H, int W) {
  const Bilin b = unnormalise(gx, gy, H, W);
  Tap t;
  t.off = b.y0 * W + b.x0;
  t.dx = b.x0 + 1 < W ? 1 : 0;
  t.dy = b.y0 + 1 < H ? W : 0;
  t.wx = b.wx;
  t.wy = b.wy;
  return t;
}
template <bool STAGED>
__device__ __forceinline__ float ld(const float* __restrict__ p) {
  return STAGED ? *p : __ldg(p);
}

// STAGED: the 3 x H x W source image is copied into shared memory with coalesced 16-byte loads, all in flight at once (one HBM
// latency per image instead of one per channel and tap round), and the gather reads it from there.
// PIPE (every row is warped, no permutation: the G-step / evaluation batch): the staged image and the flow of image n + 1 are
// fetched with cp.async into the other half of a double buffer while image n is computed -- three resident CTAs alone keep only
// ~36 KB per SM in flight, less than the bandwidth-latency product needs, and each CTA sat in a load -> sync -> compute chain.
template <int S, bool STAGED, bool PIPE>
__global__ void __launch_bounds__(256, 3) wanet_warp_fwd4_k(const float* __restrict__ x, const float* __restrict__ z,
                                                            const float* __restrict__ ident, const int* __restrict__ perm,
                                                            int num_bd, const int* __restrict__ num_bd_dev, float rescale,
                                                            float* __restrict__ out, float* __restrict__ noise_grid,
                                                            float* __restrict__ sq_partial, float* __restrict__ gl_partial,
                                                            int H, int rows) {
  pdl_entry();
  __shared__ __align__(16) WarpShared sh;
  extern __shared__ float4 simg4[];
  float* simg = (float*)simg4;
  constexpr int C = 3;
  const int W = H, HW = H * H, W4 = W >> 2, G = HW >> 2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nbd = num_bd_dev ? *num_bd_dev : num_bd;
  build_table<S>(sh, H);
  const int Q = (C * HW) >> 2;   // float4s per image
  auto fetch = [&](int r, int half) {   // PIPE: image r and its flow -> buffer half `half` (asynchronous)
    const float4* s4 = (const float4*)(x + (long long)r * C * HW);
    for (int i = tid; i < Q; i += blockDim.x) cp_async16(simg4 + half * Q + i, s4 + i);
    if (tid < 2 * S * S) cp_async4(sh.flow2 + half * 32 + tid, z + (long long)r * 2 * S * S + tid);
  };
  if (PIPE) {
    if ((int)blockIdx.x < rows) fetch(blockIdx.x, 0);
    cp_async_commit();
  }
  int it = 0;
  for (int row = blockIdx.x; row < rows; row += gridDim.x, ++it) {
    const int src = (!PIPE && perm) ? perm[row] : row;
    const float* xs = x + (long long)src * C * HW;
    float* dst = out + (long long)row * C * HW;
    if (!PIPE && row >= nbd) {  // pass-through row of the C-step batch (bit-exact copy)
      const float4* s4 = (const float4*)xs;
      float4* d4 = (float4*)dst;
      for (int i = tid; i < Q; i += blockDim.x) d4[i] = __ldg(s4 + i);
      continue;
    }
    __syncthreads();   // the table is complete / the previous image is done with the flow, sh.red and the staged image
    const float* img;
    const float* flow;
    if (PIPE) {
      const int half = it & 1;
      if (row + (int)gridDim.x < rows) fetch(row + gridDim.x, half ^ 1);
      cp_async_commit();
      cp_async_wait1();   // everything but the group just committed has landed: this image is in `half`
      img = simg + half * (C * HW);
      flow = sh.flow2 + half * 32;
    } else {
      if (STAGED) {
        const float4* s4 = (const float4*)xs;
        for (int i = tid; i < Q; i += blockDim.x) simg4[i] = __ldg(s4 + i);
      }
      if (tid < 2 * S * S) sh.flow2[tid] = z[(long long)src * 2 * S * S + tid];
      img = STAGED ? simg : xs;
      flow = sh.flow2;
    }
    __syncthreads();
    float sq = 0.f, gl1 = 0.f, gl2 = 0.f;
    for (int g = tid; g < G; g += blockDim.x) {
      const int h = g / W4, w0 = (g - h * W4) << 2;
      float nx[4], ny[4];
      noise4<S>(sh, flow, H, h, w0, nx, ny);
      const float4 idw = __ldg((const float4*)(ident + w0));
      const float idx_[4] = {idw.x, idw.y, idw.z, idw.w};
      const float idh = __ldg(ident + h);
      Tap b[4];
      float w00[4], w01[4], w10[4], w11[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float gx = fminf(fmaxf(idx_[j] * (1.f - rescale) + nx[j] * rescale, -1.f), 1.f);
        const float gy = fminf(fmaxf(idh * (1.f - rescale) + ny[j] * rescale, -1.f), 1.f);
        b[j] = make_tap(gx, gy, H, W);
        w00[j] = (1.f - b[j].wx) * (1.f - b[j].wy);
        w01[j] = b[j].wx * (1.f - b[j].wy);
        w10[j] = (1.f - b[j].wx) * b[j].wy;
        w11[j] = b[j].wx * b[j].wy;
      }
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float* pl = img + c * HW;
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float* q = pl + b[j].off;
          o[j] = ld<STAGED>(q) * w00[j] + ld<STAGED>(q + b[j].dx) * w01[j] + ld<STAGED>(q + b[j].dy) * w10[j] +
                 ld<STAGED>(q + b[j].dx + b[j].dy) * w11[j];
        }
        *(float4*)(dst + c * HW + h * W + w0) = make_float4(o[0], o[1], o[2], o[3]);
      }
      if (noise_grid) {
        float4* ng = (float4*)(noise_grid + ((long long)row * HW + h * W + w0) * 2);
        ng[0] = make_float4(nx[0], ny[0], nx[1], ny[1]);
        ng[1] = make_float4(nx[2], ny[2], nx[3], ny[3]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) sq += nx[j] * nx[j] + ny[j] * ny[j];
      if (gl_partial) {   // :213-222, see the scalar kernel
        float px_ = 0.f, py_ = 0.f;
        if (w0 > 0) noise_one<S>(sh, flow, H, h, w0 - 1, px_, py_);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          gl1 += (nx[j] - px_) * (nx[j] - px_) + (ny[j] - py_) * (ny[j] - py_);
          gl2 += nx[j] * nx[j] + (ny[j] - nx[j]) * (ny[j] - nx[j]) + ny[j] * ny[j];
          px_ = nx[j];
          py_ = ny[j];
        }
        if (w0 + 4 == W) gl1 += px_ * px_ + py_ * py_;
      }
    }
    if (sq_partial || gl_partial) {
      sq = warp_sum(sq);
      gl1 = warp_sum(gl1);
      gl2 = warp_sum(gl2);
      if (lane == 0) {
        sh.red[warp * 3 + 0] = sq;
        sh.red[warp * 3 + 1] = gl1;
        sh.red[warp * 3 + 2] = gl2;
      }
      __syncthreads();
      if (tid == 0) {
        float a = 0.f, b1 = 0.f, b2 = 0.f;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
          a += sh.red[i * 3 + 0];
          b1 += sh.red[i * 3 + 1];
          b2 += sh.red[i * 3 + 2];
        }
        if (sq_partial) sq_partial[row] = a;
        if (gl_partial) gl_partial[row] = b1 / (float)(H * (W + 2) * 4) + b2 / (float)(H * (W + 3) * 3);
      }
    }
  }
}

// PIPE here double-buffers BOTH the source image and the incoming gradient g1 (g2 must be NULL): [x | g1] of image n + 1 by cp.async.
template <int S, bool STAGED, bool PIPE>
__global__ void __launch_bounds__(256, 3) wanet_warp_bwd4_k(const float* __restrict__ x, const float* __restrict__ z,
                                                            const float* __restrict__ ident, const float* __restrict__ g1,
                                                            const float* __restrict__ g2, float rescale, float l2_scale,
                                                            float* __restrict__ dz, int H, int rows) {
  pdl_entry();
  __shared__ __align__(16) WarpShared sh;
  extern __shared__ float4 simg4[];
  float* simg = (float*)simg4;
  constexpr int NF = 2 * S * S, C = 3;
  const int W = H, HW = H * H, W4 = W >> 2, G = HW >> 2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  build_table<S>(sh, H);
  const int Q = (C * HW) >> 2;
  auto fetch = [&](int r, int half) {   // PIPE: [x | g1] of image r and its flow -> buffer half `half`
    const float4* s4 = (const float4*)(x + (long long)r * C * HW);
    const float4* q4 = (const float4*)(g1 + (long long)r * C * HW);
    for (int i = tid; i < Q; i += blockDim.x) {
      cp_async16(simg4 + half * 2 * Q + i, s4 + i);
      cp_async16(simg4 + half * 2 * Q + Q + i, q4 + i);
    }
    if (tid < NF) cp_async4(sh.flow2 + half * 32 + tid, z + (long long)r * NF + tid);
  };
  if (PIPE) {
    if ((int)blockIdx.x < rows) fetch(blockIdx.x, 0);
    cp_async_commit();
  }
  int it = 0;
  for (int n = blockIdx.x; n < rows; n += gridDim.x, ++it) {
    const float* xs = x + (long long)n * C * HW;
    __syncthreads();
    const float* img;
    const float* flow;
    const float* gsm = nullptr;
    if (PIPE) {
      const int half = it & 1;
      if (n + (int)gridDim.x < rows) fetch(n + gridDim.x, half ^ 1);
      cp_async_commit();
      cp_async_wait1();
      img = simg + half * 2 * (C * HW);
      gsm = img + C * HW;
      flow = sh.flow2 + half * 32;
    } else {
      if (STAGED) {
        const float4* s4 = (const float4*)xs;
        for (int i = tid; i < Q; i += blockDim.x) simg4[i] = __ldg(s4 + i);
      }
      if (tid < NF) sh.flow2[tid] = z[(long long)n * NF + tid];
      img = STAGED ? simg : xs;
      flow = sh.flow2;
    }
    __syncthreads();
    float acc[NF];
#pragma unroll
    for (int k = 0; k < NF; ++k) acc[k] = 0.f;
    for (int g = tid; g < G; g += blockDim.x) {
      const int h = g / W4, w0 = (g - h * W4) << 2;
      // the incoming gradients of the four pixels, all channels: issued first, independent of everything below
      float4 gq[C];
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const long long gi = ((long long)n * C + c) * HW + h * W + w0;
        gq[c] = PIPE ? *(const float4*)(gsm + c * HW + h * W + w0) : __ldg((const float4*)(g1 + gi));
        if (!PIPE && g2) {
          const float4 g2v = __ldg((const float4*)(g2 + gi));
          gq[c].x += g2v.x; gq[c].y += g2v.y; gq[c].z += g2v.z; gq[c].w += g2v.w;
        }
      }
      float nx[4], ny[4], rx[4], ry[4], dix[4], diy[4];
      noise4<S>(sh, flow, H, h, w0, nx, ny);
      const float4 idw = __ldg((const float4*)(ident + w0));
      const float idx_[4] = {idw.x, idw.y, idw.z, idw.w};
      const float idh = __ldg(ident + h);
      Tap b[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        rx[j] = idx_[j] * (1.f - rescale) + nx[j] * rescale;
        ry[j] = idh * (1.f - rescale) + ny[j] * rescale;
        b[j] = make_tap(fminf(fmaxf(rx[j], -1.f), 1.f), fminf(fmaxf(ry[j], -1.f), 1.f), H, W);
        dix[j] = diy[j] = 0.f;
      }
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float gg[4] = {gq[c].x, gq[c].y, gq[c].z, gq[c].w};
        const float* pl = img + c * HW;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float* q = pl + b[j].off;
          // outside taps count as zeros (padding_mode="zeros"): UNCONDITIONAL loads from the folded offsets, then a 0 / 1 factor
          // (a predicated load compiles to load-wait-select per value and serialises the batch: 664 -> 737 us when tried)
          const float mx = (float)b[j].dx, my = b[j].dy ? 1.f : 0.f;
          const float v00 = ld<STAGED>(q);
          const float v01 = ld<STAGED>(q + b[j].dx) * mx;
          const float v10 = ld<STAGED>(q + b[j].dy) * my;
          const float v11 = ld<STAGED>(q + b[j].dx + b[j].dy) * (mx * my);
          dix[j] = fmaf(gg[j], (v01 - v00) * (1.f - b[j].wy) + (v11 - v10) * b[j].wy, dix[j]);
          diy[j] = fmaf(gg[j], (v10 - v00) * (1.f - b[j].wx) + (v11 - v01) * b[j].wx, diy[j]);
        }
      }
      float wy[S];
#pragma unroll
      for (int py = 0; py < S; ++py) wy[py] = sh.wt[py * H + h];
#pragma unroll
      for (int px = 0; px < S; ++px) {
        const float4 wx4 = *(const float4*)(sh.wt + px * H + w0);
        const float wxv[4] = {wx4.x, wx4.y, wx4.z, wx4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          // clamp passes the gradient where -1 <= raw <= 1 (torch.clamp backward), then the blend with the identity grid
          const float dnx = ((rx[j] >= -1.f && rx[j] <= 1.f) ? dix[j] * 0.5f * (float)(W - 1) * rescale : 0.f) + l2_scale * nx[j];
          const float dny = ((ry[j] >= -1.f && ry[j] <= 1.f) ? diy[j] * 0.5f * (float)(H - 1) * rescale : 0.f) + l2_scale * ny[j];
#pragma unroll
          for (int py = 0; py < S; ++py) {
            const float k = wy[py] * wxv[j];
            acc[py * S + px] = fmaf(k, dnx, acc[py * S + px]);
            acc[S * S + py * S + px] = fmaf(k, dny, acc[S * S + py * S + px]);
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < NF; ++k) {
      const float t = warp_sum(acc[k]);
      if (lane == 0) sh.red[warp * NF + k] = t;
    }
    __syncthreads();
    if (tid < NF) {
      float t = 0.f;
      for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh.red[i * NF + tid];
      dz[(long long)n * NF + tid] = t;
    }
  }
}

static bool warp_staged(int H, size_t* bytes) {
  *bytes = (size_t)3 * H * H * sizeof(float);
  return *bytes <= 48 * 1024;   // CIFAR 12 KB, CelebA 48 KB per image; 224 x 224 gathers from global memory
}

static int warp_grid(int rows) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int g = sms * 3;   // three resident CTAs per SM (launch bounds), every CTA builds the table once
  return rows < g ? rows : g;
}

__global__ void __launch_bounds__(256) tanh_fwd_k(const float* __restrict__ x, float* __restrict__ y, long long n) {
  pdl_entry();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = tanhf(x[i]);
}

}  // namespace

extern "C" int combat_tanh_fwd(const float* x, float* y, long long n, void* stream) {
  COMBAT_ARG(x && y && n >= 0, 0);
  if (n == 0) return 0;
  pdl_launch(tanh_fwd_k, (int)min((long long)cdiv(n, 256), 1184LL), 256, 0, (cudaStream_t)stream, x, y, n);
  COMBAT_RETURN_LAUNCH("tanh_fwd");
}

extern "C" int combat_wanet_warp_fwd(const float* x, const float* z, const float* ident, const int* perm, int rows, int num_bd,
                                     const int* num_bd_dev, float grid_rescale, float* out, float* noise_grid,
                                     float* sq_partial, float* gl_partial, int C, int H, int W, int S, void* stream) {
  COMBAT_ARG(x && z && ident && out && x != out, 0);
  COMBAT_ARG(rows >= 0 && C > 0 && H > 1 && H == W && H <= WARP_MAX_HW, 4);
  COMBAT_ARG(S >= 1 && S <= WARP_MAX_S, 15);
  COMBAT_ARG(perm || !(noise_grid || sq_partial || gl_partial) || num_bd >= rows || num_bd_dev, 5);
  if (rows == 0) return 0;
  if (C == 3 && W % 4 == 0 && ((uintptr_t)x % 16) == 0 && ((uintptr_t)out % 16) == 0 && ((uintptr_t)ident % 16) == 0 &&
      (!noise_grid || ((uintptr_t)noise_grid % 16) == 0)) {
    size_t bytes;
    const bool staged = warp_staged(H, &bytes);
    // pipelined double buffer: every row warped, rows in place (the G-step / evaluation batch), two images fit next to the table
    const bool pipe = staged && !perm && !num_bd_dev && num_bd >= rows && 2 * bytes <= 96 * 1024;
#define WARP_FWD(S_)                                                                                                             \
  do {                                                                                                                           \
    if (pipe) {                                                                                                                  \
      cudaFuncSetAttribute(wanet_warp_fwd4_k<S_, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * bytes));    \
      pdl_launch(wanet_warp_fwd4_k<S_, true, true>, warp_grid(rows), 256, 2 * bytes, (cudaStream_t)stream, x, z, ident, perm,     \
                 num_bd, num_bd_dev, grid_rescale, out, noise_grid, sq_partial, gl_partial, H, rows);                            \
    } else if (staged) {                                                                                                         \
      cudaFuncSetAttribute(wanet_warp_fwd4_k<S_, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);         \
      pdl_launch(wanet_warp_fwd4_k<S_, true, false>, warp_grid(rows), 256, bytes, (cudaStream_t)stream, x, z, ident, perm,        \
                 num_bd, num_bd_dev, grid_rescale, out, noise_grid, sq_partial, gl_partial, H, rows);                            \
    } else {                                                                                                                     \
      pdl_launch(wanet_warp_fwd4_k<S_, false, false>, warp_grid(rows), 256, 0, (cudaStream_t)stream, x, z, ident, perm, num_bd,   \
                 num_bd_dev, grid_rescale, out, noise_grid, sq_partial, gl_partial, H, rows);                                    \
    }                                                                                                                            \
  } while (0)
    if (S == 1) WARP_FWD(1); else if (S == 2) WARP_FWD(2); else if (S == 3) WARP_FWD(3); else WARP_FWD(4);
#undef WARP_FWD
    COMBAT_RETURN_LAUNCH("wanet_warp_fwd");
  }
  pdl_launch(wanet_warp_fwd_k, rows, 256, 0, (cudaStream_t)stream, x, z, ident, perm, num_bd, num_bd_dev, grid_rescale, out,
             noise_grid, sq_partial, gl_partial, C, H, S);
  COMBAT_RETURN_LAUNCH("wanet_warp_fwd");
}

extern "C" int combat_wanet_warp_bwd(const float* x, const float* z, const float* ident, const float* g1, const float* g2,
                                     float grid_rescale, float l2_scale, float* dz, int rows, int C, int H, int W, int S,
                                     void* stream) {
  COMBAT_ARG(x && z && ident && g1 && dz, 0);
  COMBAT_ARG(rows >= 0 && C > 0 && H > 1 && H == W && H <= WARP_MAX_HW, 8);
  COMBAT_ARG(S >= 1 && S <= WARP_MAX_S, 12);
  if (rows == 0) return 0;
  if (C == 3 && W % 4 == 0 && ((uintptr_t)x % 16) == 0 && ((uintptr_t)g1 % 16) == 0 && ((uintptr_t)ident % 16) == 0 &&
      (!g2 || ((uintptr_t)g2 % 16) == 0)) {
    size_t bytes;
    const bool staged = warp_staged(H, &bytes);
    const bool pipe = staged && !g2 && 4 * bytes <= 96 * 1024;   // [x | g1] x 2 buffers (CIFAR: 48 KB per CTA)
#define WARP_BWD(S_)                                                                                                             \
  do {                                                                                                                           \
    if (pipe) {                                                                                                                  \
      cudaFuncSetAttribute(wanet_warp_bwd4_k<S_, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * bytes));    \
      pdl_launch(wanet_warp_bwd4_k<S_, true, true>, warp_grid(rows), 256, 4 * bytes, (cudaStream_t)stream, x, z, ident, g1, g2,   \
                 grid_rescale, l2_scale, dz, H, rows);                                                                           \
    } else if (staged) {                                                                                                         \
      cudaFuncSetAttribute(wanet_warp_bwd4_k<S_, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);         \
      pdl_launch(wanet_warp_bwd4_k<S_, true, false>, warp_grid(rows), 256, bytes, (cudaStream_t)stream, x, z, ident, g1, g2,      \
                 grid_rescale, l2_scale, dz, H, rows);                                                                           \
    } else {                                                                                                                     \
      pdl_launch(wanet_warp_bwd4_k<S_, false, false>, warp_grid(rows), 256, 0, (cudaStream_t)stream, x, z, ident, g1, g2,         \
                 grid_rescale, l2_scale, dz, H, rows);                                                                           \
    }                                                                                                                            \
  } while (0)
    if (S == 1) WARP_BWD(1); else if (S == 2) WARP_BWD(2); else if (S == 3) WARP_BWD(3); else WARP_BWD(4);
#undef WARP_BWD
    COMBAT_RETURN_LAUNCH("wanet_warp_bwd");
  }
  pdl_launch(wanet_warp_bwd_k, rows, 256, 0, (cudaStream_t)stream, x, z, ident, g1, g2, grid_rescale, l2_scale, dz, C, H, S);
  COMBAT_RETURN_LAUNCH("wanet_warp_bwd");
}
