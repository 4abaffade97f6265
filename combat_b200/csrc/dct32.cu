// HBM-bound 32x32 DCT-II / DCT-III / low-pass projection: one warp per plane, one image row (then column) per
// thread held in 32 registers, 1-D transforms as a recursive even/odd butterfly network (341 multiply-adds per
// 32-point transform instead of 1024), transposes through a private padded shared-memory tile, fully coalesced
// 16-byte global loads and stores.  Algorithmic traffic: one read + one write of the plane.
//
//   kind 1  dct_2d   (utils/dct.py:85-96)           kind 2  idct_2d  (utils/dct.py:99-111)
//   kind 3  low_freq (train_generator.py:47-55): idct_2d(mask_k * dct_2d(x)) == P x P^T, P = D^T diag(1_k) D
//           (the (x+1)/2*255 ... /255*2-1 affine of the reference cancels exactly because the DC term is kept).
#include "common.cuh"
#include "dct32_tables.h"

template <int N>
struct OddTable;
template <>
struct OddTable<32> { static __device__ __forceinline__ float at(int i) { return DCT32_T32[i]; } };
template <>
struct OddTable<16> { static __device__ __forceinline__ float at(int i) { return DCT32_T16[i]; } };
template <>
struct OddTable<8> { static __device__ __forceinline__ float at(int i) { return DCT32_T8[i]; } };
template <>
struct OddTable<4> { static __device__ __forceinline__ float at(int i) { return DCT32_T4[i]; } };
template <>
struct OddTable<2> { static __device__ __forceinline__ float at(int i) { return DCT32_T2[i]; } };

// forward: X = D_N x   (scaled so that the top-level N=32 result is orthonormal)
template <int N>
struct Dct {
  static __device__ __forceinline__ void fwd(const float (&x)[N], float (&X)[N]) {
    constexpr int H = N / 2;
    float u[H], v[H], E[H];
#pragma unroll
    for (int n = 0; n < H; ++n) {
      u[n] = x[n] + x[N - 1 - n];
      v[n] = x[n] - x[N - 1 - n];
    }
    Dct<H>::fwd(u, E);
#pragma unroll
    for (int k = 0; k < H; ++k) {
      float o = 0.f;
#pragma unroll
      for (int n = 0; n < H; ++n) o = fmaf(OddTable<N>::at(k * H + n), v[n], o);
      X[2 * k] = E[k];
      X[2 * k + 1] = o;
    }
  }
  // inverse (transpose of the forward flow graph): x = D_N^T X
  static __device__ __forceinline__ void inv(const float (&X)[N], float (&x)[N]) {
    constexpr int H = N / 2;
    float Ein[H], Oin[H], a[H];
#pragma unroll
    for (int k = 0; k < H; ++k) {
      Ein[k] = X[2 * k];
      Oin[k] = X[2 * k + 1];
    }
    Dct<H>::inv(Ein, a);
#pragma unroll
    for (int n = 0; n < H; ++n) {
      float b = 0.f;
#pragma unroll
      for (int k = 0; k < H; ++k) b = fmaf(OddTable<N>::at(k * H + n), Oin[k], b);
      x[n] = a[n] + b;
      x[N - 1 - n] = a[n] - b;
    }
  }
};
template <>
struct Dct<1> {
  static __device__ __forceinline__ void fwd(const float (&x)[1], float (&X)[1]) { X[0] = x[0] * DCT32_S0; }
  static __device__ __forceinline__ void inv(const float (&X)[1], float (&x)[1]) { x[0] = X[0] * DCT32_S0; }
};

#define TS 33  // padded tile stride: conflict-free for both row and column access

template <int IN_MODE>
__device__ __forceinline__ void load_plane_to_tile(const void* in, long long plane, float* tile, int lane) {
  if (IN_MODE == 0) {
    const float4* src = (const float4*)((const float*)in + plane * 1024);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 v = src[j * 32 + lane];  // 512 contiguous bytes per warp instruction
      const int e = (j * 32 + lane) * 4, r = e >> 5, c = e & 31;
      float* t = tile + r * TS + c;
      t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
    }
  } else if (IN_MODE == 1) {
    const uchar4* src = (const uchar4*)((const unsigned char*)in + plane * 1024);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uchar4 v = src[j * 32 + lane];
      const int e = (j * 32 + lane) * 4, r = e >> 5, c = e & 31;
      float* t = tile + r * TS + c;
      t[0] = (float)v.x; t[1] = (float)v.y; t[2] = (float)v.z; t[3] = (float)v.w;
    }
  } else {  // ((x+1)/2*255).byte(): truncation toward zero like torch's float -> uint8 cast
    const float4* src = (const float4*)((const float*)in + plane * 1024);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 v = src[j * 32 + lane];
      const int e = (j * 32 + lane) * 4, r = e >> 5, c = e & 31;
      float* t = tile + r * TS + c;
      t[0] = (float)(unsigned char)(int)((v.x + 1.0f) / 2.0f * 255.0f);
      t[1] = (float)(unsigned char)(int)((v.y + 1.0f) / 2.0f * 255.0f);
      t[2] = (float)(unsigned char)(int)((v.z + 1.0f) / 2.0f * 255.0f);
      t[3] = (float)(unsigned char)(int)((v.w + 1.0f) / 2.0f * 255.0f);
    }
  }
}

__device__ __forceinline__ void store_tile_to_plane(const float* tile, float* out, long long plane, int lane) {
  float4* dst = (float4*)(out + plane * 1024);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int e = (j * 32 + lane) * 4, r = e >> 5, c = e & 31;
    const float* t = tile + r * TS + c;
    dst[j * 32 + lane] = make_float4(t[0], t[1], t[2], t[3]);
  }
}

// KIND: 1 dct, 2 idct, 3 low-pass (keep x keep)
template <int KIND, int IN_MODE>
__global__ void __launch_bounds__(256, 2) dct32_k(const void* __restrict__ in, float* __restrict__ out, long long planes,
                                               int keep) {
  __shared__ float tiles[8][32 * TS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* tile = tiles[warp];
  const long long wstride = (long long)gridDim.x * 8;
  for (long long p = (long long)blockIdx.x * 8 + warp; p < planes; p += wstride) {
    load_plane_to_tile<IN_MODE>(in, p, tile, lane);
    __syncwarp();
    float a[32], b[32];
    // ---- pass 1: rows (thread = row `lane`)
#pragma unroll
    for (int k = 0; k < 32; ++k) a[k] = tile[lane * TS + k];
    if (KIND == 2) Dct<32>::inv(a, b); else Dct<32>::fwd(a, b);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 32; ++k) tile[lane * TS + k] = b[k];
    __syncwarp();
    // ---- pass 2: columns (thread = column `lane`)
#pragma unroll
    for (int k = 0; k < 32; ++k) a[k] = tile[k * TS + lane];
    if (KIND == 2) Dct<32>::inv(a, b); else Dct<32>::fwd(a, b);
    if (KIND == 3) {
      // b[m] = coefficient (row-frequency m, column-frequency lane): keep the top-left keep x keep block
#pragma unroll
      for (int m = 0; m < 32; ++m)
        if (m >= keep || lane >= keep) b[m] = 0.f;
      Dct<32>::inv(b, a);  // back along the columns
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 32; ++k) tile[k * TS + lane] = a[k];
      __syncwarp();
      // rows again: inverse along the row direction
#pragma unroll
      for (int k = 0; k < 32; ++k) a[k] = tile[lane * TS + k];
      Dct<32>::inv(a, b);
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 32; ++k) tile[lane * TS + k] = b[k];
    } else {
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 32; ++k) tile[k * TS + lane] = b[k];
    }
    __syncwarp();
    store_tile_to_plane(tile, out, p, lane);
    __syncwarp();
  }
}

extern "C" int combat_dct32_fast(const void* in, float* out, long long planes, int kind, int keep, int in_mode,
                                 void* stream) {
  COMBAT_ARG(in && out, 0);
  COMBAT_ARG(kind >= 1 && kind <= 3, 3);
  COMBAT_ARG(in_mode >= 0 && in_mode <= 2, 5);
  COMBAT_ARG(kind != 3 || (keep >= 1 && keep <= 32), 4);
  if (planes <= 0) return 0;
  long long blocks = (planes + 7) / 8;
  const long long cap = 148 * 16;
  int grid = (int)(blocks < cap ? blocks : cap);
  cudaStream_t st = (cudaStream_t)stream;
#define L(K, M) dct32_k<K, M><<<grid, 256, 0, st>>>(in, out, planes, keep)
  if (kind == 1) { if (in_mode == 0) L(1, 0); else if (in_mode == 1) L(1, 1); else L(1, 2); }
  else if (kind == 2) { if (in_mode == 0) L(2, 0); else if (in_mode == 1) L(2, 1); else L(2, 2); }
  else { if (in_mode == 0) L(3, 0); else if (in_mode == 1) L(3, 1); else L(3, 2); }
#undef L
  COMBAT_RETURN_LAUNCH("dct32_fast");
}
