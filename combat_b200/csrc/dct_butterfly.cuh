// 1-D orthonormal DCT-II / DCT-III of 32 or 64 points as a straight-line register network (dct32.cu runs one image row,
// then one column, per thread).  utils/dct.py:13-82 of the reference computes the same transforms through an FFT of the
// mirrored signal; dct_2d / idct_2d (:85-111) apply them along both axes.
//
//   DCT-II_N(x) = interleave( DCT-II_{N/2}(u), DCT-IV_{N/2}(v) ),  u[n] = x[n] + x[N-1-n],  v[n] = x[n] - x[N-1-n]
//   DCT-III_N   = the transposed flow graph (DCT-IV is symmetric, so the same odd part serves both directions)
//
// Odd parts of 2..8 points are dense matrix-vector products (4 + 16 + 64 multiply-adds).  Odd parts of 16 and 32 points
// go through ONE complex FFT of H/2 points between two layers of plane rotations (unit-modulus twiddles only, so the
// rounding error stays at a few ulp like the dense form): 120 instead of 256 and 296 instead of 1024 multiply-adds.
// A 64-point transform is 627 fused multiply-adds / adds instead of 1491; a 32-point one 267 instead of 403.
// Every coefficient is a compile-time constant of a fully unrolled network: FFMA with a 32-bit immediate operand.
//
// The header also compiles as plain C++ (tests/dct_butterfly_host.cpp defines DCT_FN / DCT_CONST), which is how the
// network is checked against scipy on a machine without a GPU.
#pragma once
#ifndef DCT_FN
#define DCT_FN static __device__ __forceinline__
#endif
#include "dct32_tables.h"

template <int F, int N>
struct OddTable;
#define ODD_TABLE(F, N) \
  template <>           \
  struct OddTable<F, N> { DCT_FN float at(int i) { return DCT##F##_T##N[i]; } };
ODD_TABLE(32, 32) ODD_TABLE(32, 16) ODD_TABLE(32, 8) ODD_TABLE(32, 4) ODD_TABLE(32, 2)
ODD_TABLE(64, 64) ODD_TABLE(64, 32) ODD_TABLE(64, 16) ODD_TABLE(64, 8) ODD_TABLE(64, 4) ODD_TABLE(64, 2)
#undef ODD_TABLE

template <int F, int H>
struct PreTw;  // sqrt(2/F) * exp(-i pi (4n+1) / (4H))
#define PRE_TW(F, H)                                      \
  template <>                                             \
  struct PreTw<F, H> {                                    \
    DCT_FN float c(int n) { return DCT##F##_PRE_C##H[n]; } \
    DCT_FN float s(int n) { return DCT##F##_PRE_S##H[n]; } \
  };
PRE_TW(32, 16) PRE_TW(64, 16) PRE_TW(64, 32)
#undef PRE_TW

template <int H>
struct PostTw;  // exp(-i pi k / H)
template <int Q>
struct FftTw;  // exp(-2 pi i j / Q)
#define TW(NAME, TAB, N)                          \
  template <>                                     \
  struct NAME<N> {                                \
    DCT_FN float c(int i) { return TAB##_C##N[i]; } \
    DCT_FN float s(int i) { return TAB##_S##N[i]; } \
  };
TW(PostTw, DCT_POST, 16) TW(PostTw, DCT_POST, 32) TW(FftTw, DCT_FFT, 8) TW(FftTw, DCT_FFT, 16)
#undef TW

template <int F>
DCT_FN float dc_scale() { return F == 32 ? DCT32_S0 : DCT64_S0; }

DCT_FN constexpr int dct_log2(int q) { return q <= 1 ? 0 : 1 + dct_log2(q / 2); }
// bit reversal of a `bits`-wide index (bits <= 4), branch- and loop-free so that it folds after unrolling
DCT_FN constexpr int dct_bitrev(int v, int bits) {
  return (((v & 1) << 3) | ((v & 2) << 1) | ((v & 4) >> 1) | ((v & 8) >> 3)) >> (4 - bits);
}

// in-place radix-2 decimation-in-time FFT of Q complex points held in registers; input in bit-reversed order.
// One template instantiation per stage (butterfly span M) so that every index is a compile-time constant.
template <int Q, int M>
struct FftStage {
  DCT_FN void run(float (&re)[Q], float (&im)[Q]) {
    FftStage<Q, M / 2>::run(re, im);
    constexpr int h = M / 2, step = Q / M;
#pragma unroll
    for (int g = 0; g < Q; g += M) {
#pragma unroll
      for (int j = 0; j < h; ++j) {
        const int i0 = g + j, i1 = i0 + h, tw = j * step;
        const float ar = re[i0], ai = im[i0], br = re[i1], bi = im[i1];
        if (tw == 0) {  // w = 1
          re[i0] = ar + br; im[i0] = ai + bi;
          re[i1] = ar - br; im[i1] = ai - bi;
        } else if (tw * 4 == Q) {  // w = -i
          re[i0] = ar + bi; im[i0] = ai - br;
          re[i1] = ar - bi; im[i1] = ai + br;
        } else {  // w = c - i s:  b w = (br c + bi s) + i (bi c - br s)
          const float c = FftTw<Q>::c(tw), s = FftTw<Q>::s(tw);
          re[i0] = fmaf(bi, s, fmaf(br, c, ar));
          re[i1] = fmaf(bi, -s, fmaf(br, -c, ar));
          im[i0] = fmaf(br, -s, fmaf(bi, c, ai));
          im[i1] = fmaf(br, s, fmaf(bi, -c, ai));
        }
      }
    }
  }
};
template <int Q>
struct FftStage<Q, 1> {
  DCT_FN void run(float (&)[Q], float (&)[Q]) {}
};

// o = sqrt(2/F) * C^IV_H v,  C^IV[k][n] = cos(pi (2n+1)(2k+1) / (4H)):
//   z[n] = v[2n] + i v[H-1-2n];  t[n] = z[n] exp(-i pi (4n+1)/(4H));  T = FFT_{H/2}(t);  y[k] = T[k] exp(-i pi k/H);
//   o[2k] = Re y[k],  o[H-1-2k] = -Im y[k]
template <int F, int H>
DCT_FN void dct4_fft(const float (&v)[H], float (&o)[H]) {
  constexpr int Q = H / 2, BITS = dct_log2(Q);
  static_assert(BITS <= 4, "dct_bitrev handles up to 16 points");
  float re[Q], im[Q];
#pragma unroll
  for (int n = 0; n < Q; ++n) {
    const float zr = v[2 * n], zi = v[H - 1 - 2 * n], c = PreTw<F, H>::c(n), s = PreTw<F, H>::s(n);
    const int r = dct_bitrev(n, BITS);
    re[r] = fmaf(zi, s, zr * c);
    im[r] = fmaf(zr, -s, zi * c);
  }
  FftStage<Q, Q>::run(re, im);
  o[0] = re[0];
  o[H - 1] = -im[0];
#pragma unroll
  for (int k = 1; k < Q; ++k) {
    const float c = PostTw<H>::c(k), s = PostTw<H>::s(k);
    o[2 * k] = fmaf(im[k], s, re[k] * c);
    o[H - 1 - 2 * k] = fmaf(im[k], -c, re[k] * s);
  }
}

// odd part of the N = 2H point stage of family F
template <int F, int H>
struct Dct4 {
  DCT_FN void run(const float (&v)[H], float (&o)[H]) {
#pragma unroll
    for (int k = 0; k < H; ++k) {
      float acc = 0.f;
#pragma unroll
      for (int n = 0; n < H; ++n) acc = fmaf(OddTable<F, 2 * H>::at(k * H + n), v[n], acc);
      o[k] = acc;
    }
  }
};
template <int F>
struct Dct4<F, 16> {
  DCT_FN void run(const float (&v)[16], float (&o)[16]) { dct4_fft<F, 16>(v, o); }
};
template <int F>
struct Dct4<F, 32> {
  DCT_FN void run(const float (&v)[32], float (&o)[32]) { dct4_fft<F, 32>(v, o); }
};

// forward: X = D_N x   (scaled so that the top-level N = F result is orthonormal)
template <int F, int N>
struct Dct {
  DCT_FN void fwd(const float (&x)[N], float (&X)[N]) {
    constexpr int H = N / 2;
    float u[H], v[H], E[H], O[H];
#pragma unroll
    for (int n = 0; n < H; ++n) {
      u[n] = x[n] + x[N - 1 - n];
      v[n] = x[n] - x[N - 1 - n];
    }
    Dct<F, H>::fwd(u, E);
    Dct4<F, H>::run(v, O);
#pragma unroll
    for (int k = 0; k < H; ++k) {
      X[2 * k] = E[k];
      X[2 * k + 1] = O[k];
    }
  }
  // inverse (transpose of the forward flow graph): x = D_N^T X
  DCT_FN void inv(const float (&X)[N], float (&x)[N]) {
    constexpr int H = N / 2;
    float Ein[H], Oin[H], a[H], b[H];
#pragma unroll
    for (int k = 0; k < H; ++k) {
      Ein[k] = X[2 * k];
      Oin[k] = X[2 * k + 1];
    }
    Dct<F, H>::inv(Ein, a);
    Dct4<F, H>::run(Oin, b);
#pragma unroll
    for (int n = 0; n < H; ++n) {
      x[n] = a[n] + b[n];
      x[N - 1 - n] = a[n] - b[n];
    }
  }
};
template <int F>
struct Dct<F, 1> {
  DCT_FN void fwd(const float (&x)[1], float (&X)[1]) { X[0] = x[0] * dc_scale<F>(); }
  DCT_FN void inv(const float (&X)[1], float (&x)[1]) { x[0] = X[0] * dc_scale<F>(); }
};
