// HBM-bound 32x32 and 64x64 DCT-II / DCT-III / low-pass projection: one image row (then column) per thread held in
// registers, 1-D transforms as a straight-line even/odd butterfly network with FFT-based odd parts (dct_butterfly.cuh:
// 267 multiply-adds per 32-point transform instead of 1024, 627 instead of 4096 for 64 points), ONE transpose through
// a padded shared-memory tile, coalesced global loads (16 bytes per thread) and stores (128 bytes per warp).
// Algorithmic traffic: one read + one write of the plane.
//   32x32 (CIFAR-10): one WARP per plane (8 planes per CTA, warp-level synchronisation only)
//   64x64 (CelebA)  : one 64-thread CTA per plane
//
//   kind 1  dct_2d   (utils/dct.py:85-96)           kind 2  idct_2d  (utils/dct.py:99-111)
//   kind 3  low_freq (train_generator.py:47-55): idct_2d(mask_k * dct_2d(x)) == P x P^T, P = D^T diag(1_k) D, applied
//           as P along the rows, then P along the columns
//           (the (x+1)/2*255 ... /255*2-1 affine of the reference cancels exactly because the DC term is kept).
//           The retained block size of the reference's default ratio 0.65 (keep = 20 of 32, 41 of 64) is a template
//           constant, so the compiler prunes every butterfly output / input that the mask zeroes.
#include "common.cuh"
#include "dct_butterfly.cuh"

template <int IN_MODE>
__device__ __forceinline__ void convert4(const void* in, long long idx4, float (&t)[4]) {
  if (IN_MODE == 1) {
    const uchar4 v = ((const uchar4*)in)[idx4];
    t[0] = (float)v.x; t[1] = (float)v.y; t[2] = (float)v.z; t[3] = (float)v.w;
  } else {
    const float4 v = ((const float4*)in)[idx4];
    if (IN_MODE == 0) {
      t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
    } else {  // ((x+1)/2*255).byte(): truncation toward zero like torch's float -> uint8 cast
      t[0] = (float)(unsigned char)(int)((v.x + 1.0f) / 2.0f * 255.0f);
      t[1] = (float)(unsigned char)(int)((v.y + 1.0f) / 2.0f * 255.0f);
      t[2] = (float)(unsigned char)(int)((v.z + 1.0f) / 2.0f * 255.0f);
      t[3] = (float)(unsigned char)(int)((v.w + 1.0f) / 2.0f * 255.0f);
    }
  }
}

// Shared-memory tile of one plane: row stride NP + 4 floats.  Row access by 16-byte vectors (thread t, row t) is
// conflict-free per quarter-warp ((4t + 4j) mod 32 distinct), column access by scalars (thread t, column t) always is.
template <int NP>
struct Tile {
  static constexpr int TS = NP + 4;
  static constexpr int FLOATS = NP * TS;
  // piece q (16 bytes) of the row-major plane -> its place in the padded tile
  static __device__ __forceinline__ float4* piece(float* tile, int q) { return (float4*)(tile + (q * 4 / NP) * TS + (q * 4 % NP)); }
};

// global -> tile through registers (any input mode): NP*NP/4 16-byte pieces, NP threads
template <int NP, int IN_MODE>
__device__ __forceinline__ void load_plane(const void* in, long long plane, float* tile, int tid) {
#pragma unroll
  for (int j = 0; j < NP / 4; ++j) {
    const int q = j * NP + tid;  // NP*16 contiguous bytes per instruction
    float t4[4];
    convert4<IN_MODE>(in, plane * (NP * NP / 4) + q, t4);
    *Tile<NP>::piece(tile, q) = make_float4(t4[0], t4[1], t4[2], t4[3]);
  }
}

template <int NP>
__device__ __forceinline__ void row_store(float* row, const float (&b)[NP]) {
#pragma unroll
  for (int j = 0; j < NP / 4; ++j) *(float4*)(row + 4 * j) = make_float4(b[4 * j], b[4 * j + 1], b[4 * j + 2], b[4 * j + 3]);
}

// The 1-D operator of one pass: DCT-II, DCT-III, or the low-pass projection P = D^T diag(1_keep) D.
// KEEP: compile-time upper bound of `keep` (KIND 3): outputs of the forward half at frequencies >= KEEP are never
// computed and the inverse half sees literal zeros there, so the compiler prunes both networks.
template <int NP, int KIND, int KEEP>
__device__ __forceinline__ void apply_1d(const float (&a)[NP], float (&b)[NP], int keep) {
  if (KIND == 1) {
    Dct<NP, NP>::fwd(a, b);
  } else if (KIND == 2) {
    Dct<NP, NP>::inv(a, b);
  } else {
    float c[NP];
    Dct<NP, NP>::fwd(a, c);
#pragma unroll
    for (int m = 0; m < NP; ++m)
      if (m >= KEEP || m >= keep) c[m] = 0.f;
    Dct<NP, NP>::inv(c, b);
  }
}

// The transform of the plane held in `tile` by NP threads, then its store.  Both dct_2d / idct_2d and the projection
// P X P^T are separable: pass 1 applies the 1-D operator to the rows (thread = row `tid`, 16-byte tile accesses),
// pass 2 to the columns (thread = column `tid`); the result leaves the registers as NP coalesced 4-byte stores
// (a warp writes 128 contiguous bytes of one output row per instruction) - no trip back through shared memory.
// Tile traffic per plane: fill, row read, row write, column read.
// SYNC: __syncwarp or __syncthreads; the caller synchronises between filling the tile and this call.
template <int NP, int KIND, int KEEP, typename Sync>
__device__ __forceinline__ void transform_plane(float* out, long long plane, float* tile, int tid, int keep, Sync sync) {
  constexpr int TS = Tile<NP>::TS;
  float a[NP], b[NP];
  float* row = tile + tid * TS;
  float* dst = out + plane * (NP * NP) + tid;
  // KIND 3: both passes run the SAME copy of the 1-D network (loop not unrolled).  The 64-point low-pass operator is
  // 22 KB of straight-line code and two inlined copies (43 KB) overflow the 32 KB instruction cache: ncu showed 1.8
  // warps per issue slot stalled on instruction fetch for dct64_k<3, 0, 41>; 334 -> 304 us at 16,384 images.  The plain
  // transforms fit twice and measured faster unrolled (uint8-input dct_2d: 178 vs 193 us at 32x32, 200 vs 227 at 64x64).
#pragma unroll(KIND == 3 ? 1 : 2)
  for (int pass = 0; pass < 2; ++pass) {
    if (pass == 0) {  // rows
#pragma unroll
      for (int j = 0; j < NP / 4; ++j) {
        const float4 v = *(const float4*)(row + 4 * j);
        a[4 * j] = v.x; a[4 * j + 1] = v.y; a[4 * j + 2] = v.z; a[4 * j + 3] = v.w;
      }
    } else {  // columns
#pragma unroll
      for (int k = 0; k < NP; ++k) a[k] = tile[k * TS + tid];
      sync();  // the tile may be refilled (next plane) from here on
    }
    apply_1d<NP, KIND, KEEP>(a, b, keep);
    if (pass == 0) {
      // a thread only rewrites the row it has just read: no synchronisation between its loads and these stores
      row_store<NP>(row, b);
      sync();
    } else {
#pragma unroll
      for (int k = 0; k < NP; ++k) dst[k * NP] = b[k];
    }
  }
}

struct WarpSync { __device__ __forceinline__ void operator()() const { __syncwarp(); } };
struct BlockSync { __device__ __forceinline__ void operator()() const { __syncthreads(); } };

// One plane per NP-thread group (a warp for 32x32, a 64-thread CTA for 64x64), GROUPS groups per CTA, each with its
// own tile; groups stride over the planes.  KIND: 1 dct, 2 idct, 3 low-pass (keep x keep).
template <int NP, int GROUPS, int KIND, int IN_MODE, int KEEP, typename Sync>
__device__ __forceinline__ void planes_loop(const void* in, float* out, long long planes, int keep, float* smem, Sync sync) {
  const int group = threadIdx.x / NP, tid = threadIdx.x % NP;
  float* tile = smem + group * Tile<NP>::FLOATS;
  const long long stride = (long long)gridDim.x * GROUPS;
  for (long long p = (long long)blockIdx.x * GROUPS + group; p < planes; p += stride) {
    load_plane<NP, IN_MODE>(in, p, tile, tid);
    sync();
    transform_plane<NP, KIND, KEEP>(out, p, tile, tid, keep, sync);
  }
}

template <int KIND, int IN_MODE, int KEEP>
__global__ void __launch_bounds__(256, 3) dct32_k(const void* __restrict__ in, float* __restrict__ out, long long planes,
                                                  int keep) {
  pdl_entry();
  __shared__ __align__(16) float smem[8 * Tile<32>::FLOATS];
  planes_loop<32, 8, KIND, IN_MODE, KEEP>(in, out, planes, keep, smem, WarpSync());
}

template <int KIND, int IN_MODE, int KEEP>
__global__ void __launch_bounds__(64, 10) dct64_k(const void* __restrict__ in, float* __restrict__ out, long long planes,
                                                  int keep) {
  pdl_entry();
  __shared__ __align__(16) float smem[Tile<64>::FLOATS];
  planes_loop<64, 1, KIND, IN_MODE, KEEP>(in, out, planes, keep, smem, BlockSync());
}

typedef void (*plane_kernel_t)(const void*, float*, long long, int);

// One launch.  Grid = 6 x the CTAs resident at once (or fewer if there are fewer planes), each CTA striding over the
// planes: measured on B200 at 65,536 images (profiles/r01_dct_launch_sweep.jsonl) 1x resident (a persistent grid, every
// warp in the same load / transform / store phase) reaches 84 % of the measured copy bandwidth for dct_2d at 32x32,
// 2x 91 %, 4x 98 %, 6x 98 %; a cp.async double-buffered tile at half the occupancy was never better.
template <plane_kernel_t KERNEL, int THREADS>
static void launch_planes(long long ctas_needed, const void* in, float* out, long long planes, int keep, cudaStream_t st) {
  static int cap = -1;  // per kernel instantiation (one device per process: one rank per GPU)
  if (cap < 0) {
    cap = 6 * resident_ctas(KERNEL, THREADS);
    if (cap <= 0) cap = 148 * 16;
  }
  const int grid = (int)(ctas_needed < cap ? ctas_needed : cap);
  pdl_launch(KERNEL, grid, THREADS, 0, st, in, out, planes, keep);
}

extern "C" int combat_dct32_fast(const void* in, float* out, long long planes, int kind, int keep, int in_mode,
                                 void* stream) {
  COMBAT_ARG(in && out, 0);
  COMBAT_ARG(kind >= 1 && kind <= 3, 3);
  COMBAT_ARG(in_mode >= 0 && in_mode <= 2, 5);
  COMBAT_ARG(kind != 3 || (keep >= 1 && keep <= 32), 4);
  if (planes <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
#define L(K, M, KP) launch_planes<dct32_k<K, M, KP>, 256>((planes + 7) / 8, in, out, planes, keep, st)
#define LM(K, KP) { if (in_mode == 0) L(K, 0, KP); else if (in_mode == 1) L(K, 1, KP); else L(K, 2, KP); }
  if (kind == 1) LM(1, 32)
  else if (kind == 2) LM(2, 32)
  else if (keep <= 20) LM(3, 20)
  else LM(3, 32)
#undef LM
#undef L
  COMBAT_RETURN_LAUNCH("dct32_fast");
}

extern "C" int combat_dct64_fast(const void* in, float* out, long long planes, int kind, int keep, int in_mode,
                                 void* stream) {
  COMBAT_ARG(in && out, 0);
  COMBAT_ARG(kind >= 1 && kind <= 3, 3);
  COMBAT_ARG(in_mode >= 0 && in_mode <= 2, 5);
  COMBAT_ARG(kind != 3 || (keep >= 1 && keep <= 64), 4);
  if (planes <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
#define L(K, M, KP) launch_planes<dct64_k<K, M, KP>, 64>(planes, in, out, planes, keep, st)
#define LM(K, KP) { if (in_mode == 0) L(K, 0, KP); else if (in_mode == 1) L(K, 1, KP); else L(K, 2, KP); }
  if (kind == 1) LM(1, 64)
  else if (kind == 2) LM(2, 64)
  else if (keep <= 41) LM(3, 41)
  else LM(3, 64)
#undef LM
#undef L
  COMBAT_RETURN_LAUNCH("dct64_fast");
}
