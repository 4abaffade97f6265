// tcgen05 implicit-GEMM convolution for sm_100a: TMA-fed (tiled 4-D boxes with hardware zero fill as the
// implicit padding), SWIZZLE_128B shared-memory operands, fp32 accumulators in TMEM, warp-specialised
// persistent CTAs (1 TMA warp, 1 MMA warp, 4 epilogue warps), double-buffered accumulators.
//
//   forward / dgrad :  D[pixel, co] = sum_{tap, ci} A[pixel @ tap, ci] * W[co, tap, ci]          (K-major A and B)
//   wgrad           :  D[(tap,ci), co] = sum_{pixel} X[pixel @ tap, ci] * dY[pixel, co]           (MN-major A and B)
//
// An M tile is a BW x BH x BNI box of output pixels (128 of them); the A operand of filter tap (kh,kw) is the
// same box shifted by the tap offset, fetched by ONE TMA box load whose out-of-range rows/columns the hardware
// fills with zeros.  Strided forward convs read one of four parity views of the input (tensor maps with doubled
// strides); the input-gradient of a strided conv is computed per output parity class, each class being a
// stride-1 problem over a subset of the taps.
//
// Replaces aten::conv2d / convolution_backward (cuDNN) for every conv with Cin % 64 == 0 and Cout % 64 == 0:
// classifier_models/preact_resnet.py:21-28, resnet.py:18-27, networks/models.py:275-314.
#include <cuda.h>
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"

// ---------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr) : "memory");
  return v;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// L2 prefetch of a tensor tile (no shared-memory destination, no barrier)
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* map, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accum)
      : "memory");
}
// arrives on the mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- issue discipline of the TMA and MMA warps.  These warps run their loops CONVERGED (all 32 lanes) and hand the asynchronous
// instructions to one lane with elect.sync.  Inside a divergent `if (lane == 0)` region nvcc cannot prove that descriptors, tensor
// memory addresses and barrier addresses are warp-uniform, so every UTCHMMA / UTMALDG / UTCBAR (they take UNIFORM registers) is
// wrapped in R2UR moves and an ELECT / BRA.U.ANY loop: ~93 cycles per tcgen05.mma on B200, i.e. the issuing thread -- not the
// tensor pipe or the shared-memory port -- paced every tile narrower than 256 columns (scripts/mma_rate.cu,
// profiles/r02_mma_issue.md: N = 64: 93 -> 48 cycles per MMA, N = 128: 93 -> 64 = the pipe's floor).  In converged code the same
// values live in uniform registers and the four MMAs of a K chunk are four consecutive UTCHMMA instructions.
__device__ __forceinline__ uint32_t elect_one_sync() {
  uint32_t pred = 0;
  asm volatile(
      "{\n.reg .b32 %%rx;\n.reg .pred %%px;\nelect.sync %%rx|%%px, %1;\n@%%px mov.s32 %0, 1;\n}\n"
      : "+r"(pred)
      : "r"(0xffffffffu));
  return pred;
}
__device__ __forceinline__ long long globaltimer_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// warp index / any value the compiler must treat as warp-uniform
__device__ __forceinline__ int uniform_i(int v) { return __shfl_sync(0xffffffffu, v, 0); }
__device__ __forceinline__ uint32_t uniform_u(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }

// ---- CTA pairs (cta_group::2): two CTAs of a cluster (the two SMs of a TPC) execute ONE 256-row MMA; each stages its own 128
// rows of A and HALF of B, so per SM the shared-memory traffic of the B operand (TMA writes and MMA reads) is halved.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address -> shared::cluster address of the same offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads whose completion is signalled on an mbarrier of the PAIR'S LEADER CTA (bar_cluster_addr: shared::cluster address)
__device__ __forceinline__ void tma_load_4d_2sm(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1, int c2,
                                                int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accum)
      : "memory");
}
// arrives on the mbarrier at the same offset in BOTH CTAs of the pair once all previously issued MMAs have completed
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(mask)
               : "memory");
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout), SWIZZLE_128B, version 1
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // version = 1 (Blackwell)
  d |= (uint64_t)2 << 61;  // layout type = SWIZZLE_128B
  return d;
}
// instruction descriptor: D=f32, A=B=bf16, majorness bits, N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------- parameters
#define TC_MAX_CLASSES 4
#define TC_MAX_TAPS 10  // 9 filter taps + the fused shortcut tap
#define TILE_M 128
#define KCHUNK 64  // bf16 elements per 128-byte swizzle row
#define TC_EPI_WARPS 8
#define TC_THREADS (64 + 32 * TC_EPI_WARPS)
#define EPI_CH 32                    // accumulator columns per epilogue chunk
#define EPI_ROWB (EPI_CH * 4 + 16)   // padded row pitch of the per-warp transpose tile (conflict-free 16-byte accesses)

struct TcTaps {
  int ntaps[TC_MAX_CLASSES];
  signed char view[TC_MAX_CLASSES][TC_MAX_TAPS];
  signed char dh[TC_MAX_CLASSES][TC_MAX_TAPS];
  signed char dw[TC_MAX_CLASSES][TC_MAX_TAPS];
  signed char wtap[TC_MAX_CLASSES][TC_MAX_TAPS];
  signed char wsel[TC_MAX_CLASSES][TC_MAX_TAPS];  // 1: the tap's filter comes from the second weight map (fused shortcut)
  int cls_p[TC_MAX_CLASSES], cls_q[TC_MAX_CLASSES];
};

struct TcParams {
  int N, Hc, Wc;          // pixel grid of one class (== output grid when n_classes == 1)
  int out_H, out_W, Co;   // full output tensor
  int out_scale;          // output pixel = class pixel * out_scale + (cls_p, cls_q)
  int BW, BH, BNI;        // pixel box, BW*BH*BNI == 128 (all powers of two)
  int lbw, lbh;           // log2(BW), log2(BH)
  int tiles_w, tiles_h, tiles_n, tiles_co, n_classes, total_tiles;
  int kchunks;            // Ci / 64
  // CTA-pair kernels: work item v = 2 * pair + cta_rank; a pair = two pixel tiles of ONE class and ONE output-channel tile
  int pair_mode, px_tiles_per_class, pairs_per_class, total_items;
  // row-reuse kernel (conv_tc_rr_kernel): work item = super-tile of two vertically adjacent pixel tiles (2s, 2s+1) that share
  // every weight tile; rr_rounds full rounds of gridDim.x items, then rr_rem items -- split into single tiles when they fit
  int rr_rounds, rr_rem, rr_split;
  int n_epf;              // number of epilogue-input tensors to prefetch (maps.e[0 .. n_epf)), 0: off
  const float* img;       // conv_tc_first_kernel: the float32 NCHW [N,3,H,W] input image
  // tile decode without integer division (measured: decode + addressing of the epilogue = 950 cycles per tile per warp with the
  // hardware-emulated `/` and `%` by run-time values): q = umulhi(x, m), m = floor(2^32 / d) + 1, exact while x * d < 2^32
  // (checked on the host); m = 0 encodes d = 1
  uint32_t m_co, m_w, m_h, m_n, m_ppc;
  void* out;             // bf16 or float32 (out_f32)
  const void* residual;  // bf16 or float32 (res_f32)
  const float* bias;
  int out_f32, res_f32;
  int act;
  const float* post_scale;
  const float* post_shift;
  // fused consumers of the accumulator (all optional)
  void* out2;               // bf16 NHWC: relu(v * scale2[c] + shift2[c]) -- the eval-mode BatchNorm+ReLU the next conv reads
  const float* scale2;
  const float* shift2;
  const bf16* mask;         // bf16 NHWC like out: v = mask > 0 ? v * mask_scale[c] : 0 -- backward of relu(bn_eval(.))
  const float* mask_scale;
  const bf16* post_add;     // bf16 NHWC like out, added after the mask (gradient arriving over an identity shortcut)
  // MODE 3 (train-mode relu(bn(x)) backward, reduction half): `mask` = x, mask_scale = scale; out = (x*scale+shift > 0) ? acc : 0
  const float* bnb_shift;
  const float* bnb_mean;
  const float* bnb_invstd;
  float* stats;             // optional [gridDim.x][2][Co] per-CTA partial sums / sums of squares of `out` (train-mode BatchNorm)
  long long* dbg;           // optional [gridDim.x][8] cycle counters (pipeline diagnostics, scripts/bench_conv.py --dbg)
  TcTaps taps;
};

struct TcMaps {
  CUtensorMap in[4];  // parity views of the input (only [0] when unstrided; [1] = the shortcut's gradient tensor when fused)
  CUtensorMap w;      // [Co][taps][Ci] as (Ci, taps, Co)
  CUtensorMap w2;     // fused shortcut: [Co][1][Ci]
  // epilogue inputs (mask / residual / shortcut gradient: bf16 NHWC like the output), box = one pixel tile x 64 channels.  The TMA
  // warp prefetches them into L2 two tiles ahead: the epilogue's coalesced loads of a 64-channel layer are otherwise exposed
  // HBM latency (measured: 1700 of 4180 cycles per tile of the masked 64 -> 64 input gradient).
  CUtensorMap e[2];
};

template <int BLOCK_N>
struct TcCfg {
  static constexpr int A_BYTES = TILE_M * KCHUNK * 2;    // 16 KB
  static constexpr int B_BYTES = BLOCK_N * KCHUNK * 2;   // 8 / 16 / 32 KB
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = BLOCK_N == 64 ? 7 : (BLOCK_N == 128 ? 5 : 3);
  static constexpr int TMEM_COLS = 2 * BLOCK_N;          // two accumulator buffers (128, 256 or 512: powers of two)
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + TC_EPI_WARPS * 32 * EPI_ROWB /*epilogue*/;
};

__device__ __forceinline__ uint32_t fdiv(uint32_t x, uint32_t m) { return m ? __umulhi(x, m) : x; }

__device__ __forceinline__ void decode_tile(const TcParams& p, int tile, int& cls, int& nt, int& ht, int& wt, int& cot) {
  if (p.pair_mode) {
    const uint32_t rank = tile & 1, q = (uint32_t)tile >> 1;
    const uint32_t pp = fdiv(q, p.m_co);
    cot = q - pp * p.tiles_co;
    const uint32_t c = fdiv(pp, p.m_ppc);
    cls = c;
    const uint32_t px = 2 * (pp - c * p.pairs_per_class) + rank;
    if ((int)px >= p.px_tiles_per_class) {  // odd tile count: the pair's second half is all padding (TMA zero fill, no stores)
      nt = p.tiles_n;
      ht = wt = 0;
      return;
    }
    const uint32_t r2 = fdiv(px, p.m_w);
    wt = px - r2 * p.tiles_w;
    const uint32_t r3 = fdiv(r2, p.m_h);
    ht = r2 - r3 * p.tiles_h;
    nt = r3;
    return;
  }
  uint32_t r = fdiv((uint32_t)tile, p.m_co);
  cot = tile - r * p.tiles_co;
  uint32_t r1 = fdiv(r, p.m_w);
  wt = r - r1 * p.tiles_w;
  uint32_t r2 = fdiv(r1, p.m_h);
  ht = r1 - r2 * p.tiles_h;
  uint32_t r3 = fdiv(r2, p.m_n);
  nt = r2 - r3 * p.tiles_n;
  cls = r3;
}

// k-th pixel tile of this CTA, its accumulator buffer and the parity of that buffer's barriers.  GROUP == 1: tiles blockIdx.x +
// k * gridDim.x, two buffers.  GROUP == 2 (row-reuse kernel): items of two tiles, four buffers (2 * (item & 1) + half).
template <int GROUP>
__device__ __forceinline__ bool tile_seq(const TcParams& p, int k, int& tile, int& acc, uint32_t& phase) {
  if (GROUP == 1) {
    tile = blockIdx.x + k * gridDim.x;
    acc = k & 1;
    phase = (k >> 1) & 1;
    return tile < p.total_tiles;
  }
  int i, j, st;
  if (k < 2 * p.rr_rounds) {
    i = k >> 1;
    j = k & 1;
    st = blockIdx.x + i * gridDim.x;
  } else {
    const int kk = k - 2 * p.rr_rounds;
    i = p.rr_rounds;
    if (p.rr_split) {
      if (kk > 0 || (int)blockIdx.x >= 2 * p.rr_rem) return false;
      st = p.rr_rounds * gridDim.x + (blockIdx.x >> 1);
      j = blockIdx.x & 1;
    } else {
      if (kk > 1 || (int)blockIdx.x >= p.rr_rem) return false;
      st = p.rr_rounds * gridDim.x + blockIdx.x;
      j = kk;
    }
  }
  tile = 2 * st + j;
  acc = 2 * (i & 1) + j;
  phase = (i >> 1) & 1;
  return true;
}
// i-th work item of the row-reuse kernel: super-tile index and which of its two tiles are computed (bit 0 / bit 1)
__device__ __forceinline__ bool rr_item(const TcParams& p, int i, int& st, int& jmask) {
  if (i < p.rr_rounds) {
    st = blockIdx.x + i * gridDim.x;
    jmask = 3;
    return true;
  }
  if (i > p.rr_rounds) return false;
  if (p.rr_split) {
    if ((int)blockIdx.x >= 2 * p.rr_rem) return false;
    st = p.rr_rounds * gridDim.x + (blockIdx.x >> 1);
    jmask = 1 << (blockIdx.x & 1);
    return true;
  }
  if ((int)blockIdx.x >= p.rr_rem) return false;
  st = p.rr_rounds * gridDim.x + blockIdx.x;
  jmask = 3;
  return true;
}

// ---------------------------------------------------------------------------------------------- epilogue (shared by all mainloops)
// MODE 0: out = act(acc + bias)*post_scale+post_shift + residual; out2 = bf16 relu(out*scale2+shift2)
// MODE 2: MODE 0 + per-(CTA, row quarter) partial sums / sums of squares of `out` (train-mode BatchNorm statistics)
// MODE 1: out = bf16( mask > 0 ? (acc + residual) * mask_scale : 0 ) + post_add     (residual XOR post_add, both bf16)
// MODE 3: out = bf16 g, g = (x * scale + shift > 0) ? acc : 0, + per-(CTA, row quarter) partial sums of g and g * (x - mean) * invstd
//         (the reduction of a train-mode BatchNorm+ReLU backward; x = the saved bf16 pre-normalisation tensor)
struct NoTileHook {
  __device__ __forceinline__ void operator()(int) const {}
};
// hook(k): called by every epilogue warp at the top of iteration k, before the wait on accumulator k (conv_tc_first_kernel builds
// the A operand of tile k + 1 there)
template <int BLOCK_N, int MODE, bool PAIR = false, int GROUP = 1, typename Hook = NoTileHook>
__device__ __forceinline__ void tc_epilogue(const TcParams& p, uint8_t* stg_base, uint32_t tmem_base, uint64_t* tfull_bar,
                                            uint64_t* tempty_bar, int warp, int lane, Hook hook = Hook()) {
  // PAIR: the accumulator-empty barriers live in the pair's leader CTA (its MMA thread waits for the epilogue warps of BOTH CTAs)
  uint32_t tempty_leader[2] = {0, 0};
  if (PAIR) {
    tempty_leader[0] = mapa_u32(smem_u32(&tempty_bar[0]), 0);
    tempty_leader[1] = mapa_u32(smem_u32(&tempty_bar[1]), 0);
  }
  static_assert(!(PAIR && GROUP != 1), "CTA pairs use two accumulator buffers");
  // ===================== epilogue: TMEM -> registers -> shared (transpose) -> coalesced global =====================
  // 8 warps: warp w owns TMEM lanes [32*(w%4), +32) (hardware rule) and the column half (w-2)/4 of the tile.
  // tcgen05.ld hands a thread one accumulator ROW (pixel); writing rows straight to NHWC memory makes every store
  // instruction touch 32 different lines.  Each warp therefore transposes its 32 x 32 chunk through a padded
  // shared-memory tile and continues in a layout where a warp instruction covers 4 rows x 128 contiguous bytes:
  // residual / mask / shortcut-gradient loads and all stores are full-line accesses.  Those loads are ISSUED BEFORE
  // the wait on the accumulator, so their latency hides under the tile's MMAs.
  const int quarter = warp & 3;
  const int half = (warp - 2) >> 2;
  constexpr int NCH = BLOCK_N / 64;  // 32-column chunks per warp
  // 64- / 128-wide tiles prefetch the epilogue inputs of the whole tile (64-wide: one tile ahead); 256-wide tiles have K >=
  // 2304 of MMA work per tile to hide the epilogue behind and load per chunk instead (the prefetch would need 128 registers)
  constexpr bool PREFETCH = NCH <= 2;
  constexpr int NPF = PREFETCH ? NCH : 1;
  const uint32_t stg = smem_u32(stg_base) + (warp - 2) * (32 * EPI_ROWB);  // explicit shared-space address
  constexpr bool FWD = MODE == 0 || MODE == 2;    // MODE 2 = MODE 0 + train-mode BatchNorm statistics of `out`
  constexpr bool BNB = MODE == 3;
  constexpr bool STATS = MODE == 2 || MODE == 3;
  // running per-column sum / sum of squares of this warp's rows over ALL its tiles (the launcher guarantees that the
  // output-channel tile of a CTA never changes: gridDim.x % tiles_co == 0); written once at the end
  float4 ssum[NCH], ssq[NCH];
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) ssum[ch] = ssq[ch] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int sub = lane >> 3;     // row within a group of 4
  const int cseg = lane & 7;     // 16-byte column segment: columns cseg*4 .. cseg*4+3 of the chunk
  // per-tile addressing + prefetch of the epilogue inputs, software-pipelined ONE TILE AHEAD: the loads of tile t+1 are
  // in flight while tile t is transposed, combined and stored
  struct TileCtx {
    uint32_t ob[8];           // element offset of (row i, this lane's first channel) in the NHWC output (< 2^32 elements)
    uint32_t okmask;          // rows inside the tensor
    int col0;
    bool has_acc;
    float4 pr[NPF][8];        // MODE 0: float32 residual | bf16 residual in .x,.y;  MODE 1: mask in .x,.y, addend in .z,.w
  };
  // epilogue inputs of one 32-column chunk (coalesced, unconditional: rows outside the tensor read element 0)
  auto load_inputs = [&](const TileCtx& t, int ch, float4 (&dst)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t o = (t.okmask >> i & 1) ? t.ob[i] + ch * EPI_CH : 0u;
      if (BNB) {
        *(uint2*)&dst[i].x = __ldg((const uint2*)(p.mask + o));
      } else if (MODE != 1) {
        if (p.residual) {
          if (p.res_f32) dst[i] = __ldg((const float4*)((const float*)p.residual + o));
          else *(uint2*)&dst[i].x = __ldg((const uint2*)((const bf16*)p.residual + o));
        }
      } else {
        *(uint2*)&dst[i].x = __ldg((const uint2*)(p.mask + o));
        const bf16* addp = p.residual ? (const bf16*)p.residual : p.post_add;
        if (addp) *(uint2*)&dst[i].z = __ldg((const uint2*)(addp + o));
      }
    }
  };
  auto prepare = [&](int tile, TileCtx& t) {
    int cls, nt, ht, wt, cot;
    decode_tile(p, tile, cls, nt, ht, wt, cot);
    t.has_acc = p.taps.ntaps[cls] > 0;
    t.col0 = cot * BLOCK_N + half * (BLOCK_N / 2) + cseg * 4;  // this lane's first channel of chunk 0
    t.okmask = 0;
    const uint32_t cp = p.taps.cls_p[cls], cq = p.taps.cls_q[cls];
    // the 8 rows this lane owns in the coalesced phase: row = quarter*32 + 4*i + sub
    if (p.BW >= 32) {
      // 32 | BW: the warp's 32 pixels are one run inside ONE image row -- one address, then a constant stride (the general
      // loop below costs ~530 cycles per tile per warp, which is exposed whenever the epilogue is the bound)
      const int row0 = quarter * 32 + sub;
      const int wi0 = row0 & (p.BW - 1);
      const int r2 = row0 >> p.lbw;
      const int hi = r2 & (p.BH - 1);
      const int ni = r2 >> p.lbh;
      const int n = nt * p.BNI + ni, a = ht * p.BH + hi, b0 = wt * p.BW + wi0;
      const bool vrow = n < p.N && a < p.Hc;
      const uint32_t oh = a * p.out_scale + cp, ow0 = b0 * p.out_scale + cq;
      const uint32_t o0 = (((uint32_t)n * (uint32_t)p.out_H + oh) * (uint32_t)p.out_W + ow0) * (uint32_t)p.Co + (uint32_t)t.col0;
      const uint32_t step = 4u * (uint32_t)p.out_scale * (uint32_t)p.Co;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        t.ob[i] = o0 + i * step;
        t.okmask |= ((vrow && b0 + 4 * i < p.Wc) ? 1u : 0u) << i;
      }
    } else
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = quarter * 32 + 4 * i + sub;
      const int wi = row & (p.BW - 1);
      const int r2 = row >> p.lbw;
      const int hi = r2 & (p.BH - 1);
      const int ni = r2 >> p.lbh;
      const int n = nt * p.BNI + ni, a = ht * p.BH + hi, b = wt * p.BW + wi;
      const bool valid = n < p.N && a < p.Hc && b < p.Wc;
      const uint32_t oh = a * p.out_scale + cp, ow = b * p.out_scale + cq;
      // element offset < 2^32 (checked on the host): modular 32-bit arithmetic gives it directly (rows outside the tensor are
      // masked, their value is irrelevant)
      t.ob[i] = (((uint32_t)n * (uint32_t)p.out_H + oh) * (uint32_t)p.out_W + ow) * (uint32_t)p.Co + (uint32_t)t.col0;
      t.okmask |= (valid ? 1u : 0u) << i;
    }
    // Unconditional loads (rows outside the tensor read element 0 and are discarded at the store): a predicated load
    // into a zero-initialised register becomes "load to a temporary, wait, move" and serialises the whole batch.
    if (PREFETCH) {
#pragma unroll
      for (int ch = 0; ch < NPF; ++ch) load_inputs(t, ch, t.pr[ch]);
    }
  };
  // one tile ahead only where the register file allows it (64-wide tiles: 32 prefetch registers per tile)
  constexpr bool AHEAD = NCH == 1;
  long long e_wait = 0, e_body = 0;
  TileCtx cur, nxt;
  int tile, acc, tile_n, acc_n;
  uint32_t acc_phase, phase_n;
  // GROUP == 1: tiles of a class without taps (strided input gradients) use no accumulator, so the buffer index / parity
  // advance only on tiles that have one (as in the MMA warp)
  int acc_run = 0;
  uint32_t phase_run = 0;
  if (AHEAD && tile_seq<GROUP>(p, 0, tile, acc, acc_phase)) prepare(tile, cur);
  for (int k = 0; tile_seq<GROUP>(p, k, tile, acc, acc_phase); ++k) {
    if (GROUP == 1) { acc = acc_run; acc_phase = phase_run; }
    hook(k);
    const long long tp0 = (p.dbg && warp == 2 && lane == 0) ? clock64() : 0;
    const bool more = AHEAD && tile_seq<GROUP>(p, k + 1, tile_n, acc_n, phase_n);
    if (AHEAD) {
      if (more) prepare(tile_n, nxt);
    } else {
      prepare(tile, cur);
    }
    const bool has_acc = cur.has_acc;
    const int col0 = cur.col0;
    const uint32_t okmask = cur.okmask;
    uint32_t (&ob)[8] = cur.ob;
    float4 (&pr_tile)[NPF][8] = cur.pr;
    float4 pr_pipe[2][8];
    const bool dbg_w = p.dbg && warp == 2 && lane == 0;
    const long long te0 = dbg_w ? clock64() : 0;
    if (dbg_w && p.img) p.dbg[blockIdx.x * 8 + 3] += te0 - tp0;   // (first-conv kernel only: slot 3 is free there) addressing / prefetch
    if (has_acc) {
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
    }
    const long long te1 = dbg_w ? clock64() : 0;
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BLOCK_N + half * (BLOCK_N / 2);
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      const int c0 = ch * EPI_CH;
      const int colg = col0 + c0;
      // 256-wide tiles: inputs of chunk ch+1 are requested before chunk ch is processed (two register buffers)
      if (!PREFETCH) {
        if (ch == 0) load_inputs(cur, 0, pr_pipe[0]);
        if (ch + 1 < NCH) load_inputs(cur, ch + 1, pr_pipe[(ch + 1) & 1]);
      }
      float4 (&prc)[8] = PREFETCH ? pr_tile[PREFETCH ? ch : 0] : pr_pipe[ch & 1];
      float4 f[8];
      if (has_acc) {
        uint32_t v[32];
        tmem_ld16(taddr + c0, *(uint32_t(*)[16])&v[0]);
        tmem_ld16(taddr + c0 + 16, *(uint32_t(*)[16])&v[16]);
        tmem_ld_wait();
        const uint32_t wr = stg + lane * EPI_ROWB;
#pragma unroll
        for (int j = 0; j < 8; ++j) sts128(wr + j * 16, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = lds128(stg + (4 * i + sub) * EPI_ROWB + cseg * 16);
        __syncwarp();
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (FWD) {
        if (p.bias) {
          const float4 bb = *(const float4*)(p.bias + colg);
#pragma unroll
          for (int i = 0; i < 8; ++i) { f[i].x += bb.x; f[i].y += bb.y; f[i].z += bb.z; f[i].w += bb.w; }
        }
        if (p.act == 2) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            // ELU on the SFU: exp(x) - 1 has an ABSOLUTE error of ~1e-7 for x <= 0, far below the bf16 / float32-accumulator
            // noise of the tensor-core path (the float32 CUDA-core path keeps expm1f)
            f[i].x = f[i].x > 0.f ? f[i].x : __expf(f[i].x) - 1.f; f[i].y = f[i].y > 0.f ? f[i].y : __expf(f[i].y) - 1.f;
            f[i].z = f[i].z > 0.f ? f[i].z : __expf(f[i].z) - 1.f; f[i].w = f[i].w > 0.f ? f[i].w : __expf(f[i].w) - 1.f;
          }
        }
        if (p.post_scale) {
          const float4 ss = *(const float4*)(p.post_scale + colg), hh = *(const float4*)(p.post_shift + colg);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            f[i].x = fmaf(f[i].x, ss.x, hh.x); f[i].y = fmaf(f[i].y, ss.y, hh.y);
            f[i].z = fmaf(f[i].z, ss.z, hh.z); f[i].w = fmaf(f[i].w, ss.w, hh.w);
          }
        }
        if (p.residual) {
          if (p.res_f32) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { f[i].x += prc[i].x; f[i].y += prc[i].y; f[i].z += prc[i].z; f[i].w += prc[i].w; }
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint32_t u0 = __float_as_uint(prc[i].x), u1 = __float_as_uint(prc[i].y);
              const float2 r0 = __bfloat1622float2(*(const __nv_bfloat162*)&u0), r1 = __bfloat1622float2(*(const __nv_bfloat162*)&u1);
              f[i].x += r0.x; f[i].y += r0.y; f[i].z += r1.x; f[i].w += r1.y;
            }
          }
        }
      } else if (BNB) {
        const float4 sc = *(const float4*)(p.mask_scale + colg), sh = *(const float4*)(p.bnb_shift + colg);
        const float4 mu = *(const float4*)(p.bnb_mean + colg), is = *(const float4*)(p.bnb_invstd + colg);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t x0u = __float_as_uint(prc[i].x), x1u = __float_as_uint(prc[i].y);
          const float2 x0 = __bfloat1622float2(*(const __nv_bfloat162*)&x0u), x1 = __bfloat1622float2(*(const __nv_bfloat162*)&x1u);
          // the forward's affine_act computed fmaf(x, scale, shift) from the same stored x: the same predicate
          f[i].x = fmaf(x0.x, sc.x, sh.x) > 0.f ? f[i].x : 0.f; f[i].y = fmaf(x0.y, sc.y, sh.y) > 0.f ? f[i].y : 0.f;
          f[i].z = fmaf(x1.x, sc.z, sh.z) > 0.f ? f[i].z : 0.f; f[i].w = fmaf(x1.y, sc.w, sh.w) > 0.f ? f[i].w : 0.f;
          if (okmask >> i & 1) {
            ssum[ch].x += f[i].x; ssum[ch].y += f[i].y; ssum[ch].z += f[i].z; ssum[ch].w += f[i].w;
            ssq[ch].x = fmaf(f[i].x, (x0.x - mu.x) * is.x, ssq[ch].x); ssq[ch].y = fmaf(f[i].y, (x0.y - mu.y) * is.y, ssq[ch].y);
            ssq[ch].z = fmaf(f[i].z, (x1.x - mu.z) * is.z, ssq[ch].z); ssq[ch].w = fmaf(f[i].w, (x1.y - mu.w) * is.w, ssq[ch].w);
          }
        }
      } else {
        const float4 ms = *(const float4*)(p.mask_scale + colg);
        const bool pre_add = p.residual != nullptr;
        const bool any_add = pre_add || p.post_add != nullptr;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t m0u = __float_as_uint(prc[i].x), m1u = __float_as_uint(prc[i].y);
          const uint32_t a0u = any_add ? __float_as_uint(prc[i].z) : 0u, a1u = any_add ? __float_as_uint(prc[i].w) : 0u;
          const float2 m0 = __bfloat1622float2(*(const __nv_bfloat162*)&m0u), m1 = __bfloat1622float2(*(const __nv_bfloat162*)&m1u);
          const float2 a0 = __bfloat1622float2(*(const __nv_bfloat162*)&a0u), a1 = __bfloat1622float2(*(const __nv_bfloat162*)&a1u);
          if (pre_add) { f[i].x += a0.x; f[i].y += a0.y; f[i].z += a1.x; f[i].w += a1.y; }
          f[i].x = m0.x > 0.f ? f[i].x * ms.x : 0.f; f[i].y = m0.y > 0.f ? f[i].y * ms.y : 0.f;
          f[i].z = m1.x > 0.f ? f[i].z * ms.z : 0.f; f[i].w = m1.y > 0.f ? f[i].w * ms.w : 0.f;
          if (!pre_add) { f[i].x += a0.x; f[i].y += a0.y; f[i].z += a1.x; f[i].w += a1.y; }
        }
      }
      if (STATS && !BNB) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (okmask >> i & 1) {
            ssum[ch].x += f[i].x; ssum[ch].y += f[i].y; ssum[ch].z += f[i].z; ssum[ch].w += f[i].w;
            ssq[ch].x = fmaf(f[i].x, f[i].x, ssq[ch].x); ssq[ch].y = fmaf(f[i].y, f[i].y, ssq[ch].y);
            ssq[ch].z = fmaf(f[i].z, f[i].z, ssq[ch].z); ssq[ch].w = fmaf(f[i].w, f[i].w, ssq[ch].w);
          }
      }
      if (p.out) {
        if (p.out_f32) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (okmask >> i & 1) *(float4*)((float*)p.out + ob[i] + c0) = f[i];
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (okmask >> i & 1) {
              const __nv_bfloat162 h0 = __floats2bfloat162_rn(f[i].x, f[i].y), h1 = __floats2bfloat162_rn(f[i].z, f[i].w);
              *(uint2*)((bf16*)p.out + ob[i] + c0) = make_uint2(*(const uint32_t*)&h0, *(const uint32_t*)&h1);
            }
        }
      }
      if (FWD && p.out2) {
        const float4 ss = *(const float4*)(p.scale2 + colg), hh = *(const float4*)(p.shift2 + colg);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (okmask >> i & 1) {
            const __nv_bfloat162 h0 = __floats2bfloat162_rn(fmaxf(fmaf(f[i].x, ss.x, hh.x), 0.f), fmaxf(fmaf(f[i].y, ss.y, hh.y), 0.f));
            const __nv_bfloat162 h1 = __floats2bfloat162_rn(fmaxf(fmaf(f[i].z, ss.z, hh.z), 0.f), fmaxf(fmaf(f[i].w, ss.w, hh.w), 0.f));
            *(uint2*)((bf16*)p.out2 + ob[i] + c0) = make_uint2(*(const uint32_t*)&h0, *(const uint32_t*)&h1);
          }
      }
    }
    if (has_acc) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_cluster(tempty_leader[acc & 1]);
        else mbar_arrive(&tempty_bar[acc]);
      }
      if (++acc_run == 2) { acc_run = 0; phase_run ^= 1; }
    }
    if (dbg_w) {
      e_wait += te1 - te0;          // epilogue warp waiting for the accumulator
      e_body += clock64() - te1;    // epilogue body
    }
    if (more) cur = nxt;
  }
  if (p.dbg && warp == 2 && lane == 0) {
    p.dbg[blockIdx.x * 8 + 4] += e_wait;
    p.dbg[blockIdx.x * 8 + 5] += e_body;
  }
  if (STATS) {
    // partial block = (CTA, quarter): [2][Co]; lanes with sub == 0 hold the sums of the warp's 32 rows after the reduction
    const int cot = (PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x) % p.tiles_co;
    float* dst = p.stats + ((long long)blockIdx.x * 4 + quarter) * 2 * p.Co + cot * BLOCK_N + half * (BLOCK_N / 2) + cseg * 4;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      float4 sm = ssum[ch], sq = ssq[ch];
#pragma unroll
      for (int o = 8; o <= 16; o <<= 1) {
        sm.x += __shfl_xor_sync(0xffffffffu, sm.x, o); sm.y += __shfl_xor_sync(0xffffffffu, sm.y, o);
        sm.z += __shfl_xor_sync(0xffffffffu, sm.z, o); sm.w += __shfl_xor_sync(0xffffffffu, sm.w, o);
        sq.x += __shfl_xor_sync(0xffffffffu, sq.x, o); sq.y += __shfl_xor_sync(0xffffffffu, sq.y, o);
        sq.z += __shfl_xor_sync(0xffffffffu, sq.z, o); sq.w += __shfl_xor_sync(0xffffffffu, sq.w, o);
      }
      if (sub == 0) {
        *(float4*)(dst + ch * EPI_CH) = sm;
        *(float4*)(dst + p.Co + ch * EPI_CH) = sq;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- fwd / dgrad kernel
template <int BLOCK_N, int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const __grid_constant__ TcMaps maps, const TcParams p) {
  pdl_launch_dependents();  // the next kernel may be scheduled; this one waits for its predecessor after its prologue
  const long long dbg_c0 = p.dbg ? clock64() : 0, dbg_g0 = p.dbg ? globaltimer_ns() : 0;  // whole-kernel cycles / ns of this CTA
  using Cfg = TcCfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);  // SWIZZLE_128B atoms need 1024 B alignment
  uint64_t* bars = (uint64_t*)(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::STAGES;
  uint64_t* tfull_bar = bars + 2 * Cfg::STAGES;
  uint64_t* tempty_bar = bars + 2 * Cfg::STAGES + 2;
  uint32_t* tmem_ptr_smem = (uint32_t*)(bars + 2 * Cfg::STAGES + 4);

  const int warp = uniform_i(threadIdx.x >> 5), lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.w);
    tma_prefetch_desc(&maps.in[0]);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], TC_EPI_WARPS);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = uniform_u(*tmem_ptr_smem);
  pdl_wait();  // barriers initialised, tensor memory allocated: nothing above touched global memory

  if (warp == 0) {
    // ===================== TMA producer (converged warp, one elected lane issues) =====================
    int stage = 0;
    uint32_t phase = 0;
    long long w_empty = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      int cls, nt, ht, wt, cot;
      decode_tile(p, tile, cls, nt, ht, wt, cot);
      const int ntaps = p.taps.ntaps[cls];
      for (int t = 0; t < ntaps; ++t) {
        const CUtensorMap* amap = &maps.in[p.taps.view[cls][t]];
        const int cw = wt * p.BW + p.taps.dw[cls][t];
        const int ch = ht * p.BH + p.taps.dh[cls][t];
        const int cn = nt * p.BNI;
        const int wtap = p.taps.wtap[cls][t];
        const CUtensorMap* wmap = p.taps.wsel[cls][t] ? &maps.w2 : &maps.w;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          const long long t0 = p.dbg ? clock64() : 0;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (p.dbg) w_empty += clock64() - t0;   // producer waiting for a free slot
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          if (elect_one_sync()) {
            mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
            tma_load_4d(sa, amap, &full_bar[stage], kc * KCHUNK, cw, ch, cn);
            tma_load_3d(sb, wmap, &full_bar[stage], kc * KCHUNK, wtap, cot * BLOCK_N);
          }
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    if (p.dbg && lane == 0) p.dbg[blockIdx.x * 8 + 0] += w_empty;
  } else if (warp == 1) {
    // ===================== MMA issuer (converged warp, one elected lane issues) =====================
    constexpr uint32_t idesc = make_idesc(TILE_M, BLOCK_N, 0, 0);
    const uint32_t smem_base = smem_u32(smem);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    long long w_tempty = 0, w_full = 0, t_last = 0;
    const long long tstart = p.dbg ? clock64() : 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      if (p.dbg) t_last = clock64() - tstart;  // (time up to the start of the last tile)
      int cls, nt, ht, wt, cot;
      decode_tile(p, tile, cls, nt, ht, wt, cot);
      const int kiters = p.taps.ntaps[cls] * p.kchunks;
      if (kiters == 0) continue;
      long long t0 = p.dbg ? clock64() : 0;
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      if (p.dbg) w_tempty += clock64() - t0;     // MMA warp waiting for the epilogue (accumulator buffer)
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
      for (int it = 0; it < kiters; ++it) {
        t0 = p.dbg ? clock64() : 0;
        mbar_wait(&full_bar[stage], phase);
        if (p.dbg) w_full += clock64() - t0;     // MMA warp waiting for operands (TMA)
        tc_fence_after();
        const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
        const uint32_t sb = sa + Cfg::A_BYTES;
        const uint64_t adesc = make_smem_desc(sa, 16, 1024);
        const uint64_t bdesc = make_smem_desc(sb, 16, 1024);
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < KCHUNK / 16; ++k) {
            // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle row: +2 in the (addr >> 4) field
            umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (it | k) != 0);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
        }
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
      if (elect_one_sync()) umma_commit(&tfull_bar[acc]);  // accumulator complete
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (p.dbg && lane == 0) {
      p.dbg[blockIdx.x * 8 + 1] += w_tempty;
      p.dbg[blockIdx.x * 8 + 2] += w_full;
      p.dbg[blockIdx.x * 8 + 3] = t_last;
    }
  } else {
    tc_epilogue<BLOCK_N, MODE>(p, smem + Cfg::STAGES * Cfg::STAGE_BYTES + 256, tmem_base, tfull_bar, tempty_bar, warp, lane);
  }
  tc_fence_before();
  __syncthreads();
  if (p.dbg && threadIdx.x == 0) {
    p.dbg[blockIdx.x * 8 + 6] += clock64() - dbg_c0;
    p.dbg[blockIdx.x * 8 + 7] += globaltimer_ns() - dbg_g0;
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------- fwd / dgrad kernel, CTA pairs
// Same pipeline as conv_tc_kernel, executed by a cluster of two CTAs (the two SMs of a TPC) as ONE 256 x BLOCK_N MMA per K step
// (tcgen05.mma.cta_group::2): CTA r of the pair owns pixel tile 2*pair + r (its 128 accumulator rows live in its own TMEM) and
// stages rows [r * BLOCK_N/2, +BLOCK_N/2) of the weight tile; the tensor cores of both SMs read both halves.  Per SM and K step
// of 64: 16 KB of A + BLOCK_N/2 * 128 B of B through the shared-memory port (TMA write + MMA read) instead of 16 KB + BLOCK_N *
// 128 B -- the port (128 B/clk, shared with the epilogue transposes) is what bounds the single-CTA kernel.
// Barriers: the leader's (rank 0) full barriers collect the TMA bytes of BOTH CTAs; its MMA thread frees smem slots and
// publishes accumulators with multicast commits (one arrive in each CTA); the epilogue warps of both CTAs release an
// accumulator buffer by arriving on the LEADER's barrier.
template <int BLOCK_N>
struct TcPairCfg {
  static constexpr int A_BYTES = TILE_M * KCHUNK * 2;            // 16 KB
  static constexpr int B_BYTES = (BLOCK_N / 2) * KCHUNK * 2;     // this CTA's half: 8 / 16 KB
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = BLOCK_N == 256 ? 5 : 7;
  static constexpr int TMEM_COLS = 2 * BLOCK_N;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256 + TC_EPI_WARPS * 32 * EPI_ROWB;
};

template <int BLOCK_N, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
    conv_tc_pair_kernel(const __grid_constant__ TcMaps maps, const TcParams p) {
  pdl_launch_dependents();  // the next kernel may be scheduled; this one waits for its predecessor after its prologue
  const long long dbg_c0 = p.dbg ? clock64() : 0, dbg_g0 = p.dbg ? globaltimer_ns() : 0;  // whole-kernel cycles / ns of this CTA
  using Cfg = TcPairCfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::STAGES;
  uint64_t* tfull_bar = bars + 2 * Cfg::STAGES;
  uint64_t* tempty_bar = bars + 2 * Cfg::STAGES + 2;
  uint32_t* tmem_ptr_smem = (uint32_t*)(bars + 2 * Cfg::STAGES + 4);

  const int warp = uniform_i(threadIdx.x >> 5), lane = threadIdx.x & 31;
  const uint32_t rank = uniform_u(cluster_ctarank());

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.w);
    tma_prefetch_desc(&maps.in[0]);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);    // leader: its own arrive.expect_tx; the bytes of both CTAs complete the phase
      mbar_init(&empty_bar[s], 1);   // one multicast commit arrive per use
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 2 * TC_EPI_WARPS);  // leader: the epilogue warps of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm(tmem_ptr_smem, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers are initialised before any remote arrive / multicast commit / remote complete_tx
  tc_fence_after();
  const uint32_t tmem_base = uniform_u(*tmem_ptr_smem);
  pdl_wait();  // barriers initialised, tensor memory allocated: nothing above touched global memory

  if (warp == 0) {
    // ===================== TMA producer (both CTAs; converged warp, one elected lane issues) =====================
    int stage = 0;
    uint32_t phase = 0;
    long long w_empty = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      int cls, nt, ht, wt, cot;
      decode_tile(p, tile, cls, nt, ht, wt, cot);
      const int ntaps = p.taps.ntaps[cls];
      for (int t = 0; t < ntaps; ++t) {
        const CUtensorMap* amap = &maps.in[p.taps.view[cls][t]];
        const int cw = wt * p.BW + p.taps.dw[cls][t];
        const int ch = ht * p.BH + p.taps.dh[cls][t];
        const int cn = nt * p.BNI;
        const int wtap = p.taps.wtap[cls][t];
        const CUtensorMap* wmap = p.taps.wsel[cls][t] ? &maps.w2 : &maps.w;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          const long long t0 = p.dbg ? clock64() : 0;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (p.dbg) w_empty += clock64() - t0;   // producer waiting for a free slot
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          const uint32_t lead_full = mapa_u32(smem_u32(&full_bar[stage]), 0);
          if (elect_one_sync()) {
            if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
            tma_load_4d_2sm(sa, amap, lead_full, kc * KCHUNK, cw, ch, cn);
            tma_load_3d_2sm(sb, wmap, lead_full, kc * KCHUNK, wtap, cot * BLOCK_N + (int)rank * (BLOCK_N / 2));
          }
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    if (p.dbg && lane == 0) p.dbg[blockIdx.x * 8 + 0] += w_empty;
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only; converged warp, one elected lane issues) =====================
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc(2 * TILE_M, BLOCK_N, 0, 0);
      const uint32_t smem_base = smem_u32(smem);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      long long w_tempty = 0, w_full = 0, t_last = 0;
      const long long tstart = p.dbg ? clock64() : 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        if (p.dbg) t_last = clock64() - tstart;  // (time up to the start of the last tile)
        int cls, nt, ht, wt, cot;
        decode_tile(p, tile, cls, nt, ht, wt, cot);
        const int kiters = p.taps.ntaps[cls] * p.kchunks;
        if (kiters == 0) continue;
        long long t0 = p.dbg ? clock64() : 0;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        if (p.dbg) w_tempty += clock64() - t0;     // MMA warp waiting for the epilogue (accumulator buffer)
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int it = 0; it < kiters; ++it) {
          t0 = p.dbg ? clock64() : 0;
          mbar_wait(&full_bar[stage], phase);
          if (p.dbg) w_full += clock64() - t0;     // MMA warp waiting for operands (TMA)
          tc_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t sb = sa + Cfg::A_BYTES;
          const uint64_t adesc = make_smem_desc(sa, 16, 1024);
          const uint64_t bdesc = make_smem_desc(sb, 16, 1024);
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < KCHUNK / 16; ++k) umma_bf16_2sm(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (it | k) != 0);
            umma_commit_2sm(&empty_bar[stage]);  // frees the slot in both CTAs
          }
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
        if (elect_one_sync()) umma_commit_2sm(&tfull_bar[acc]);  // accumulator complete, both CTAs
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (p.dbg && lane == 0) {
        p.dbg[blockIdx.x * 8 + 1] += w_tempty;
        p.dbg[blockIdx.x * 8 + 2] += w_full;
        p.dbg[blockIdx.x * 8 + 3] = t_last;
      }
    }
  } else {
    tc_epilogue<BLOCK_N, MODE, true>(p, smem + Cfg::STAGES * Cfg::STAGE_BYTES + 256, tmem_base, tfull_bar, tempty_bar, warp, lane);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer may still be reading this CTA's shared memory / writing its tensor memory until here
  if (p.dbg && threadIdx.x == 0) {
    p.dbg[blockIdx.x * 8 + 6] += clock64() - dbg_c0;
    p.dbg[blockIdx.x * 8 + 7] += globaltimer_ns() - dbg_g0;
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------- 64 -> 64 channels, 3x3, stride 1
// The widest feature maps (32x32x64: PreAct layer1, the generator's outer convs) are bound by the L2 -> shared-memory
// operand stream of the generic kernel above, which reloads every input pixel once per filter tap and the whole filter
// once per tile (216 KB per 128-pixel tile).  Here
//   * all 9 taps of the 64x64 filter (72 KB) stay resident in shared memory for the lifetime of the CTA;
//   * a tile is BH full-width image rows; for each horizontal tap offset dw one TMA box of BH+2 rows is loaded, and the
//     three vertical taps read it at row offsets 0, W, 2W (whole 1024-byte swizzle atoms, so the shared-memory
//     descriptors stay aligned): 3 loads of (BH+2)*W pixels instead of 9 loads of BH*W -- 72 KB per tile.
struct Tc64Cfg {
  static constexpr int W_BYTES = 9 * 64 * KCHUNK * 2;  // 72 KB resident filter
  static constexpr int STAGES = 4;  // upper bound; the launch picks how many fit (n_stages)
  static constexpr int SMEM_BYTES_FIXED = W_BYTES + 1024 /*align*/ + 256 /*barriers*/ + TC_EPI_WARPS * 32 * EPI_ROWB;
};

template <int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc64_kernel(const __grid_constant__ TcMaps maps, const TcParams p,
                                                                  const int stage_bytes, const int n_stages) {
  pdl_launch_dependents();  // the next kernel may be scheduled; this one waits for its predecessor after its prologue
  const long long dbg_c0 = p.dbg ? clock64() : 0, dbg_g0 = p.dbg ? globaltimer_ns() : 0;  // whole-kernel cycles / ns of this CTA
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* wsm = smem;                               // [9 taps][64 co][64 ci] bf16, SWIZZLE_128B
  uint8_t* stages = smem + Tc64Cfg::W_BYTES;         // STAGES x stage_bytes ((BH+2)*BW pixels x 128 B, 1024-aligned)
  uint8_t* tail = stages + n_stages * stage_bytes;
  uint64_t* bars = (uint64_t*)tail;
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Tc64Cfg::STAGES;
  uint64_t* tfull_bar = bars + 2 * Tc64Cfg::STAGES;
  uint64_t* tempty_bar = bars + 2 * Tc64Cfg::STAGES + 2;
  uint64_t* w_bar = bars + 2 * Tc64Cfg::STAGES + 4;
  uint32_t* tmem_ptr_smem = (uint32_t*)(bars + 2 * Tc64Cfg::STAGES + 5);
  const int warp = uniform_i(threadIdx.x >> 5), lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.w);
    tma_prefetch_desc(&maps.in[0]);
    for (int s = 0; s < Tc64Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], TC_EPI_WARPS);
    }
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = uniform_u(*tmem_ptr_smem);
  pdl_wait();  // barriers initialised, tensor memory allocated: nothing above touched global memory

  if (warp == 0) {
    // TMA producer: converged warp, one elected lane issues
    if (elect_one_sync()) {
      mbar_expect_tx(w_bar, Tc64Cfg::W_BYTES);
      for (int t = 0; t < 9; ++t) tma_load_3d(wsm + t * 8192, &maps.w, w_bar, 0, t, 0);
    }
    int stage = 0;
    uint32_t phase = 0;
    long long w_empty = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      int cls, nt, ht, wt, cot;
      decode_tile(p, tile, cls, nt, ht, wt, cot);
      for (int dw = -1; dw <= 1; ++dw) {
        const long long t0 = p.dbg ? clock64() : 0;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (p.dbg) w_empty += clock64() - t0;
        if (elect_one_sync()) {
          mbar_expect_tx(&full_bar[stage], stage_bytes);
          tma_load_4d(stages + stage * stage_bytes, &maps.in[0], &full_bar[stage], 0, dw, ht * p.BH - 1, nt);
        }
        if (++stage == n_stages) { stage = 0; phase ^= 1; }
      }
      const int tile2 = tile + 2 * gridDim.x;   // epilogue inputs of the tile after next -> L2
      if (p.n_epf > 0 && tile2 < p.total_tiles) {
        decode_tile(p, tile2, cls, nt, ht, wt, cot);
        if (elect_one_sync()) {
          tma_prefetch_4d(&maps.e[0], 0, 0, ht * p.BH, nt);
          if (p.n_epf > 1) tma_prefetch_4d(&maps.e[1], 0, 0, ht * p.BH, nt);
        }
      }
    }
    if (p.dbg && lane == 0) p.dbg[blockIdx.x * 8 + 0] += w_empty;
  } else if (warp == 1) {
    // MMA issuer: converged warp, one elected lane issues
    constexpr uint32_t idesc = make_idesc(TILE_M, 64, 0, 0);
    long long w_tempty = 0, w_full = 0;
    mbar_wait(w_bar, 0);
    tc_fence_after();
    const uint32_t wbase = smem_u32(wsm);
    const uint32_t sbase = smem_u32(stages);
    const uint32_t row_step = (uint32_t)p.BW * 128;  // one image row of the staged box
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    const long long tstart = p.dbg ? clock64() : 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      long long t0 = p.dbg ? clock64() : 0;
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      if (p.dbg) w_tempty += clock64() - t0;
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * 64;
      for (int dwi = 0; dwi < 3; ++dwi) {
        t0 = p.dbg ? clock64() : 0;
        mbar_wait(&full_bar[stage], phase);
        if (p.dbg) w_full += clock64() - t0;
        tc_fence_after();
        const uint32_t sa = sbase + stage * stage_bytes;
        if (elect_one_sync()) {
#pragma unroll
          for (int dhi = 0; dhi < 3; ++dhi) {
            const uint64_t adesc = make_smem_desc(sa + dhi * row_step, 16, 1024);
            const uint64_t bdesc = make_smem_desc(wbase + (dhi * 3 + dwi) * 8192, 16, 1024);
#pragma unroll
            for (int k = 0; k < KCHUNK / 16; ++k) umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (dwi | dhi | k) != 0);
          }
          umma_commit(&empty_bar[stage]);
        }
        if (++stage == n_stages) { stage = 0; phase ^= 1; }
      }
      if (elect_one_sync()) umma_commit(&tfull_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (p.dbg && lane == 0) {
      p.dbg[blockIdx.x * 8 + 1] += w_tempty;
      p.dbg[blockIdx.x * 8 + 2] += w_full;
      p.dbg[blockIdx.x * 8 + 3] += clock64() - tstart;
    }
  } else {
    tc_epilogue<64, MODE>(p, tail + 256, tmem_base, tfull_bar, tempty_bar, warp, lane);
  }
  tc_fence_before();
  __syncthreads();
  if (p.dbg && threadIdx.x == 0) {
    p.dbg[blockIdx.x * 8 + 6] += clock64() - dbg_c0;
    p.dbg[blockIdx.x * 8 + 7] += globaltimer_ns() - dbg_g0;
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

// ---------------------------------------------------------------------------------------------- 128 output channels, 3x3, stride 1
// With the MMA issue fixed, the generic 128-wide kernel is bound by its operand stream (576 KB of TMA loads per 128-pixel tile:
// ~75 B/clk per SM is what L2 -> shared memory sustains with 160 KB in flight).  This kernel loads less:
//   * row re-use (as conv_tc64_kernel): per horizontal tap offset and K chunk ONE box of the tile's rows plus a halo row above and
//     below; the three vertical taps read it at row offsets 0, W, 2W;
//   * a work item is a SUPER-TILE of two vertically adjacent 128-pixel tiles (one (2 BH + 2)-row box serves both) that SHARE every
//     weight tile: each 16 KB weight tile is loaded once per 256 pixels and feeds two accumulators.
// Per 128 pixels (128 -> 128 @16x16): 108 KB of activations + 144 KB of weights instead of 288 + 288.  Four accumulator buffers
// (2 items x 2 tiles, 512 TMEM columns); separate rings for activation boxes and weight tiles.
struct TcRRCfg {
  static constexpr int B_BYTES = 128 * KCHUNK * 2;  // 16 KB weight tile (128 output channels x 64 input channels)
  static constexpr int SA = 3;                      // activation boxes in flight (upper bound)
  static constexpr int SB = 6;                      // weight tiles in flight (upper bound)
  static constexpr int SMEM_FIXED = 1024 /*align*/ + 256 /*barriers*/ + TC_EPI_WARPS * 32 * EPI_ROWB;
};

template <int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_rr_kernel(const __grid_constant__ TcMaps maps, const TcParams p, const int a_bytes,
                                                                   const int n_a, const int n_b) {
  pdl_launch_dependents();
  const long long dbg_c0 = p.dbg ? clock64() : 0, dbg_g0 = p.dbg ? globaltimer_ns() : 0;  // whole-kernel cycles / ns of this CTA
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* a_ring = smem;                              // n_a x a_bytes: (2 BH + 2) rows x BW pixels x 128 B, SWIZZLE_128B
  uint8_t* b_ring = smem + n_a * a_bytes;              // n_b x 16 KB
  uint8_t* tail = b_ring + n_b * TcRRCfg::B_BYTES;
  uint64_t* bars = (uint64_t*)tail;
  uint64_t* a_full = bars;
  uint64_t* a_empty = bars + TcRRCfg::SA;
  uint64_t* b_full = bars + 2 * TcRRCfg::SA;
  uint64_t* b_empty = bars + 2 * TcRRCfg::SA + TcRRCfg::SB;
  uint64_t* tfull_bar = bars + 2 * TcRRCfg::SA + 2 * TcRRCfg::SB;       // 4
  uint64_t* tempty_bar = tfull_bar + 4;                                 // 4
  uint32_t* tmem_ptr_smem = (uint32_t*)(tempty_bar + 4);
  const int warp = uniform_i(threadIdx.x >> 5), lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.w);
    tma_prefetch_desc(&maps.in[0]);
    for (int s = 0; s < TcRRCfg::SA; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < TcRRCfg::SB; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int a = 0; a < 4; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], TC_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = uniform_u(*tmem_ptr_smem);
  pdl_wait();

  const int supers_h = p.tiles_h >> 1;  // super-tiles per image
  if (warp == 0) {
    // ===================== TMA producer =====================
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    long long w_empty = 0;
    int st, jmask;
    for (int i = 0; rr_item(p, i, st, jmask); ++i) {
      const int n = st / supers_h, hs = st - n * supers_h;
      for (int dw = -1; dw <= 1; ++dw)
        for (int kc = 0; kc < p.kchunks; ++kc) {
          long long t0 = p.dbg ? clock64() : 0;
          mbar_wait(&a_empty[sa], pa ^ 1);
          if (p.dbg) w_empty += clock64() - t0;
          if (elect_one_sync()) {
            mbar_expect_tx(&a_full[sa], a_bytes);
            tma_load_4d(a_ring + sa * a_bytes, &maps.in[0], &a_full[sa], kc * KCHUNK, dw, hs * 2 * p.BH - 1, n);
          }
          if (++sa == n_a) { sa = 0; pa ^= 1; }
          for (int dh = 0; dh < 3; ++dh) {
            t0 = p.dbg ? clock64() : 0;
            mbar_wait(&b_empty[sb], pb ^ 1);
            if (p.dbg) w_empty += clock64() - t0;
            if (elect_one_sync()) {
              mbar_expect_tx(&b_full[sb], TcRRCfg::B_BYTES);
              tma_load_3d(b_ring + sb * TcRRCfg::B_BYTES, &maps.w, &b_full[sb], kc * KCHUNK, dh * 3 + dw + 1, 0);
            }
            if (++sb == n_b) { sb = 0; pb ^= 1; }
          }
        }
      int st2, jm2;
      if (p.n_epf > 0 && rr_item(p, i + 1, st2, jm2)) {   // epilogue inputs of the next item (two tiles, two 64-channel halves) -> L2
        const int n2 = st2 / supers_h, hs2 = st2 - n2 * supers_h;
        if (elect_one_sync()) {
          for (int e = 0; e < p.n_epf; ++e)
            for (int j = 0; j < 2; ++j)
              if (jm2 >> j & 1) {
                tma_prefetch_4d(&maps.e[e], 0, 0, (2 * hs2 + j) * p.BH, n2);
                tma_prefetch_4d(&maps.e[e], 64, 0, (2 * hs2 + j) * p.BH, n2);
              }
        }
      }
    }
    if (p.dbg && lane == 0) p.dbg[blockIdx.x * 8 + 0] += w_empty;
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc(TILE_M, 128, 0, 0);
    const uint32_t a_base = smem_u32(a_ring), b_base = smem_u32(b_ring);
    const uint32_t row_step = (uint32_t)p.BW * 128;   // one image row of the staged box
    const uint32_t half_step = (uint32_t)p.BH * row_step;
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    long long w_tempty = 0, w_full = 0, t_last = 0;
    const long long tstart = p.dbg ? clock64() : 0;
    int st, jmask;
    for (int i = 0; rr_item(p, i, st, jmask); ++i) {
      if (p.dbg) t_last = clock64() - tstart;
      const int buf = i & 1;
      const uint32_t ph = (i >> 1) & 1;
      long long t0 = p.dbg ? clock64() : 0;
      if (jmask & 1) mbar_wait(&tempty_bar[2 * buf], ph ^ 1);
      if (jmask & 2) mbar_wait(&tempty_bar[2 * buf + 1], ph ^ 1);
      if (p.dbg) w_tempty += clock64() - t0;
      tc_fence_after();
      const uint32_t d0 = tmem_base + (2 * buf) * 128, d1 = d0 + 128;
      bool first = true;
      for (int dwi = 0; dwi < 3; ++dwi)
        for (int kc = 0; kc < p.kchunks; ++kc) {
          t0 = p.dbg ? clock64() : 0;
          mbar_wait(&a_full[sa], pa);
          if (p.dbg) w_full += clock64() - t0;
          const uint32_t abox = a_base + sa * a_bytes;
          for (int dhi = 0; dhi < 3; ++dhi) {
            t0 = p.dbg ? clock64() : 0;
            mbar_wait(&b_full[sb], pb);
            if (p.dbg) w_full += clock64() - t0;
            tc_fence_after();
            const uint64_t bdesc = make_smem_desc(b_base + sb * TcRRCfg::B_BYTES, 16, 1024);
            const uint64_t adesc0 = make_smem_desc(abox + dhi * row_step, 16, 1024);
            const uint64_t adesc1 = make_smem_desc(abox + dhi * row_step + half_step, 16, 1024);
            if (elect_one_sync()) {
              if (jmask & 1) {
#pragma unroll
                for (int k = 0; k < KCHUNK / 16; ++k) umma_bf16(d0, adesc0 + 2 * k, bdesc + 2 * k, idesc, !(first && k == 0));
              }
              if (jmask & 2) {
#pragma unroll
                for (int k = 0; k < KCHUNK / 16; ++k) umma_bf16(d1, adesc1 + 2 * k, bdesc + 2 * k, idesc, !(first && k == 0));
              }
              umma_commit(&b_empty[sb]);
            }
            first = false;
            if (++sb == n_b) { sb = 0; pb ^= 1; }
          }
          if (elect_one_sync()) umma_commit(&a_empty[sa]);
          if (++sa == n_a) { sa = 0; pa ^= 1; }
        }
      if (elect_one_sync()) {
        if (jmask & 1) umma_commit(&tfull_bar[2 * buf]);
        if (jmask & 2) umma_commit(&tfull_bar[2 * buf + 1]);
      }
    }
    if (p.dbg && lane == 0) {
      p.dbg[blockIdx.x * 8 + 1] += w_tempty;
      p.dbg[blockIdx.x * 8 + 2] += w_full;
      p.dbg[blockIdx.x * 8 + 3] = t_last;
    }
  } else {
    tc_epilogue<128, MODE, false, 2>(p, tail + 256, tmem_base, tfull_bar, tempty_bar, warp, lane);
  }
  tc_fence_before();
  __syncthreads();
  if (p.dbg && threadIdx.x == 0) {
    p.dbg[blockIdx.x * 8 + 6] += clock64() - dbg_c0;
    p.dbg[blockIdx.x * 8 + 7] += globaltimer_ns() - dbg_g0;
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = (PFN_encodeTiled)ptr;
  }
  return fn;
}

// NHWC bf16 activation (possibly a parity view): dims (C, W, H, N), element strides given for w/h/n
static int make_act_map(CUtensorMap* m, const void* base, int C, int W, int H, int N, long long sw, long long sh, long long sn,
                        int BW, int BH, int BNI) {  // BH: rows of the box (BH+2 for the row-reuse kernel)
  PFN_encodeTiled enc = get_encode();
  if (!enc) return -1;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)sw * 2, (cuuint64_t)sh * 2, (cuuint64_t)sn * 2};
  cuuint32_t box[4] = {KCHUNK, (cuuint32_t)BW, (cuuint32_t)BH, (cuuint32_t)BNI};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -(int)r - 2000;
}

// weights [rows][taps][K] bf16 as (K, taps, rows); box (64, 1, box_rows)
static int make_w_map(CUtensorMap* m, const void* base, int K, int taps, int rows, int box_rows) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return -1;
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)taps, (cuuint64_t)rows};
  cuuint64_t strides[2] = {(cuuint64_t)K * 2, (cuuint64_t)K * taps * 2};
  cuuint32_t box[3] = {KCHUNK, 1, (cuuint32_t)box_rows};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -(int)r - 2000;
}

static int p2ceil(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

static void pick_box(int Hc, int Wc, int pixels, int* BW, int* BH, int* BNI) {
  int bw = p2ceil(Wc);
  if (bw > pixels) bw = pixels;
  int bh = p2ceil(Hc);
  if (bh > pixels / bw) bh = pixels / bw;
  *BW = bw;
  *BH = bh;
  *BNI = pixels / (bw * bh);
}

static int floordiv2(int v) { return v >= 0 ? v / 2 : -((-v + 1) / 2); }

// multiplier of the division-free tile decode (fdiv); 0 encodes d == 1
static uint32_t fdiv_magic(int d) { return d <= 1 ? 0u : (uint32_t)((1ull << 32) / (unsigned long long)d + 1ull); }
// fills the multipliers; false if the exactness condition x * d < 2^32 could fail for this problem
static bool set_fdiv(TcParams& p) {
  p.m_co = fdiv_magic(p.tiles_co);
  p.m_w = fdiv_magic(p.tiles_w);
  p.m_h = fdiv_magic(p.tiles_h);
  p.m_n = fdiv_magic(p.tiles_n);
  p.m_ppc = fdiv_magic(p.pairs_per_class);
  long long dmax = p.tiles_co;
  if (p.tiles_w > dmax) dmax = p.tiles_w;
  if (p.tiles_h > dmax) dmax = p.tiles_h;
  if (p.tiles_n > dmax) dmax = p.tiles_n;
  if (p.pairs_per_class > dmax) dmax = p.pairs_per_class;
  return (long long)(p.total_tiles > 0 ? p.total_tiles : 1) * dmax < (1LL << 32);
}

extern "C" int combat_conv_tc_supported(const combat_conv_tc_desc* d) {
  if (!d) return 0;
  if (d->Ci % 64 || d->Co % 64) return 0;
  if (d->KH != d->KW || (d->KH != 1 && d->KH != 3)) return 0;
  if (!((d->stride == 1 && d->up == 1) || (d->stride == 2 && d->up == 1) || (d->stride == 1 && d->up == 2))) return 0;
  if (d->stride == 2 && ((d->Hi | d->Wi) & 1)) return 0;
  if (d->up == 2 && ((d->Ho | d->Wo) & 1)) return 0;
  return get_encode() != nullptr;
}

static int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// clusters of two CTAs that can be resident at once (one CTA per SM: the two SMs of a TPC); queried once
static int max_pair_clusters() {
  static int n = 0;
  if (!n) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * num_sms());
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = TcPairCfg<256>::SMEM_BYTES;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2;
    attr.val.clusterDim.y = attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    cudaFuncSetAttribute(conv_tc_pair_kernel<256, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcPairCfg<256>::SMEM_BYTES);
    int q = 0;
    if (cudaOccupancyMaxActiveClusters(&q, conv_tc_pair_kernel<256, 0>, &cfg) != cudaSuccess || q <= 0) {
      cudaGetLastError();
      q = num_sms() / 2;
    }
    n = q < num_sms() / 2 ? q : num_sms() / 2;
  }
  return n;
}

static int g_last_grid = 0;
extern "C" int combat_conv_tc_last_grid(void) { return g_last_grid; }

template <int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_first_kernel(const __grid_constant__ TcMaps maps, const TcParams p);  // below

extern "C" int combat_conv_tc(const combat_conv_tc_desc* d, void* stream) {
  COMBAT_ARG(d && d->in && d->w && (d->out || d->out2), 0);
  COMBAT_ARG(combat_conv_tc_supported(d), 0);
  COMBAT_ARG((long long)d->N * d->Ho * d->Wo * d->Co < (1LL << 32), 0);  // 32-bit element offsets in the epilogue
  COMBAT_ARG(!d->out2 || (d->scale2 && d->shift2), 0);
  COMBAT_ARG(!d->mask || d->mask_scale, 0);
  // masked (backward) epilogue: bf16 in/out, at most one addend, no forward-only features
  COMBAT_ARG(!d->mask || (d->out && !d->out_f32 && !d->res_f32 && !(d->residual && d->post_add) && !d->out2 && !d->bias &&
                          !d->act && !d->post_scale), 0);
  COMBAT_ARG(d->mask || !d->post_add, 0);
  COMBAT_ARG(!d->in2 == !d->w2, 0);
  COMBAT_ARG(!d->in2 || (d->up == 2 && d->KH == 3 && d->pad == 1), 0);  // pad = k - 1 - pad of the forward conv (3x3, pad 1)
  TcParams p;
  memset(&p, 0, sizeof(p));
  TcMaps maps;
  memset(&maps, 0, sizeof(maps));
  const int KH = d->KH, KW = d->KW, pad = d->pad;
  p.N = d->N;
  p.out_H = d->Ho;
  p.out_W = d->Wo;
  p.Co = d->Co;
  p.kchunks = d->Ci / KCHUNK;
  p.out = d->out;
  p.residual = d->residual;
  p.bias = d->bias;
  p.out_f32 = d->out_f32;
  p.res_f32 = d->res_f32;
  p.act = d->act;
  p.post_scale = d->post_scale;
  p.post_shift = d->post_shift;
  p.out2 = d->out2;
  p.scale2 = d->scale2;
  p.shift2 = d->shift2;
  p.mask = (const bf16*)d->mask;
  p.mask_scale = d->mask_scale;
  p.post_add = (const bf16*)d->post_add;
  const bool bnb = d->bnb_x != nullptr;
  if (bnb) {
    // fused train-mode BatchNorm backward reduction: bf16 out, no other epilogue feature, partial sums wanted
    COMBAT_ARG(d->bnb_scale && d->bnb_shift && d->bnb_mean && d->bnb_invstd && d->stats && d->out && !d->out_f32, 0);
    COMBAT_ARG(!d->mask && !d->post_add && !d->residual && !d->out2 && !d->bias && !d->act && !d->post_scale, 0);
    p.mask = (const bf16*)d->bnb_x;
    p.mask_scale = d->bnb_scale;
    p.bnb_shift = d->bnb_shift;
    p.bnb_mean = d->bnb_mean;
    p.bnb_invstd = d->bnb_invstd;
  }
  p.dbg = (getenv("COMBAT_TC_DBG") && !d->bnb_x) ? (long long*)d->stats : nullptr;
  p.stats = p.dbg ? nullptr : d->stats;
  if (d->in_nchw3) {
    // first conv: float32 NCHW image, operand built in shared memory by the epilogue warps (conv_tc_first_kernel)
    COMBAT_ARG(d->Co == 64 && d->Ci == 64 && KH == 3 && pad == 1 && d->stride == 1 && d->up == 1 && d->Ho == d->Hi && d->Wo == d->Wi, 0);
    COMBAT_ARG(!d->mask && !d->residual && !d->in2 && !bnb && !d->post_add, 0);
    p.n_classes = 1;
    p.out_scale = 1;
    p.Hc = d->Ho;
    p.Wc = d->Wo;
    pick_box(p.Hc, p.Wc, TILE_M, &p.BW, &p.BH, &p.BNI);
    COMBAT_ARG(p.BW == d->Wo && p.BNI == 1, 0);   // tiles of whole image rows (W a power of two <= 128)
    p.taps.ntaps[0] = 1;
    p.tiles_co = 1;
    p.tiles_w = 1;
    p.tiles_h = cdiv(p.Hc, p.BH);
    p.tiles_n = d->N;
    p.total_tiles = p.tiles_h * p.tiles_n;
    while ((1 << p.lbw) < p.BW) ++p.lbw;
    while ((1 << p.lbh) < p.BH) ++p.lbh;
    p.img = (const float*)d->in;
    COMBAT_ARG(set_fdiv(p), 0);
    int rc1 = make_w_map(&maps.w, d->w, 64, 1, 64, 64);
    if (rc1) return rc1;
    const int grid1 = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
    g_last_grid = grid1;
    cudaStream_t st1 = (cudaStream_t)stream;
    if (p.stats) {
      cudaMemsetAsync(p.stats, 0, (size_t)grid1 * 4 * 2 * d->Co * sizeof(float), st1);
      g_last_grid = grid1 * 4;
    }
    const int smem1 = 8192 + 2 * 16384 + 256 + TC_EPI_WARPS * 32 * EPI_ROWB + 1024;
    if (p.stats) {
      cudaFuncSetAttribute(conv_tc_first_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1);
      pdl_launch(conv_tc_first_kernel<2>, grid1, TC_THREADS, smem1, st1, maps, p);
    } else {
      cudaFuncSetAttribute(conv_tc_first_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1);
      pdl_launch(conv_tc_first_kernel<0>, grid1, TC_THREADS, smem1, st1, maps, p);
    }
    COMBAT_RETURN_LAUNCH("conv_tc_first");
  }
  // 256-wide tiles for the deep layers: per MMA the 128-row A operand is read once for 256 instead of 128 output channels
  // (the shared-memory operand path is what bounds the 128-wide kernel); they need K large enough to hide the epilogue
  const bool wide = d->Co % 256 == 0 && d->Ci * d->KH * d->KW >= 1152 && !getenv("COMBAT_NO_BN256");
  const int BLOCK_N = wide ? 256 : ((d->Co % 128 == 0) ? 128 : 64);
  p.tiles_co = d->Co / BLOCK_N;
  int rc;
  if (d->up == 1) {
    p.n_classes = 1;
    p.out_scale = 1;
    p.Hc = d->Ho;
    p.Wc = d->Wo;
    pick_box(p.Hc, p.Wc, TILE_M, &p.BW, &p.BH, &p.BNI);
    const int s = d->stride;
    int nt = 0;
    for (int kh = 0; kh < KH; ++kh)
      for (int kw = 0; kw < KW; ++kw) {
        const int oh = kh - pad, ow = kw - pad;
        int ph = 0, pw = 0, dh = oh, dw = ow;
        if (s == 2) {
          ph = ((oh % 2) + 2) % 2;
          pw = ((ow % 2) + 2) % 2;
          dh = floordiv2(oh);
          dw = floordiv2(ow);
        }
        p.taps.view[0][nt] = (signed char)(ph * 2 + pw);
        p.taps.dh[0][nt] = (signed char)dh;
        p.taps.dw[0][nt] = (signed char)dw;
        p.taps.wtap[0][nt] = (signed char)(kh * KW + kw);
        ++nt;
      }
    p.taps.ntaps[0] = nt;
    const long long C = d->Ci, W = d->Wi, H = d->Hi;
    if (s == 1) {
      rc = make_act_map(&maps.in[0], d->in, d->Ci, d->Wi, d->Hi, d->N, C, W * C, H * W * C, p.BW, p.BH, p.BNI);
      if (rc) return rc;
    } else {
      for (int ph = 0; ph < 2; ++ph)
        for (int pw = 0; pw < 2; ++pw) {
          const bf16* base = (const bf16*)d->in + ((long long)ph * W + pw) * C;
          rc = make_act_map(&maps.in[ph * 2 + pw], base, d->Ci, (d->Wi - pw + 1) / 2, (d->Hi - ph + 1) / 2, d->N, 2 * C,
                            2 * W * C, H * W * C, p.BW, p.BH, p.BNI);
          if (rc) return rc;
        }
    }
  } else {
    // input gradient of a stride-2 conv: one stride-1 problem per output parity class
    p.n_classes = 4;
    p.out_scale = 2;
    p.Hc = d->Ho / 2;
    p.Wc = d->Wo / 2;
    pick_box(p.Hc, p.Wc, TILE_M, &p.BW, &p.BH, &p.BNI);
    for (int cp = 0; cp < 2; ++cp)
      for (int cq = 0; cq < 2; ++cq) {
        const int cls = cp * 2 + cq;
        p.taps.cls_p[cls] = cp;
        p.taps.cls_q[cls] = cq;
        int nt = 0;
        for (int kh = 0; kh < KH; ++kh)
          for (int kw = 0; kw < KW; ++kw) {
            const int nh = cp - pad + kh, nw = cq - pad + kw;  // (2a + p - pad + kh) / 2 = a + nh / 2
            if ((nh & 1) || (nw & 1)) continue;
            p.taps.view[cls][nt] = 0;
            p.taps.dh[cls][nt] = (signed char)(nh / 2);
            p.taps.dw[cls][nt] = (signed char)(nw / 2);
            p.taps.wtap[cls][nt] = (signed char)(kh * KW + kw);
            ++nt;
          }
        if (cls == 0 && d->in2) {  // fused 1x1 stride-2 shortcut gradient: dx[2a, 2b] += w2^T in2[a, b]
          p.taps.view[cls][nt] = 1;
          p.taps.dh[cls][nt] = 0;
          p.taps.dw[cls][nt] = 0;
          p.taps.wtap[cls][nt] = 0;
          p.taps.wsel[cls][nt] = 1;
          ++nt;
        }
        p.taps.ntaps[cls] = nt;
      }
    const long long C = d->Ci, W = d->Wi, H = d->Hi;
    rc = make_act_map(&maps.in[0], d->in, d->Ci, d->Wi, d->Hi, d->N, C, W * C, H * W * C, p.BW, p.BH, p.BNI);
    if (rc) return rc;
    if (d->in2) {
      rc = make_act_map(&maps.in[1], d->in2, d->Ci, d->Wi, d->Hi, d->N, C, W * C, H * W * C, p.BW, p.BH, p.BNI);
      if (rc) return rc;
    }
  }
  // 64 -> 64, 3x3, stride 1, tiles of whole image rows (8 | W so that a row is a whole number of swizzle atoms)
  const int stage_bytes64 = (p.BH + 2) * p.BW * 128;
  int n_stages64 = (232448 - Tc64Cfg::SMEM_BYTES_FIXED) / stage_bytes64;
  if (n_stages64 > Tc64Cfg::STAGES) n_stages64 = Tc64Cfg::STAGES;
  const bool use64 = d->Ci == 64 && d->Co == 64 && KH == 3 && pad == 1 && d->stride == 1 && d->up == 1 && p.BW == d->Wo &&
                     p.BNI == 1 && (p.BW % 8) == 0 && n_stages64 >= 2 && !getenv("COMBAT_NO_TC64");
  if (use64) {
    const long long C = d->Ci, W = d->Wi, H = d->Hi;
    rc = make_act_map(&maps.in[0], d->in, d->Ci, d->Wi, d->Hi, d->N, C, W * C, H * W * C, p.BW, p.BH + 2, 1);
    if (rc) return rc;
  }
  // 128 output channels, 3x3, stride 1, tiles of whole image rows: row re-use + two tiles per weight tile (conv_tc_rr_kernel)
  int rr_a_bytes = 0, rr_na = 0, rr_nb = 0;
  bool use_rr = !use64 && d->Co == 128 && KH == 3 && pad == 1 && d->stride == 1 && d->up == 1 && !d->in2 && p.BW == d->Wo &&
                p.BNI == 1 && (p.BW % 8) == 0 && (p.Hc % (2 * p.BH)) == 0 && !getenv("COMBAT_NO_RR");
  if (use_rr) {
    rr_a_bytes = (2 * p.BH + 2) * p.BW * 128;
    const int budget = 232448 - TcRRCfg::SMEM_FIXED;
    rr_na = TcRRCfg::SA;
    rr_nb = (budget - rr_na * rr_a_bytes) / TcRRCfg::B_BYTES;
    if (rr_nb < 3) {
      rr_na = 2;
      rr_nb = (budget - rr_na * rr_a_bytes) / TcRRCfg::B_BYTES;
    }
    if (rr_nb > TcRRCfg::SB) rr_nb = TcRRCfg::SB;
    if (rr_nb < 3) use_rr = false;
  }
  if (use_rr) {
    const long long C = d->Ci, W = d->Wi, H = d->Hi;
    rc = make_act_map(&maps.in[0], d->in, d->Ci, d->Wi, d->Hi, d->N, C, W * C, H * W * C, p.BW, 2 * p.BH + 2, 1);
    if (rc) return rc;
  }
  // CTA pairs (cta_group::2): each CTA of a pair stages half of the weight tile.  Measured inside the step (B200, batch 512):
  // 256-wide tiles gain 13-15 % (256->256 @8x8 fwd 1042 -> 1204 TFLOP/s, 512->512 @4x4 1107 -> 1272, their input gradients 900 ->
  // 1005); 128-wide tiles LOSE 6-9 % (half-size MMAs, twice the cluster-scope barrier round trips per byte) and keep the
  // single-CTA kernel unless COMBAT_PAIR128 is set.
  const bool pair = !use64 && !use_rr && (BLOCK_N == 256 || (BLOCK_N == 128 && getenv("COMBAT_PAIR128"))) && !getenv("COMBAT_NO_PAIR");
  rc = make_w_map(&maps.w, d->w, d->Ci, KH * KW, d->Co, pair ? BLOCK_N / 2 : BLOCK_N);
  if (rc) return rc;
  if (d->in2) {
    rc = make_w_map(&maps.w2, d->w2, d->Ci, 1, d->Co, pair ? BLOCK_N / 2 : BLOCK_N);
    if (rc) return rc;
  }
  // measured (B200, batch 512): masked 64 -> 64 input gradient 75.3 -> 69.1 us; the 128-channel row-reuse kernel loses 1 us with
  // it (its epilogue is not the bound), so only the 64-channel kernel prefetches unless COMBAT_EPF_RR is set
  if ((use64 || (use_rr && getenv("COMBAT_EPF_RR"))) && !getenv("COMBAT_NO_EPF")) {
    // epilogue-input prefetch (bf16 tensors shaped like the output, whole-row tiles of one image)
    const void* ein[2] = {nullptr, nullptr};
    int ne = 0;
    if (d->bnb_x) ein[ne++] = d->bnb_x;
    if (d->mask) ein[ne++] = d->mask;
    if (d->residual && !d->res_f32) ein[ne++] = d->residual;
    if (d->post_add && ne < 2) ein[ne++] = d->post_add;
    for (int e = 0; e < ne; ++e) {
      const long long Cc = d->Co;
      rc = make_act_map(&maps.e[e], ein[e], d->Co, d->Wo, d->Ho, d->N, Cc, (long long)d->Wo * Cc, (long long)d->Ho * d->Wo * Cc, p.BW,
                        p.BH, 1);
      if (rc) return rc;
    }
    p.n_epf = ne;
  }
  p.lbw = 0;
  while ((1 << p.lbw) < p.BW) ++p.lbw;
  p.lbh = 0;
  while ((1 << p.lbh) < p.BH) ++p.lbh;
  p.tiles_w = cdiv(p.Wc, p.BW);
  p.tiles_h = cdiv(p.Hc, p.BH);
  p.tiles_n = cdiv(p.N, p.BNI);
  p.total_tiles = p.n_classes * p.tiles_n * p.tiles_h * p.tiles_w * p.tiles_co;
  int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  if (pair) {
    p.pair_mode = 1;
    p.px_tiles_per_class = p.tiles_n * p.tiles_h * p.tiles_w;
    p.pairs_per_class = (p.px_tiles_per_class + 1) / 2;
    const int total_pairs = p.n_classes * p.pairs_per_class * p.tiles_co;
    p.total_items = p.total_tiles = 2 * total_pairs;   // the kernels' tile loops run over work items v = 2 * pair + cta_rank
    int nclusters = total_pairs < max_pair_clusters() ? total_pairs : max_pair_clusters();
    if (d->stats && nclusters >= p.tiles_co) nclusters -= nclusters % p.tiles_co;  // a CTA's channel tile must never change
    grid = 2 * nclusters;
  }
  if (use_rr) {
    const int n_super = p.total_tiles / 2;   // tiles_co == 1, tiles_w == 1, tiles_h even: tiles 2s and 2s+1 are one image's neighbours
    grid = n_super < num_sms() ? n_super : num_sms();
    p.rr_rounds = n_super / grid;
    p.rr_rem = n_super - p.rr_rounds * grid;
    p.rr_split = p.rr_rem > 0 && 2 * p.rr_rem <= grid;
  }
  g_last_grid = grid;
  COMBAT_ARG(set_fdiv(p), 0);
  COMBAT_ARG(!p.stats || (d->Co <= 512 && !d->mask), 0);
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH_P(BN, MD)                                                                                                        \
  {                                                                                                                             \
    cudaFuncSetAttribute(conv_tc_pair_kernel<BN, MD>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcPairCfg<BN>::SMEM_BYTES);  \
    pdl_launch(conv_tc_pair_kernel<BN, MD>, grid, TC_THREADS, TcPairCfg<BN>::SMEM_BYTES, st, maps, p);                                  \
  }
#define LAUNCH_C(BN, MD)                                                                                                  \
  {                                                                                                                       \
    cudaFuncSetAttribute(conv_tc_kernel<BN, MD>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<BN>::SMEM_BYTES);     \
    pdl_launch(conv_tc_kernel<BN, MD>, grid, TC_THREADS, TcCfg<BN>::SMEM_BYTES, st, maps, p);                                     \
  }
  const int mode = bnb ? 3 : (d->mask ? 1 : (p.stats ? 2 : 0));
  if (p.stats) {  // partial blocks of (CTA, quarter): zero-filled, every warp writes only its own column range
    COMBAT_ARG(grid % p.tiles_co == 0, 0);
    cudaMemsetAsync(p.stats, 0, (size_t)grid * 4 * 2 * d->Co * sizeof(float), st);
    g_last_grid = grid * 4;
  }
  if (use64) {
    const int smem_bytes = Tc64Cfg::SMEM_BYTES_FIXED + n_stages64 * stage_bytes64;
    if (mode == 3) {
      cudaFuncSetAttribute(conv_tc64_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
      pdl_launch(conv_tc64_kernel<3>, grid, TC_THREADS, smem_bytes, st, maps, p, stage_bytes64, n_stages64);
    } else if (mode == 1) {
      cudaFuncSetAttribute(conv_tc64_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
      pdl_launch(conv_tc64_kernel<1>, grid, TC_THREADS, smem_bytes, st, maps, p, stage_bytes64, n_stages64);
    } else if (mode == 2) {
      cudaFuncSetAttribute(conv_tc64_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
      pdl_launch(conv_tc64_kernel<2>, grid, TC_THREADS, smem_bytes, st, maps, p, stage_bytes64, n_stages64);
    } else {
      cudaFuncSetAttribute(conv_tc64_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
      pdl_launch(conv_tc64_kernel<0>, grid, TC_THREADS, smem_bytes, st, maps, p, stage_bytes64, n_stages64);
    }
    COMBAT_RETURN_LAUNCH("conv_tc64");
  }
  if (use_rr) {
    const int smem_bytes = TcRRCfg::SMEM_FIXED + rr_na * rr_a_bytes + rr_nb * TcRRCfg::B_BYTES;
    if (mode == 3) {
      cudaFuncSetAttribute(conv_tc_rr_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
      pdl_launch(conv_tc_rr_kernel<3>, grid, TC_THREADS, smem_bytes, st, maps, p, rr_a_bytes, rr_na, rr_nb);
    } else if (mode == 1) {
      cudaFuncSetAttribute(conv_tc_rr_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
      pdl_launch(conv_tc_rr_kernel<1>, grid, TC_THREADS, smem_bytes, st, maps, p, rr_a_bytes, rr_na, rr_nb);
    } else if (mode == 2) {
      cudaFuncSetAttribute(conv_tc_rr_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
      pdl_launch(conv_tc_rr_kernel<2>, grid, TC_THREADS, smem_bytes, st, maps, p, rr_a_bytes, rr_na, rr_nb);
    } else {
      cudaFuncSetAttribute(conv_tc_rr_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
      pdl_launch(conv_tc_rr_kernel<0>, grid, TC_THREADS, smem_bytes, st, maps, p, rr_a_bytes, rr_na, rr_nb);
    }
    COMBAT_RETURN_LAUNCH("conv_tc_rr");
  }
  if (pair) {
    if (BLOCK_N == 256) { if (mode == 3) LAUNCH_P(256, 3) else if (mode == 1) LAUNCH_P(256, 1) else if (mode == 2) LAUNCH_P(256, 2) else LAUNCH_P(256, 0) }
    else { if (mode == 3) LAUNCH_P(128, 3) else if (mode == 1) LAUNCH_P(128, 1) else if (mode == 2) LAUNCH_P(128, 2) else LAUNCH_P(128, 0) }
    COMBAT_RETURN_LAUNCH("conv_tc_pair");
  }
  if (BLOCK_N == 256) { if (mode == 3) LAUNCH_C(256, 3) else if (mode == 1) LAUNCH_C(256, 1) else if (mode == 2) LAUNCH_C(256, 2) else LAUNCH_C(256, 0) }
  else if (BLOCK_N == 128) { if (mode == 3) LAUNCH_C(128, 3) else if (mode == 1) LAUNCH_C(128, 1) else if (mode == 2) LAUNCH_C(128, 2) else LAUNCH_C(128, 0) }
  else { if (mode == 3) LAUNCH_C(64, 3) else if (mode == 1) LAUNCH_C(64, 1) else if (mode == 2) LAUNCH_C(64, 2) else LAUNCH_C(64, 0) }
#undef LAUNCH_C
#undef LAUNCH_P
  COMBAT_RETURN_LAUNCH("conv_tc");
}

// ---------------------------------------------------------------------------------------------- 3 -> 64 channels, 3x3, stride 1
// The first conv of the classifiers / the detector (preact_resnet.py:77, resnet.py:73, model.py:12) read a float32 NCHW image:
// K = 27 cannot be fetched by TMA as a K-major operand, and on CUDA cores the layer was FP32-FMA bound (conv_cin3_k: 115 us for a
// 1024 x 32 x 32 batch, 6 % of the step).  Here the eight epilogue warps BUILD the A operand in shared memory -- 128 pixels x 64
// bf16 columns [hi(27) | 0 | lo(27) | 0] with hi = bf16(x), lo = bf16(x - hi) (the image enters with ~16 mantissa bits; the filter
// is stored twice, NetBase._w64_for) in the SWIZZLE_128B K-major layout TMA would have produced (16-byte chunk index XOR row & 7),
// publish it to the async proxy (fence.proxy.async + mbarrier) and the MMA warp issues four N = 64 MMAs per tile.  Each warp builds
// the operand of tile k + 1 and then runs the ordinary epilogue (all modes: fused eval BatchNorm+ReLU second output, train-mode
// statistics, ELU + affine of the detector) of tile k, so the kernel is bound by its output stream.
template <int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_first_kernel(const __grid_constant__ TcMaps maps, const TcParams p) {
  pdl_launch_dependents();
  const long long dbg_c0 = p.dbg ? clock64() : 0, dbg_g0 = p.dbg ? globaltimer_ns() : 0;  // whole-kernel cycles / ns of this CTA
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* wsm = smem;                 // [64 co][64 k] bf16, SWIZZLE_128B: 8 KB
  uint8_t* abuf = smem + 8192;         // 2 x [128 pixels][64 k] bf16: 2 x 16 KB
  uint8_t* tail = abuf + 2 * 16384;
  uint64_t* bars = (uint64_t*)tail;
  uint64_t* a_full = bars;             // 2: one arrive per building warp
  uint64_t* a_empty = bars + 2;        // 2: MMA commit
  uint64_t* tfull_bar = bars + 4;
  uint64_t* tempty_bar = bars + 6;
  uint64_t* w_bar = bars + 8;
  uint32_t* tmem_ptr_smem = (uint32_t*)(bars + 9);
  const int warp = uniform_i(threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.w);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&a_full[s], TC_EPI_WARPS);
      mbar_init(&a_empty[s], 1);
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], TC_EPI_WARPS);
    }
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = uniform_u(*tmem_ptr_smem);
  pdl_wait();

  if (warp == 0) {
    if (elect_one_sync()) {
      mbar_expect_tx(w_bar, 8192);
      tma_load_3d(wsm, &maps.w, w_bar, 0, 0, 0);
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc(TILE_M, 64, 0, 0);
    mbar_wait(w_bar, 0);
    const uint32_t wbase = smem_u32(wsm), abase = smem_u32(abuf);
    int k = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++k) {
      const int s = k & 1;
      const uint32_t ph = (k >> 1) & 1;
      mbar_wait(&tempty_bar[s], ph ^ 1);
      mbar_wait(&a_full[s], ph);
      tc_fence_after();
      const uint64_t adesc = make_smem_desc(abase + s * 16384, 16, 1024);
      const uint64_t bdesc = make_smem_desc(wbase, 16, 1024);
      if (elect_one_sync()) {
#pragma unroll
        for (int kk = 0; kk < KCHUNK / 16; ++kk) umma_bf16(tmem_base + s * 64, adesc + 2 * kk, bdesc + 2 * kk, idesc, kk != 0);
        umma_commit(&a_empty[s]);
        umma_commit(&tfull_bar[s]);
      }
    }
  } else {
    // builder of tile index kb (0-based position in this CTA's tile sequence): thread -> (pixel row, K half).  Warps 2..5 build
    // k = 0..15, warps 6..9 k = 16..31 (k = tap * 3 + ci < 27, zero above) of BOTH the hi and the lo columns: every image value
    // is loaded once; taps and channels are compile-time constants inside each branch.
    const int eidx = (warp - 2) * 32 + lane;
    const int row = eidx & 127;
    const int khalf = uniform_i(eidx >> 7);
    const int hi_r = row >> p.lbw, wi = row & (p.BW - 1);
    const long long HW = (long long)p.Hc * p.Wc;
    auto build = [&](int kb) {
      const int tile = blockIdx.x + kb * gridDim.x;
      if (tile >= p.total_tiles) return;
      const int s = kb & 1;
      const bool dbg_b = p.dbg && warp == 2 && lane == 0;
      const long long tb0 = dbg_b ? clock64() : 0;
      mbar_wait(&a_empty[s], ((kb >> 1) & 1) ^ 1);
      const long long tb1 = dbg_b ? clock64() : 0;
      const int ht = tile % p.tiles_h, n = tile / p.tiles_h;
      const int h = ht * p.BH + hi_r;
      // UNCONDITIONAL loads from clamped coordinates, zeroed afterwards: a predicated load into a pre-zeroed register is compiled to
      // "load, wait, select" per value and serialises the whole batch (first version: ~3000 cycles per tile in this lambda)
      const float* xn = p.img + (long long)n * 3 * HW;
      int roff[3], coff[3];
      bool vh[3], vw[3];
#pragma unroll
      for (int dd = 0; dd < 3; ++dd) {
        const int ih = h + dd - 1, iw = wi + dd - 1;
        vh[dd] = ih >= 0 && ih < p.Hc;
        vw[dd] = iw >= 0 && iw < p.Wc;
        roff[dd] = min(max(ih, 0), p.Hc - 1) * p.Wc;
        coff[dd] = min(max(iw, 0), p.Wc - 1);
      }
      uint32_t hi_pk[8], lo_pk[8];
      auto gather = [&](auto KH) {
        constexpr int K0 = decltype(KH)::value * 16;
        float raw[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int kq = K0 + j;
          if (kq < 27) {
            const int tap = kq / 3, ci = kq - tap * 3, dh = tap / 3, dw = tap - dh * 3;
            raw[j] = __ldg(xn + ci * HW + roff[dh] + coff[dw]);
          } else {
            raw[j] = 0.f;
          }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float v2[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int kq = K0 + 2 * q + e;
            float v = 0.f;
            if (kq < 27) {
              const int tap = kq / 3, dh = tap / 3, dw = tap - dh * 3;
              v = (vh[dh] && vw[dw]) ? raw[2 * q + e] : 0.f;
            }
            v2[e] = v;
          }
          const __nv_bfloat162 hv = __floats2bfloat162_rn(v2[0], v2[1]);
          const float2 hf = __bfloat1622float2(hv);
          const __nv_bfloat162 lv = __floats2bfloat162_rn(v2[0] - hf.x, v2[1] - hf.y);
          hi_pk[q] = *(const uint32_t*)&hv;
          lo_pk[q] = *(const uint32_t*)&lv;
        }
      };
      if (khalf == 0) gather(std::integral_constant<int, 0>());
      else gather(std::integral_constant<int, 1>());
      const long long tb2 = dbg_b ? clock64() : 0;
      const uint32_t dst = smem_u32(abuf + s * 16384) + row * 128;
      const int sw = row & 7;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        sts128(dst + (((2 * khalf + c) ^ sw) << 4), hi_pk[4 * c], hi_pk[4 * c + 1], hi_pk[4 * c + 2], hi_pk[4 * c + 3]);
        sts128(dst + (((4 + 2 * khalf + c) ^ sw) << 4), lo_pk[4 * c], lo_pk[4 * c + 1], lo_pk[4 * c + 2], lo_pk[4 * c + 3]);
      }
      fence_proxy_async();   // generic-proxy writes -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full[s]);
      if (dbg_b) {   // builder: wait for the free operand buffer | gather + split | store + publish
        p.dbg[blockIdx.x * 8 + 0] += tb1 - tb0;
        p.dbg[blockIdx.x * 8 + 1] += tb2 - tb1;
        p.dbg[blockIdx.x * 8 + 2] += clock64() - tb2;
      }
    };
    build(0);
    tc_epilogue<64, MODE>(p, tail + 256, tmem_base, tfull_bar, tempty_bar, warp, lane, [&](int k) { build(k + 1); });
  }
  tc_fence_before();
  __syncthreads();
  if (p.dbg && threadIdx.x == 0) {
    p.dbg[blockIdx.x * 8 + 6] += clock64() - dbg_c0;
    p.dbg[blockIdx.x * 8 + 7] += globaltimer_ns() - dbg_g0;
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

// ---------------------------------------------------------------------------------------------- 64 -> 3 channels, 3x3, stride 1
// The image-side convs with 3 OUTPUT channels (the generator's last conv and the input gradient of the classifiers' first conv:
// networks/models.py:316, preact_resnet.py:77 backward) ran on CUDA cores at ~106 us per 512 x 32 x 32 batch (FP32-FMA bound,
// 2.8 % of the step).  On the tensor pipe they are pure data movement: conv_tc64_kernel's mainloop (resident filter, one
// (BH+2)-row box per horizontal tap, vertical taps as row offsets) with N = 16 accumulator columns (3 used; filter rows 3..15 are
// TMA zero fill), and an epilogue that needs NO transposition: a thread owns a pixel, consecutive lanes are consecutive pixels of
// an image row, so the three NCHW float32 planes are written with coalesced 128-byte stores straight from tcgen05.ld.
struct TcO3Params {
  int N, H, W;
  int BW, BH, tiles_h, total_tiles;
  int act;             // 0 none, 1 tanh
  const float* bias;   // [3] or NULL
  float* out;          // NCHW float32 [N,3,H,W]
};
struct TcO3Maps {
  CUtensorMap in;  // box (64, BW, BH+2, 1)
  CUtensorMap w;   // [3 rows][9 taps][64]: box (64, 1, 16)
};
#define O3_STAGES 6

__global__ void __launch_bounds__(192, 1) conv_tc_cout3_kernel(const __grid_constant__ TcO3Maps maps, const TcO3Params p, const int stage_bytes,
                                                               const int n_stages) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* wsm = smem;                      // [9 taps][16 rows][64 ci] bf16, SWIZZLE_128B (2 KB per tap)
  uint8_t* stages = smem + 9 * 2048;        // 18 KB is a multiple of 1024
  uint64_t* bars = (uint64_t*)(stages + n_stages * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + O3_STAGES;
  uint64_t* tfull_bar = bars + 2 * O3_STAGES;
  uint64_t* tempty_bar = bars + 2 * O3_STAGES + 2;
  uint64_t* w_bar = bars + 2 * O3_STAGES + 4;
  uint32_t* tmem_ptr_smem = (uint32_t*)(bars + 2 * O3_STAGES + 5);
  const int warp = uniform_i(threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.w);
    tma_prefetch_desc(&maps.in);
    for (int s = 0; s < O3_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 4);
    }
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = uniform_u(*tmem_ptr_smem);
  pdl_wait();

  if (warp == 0) {
    if (elect_one_sync()) {
      mbar_expect_tx(w_bar, 9 * 2048);
      for (int t = 0; t < 9; ++t) tma_load_3d(wsm + t * 2048, &maps.w, w_bar, 0, t, 0);
    }
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int ht = tile % p.tiles_h, n = tile / p.tiles_h;
      for (int dw = -1; dw <= 1; ++dw) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one_sync()) {
          mbar_expect_tx(&full_bar[stage], stage_bytes);
          tma_load_4d(stages + stage * stage_bytes, &maps.in, &full_bar[stage], 0, dw, ht * p.BH - 1, n);
        }
        if (++stage == n_stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc(TILE_M, 16, 0, 0);
    mbar_wait(w_bar, 0);
    tc_fence_after();
    const uint32_t wbase = smem_u32(wsm), sbase = smem_u32(stages);
    const uint32_t row_step = (uint32_t)p.BW * 128;
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * 16;
      for (int dwi = 0; dwi < 3; ++dwi) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = sbase + stage * stage_bytes;
        if (elect_one_sync()) {
#pragma unroll
          for (int dhi = 0; dhi < 3; ++dhi) {
            const uint64_t adesc = make_smem_desc(sa + dhi * row_step, 16, 1024);
            const uint64_t bdesc = make_smem_desc(wbase + (dhi * 3 + dwi) * 2048, 16, 1024);
#pragma unroll
            for (int k = 0; k < KCHUNK / 16; ++k) umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (dwi | dhi | k) != 0);
          }
          umma_commit(&empty_bar[stage]);
        }
        if (++stage == n_stages) { stage = 0; phase ^= 1; }
      }
      if (elect_one_sync()) umma_commit(&tfull_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // epilogue: warp w owns TMEM lanes [32 * (w % 4), +32) = 32 consecutive pixels of the tile
    const int quarter = warp & 3;
    const float b0 = p.bias ? p.bias[0] : 0.f, b1 = p.bias ? p.bias[1] : 0.f, b2 = p.bias ? p.bias[2] : 0.f;
    const long long HW = (long long)p.H * p.W;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int ht = tile % p.tiles_h, n = tile / p.tiles_h;
      const int row = quarter * 32 + lane;
      const int hi = row / p.BW, wi = row - hi * p.BW;
      const int h = ht * p.BH + hi;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      uint32_t v[16];
      tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * 16, v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      if (h < p.H) {
        float y0 = __uint_as_float(v[0]) + b0, y1 = __uint_as_float(v[1]) + b1, y2 = __uint_as_float(v[2]) + b2;
        if (p.act == 1) { y0 = tanhf(y0); y1 = tanhf(y1); y2 = tanhf(y2); }
        float* o = p.out + (long long)n * 3 * HW + (long long)h * p.W + wi;
        o[0] = y0;
        o[HW] = y1;
        o[2 * HW] = y2;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 32);
  }
}

extern "C" int combat_conv_tc_cout3_supported(int N, int H, int W) {
  int bw, bh, bni;
  if (N <= 0 || H <= 0 || W <= 0) return 0;
  pick_box(H, W, TILE_M, &bw, &bh, &bni);
  const int stage_bytes = (bh + 2) * bw * 128;
  return bw == W && bni == 1 && (bw % 8) == 0 && 2 * stage_bytes + 9 * 2048 + 1024 + 256 <= 232448;
}

extern "C" int combat_conv_tc_cout3(const void* in, const void* w, const float* bias, float* out, int N, int H, int W, int act,
                                    void* stream) {
  COMBAT_ARG(in && w && out && (act == 0 || act == 1), 0);
  COMBAT_ARG(combat_conv_tc_cout3_supported(N, H, W), 0);
  TcO3Params p;
  memset(&p, 0, sizeof(p));
  TcO3Maps maps;
  memset(&maps, 0, sizeof(maps));
  int bni;
  pick_box(H, W, TILE_M, &p.BW, &p.BH, &bni);
  p.N = N; p.H = H; p.W = W;
  p.tiles_h = cdiv(H, p.BH);
  p.total_tiles = p.tiles_h * N;
  p.act = act;
  p.bias = bias;
  p.out = out;
  const int stage_bytes = (p.BH + 2) * p.BW * 128;
  int n_stages = (232448 - 9 * 2048 - 1024 - 256) / stage_bytes;
  if (n_stages > O3_STAGES) n_stages = O3_STAGES;
  int rc = make_act_map(&maps.in, in, 64, W, H, N, 64, (long long)W * 64, (long long)H * W * 64, p.BW, p.BH + 2, 1);
  if (rc) return rc;
  rc = make_w_map(&maps.w, w, 64, 9, 3, 16);
  if (rc) return rc;
  const int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  const int smem_bytes = 9 * 2048 + n_stages * stage_bytes + 1024 + 256;
  cudaFuncSetAttribute(conv_tc_cout3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  pdl_launch(conv_tc_cout3_kernel, grid, 192, smem_bytes, (cudaStream_t)stream, maps, p, stage_bytes, n_stages);
  COMBAT_RETURN_LAUNCH("conv_tc_cout3");
}

// ---------------------------------------------------------------------------------------------- wgrad kernel
// D[(tap,ci), co] = sum_pixels X[pixel @ tap, ci] * dY[pixel, co]; both operands MN-major (pixels are the K dim).
// One CTA = (pair of 64-row (tap, ci-chunk) blocks) x (BLOCK_N output channels) x (one split of the pixel range);
// partial sums are reduced with coalesced fp32 red.global.add into the channels-last (OHWI) gradient buffer.
struct TcWgradParams {
  int N, Ho, Wo, Co, Ci, taps;
  int BW, BH, BNI;  // 64-pixel box
  int tiles_w, tiles_h, tiles_n, total_boxes, boxes_per_split;
  int kchunks, row_chunks, m_tiles, n_tiles;
  float* dw;  // [Co][taps][Ci] fp32
  TcTaps taps_tbl;  // class 0 only
};
struct TcWgradMaps {
  CUtensorMap x[4];
  CUtensorMap dy;
};

template <int BLOCK_N>
struct TcWCfg {
  static constexpr int A_BYTES = 2 * 64 * KCHUNK * 2;             // two 64-pixel x 64-channel boxes: 16 KB
  static constexpr int B_BYTES = (BLOCK_N / 64) * 64 * KCHUNK * 2;  // 8 KB per 64 output channels
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = BLOCK_N == 256 ? 4 : (BLOCK_N == 128 ? 6 : 8);
  static constexpr int TMEM_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

template <int BLOCK_N>
__global__ void __launch_bounds__(192, 1) conv_tc_wgrad_kernel(const __grid_constant__ TcWgradMaps maps, const TcWgradParams p) {
  pdl_launch_dependents();  // the next kernel may be scheduled; this one waits for its predecessor after its prologue
  using Cfg = TcWCfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::STAGES;
  uint64_t* tfull_bar = bars + 2 * Cfg::STAGES;
  uint32_t* tmem_ptr_smem = (uint32_t*)(bars + 2 * Cfg::STAGES + 1);
  const int warp = uniform_i(threadIdx.x >> 5), lane = threadIdx.x & 31;

  const int m_tile = blockIdx.x % p.m_tiles, n_tile = blockIdx.x / p.m_tiles;
  const int box0 = blockIdx.y * p.boxes_per_split;
  int box1 = box0 + p.boxes_per_split;
  if (box1 > p.total_boxes) box1 = p.total_boxes;
  const int nbox = box1 - box0;  // >= 1 by construction

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.dy);
    tma_prefetch_desc(&maps.x[0]);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = uniform_u(*tmem_ptr_smem);
  pdl_wait();  // barriers initialised, tensor memory allocated: nothing above touched global memory

  // the two 64-row blocks of this M tile
  int rc[2], tap[2], cic[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    rc[j] = m_tile * 2 + j;
    const int r = rc[j] < p.row_chunks ? rc[j] : 0;
    tap[j] = r / p.kchunks;
    cic[j] = r % p.kchunks;
  }

  if (warp == 0) {
    // TMA producer: converged warp, one elected lane issues
    int stage = 0;
    uint32_t phase = 0;
    for (int b = box0; b < box1; ++b) {
      const int wt = b % p.tiles_w;
      const int r2 = b / p.tiles_w;
      const int ht = r2 % p.tiles_h, nt = r2 / p.tiles_h;
      mbar_wait(&empty_bar[stage], phase ^ 1);
      uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
      uint8_t* sb = sa + Cfg::A_BYTES;
      if (elect_one_sync()) {
        mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int t = tap[j];
          tma_load_4d(sa + j * 8192, &maps.x[p.taps_tbl.view[0][t]], &full_bar[stage], cic[j] * KCHUNK,
                      wt * p.BW + p.taps_tbl.dw[0][t], ht * p.BH + p.taps_tbl.dh[0][t], nt * p.BNI);
        }
#pragma unroll
        for (int j = 0; j < BLOCK_N / 64; ++j)
          tma_load_4d(sb + j * 8192, &maps.dy, &full_bar[stage], n_tile * BLOCK_N + j * 64, wt * p.BW, ht * p.BH, nt * p.BNI);
      }
      if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // MMA issuer: converged warp, one elected lane issues
    constexpr uint32_t idesc = make_idesc(TILE_M, BLOCK_N, 1, 1);
    const uint32_t smem_base = smem_u32(smem);
    int stage = 0;
    uint32_t phase = 0;
    for (int b = 0; b < nbox; ++b) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
      const uint32_t sb = sa + Cfg::A_BYTES;
      // MN-major SWIZZLE_128B: LBO = distance between 64-element MN chunks (8 KB), SBO = 8 K-rows (1 KB)
      const uint64_t adesc = make_smem_desc(sa, 8192, 1024);
      const uint64_t bdesc = make_smem_desc(sb, 8192, 1024);
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)  // 16 pixels (K rows of 128 B) per MMA: +2048 B = +128 in the (addr >> 4) field
          umma_bf16(tmem_base, adesc + 128 * k, bdesc + 128 * k, idesc, (b | k) != 0);
        umma_commit(&empty_bar[stage]);
      }
      if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
    }
    if (elect_one_sync()) umma_commit(tfull_bar);
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int j = row >> 6;
    const int ci = cic[j] * KCHUNK + (row & 63);
    const bool valid = rc[j] < p.row_chunks;
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    float* base = p.dw + (long long)tap[j] * p.Ci + ci;
    const long long co_stride = (long long)p.taps * p.Ci;
#pragma unroll 1
    for (int c0 = 0; c0 < BLOCK_N; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(taddr + c0, v);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const int co = n_tile * BLOCK_N + c0 + q;
          atomicAdd(base + co * co_stride, __uint_as_float(v[q]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------- wgrad, 64 -> 64 channels, 3x3, stride 1
// The generic weight-gradient kernel above gives every pair of filter taps its own CTAs, so the layer input is streamed
// from L2 ten times and dY five times (986 MB for a 512 x 32 x 32 x 64 layer: L2-bandwidth bound at ~350 TFLOP/s).
// Here one CTA owns ALL nine taps for its share of the pixels: per 128-pixel tile it loads the three horizontally shifted
// (BH+2)-row boxes of X (as conv_tc64_kernel) and one box of dY, and accumulates the whole 576 x 64 gradient in five
// 128 x 64 TMEM accumulators (tap pairs; the vertical taps are row offsets into the same box).  The accumulators live
// for the whole kernel; one pass of coalesced float32 red.global.add at the end.
struct TcW64Params {
  int N, H, W;            // output == input plane (stride 1)
  int BW, BH;             // tile = BH full rows
  int tiles_h, total_tiles;
  int copy_bytes;         // (BH+2) * BW * 128
  float* dw;              // [64 co][9 taps][64 ci] float32
};
struct TcW64Maps {
  CUtensorMap x;   // box (64, BW, BH+2, 1)
  CUtensorMap dy;  // box (64, BW, BH, 1)
};

__global__ void __launch_bounds__(192, 1) conv_tc_wgrad64_kernel(const __grid_constant__ TcW64Maps maps, const TcW64Params p) {
  pdl_launch_dependents();  // the next kernel may be scheduled; this one waits for its predecessor after its prologue
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int stage_bytes = 3 * p.copy_bytes + 128 * 128;  // three X copies + the 128-pixel dY box
  uint64_t* bars = (uint64_t*)(smem + 2 * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + 2;
  uint64_t* tfull_bar = bars + 4;
  uint32_t* tmem_ptr_smem = (uint32_t*)(bars + 5);
  const int warp = uniform_i(threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.x);
    tma_prefetch_desc(&maps.dy);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = uniform_u(*tmem_ptr_smem);
  pdl_wait();  // barriers initialised, tensor memory allocated: nothing above touched global memory

  if (warp == 0) {
    // TMA producer: converged warp, one elected lane issues
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int ht = tile % p.tiles_h, n = tile / p.tiles_h;
      mbar_wait(&empty_bar[stage], phase ^ 1);
      uint8_t* sa = smem + stage * stage_bytes;
      if (elect_one_sync()) {
        mbar_expect_tx(&full_bar[stage], stage_bytes);
        for (int dwi = 0; dwi < 3; ++dwi) tma_load_4d(sa + dwi * p.copy_bytes, &maps.x, &full_bar[stage], 0, dwi - 1, ht * p.BH - 1, n);
        tma_load_4d(sa + 3 * p.copy_bytes, &maps.dy, &full_bar[stage], 0, 0, ht * p.BH, n);
      }
      if (++stage == 2) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // MMA issuer: converged warp, one elected lane issues
    constexpr uint32_t idesc = make_idesc(TILE_M, 64, 1, 1);
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t row_step = (uint32_t)p.BW * 128;
    int stage = 0;
    uint32_t phase = 0;
    bool first = true;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      const uint32_t sa = smem_base + stage * stage_bytes;
      const uint32_t sb = sa + 3 * p.copy_bytes;
      // taps in address order: index i = dwi*3 + dhi at sa + dwi*copy + dhi*row_step; accumulator a pairs (2a, 2a+1),
      // the last one pairs (7, 8) again and only its upper 64 rows (tap index 8) are used
      if (elect_one_sync()) {
#pragma unroll
        for (int a = 0; a < 5; ++a) {
          const int i0 = a < 4 ? 2 * a : 7, i1 = i0 + 1;
          const uint32_t addr0 = sa + (i0 / 3) * p.copy_bytes + (i0 % 3) * row_step;
          const uint32_t addr1 = sa + (i1 / 3) * p.copy_bytes + (i1 % 3) * row_step;
          const uint64_t adesc = make_smem_desc(addr0, addr1 - addr0, 1024);
          const uint64_t bdesc = make_smem_desc(sb, 8192, 1024);
#pragma unroll
          for (int k = 0; k < 8; ++k)  // 16 pixels (K rows of 128 B) per MMA: +2048 B = +128 in the (addr >> 4) field
            umma_bf16(tmem_base + a * 64, adesc + 128 * k, bdesc + 128 * k, idesc, !(first && k == 0));
        }
        umma_commit(&empty_bar[stage]);
      }
      first = false;
      if (++stage == 2) { stage = 0; phase ^= 1; }
    }
    if (elect_one_sync()) umma_commit(tfull_bar);
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    const long long co_stride = 9LL * 64;
#pragma unroll 1
    for (int a = 0; a < 5; ++a) {
      const int i = (a < 4 ? 2 * a : 7) + (row >> 6);   // address-order tap index of this row's chunk
      const bool used = a < 4 || row >= 64;
      const int dwi = i / 3, dhi = i % 3;
      float* base = p.dw + (long long)(dhi * 3 + dwi) * 64 + (row & 63);  // tap (kh = dhi, kw = dwi), ci = row & 63
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + a * 64;
#pragma unroll 1
      for (int c0 = 0; c0 < 64; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + c0, v);
        tmem_ld_wait();
        if (used) {
#pragma unroll
          for (int q = 0; q < 16; ++q) atomicAdd(base + (c0 + q) * co_stride, __uint_as_float(v[q]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

extern "C" int combat_conv_tc_wgrad(const combat_conv_tc_desc* d, const void* dy, float* dw_ohwi, void* stream) {
  COMBAT_ARG(d && d->in && dy && dw_ohwi, 0);
  COMBAT_ARG(combat_conv_tc_supported(d) && d->up == 1, 0);
  TcWgradParams p;
  memset(&p, 0, sizeof(p));
  TcWgradMaps maps;
  memset(&maps, 0, sizeof(maps));
  const int KH = d->KH, KW = d->KW, pad = d->pad, s = d->stride;
  if (d->Ci == 64 && d->Co == 64 && KH == 3 && pad == 1 && s == 1 && !getenv("COMBAT_NO_TC64")) {
    // all nine taps in one CTA (conv_tc_wgrad64_kernel): tiles of BH full-width rows of one image
    TcW64Params q;
    memset(&q, 0, sizeof(q));
    int bni;
    pick_box(d->Ho, d->Wo, TILE_M, &q.BW, &q.BH, &bni);
    q.copy_bytes = (q.BH + 2) * q.BW * 128;
    const int smem_bytes = 2 * (3 * q.copy_bytes + 128 * 128) + 1024 + 256;
    if (q.BW == d->Wo && bni == 1 && (q.BW % 8) == 0 && smem_bytes <= 232448) {
      TcW64Maps wm;
      memset(&wm, 0, sizeof(wm));
      const long long C = 64, W = d->Wi, H = d->Hi;
      int rc = make_act_map(&wm.x, d->in, 64, d->Wi, d->Hi, d->N, C, W * C, H * W * C, q.BW, q.BH + 2, 1);
      if (rc) return rc;
      rc = make_act_map(&wm.dy, dy, 64, d->Wo, d->Ho, d->N, C, W * C, H * W * C, q.BW, q.BH, 1);
      if (rc) return rc;
      q.N = d->N; q.H = d->Ho; q.W = d->Wo;
      q.tiles_h = cdiv(d->Ho, q.BH);
      q.total_tiles = q.tiles_h * d->N;
      q.dw = dw_ohwi;
      const int grid = q.total_tiles < num_sms() ? q.total_tiles : num_sms();
      cudaFuncSetAttribute(conv_tc_wgrad64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
      pdl_launch(conv_tc_wgrad64_kernel, grid, 192, smem_bytes, (cudaStream_t)stream, wm, q);
      COMBAT_RETURN_LAUNCH("conv_tc_wgrad64");
    }
  }
  p.N = d->N; p.Ho = d->Ho; p.Wo = d->Wo; p.Co = d->Co; p.Ci = d->Ci; p.taps = KH * KW;
  p.kchunks = d->Ci / KCHUNK;
  p.row_chunks = p.taps * p.kchunks;
  p.m_tiles = (p.row_chunks + 1) / 2;
  const int BLOCK_N = d->Co % 256 == 0 ? 256 : (d->Co % 128 == 0 ? 128 : 64);
  p.n_tiles = d->Co / BLOCK_N;
  p.dw = dw_ohwi;
  pick_box(d->Ho, d->Wo, 64, &p.BW, &p.BH, &p.BNI);
  int nt = 0;
  for (int kh = 0; kh < KH; ++kh)
    for (int kw = 0; kw < KW; ++kw) {
      const int oh = kh - pad, ow = kw - pad;
      int ph = 0, pw = 0, dh = oh, dwv = ow;
      if (s == 2) {
        ph = ((oh % 2) + 2) % 2; pw = ((ow % 2) + 2) % 2;
        dh = floordiv2(oh); dwv = floordiv2(ow);
      }
      p.taps_tbl.view[0][nt] = (signed char)(ph * 2 + pw);
      p.taps_tbl.dh[0][nt] = (signed char)dh;
      p.taps_tbl.dw[0][nt] = (signed char)dwv;
      ++nt;
    }
  int rc;
  const long long C = d->Ci, W = d->Wi, H = d->Hi;
  if (s == 1) {
    rc = make_act_map(&maps.x[0], d->in, d->Ci, d->Wi, d->Hi, d->N, C, W * C, H * W * C, p.BW, p.BH, p.BNI);
    if (rc) return rc;
  } else {
    for (int ph = 0; ph < 2; ++ph)
      for (int pw = 0; pw < 2; ++pw) {
        const bf16* base = (const bf16*)d->in + ((long long)ph * W + pw) * C;
        rc = make_act_map(&maps.x[ph * 2 + pw], base, d->Ci, (d->Wi - pw + 1) / 2, (d->Hi - ph + 1) / 2, d->N, 2 * C, 2 * W * C,
                          H * W * C, p.BW, p.BH, p.BNI);
        if (rc) return rc;
      }
  }
  const long long Co = d->Co;
  rc = make_act_map(&maps.dy, dy, d->Co, d->Wo, d->Ho, d->N, Co, (long long)d->Wo * Co, (long long)d->Ho * d->Wo * Co, p.BW, p.BH,
                    p.BNI);
  if (rc) return rc;
  p.tiles_w = cdiv(d->Wo, p.BW);
  p.tiles_h = cdiv(d->Ho, p.BH);
  p.tiles_n = cdiv(d->N, p.BNI);
  p.total_boxes = p.tiles_w * p.tiles_h * p.tiles_n;
  const int ctas_mn = p.m_tiles * p.n_tiles;
  // ONE wave of CTAs (the shared-memory ring leaves one CTA per SM): the pixel range is split so that m_tiles * n_tiles * splits
  // is the largest multiple of the (M, N) tile count that fits the SMs.  (Round 1 aimed at two CTAs per SM with
  // ceil(2 * SMs / tiles) splits: 297 / 306 / 360 CTAs for the 128 / 256 / 512-channel layers = a third, almost empty wave.)
  int splits = num_sms() / ctas_mn;
  if (getenv("COMBAT_WGRAD_SPLITS2")) splits = cdiv(2 * num_sms(), ctas_mn);
  if (splits > p.total_boxes) splits = p.total_boxes;
  if (splits < 1) splits = 1;
  p.boxes_per_split = cdiv(p.total_boxes, splits);
  splits = cdiv(p.total_boxes, p.boxes_per_split);
  dim3 grid(ctas_mn, splits);
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH_W(BN)                                                                                              \
  {                                                                                                               \
    cudaFuncSetAttribute(conv_tc_wgrad_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcWCfg<BN>::SMEM_BYTES); \
    pdl_launch(conv_tc_wgrad_kernel<BN>, grid, 192, TcWCfg<BN>::SMEM_BYTES, st, maps, p);                                 \
  }
  if (BLOCK_N == 256) LAUNCH_W(256) else if (BLOCK_N == 128) LAUNCH_W(128) else LAUNCH_W(64)
#undef LAUNCH_W
  COMBAT_RETURN_LAUNCH("conv_tc_wgrad");
}
