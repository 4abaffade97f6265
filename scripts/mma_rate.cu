// Micro-benchmark behind DESIGN.md section 4.1: what paces tcgen05.mma (SS mode, bf16, K-major SWIZZLE_128B operands) on B200?
//   part 1  MMA-only rate vs (cta_group, N): operands static in shared memory, no other traffic
//   part 2  the same with a free-running TMA producer writing 16 KB boxes into OTHER shared-memory buffers of the same SM
//           (how the tensor pipe and the TMA writes share the shared-memory port), and TMA alone
//   part 3  does a K-major SWIZZLE_128B descriptor whose start address is shifted by s x 128 B (one pixel row of a staged box)
//           read rows s .. s+127?  (would let ONE box serve the three horizontal taps of a 3x3 conv)
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/_bin/mma_rate scripts/mma_rate.cu -lcuda
// Run  :  scripts/_bin/mma_rate            (prints one line per case; exit code 0)
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

typedef __nv_bfloat16 bf16;

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e__ = (x);                                                                      \
    if (e__ != cudaSuccess) {                                                                   \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__);          \
      exit(1);                                                                                  \
    }                                                                                           \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_u32(dst)),
               "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
template <int CG>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  if (CG == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
  }
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
  else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
}
template <int CG>
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  if (CG == 1)
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accum)
        : "memory");
  else
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accum)
        : "memory");
}
template <int CG>
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  if (CG == 1) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  } else {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(mask)
                 : "memory");
  }
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// one lane of a converged warp (CUTLASS's elect_one_sync)
__device__ __forceinline__ uint32_t elect_one_sync() {
  uint32_t pred = 0;
  asm volatile(
      "{\n.reg .b32 %%rx;\n.reg .pred %%px;\nelect.sync %%rx|%%px, %1;\n@%%px mov.s32 %0, 1;\n}\n"
      : "+r"(pred)
      : "r"(0xffffffffu));
  return pred;
}

struct RateParams {
  int N;          // MMA N (full, both CTAs of a pair together)
  int groups;     // groups of 4 MMAs (K = 64)
  int flags;      // F_* below
  int tma_rows;   // rows of the TMA source tensor
  int stages;     // F_PIPE: pipeline depth
  long long* out; // [grid][4]: mma cycles, mma count, tma cycles, tma boxes
};
#define F_NO_MMA 1      // TMA only (for `groups` boxes)
#define F_TMA_FREE 2    // free-running TMA producer into buffers the MMA does not read
#define F_COMMIT 4      // tcgen05.commit after every group of 4 MMAs (ring of 8 mbarriers, waited on before re-use)
#define F_PIPE 8        // conv-like operand pipeline: TMA -> full barrier -> 4 MMAs -> commit -> empty barrier (cta_group 1 only)
#define F_TMEMLD 16     // four warps loop tcgen05.ld over the accumulator columns (epilogue-like tensor-memory reads)
#define F_LSU 32        // four warps loop the st.shared.v4 / ld.shared.v4 transposition of the conv epilogue on a private tile
#define F_UNIFORM 128  // the MMA warp runs its loop CONVERGED (all 32 lanes) and issues through elect.sync: operands stay in uniform registers
#define F_GMEM 64       // four warps stream 16-byte global stores (epilogue-like output traffic)

#define A_BUF 16384
#define B_BUF 32768
#define T_BUF 16384
#define T_RING 4
#define OPER_BYTES (200 * 1024)
#define LSU_ROWB 144

__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr) : "memory");
  return v;
}

template <int CG>
__global__ void __launch_bounds__(192, 1) mma_rate_k(const __grid_constant__ CUtensorMap map, const __grid_constant__ CUtensorMap bmap,
                                                     const RateParams p, float* sink) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = smem;                       // static operands: 2 x 16 KB
  uint8_t* sb = smem + 2 * A_BUF;           // 2 x 32 KB
  uint8_t* st = sb + 2 * B_BUF;             // T_RING x 16 KB
  uint8_t* lsu = smem + OPER_BYTES;         // 4 warps x 32 x LSU_ROWB
  uint64_t* bars = (uint64_t*)(lsu + 4 * 32 * LSU_ROWB);
  uint64_t* done_bar = bars;                // MMA completion
  uint64_t* tfull = bars + 1;               // T_RING
  uint64_t* cbar = bars + 1 + T_RING;       // 8: commit ring
  uint64_t* pfull = cbar + 8;               // 8: pipeline full
  uint64_t* pempty = pfull + 8;             // 8: pipeline empty
  uint32_t* tmem_ptr = (uint32_t*)(pempty + 8);
  volatile int* stop = (volatile int*)(tmem_ptr + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;
  const int b_local = (CG == 2 ? p.N / 2 : p.N) * 128;       // bytes of this CTA's B tile per K chunk
  const int stage_bytes = A_BUF + ((b_local + 1023) & ~1023);
  // operands: small pseudo-random bf16 values
  for (int i = threadIdx.x; i < OPER_BYTES / 2; i += blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
    ((bf16*)smem)[i] = __float2bfloat16_rn(((int)(h >> 24) - 128) * (1.f / 256.f));
  }
  if (threadIdx.x == 0) {
    mbar_init(done_bar, 1);
    for (int s = 0; s < T_RING; ++s) mbar_init(&tfull[s], 1);
    for (int s = 0; s < 8; ++s) {
      mbar_init(&cbar[s], 1);
      mbar_init(&pfull[s], 1);
      mbar_init(&pempty[s], 1);
    }
    *stop = 0;
    fence_barrier_init();
  }
  fence_proxy_async();
  if (warp == 1) tmem_alloc<CG>(tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  long long* o = p.out + (long long)blockIdx.x * 4;
  const int warp_u = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const uint32_t rank_u = __shfl_sync(0xffffffffu, rank, 0);
  if (warp_u == 1 && (p.flags & F_UNIFORM)) {
    // whole warp, converged; only the elected lane issues
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    if (rank_u == 0) {
      const uint32_t idesc = make_idesc(128 * CG, p.N);
      const uint32_t sa_u = smem_u32(sa), sb_u = smem_u32(sb);
      const long long t0 = clock64();
      for (int g = 0; g < p.groups; ++g) {
        if ((p.flags & F_COMMIT) && g >= 8) mbar_wait(&cbar[g & 7], ((g >> 3) - 1) & 1);
        const uint64_t adesc = make_smem_desc(sa_u + (g & 1) * A_BUF, 16, 1024, 0);
        const uint64_t bdesc = make_smem_desc(sb_u + (g & 1) * B_BUF, 16, 1024, 0);
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma<CG>(tmem_u + (g & 1) * 256, adesc + 2 * k, bdesc + 2 * k, idesc, k != 0);
          if (p.flags & F_COMMIT) umma_commit<CG>(&cbar[g & 7]);
        }
        __syncwarp();
      }
      if (elect_one_sync()) umma_commit<CG>(done_bar);
      __syncwarp();
      mbar_wait(done_bar, 0);
      const long long t1 = clock64();
      if (lane == 0) {
        o[0] = t1 - t0;
        o[1] = 4LL * p.groups;
      }
    } else {
      mbar_wait(done_bar, 0);
    }
    if (lane == 0) *stop = 1;
  } else if (warp == 1 && lane == 0) {
    if (!(p.flags & F_NO_MMA) && rank == 0) {
      const uint32_t idesc = make_idesc(128 * CG, p.N);
      const long long t0 = clock64();
      if (p.flags & F_PIPE) {
        int stage = 0;
        uint32_t phase = 0;
        for (int g = 0; g < p.groups; ++g) {
          mbar_wait(&pfull[stage], phase);
          tc_fence_after();
          const uint32_t s0 = smem_u32(smem + stage * stage_bytes);
          const uint64_t adesc = make_smem_desc(s0, 16, 1024, 0);
          const uint64_t bdesc = make_smem_desc(s0 + A_BUF, 16, 1024, 0);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma<CG>(tmem_base + (g & 1) * 256, adesc + 2 * k, bdesc + 2 * k, idesc, k != 0);
          umma_commit<CG>(&pempty[stage]);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      } else {
        for (int g = 0; g < p.groups; ++g) {
          if ((p.flags & F_COMMIT) && g >= 8) mbar_wait(&cbar[g & 7], ((g >> 3) - 1) & 1);
          const uint64_t adesc = make_smem_desc(smem_u32(sa + (g & 1) * A_BUF), 16, 1024, 0);
          const uint64_t bdesc = make_smem_desc(smem_u32(sb + (g & 1) * B_BUF), 16, 1024, 0);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma<CG>(tmem_base + (g & 1) * 256, adesc + 2 * k, bdesc + 2 * k, idesc, k != 0);
          if (p.flags & F_COMMIT) umma_commit<CG>(&cbar[g & 7]);
        }
      }
      umma_commit<CG>(done_bar);
      mbar_wait(done_bar, 0);
      const long long t1 = clock64();
      o[0] = t1 - t0;
      o[1] = 4LL * p.groups;
    } else if (!(p.flags & F_NO_MMA)) {
      mbar_wait(done_bar, 0);  // follower: the multicast commit arrives here too
    }
    *stop = 1;
  } else if (warp == 0 && lane == 0 && (p.flags & F_PIPE)) {
    // conv-like producer: A box (128 rows x 128 B) + B box (b_local bytes) per group
    const int nbox = p.tma_rows / 256;
    int row = (blockIdx.x * 37) % nbox;
    int stage = 0;
    uint32_t phase = 0;
    const long long t0 = clock64();
    for (int g = 0; g < p.groups; ++g) {
      mbar_wait(&pempty[stage], phase ^ 1);
      uint8_t* s0 = smem + stage * stage_bytes;
      mbar_expect_tx(&pfull[stage], A_BUF + b_local);
      tma_load_2d(s0, &map, &pfull[stage], 0, row * 256);
      tma_load_2d(s0 + A_BUF, &bmap, &pfull[stage], 0, row * 256);
      row = (row + 1) % nbox;
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
    o[2] = clock64() - t0;
    o[3] = ((long long)p.groups * (A_BUF + b_local)) / T_BUF;
  } else if (warp == 0 && lane == 0 && (p.flags & (F_TMA_FREE | F_NO_MMA))) {
    // free-running TMA producer: ring of T_RING boxes, re-issued as soon as each lands
    const long long t0 = clock64();
    long long boxes = 0;
    const int nbox = p.tma_rows / 128;
    int row = (blockIdx.x * 37) % nbox;
    for (int s = 0; s < T_RING; ++s) {
      mbar_expect_tx(&tfull[s], T_BUF);
      tma_load_2d(st + s * T_BUF, &map, &tfull[s], 0, row * 128);
      row = (row + 1) % nbox;
    }
    uint32_t phase = 0;
    bool run = true;
    while (run) {
      for (int s = 0; s < T_RING; ++s) {
        mbar_wait(&tfull[s], phase);
        ++boxes;
        run = (p.flags & F_NO_MMA) ? boxes < p.groups : (*stop == 0);
        if (!run) {
          // drain the rest of the ring before leaving (the barriers must not be hit after exit)
          for (int s2 = s + 1; s2 < T_RING; ++s2) mbar_wait(&tfull[s2], phase);
          for (int s2 = 0; s2 < s; ++s2) mbar_wait(&tfull[s2], phase ^ 1);
          break;
        }
        mbar_expect_tx(&tfull[s], T_BUF);
        tma_load_2d(st + s * T_BUF, &map, &tfull[s], 0, row * 128);
        row = (row + 1) % nbox;
      }
      phase ^= 1;
    }
    o[2] = clock64() - t0;
    o[3] = boxes;
  } else if (warp >= 2 && (p.flags & (F_TMEMLD | F_LSU | F_GMEM))) {
    const int q = warp - 2;  // TMEM lane quarter = warp % 4 (warps 2..5 -> 2, 3, 0, 1)
    const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t tile = smem_u32(lsu + q * 32 * LSU_ROWB);
    const int sub = lane >> 3, cseg = lane & 7;
    float acc = 0.f;
    uint32_t col = 0;
    float4* gdst = (float4*)sink + ((size_t)blockIdx.x * 4 + q) * 65536;  // 1 MB per warp, re-written round robin
    uint32_t gi = 0;
    while (*stop == 0) {
      uint32_t v[32];
      if (p.flags & F_TMEMLD) {
        tmem_ld16(taddr + col, *(uint32_t(*)[16]) & v[0]);
        tmem_ld16(taddr + col + 16, *(uint32_t(*)[16]) & v[16]);
        tmem_ld_wait();
        col = (col + 32) & 511;
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = lane + j;
      }
      float4 f[8];
      if (p.flags & F_LSU) {
        const uint32_t wr = tile + lane * LSU_ROWB;
#pragma unroll
        for (int j = 0; j < 8; ++j) sts128(wr + j * 16, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = lds128(tile + (4 * i + sub) * LSU_ROWB + cseg * 16);
        __syncwarp();
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
      }
      if (p.flags & F_GMEM) {
#pragma unroll
        for (int i = 0; i < 8; ++i) gdst[(gi + i * 32 + lane) & 65535] = f[i];
        gi += 256;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc += f[i].x + f[i].y + f[i].z + f[i].w;
      }
    }
    if (acc == 12345.678f) sink[0] = acc;
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<CG>(tmem_base, 512);
  }
}

// ---- part 3: shifted start address.  A tile: 256 rows x 64 bf16 (128 B per row) stored the way TMA SWIZZLE_128B stores it
// (16-byte chunk index XOR (row & 7)), value(r, c) = (7 r + 3 c) % 251.  B = 64 x 64 identity.  D = A rows [s, s+128).
__global__ void __launch_bounds__(128, 1) shift_test_k(int shift_rows, int base_off, float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = smem;              // 256 x 128 B
  uint8_t* sb = smem + 32768;      // 64 x 128 B
  uint64_t* done_bar = (uint64_t*)(sb + 8192);
  uint32_t* tmem_ptr = (uint32_t*)(done_bar + 1);
  for (int i = threadIdx.x; i < 256 * 64; i += blockDim.x) {
    const int r = i >> 6, c = i & 63;
    const int chunk = (c >> 3) ^ (r & 7);
    ((bf16*)sa)[r * 64 + chunk * 8 + (c & 7)] = __float2bfloat16_rn((float)((7 * r + 3 * c) % 251));
  }
  for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
    const int r = i >> 6, c = i & 63;
    const int chunk = (c >> 3) ^ (r & 7);
    ((bf16*)sb)[r * 64 + chunk * 8 + (c & 7)] = __float2bfloat16_rn(r == c ? 1.f : 0.f);
  }
  if (threadIdx.x == 0) {
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  fence_proxy_async();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 1) tmem_alloc<1>(tmem_ptr, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (warp == 1 && lane == 0) {
    const uint32_t idesc = make_idesc(128, 64);
    const uint64_t adesc = make_smem_desc(smem_u32(sa) + shift_rows * 128, 16, 1024, base_off);
    const uint64_t bdesc = make_smem_desc(smem_u32(sb), 16, 1024, 0);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma<1>(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, k != 0);
    umma_commit<1>(done_bar);
  }
  mbar_wait(done_bar, 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < 64; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int q = 0; q < 16; ++q) out[row * 64 + c0 + q] = __uint_as_float(v[q]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 64);
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int CG>
static void run_rate(const CUtensorMap& map, const CUtensorMap& bmap, int N, int flags, int groups, int rows, long long* d_out, float* sink,
                     int sms, int stages = 4) {
  RateParams p;
  p.N = N; p.groups = groups; p.flags = flags; p.tma_rows = rows; p.out = d_out; p.stages = stages;
  const int smem_bytes = OPER_BYTES + 4 * 32 * LSU_ROWB + 1024 + 512;
  CK(cudaFuncSetAttribute(mma_rate_k<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  const int grid = CG == 2 ? (sms / 2) * 2 : sms;
  CK(cudaMemset(d_out, 0, sizeof(long long) * 4 * grid));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = smem_bytes;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = CG;
  attr.val.clusterDim.y = attr.val.clusterDim.z = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {  // second run is the one reported (clocks up, L2 warm)
    CK(cudaLaunchKernelEx(&cfg, mma_rate_k<CG>, map, bmap, p, sink));
    CK(cudaDeviceSynchronize());
  }
  std::vector<long long> h(4 * grid);
  CK(cudaMemcpy(h.data(), d_out, sizeof(long long) * 4 * grid, cudaMemcpyDeviceToHost));
  double mc = 0, mn = 0, tc = 0, tb = 0;
  int nm = 0, nt = 0;
  for (int b = 0; b < grid; ++b) {
    if (h[4 * b + 1] > 0) { mc += (double)h[4 * b]; mn += (double)h[4 * b + 1]; ++nm; }
    if (h[4 * b + 3] > 0) { tc += (double)h[4 * b + 2]; tb += (double)h[4 * b + 3]; ++nt; }
  }
  const double floor_cyc = 128.0 * N / 256.0;  // cycles per K=16 dispatch at the tensor-pipe peak (per SM; a pair runs M=256 in the same time)
  char fl[160] = "";
  if (flags & F_NO_MMA) strcat(fl, " tma-only");
  if (flags & F_TMA_FREE) strcat(fl, " +tma(free)");
  if (flags & F_COMMIT) strcat(fl, " +commit/4");
  if (flags & F_PIPE) snprintf(fl + strlen(fl), 32, " pipeline(%d)", stages);
  if (flags & F_TMEMLD) strcat(fl, " +tmem.ld");
  if (flags & F_LSU) strcat(fl, " +sts/lds");
  if (flags & F_GMEM) strcat(fl, " +stg");
  if (flags & F_UNIFORM) strcat(fl, " [uniform issue]");
  printf("cta_group %d  N %3d %-38s: ", CG, N, fl[0] ? fl : " mma only");
  if (nm) {
    const double cyc = mc / mn;
    const double local_b = CG == 2 ? N / 2 : N;
    printf("%7.1f cyc/MMA (floor %5.1f, %5.1f %% of peak), MMA smem reads %6.1f B/clk/SM", cyc, floor_cyc, 100.0 * floor_cyc / cyc,
           (128 + local_b) * 32.0 / cyc);
  }
  if (nt) printf("  | TMA writes %6.1f B/clk/SM", tb * T_BUF / tc);
  printf("\n");
  fflush(stdout);
}

static CUtensorMap make_map(void* fn, void* src, int rows, int box_rows) {
  CUtensorMap map;
  cuuint64_t dims[2] = {64, (cuuint64_t)rows};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = ((PFN_encodeTiled)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, src, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("tensor map encode failed %d\n", (int)r); exit(1); }
  return map;
}

int main() {
  int dev = 0, sms = 0;
  CK(cudaSetDevice(dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  // TMA source: 32 MB of bf16 rows of 64 channels (L2 resident after the warm-up run)
  const int rows = 262144;
  bf16* src;
  CK(cudaMalloc(&src, (size_t)rows * 128));
  CK(cudaMemset(src, 0, (size_t)rows * 128));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  CUtensorMap map = make_map(fn, src, rows, 128);
  long long* d_out;
  CK(cudaMalloc(&d_out, sizeof(long long) * 4 * 256));
  float* sink;
  CK(cudaMalloc(&sink, (size_t)148 * 4 * 65536 * 16));
  printf("# SMs %d; cycles are SM clocks (clock64) on the MMA-issuing thread, averaged over CTAs\n", sms);
  const int Ns[] = {64, 128, 256};
  const int only = getenv("MMA_RATE_PART") ? atoi(getenv("MMA_RATE_PART")) : 0;
  if (only == 0 || only == 1) {
    for (int N : Ns) {
      CUtensorMap bmap1 = make_map(fn, src, rows, N), bmap2 = make_map(fn, src, rows, N / 2);
      const int fl[] = {0, F_UNIFORM, F_UNIFORM | F_COMMIT, F_UNIFORM | F_COMMIT | F_TMA_FREE | F_TMEMLD | F_LSU | F_GMEM, F_TMA_FREE, F_COMMIT, F_TMEMLD, F_LSU, F_TMEMLD | F_LSU, F_GMEM, F_TMEMLD | F_LSU | F_GMEM,
                        F_COMMIT | F_TMA_FREE | F_TMEMLD | F_LSU | F_GMEM};
      for (int f : fl) {
        run_rate<1>(map, bmap1, N, f, 4000, rows, d_out, sink, sms);
        run_rate<2>(map, bmap2, N, f, 4000, rows, d_out, sink, sms);
      }
      for (int stages = 2; stages <= 4; ++stages)
        if (stages * (A_BUF + N * 128) <= OPER_BYTES) {
          run_rate<1>(map, bmap1, N, F_PIPE, 4000, rows, d_out, sink, sms, stages);
          run_rate<1>(map, bmap1, N, F_PIPE | F_TMEMLD | F_LSU | F_GMEM, 4000, rows, d_out, sink, sms, stages);
        }
    }
    CUtensorMap bmap1 = make_map(fn, src, rows, 64);
    run_rate<1>(map, bmap1, 64, F_NO_MMA, 20000, rows, d_out, sink, sms);
  }
  if (only == 1) return 0;

  // part 3
  float* d_d;
  CK(cudaMalloc(&d_d, 128 * 64 * sizeof(float)));
  CK(cudaFuncSetAttribute(shift_test_k, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 + 8192 + 1024 + 64));
  std::vector<float> h(128 * 64);
  for (int s = 0; s <= 9; ++s)
    for (int bo = 0; bo < 2; ++bo) {
      const int base_off = bo ? (s & 7) : 0;
      if (bo && base_off == 0) continue;
      CK(cudaMemset(d_d, 0xff, 128 * 64 * sizeof(float)));
      shift_test_k<<<1, 128, 32768 + 8192 + 1024 + 64>>>(s, base_off, d_d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("shift %d base_offset %d : CUDA error %s\n", s, base_off, cudaGetErrorString(e)); return 2; }
      CK(cudaMemcpy(h.data(), d_d, 128 * 64 * sizeof(float), cudaMemcpyDeviceToHost));
      int ok_rows = 0, first_bad = -1;
      for (int m = 0; m < 128; ++m) {
        bool ok = true;
        for (int c = 0; c < 64; ++c) ok = ok && h[m * 64 + c] == (float)((7 * (m + s) + 3 * c) % 251);
        ok_rows += ok;
        if (!ok && first_bad < 0) first_bad = m;
      }
      printf("shift %d rows, base_offset field %d : %3d / 128 rows equal A[m + shift]", s, base_off, ok_rows);
      if (first_bad >= 0) {
        // which source row did the first wrong row come from (if any)?
        int src_row = -1;
        for (int r2 = 0; r2 < 256 && src_row < 0; ++r2) {
          bool eq = true;
          for (int c = 0; c < 64; ++c) eq = eq && h[first_bad * 64 + c] == (float)((7 * r2 + 3 * c) % 251);
          if (eq) src_row = r2;
        }
        printf("  (first wrong row m=%d holds source row %d; D[m][0..3] = %g %g %g %g)", first_bad, src_row, h[first_bad * 64], h[first_bad * 64 + 1],
               h[first_bad * 64 + 2], h[first_bad * 64 + 3]);
      }
      printf("\n");
    }
  return 0;
}
