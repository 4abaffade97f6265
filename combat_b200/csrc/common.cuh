// Shared helpers for the combat_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/combat_b200.h"

typedef __nv_bfloat16 bf16;

extern long long g_combat_launches;
extern char g_combat_err[256];
void combat_set_err(const char* what, cudaError_t e);

#define COMBAT_COUNT_LAUNCH() (++g_combat_launches)

// launch-and-check used by every entry point: peeks (does not clear a sticky error from elsewhere)
#define COMBAT_RETURN_LAUNCH(name)               \
  do {                                           \
    COMBAT_COUNT_LAUNCH();                       \
    cudaError_t e__ = cudaGetLastError();        \
    if (e__ != cudaSuccess) {                    \
      combat_set_err(name, e__);                 \
      return -(int)e__;                          \
    }                                            \
    return 0;                                    \
  } while (0)

#define COMBAT_CHECK_LAUNCH(name)                \
  do {                                           \
    COMBAT_COUNT_LAUNCH();                       \
    cudaError_t e__ = cudaGetLastError();        \
    if (e__ != cudaSuccess) {                    \
      combat_set_err(name, e__);                 \
      return -(int)e__;                          \
    }                                            \
  } while (0)

#define COMBAT_ARG(cond, k)                      \
  do {                                           \
    if (!(cond)) {                               \
      combat_set_err("bad argument: " #cond, cudaErrorInvalidValue); \
      return -1000 - (k);                        \
    }                                            \
  } while (0)

// ---- programmatic dependent launch (PDL), opt-in with COMBAT_PDL=1 (measured slower inside the step, see lib.cu).  With it every
// kernel of the library is launched with cudaLaunchAttributeProgrammaticStreamSerialization and begins with pdl_entry(): `griddepcontrol.launch_dependents` lets the
// NEXT kernel of the stream / captured graph be scheduled while this one is still running (its CTAs become resident as SM
// resources free up), `griddepcontrol.wait` blocks until every prerequisite grid has completed and its memory is visible -- so
// nothing a kernel reads or writes can race with its predecessor, and what overlaps is the launch latency, CTA rasterisation
// and (tcgen05 kernels) the barrier / tensor-memory prologue.  Without the attribute both instructions are no-ops.
// (-DCOMBAT_PDL_LATE builds the library without the explicit trigger: dependents are then released when the primary grid completes.
// Measured after the round-2 kernel work: 9.87 ms without PDL, 10.32 with the trigger at kernel entry, 10.08 without the trigger.)
#ifdef COMBAT_PDL_LATE
__device__ __forceinline__ void pdl_launch_dependents() {}
#else
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_entry() {
  pdl_launch_dependents();
  pdl_wait();
}
bool combat_pdl_enabled();

template <typename... P, typename... A>
static inline cudaError_t pdl_launch(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = combat_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(args)...);
}

template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }

template <typename T>
__device__ __forceinline__ T from_f(float v);
template <>
__device__ __forceinline__ float from_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// dtype-erased scalar load/store (generic strided SIMT paths)
__device__ __forceinline__ float ld_any(const void* p, long long i, int dtype) {
  return dtype == COMBAT_F32 ? ((const float*)p)[i] : __bfloat162float(((const bf16*)p)[i]);
}
__device__ __forceinline__ void st_any(void* p, long long i, int dtype, float v) {
  if (dtype == COMBAT_F32)
    ((float*)p)[i] = v;
  else
    ((bf16*)p)[i] = __float2bfloat16_rn(v);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// Grid of a grid-stride kernel: every CTA resident at once (SM count x CTAs per SM at this kernel's register / shared
// memory footprint), so that there is no partial last wave.  0 if the query fails (the caller falls back to a cap).
template <typename Kernel>
static inline int resident_ctas(Kernel kernel, int threads, size_t dyn_smem = 0) {
  int dev = 0, sms = 0, per_sm = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, dyn_smem) != cudaSuccess) return 0;
  return sms * per_sm;
}

#define DISPATCH_DTYPE(dtype, ...)        \
  if ((dtype) == COMBAT_F32) {            \
    typedef float T;                      \
    __VA_ARGS__                           \
  } else {                                \
    typedef bf16 T;                       \
    __VA_ARGS__                           \
  }
