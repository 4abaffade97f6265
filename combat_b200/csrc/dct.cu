// Batched full-plane 2-D transforms  out[p] = L * X[p] * R^T  (DCT-II, DCT-III, low-pass projection).
// Replaces utils/dct.py:13-111 (FFT route) and train_generator.py:47-55 (low_freq) of the reference.
//
// Two generic kernels (any N, any L/R):
//   plane_transform_smem  N <= 64 : one plane per CTA iteration, X/T/L/R^T staged in shared memory
//   plane_mm              N  > 64 : two global passes through a caller workspace
// and the HBM-bound register-butterfly kernels for N == 32 live in dct32.cu.
#include "common.cuh"

template <int MODE>
__device__ __forceinline__ float load_in(const void* in, long long i) {
  if (MODE == 0) return ((const float*)in)[i];
  if (MODE == 1) return (float)((const unsigned char*)in)[i];
  // MODE 2: ((x+1)/2*255).byte() -- truncation toward zero, as torch's float->uint8 cast
  float v = (((const float*)in)[i] + 1.0f) / 2.0f * 255.0f;
  return (float)(unsigned char)(int)v;
}

// smem layout (floats): Ls[N*N] | Rt[N*(N+1)] (R transposed) | Xs[N*(N+4)] | Ts[N*(N+1)]
template <int MODE>
__global__ void __launch_bounds__(256) plane_transform_smem(const void* __restrict__ in, float* __restrict__ out,
                                                            const float* __restrict__ L, const float* __restrict__ R,
                                                            long long planes, int N) {
  pdl_entry();
  extern __shared__ float sm[];
  const int XS = N + 4;  // keeps rows 16B aligned for float4 broadcast loads (N % 4 == 0 enforced by host)
  float* Ls = sm;
  float* Rt = Ls + N * N;
  float* Xs = Rt + N * (N + 1);
  float* Ts = Xs + N * XS;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int e = tid; e < N * N; e += nt) {
    int r = e / N, c = e % N;
    Ls[e] = L[e];
    Rt[c * (N + 1) + r] = R[e];
  }
  const int NN = N * N;
  for (long long p = blockIdx.x; p < planes; p += gridDim.x) {
    __syncthreads();
    for (int e = tid; e < NN; e += nt) Xs[(e / N) * XS + (e % N)] = load_in<MODE>(in, p * NN + e);
    __syncthreads();
    // T = X * R^T : T[i][j] = sum_k X[i][k] * R[j][k]
    for (int e = tid; e < NN; e += nt) {
      int i = e / N, j = e % N;
      float acc = 0.f;
      const float4* xr = (const float4*)(Xs + i * XS);
      for (int k4 = 0; k4 < N / 4; ++k4) {
        float4 xv = xr[k4];
        const float* rt = Rt + (k4 * 4) * (N + 1) + j;
        acc = fmaf(xv.x, rt[0], acc);
        acc = fmaf(xv.y, rt[N + 1], acc);
        acc = fmaf(xv.z, rt[2 * (N + 1)], acc);
        acc = fmaf(xv.w, rt[3 * (N + 1)], acc);
      }
      Ts[i * (N + 1) + j] = acc;
    }
    __syncthreads();
    // out = L * T : out[i][j] = sum_k L[i][k] * T[k][j]
    for (int e = tid; e < NN; e += nt) {
      int i = e / N, j = e % N;
      float acc = 0.f;
      const float4* lr = (const float4*)(Ls + i * N);
      for (int k4 = 0; k4 < N / 4; ++k4) {
        float4 lv = lr[k4];
        const float* tc = Ts + (k4 * 4) * (N + 1) + j;
        acc = fmaf(lv.x, tc[0], acc);
        acc = fmaf(lv.y, tc[N + 1], acc);
        acc = fmaf(lv.z, tc[2 * (N + 1)], acc);
        acc = fmaf(lv.w, tc[3 * (N + 1)], acc);
      }
      out[p * NN + e] = acc;
    }
  }
}

// generic fallback: SIDE 0: out[p] = X[p] * M^T ; SIDE 1: out[p] = M * X[p].  32x32 tiles.
template <int MODE, int SIDE>
__global__ void __launch_bounds__(256) plane_mm(const void* __restrict__ in, float* __restrict__ out,
                                                const float* __restrict__ M, int N) {
  pdl_entry();
  __shared__ float As[32][33], Bs[32][33];
  const long long p = blockIdx.z;
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const long long base = p * (long long)N * N;
  for (int k0 = 0; k0 < N; k0 += 32) {
    for (int r = ty; r < 32; r += 8) {
      int i = i0 + r, k = k0 + tx, j = j0 + r;
      if (SIDE == 0) {  // A = X[i][k], B = M[j][k]
        As[r][tx] = (i < N && k < N) ? load_in<MODE>(in, base + (long long)i * N + k) : 0.f;
        Bs[r][tx] = (j < N && k < N) ? M[(long long)j * N + k] : 0.f;
      } else {  // A = M[i][k], B^T: X[k][j] stored as Bs[kk][jj]
        As[r][tx] = (i < N && k < N) ? M[(long long)i * N + k] : 0.f;
        int kk = k0 + r, jj = j0 + tx;
        Bs[r][tx] = (kk < N && jj < N) ? load_in<MODE>(in, base + (long long)kk * N + jj) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int r = ty + 8 * q;
      float a = 0.f;
      for (int k = 0; k < 32; ++k) a = fmaf(As[r][k], SIDE == 0 ? Bs[tx][k] : Bs[k][tx], a);
      acc[q] += a;
    }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    int i = i0 + ty + 8 * q, j = j0 + tx;
    if (i < N && j < N) out[base + (long long)i * N + j] = acc[q];
  }
}

extern "C" int combat_plane_transform(const void* in, float* out, const float* L, const float* R, long long planes, int N,
                                      int in_mode, float* workspace, void* stream) {
  COMBAT_ARG(in && out && L && R, 0);
  COMBAT_ARG(N > 0 && in_mode >= 0 && in_mode <= 2, 5);
  if (planes <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (N <= 64 && N % 4 == 0) {
    size_t smem = sizeof(float) * (size_t)(N * N + N * (N + 1) + N * (N + 4) + N * (N + 1));
    int grid = (int)(planes < 148 * 8 ? planes : 148 * 8);
#define LAUNCH(MODE)                                                                                          \
  {                                                                                                           \
    cudaFuncSetAttribute(plane_transform_smem<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    pdl_launch(plane_transform_smem<MODE>, grid, 256, smem, st, in, out, L, R, planes, N);                            \
  }
    if (in_mode == 0) LAUNCH(0) else if (in_mode == 1) LAUNCH(1) else LAUNCH(2)
#undef LAUNCH
    COMBAT_RETURN_LAUNCH("plane_transform_smem");
  }
  COMBAT_ARG(workspace != nullptr, 7);
  COMBAT_ARG(planes <= 65535, 4);
  dim3 grid(cdiv(N, 32), cdiv(N, 32), (unsigned)planes);
  if (in_mode == 0)
    pdl_launch(plane_mm<0, 0>, grid, 256, 0, st, in, workspace, R, N);
  else if (in_mode == 1)
    pdl_launch(plane_mm<1, 0>, grid, 256, 0, st, in, workspace, R, N);
  else
    pdl_launch(plane_mm<2, 0>, grid, 256, 0, st, in, workspace, R, N);
  COMBAT_CHECK_LAUNCH("plane_mm<0>");
  pdl_launch(plane_mm<0, 1>, grid, 256, 0, st, workspace, out, L, N);
  COMBAT_RETURN_LAUNCH("plane_mm<1>");
}
