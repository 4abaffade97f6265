// Cross-entropy (mean) forward + backward + argmax metrics as warp-shuffle reductions, and the
// deterministic final reduction of the fused MSE partial sums.
// Replaces torch.nn.CrossEntropyLoss / MSELoss at train_generator.py:162-164,207,231,234,251 and the
// argmax/eq/sum metric ops at :262-267.
#include "common.cuh"

// one warp per sample row; a single CTA so that the mean is reduced in a fixed order (deterministic)
__global__ void __launch_bounds__(1024) cross_entropy_k(const float* __restrict__ logits, const long long* __restrict__ tgt,
                                                        const long long* __restrict__ tgt2, int B, int C, float gscale,
                                                        float* __restrict__ loss_out, float* __restrict__ dlogits,
                                                        int* __restrict__ counts) {
  pdl_entry();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  float loss_acc = 0.f;
  int c1 = 0, c2 = 0;
  const float invB = 1.f / (float)B;
  if (C <= 16) {
    // few classes (10 / 8 on this path): one THREAD per row, all loads independent -- the warp-per-row form below spends
    // three dependent global round trips per row on ten values
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
      const float* row = logits + (long long)b * C;
      float v[16];
      float m = -INFINITY;
      int am = 0;
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        v[c] = c < C ? row[c] : -INFINITY;
        if (v[c] > m) { m = v[c]; am = c; }  // first index wins ties
      }
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < 16; ++c)
        if (c < C) s += expf(v[c] - m);
      const float lse = m + logf(s);
      const int t = (int)tgt[b];  // t < 0: row ignored (no loss, no gradient, never counted) -- eval() masks rows this way
      if (dlogits) {
#pragma unroll
        for (int c = 0; c < 16; ++c)
          if (c < C) dlogits[(long long)b * C + c] = t < 0 ? 0.f : gscale * invB * (expf(v[c] - lse) - (c == t ? 1.f : 0.f));
      }
      float vt = lse;
#pragma unroll
      for (int c = 0; c < 16; ++c)
        if (c == t) vt = v[c];
      loss_acc += lse - vt;
      c1 += (am == t);
      if (tgt2) c2 += (am == (int)tgt2[b]);
    }
    // fixed-order reduction: lanes -> warps -> thread 0
    loss_acc = warp_sum(loss_acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      c1 += __shfl_xor_sync(0xffffffffu, c1, o);
      c2 += __shfl_xor_sync(0xffffffffu, c2, o);
    }
  } else
  for (int b = warp; b < B; b += nwarp) {
    const float* row = logits + (long long)b * C;
    float m = -INFINITY;
    int am = 0;
    for (int c = lane; c < C; c += 32) {
      float v = row[c];
      if (v > m) { m = v; am = c; }
    }
    // warp argmax, first index wins ties (torch.argmax semantics on equal values: lowest index)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float om = __shfl_xor_sync(0xffffffffu, m, o);
      int oa = __shfl_xor_sync(0xffffffffu, am, o);
      if (om > m || (om == m && oa < am)) { m = om; am = oa; }
    }
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += expf(row[c] - m);
    s = warp_sum(s);
    const float lse = m + logf(s);
    const int t = (int)tgt[b];  // t < 0: row ignored
    if (dlogits) {
      for (int c = lane; c < C; c += 32) {
        float p = expf(row[c] - lse);
        dlogits[(long long)b * C + c] = t < 0 ? 0.f : gscale * invB * (p - (c == t ? 1.f : 0.f));
      }
    }
    if (lane == 0) {
      if (t >= 0) loss_acc += lse - row[t];
      c1 += (am == t);
      if (tgt2) c2 += (am == (int)tgt2[b]);
    }
  }
  __shared__ float sl[32];
  __shared__ int s1[32], s2[32];
  if (lane == 0) { sl[warp] = loss_acc; s1[warp] = c1; s2[warp] = c2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float L = 0.f;
    int a = 0, bb = 0;
    for (int w = 0; w < nwarp; ++w) { L += sl[w]; a += s1[w]; bb += s2[w]; }
    if (loss_out) loss_out[0] = L * invB;
    if (counts) { counts[0] = a; counts[1] = bb; }
  }
}

__global__ void __launch_bounds__(256) sum_scale_k(const float* __restrict__ partial, int n, float scale,
                                                   float* __restrict__ out) {
  pdl_entry();
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += (double)partial[i];
  __shared__ double red[256];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (float)(red[0] * (double)scale);
}

extern "C" int combat_cross_entropy(const float* logits, const long long* targets, const long long* targets2, int B, int C,
                                    float grad_scale, float* loss_out, float* dlogits, int* counts_out, void* stream) {
  COMBAT_ARG(logits && targets, 0);
  COMBAT_ARG(B > 0 && C > 0, 3);
  pdl_launch(cross_entropy_k, 1, 1024, 0, (cudaStream_t)stream, logits, targets, targets2, B, C, grad_scale, loss_out, dlogits,
                                                        counts_out);
  COMBAT_RETURN_LAUNCH("cross_entropy");
}

extern "C" int combat_sum_scale(const float* partial, int n, float scale, float* out, void* stream) {
  COMBAT_ARG(partial && out && n > 0, 0);
  pdl_launch(sum_scale_k, 1, 256, 0, (cudaStream_t)stream, partial, n, scale, out);
  COMBAT_RETURN_LAUNCH("sum_scale");
}
