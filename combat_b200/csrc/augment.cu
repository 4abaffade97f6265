// PostTensorTransform (utils/dataloader.py:45-60 of the reference) as ONE gather kernel and its adjoint:
//   kornia RandomCrop(size=(H,W), padding=pad)  -> zero-pad by `pad`, take the H x W window at (ys, xs)          (:50-52)
//   kornia RandomRotation(degrees)              -> inverse-affine bilinear resampling about ((W-1)/2, (H-1)/2),
//                                                  zeros outside, align_corners=True                              (:53)
//   kornia RandomHorizontalFlip                 -> per-sample mirror of the columns                               (:55)
// applied in that order.  All random decisions (the two whole-batch `random.random()` gates of ProbTransform :17, the
// per-sample crop offsets, angles and flip bits) are drawn on the HOST (combat_b200/utils/dataloader.py) and arrive as
// one 8-float parameter row per image:
//   [0] crop shift x = xs - pad   [1] crop shift y = ys - pad      (0, 0 when the crop gate is off)
//   [2] cos(angle)                [3] sin(angle)                   [4] rotation on (0 / 1)
//   [5] flip (0 / 1)              [6], [7] unused
// out(n, c, y, x) = R(xf, y) with xf = flip ? W-1-x : x,
//   R(u, v)  = bilinear sample of the cropped image Cr at (sx, sy) = (a*(u-cx) - b*(v-cy) + cx, b*(u-cx) + a*(v-cy) + cy),
//   Cr(i, j) = 0 <= i < W, 0 <= j < H ? in(i + shift_x, j + shift_y) (zero outside the image) : 0.
// Images are NCHW float32 (the reference's tensors).  HBM-bound: 2 * C*H*W*4 bytes per image, a few MB per call.
#include "common.cuh"

namespace {

struct TfRow {
  float sx, sy, ca, sa, rot, flip;
};

__device__ __forceinline__ TfRow load_row(const float* __restrict__ params, int n) {
  const float4 a = __ldg((const float4*)(params + (long long)n * 8));
  const float2 b = __ldg((const float2*)(params + (long long)n * 8 + 4));
  TfRow r;
  r.sx = a.x; r.sy = a.y; r.ca = a.z; r.sa = a.w; r.rot = b.x; r.flip = b.y;
  return r;
}

// the (up to) four taps of output pixel (x, y): source offsets inside one H x W plane (or -1) and their weights
struct Taps {
  int off[4];
  float w[4];
};

__device__ __forceinline__ Taps make_taps(const TfRow& r, int x, int y, int H, int W) {
  Taps t;
  const int xf = r.flip != 0.f ? W - 1 - x : x;
  const int shx = (int)r.sx, shy = (int)r.sy;
  auto src_off = [&](int i, int j) -> int {  // cropped-image pixel (i, j) -> offset in the input plane
    if (i < 0 || i >= W || j < 0 || j >= H) return -1;
    const int px = i + shx, py = j + shy;
    if (px < 0 || px >= W || py < 0 || py >= H) return -1;
    return py * W + px;
  };
  if (r.rot == 0.f) {
    t.off[0] = src_off(xf, y);
    t.w[0] = 1.f;
    t.off[1] = t.off[2] = t.off[3] = -1;
    t.w[1] = t.w[2] = t.w[3] = 0.f;
    return t;
  }
  const float cx = 0.5f * (float)(W - 1), cy = 0.5f * (float)(H - 1);
  const float u = (float)xf - cx, v = (float)y - cy;
  const float fx = r.ca * u - r.sa * v + cx;
  const float fy = r.sa * u + r.ca * v + cy;
  const float x0f = floorf(fx), y0f = floorf(fy);
  const int x0 = (int)x0f, y0 = (int)y0f;
  const float wx1 = fx - x0f, wy1 = fy - y0f;
  const float wx0 = 1.f - wx1, wy0 = 1.f - wy1;
  t.off[0] = src_off(x0, y0);         t.w[0] = wx0 * wy0;
  t.off[1] = src_off(x0 + 1, y0);     t.w[1] = wx1 * wy0;
  t.off[2] = src_off(x0, y0 + 1);     t.w[2] = wx0 * wy1;
  t.off[3] = src_off(x0 + 1, y0 + 1); t.w[3] = wx1 * wy1;
  return t;
}

__global__ void __launch_bounds__(256) post_transform_fwd_k(const float* __restrict__ in, float* __restrict__ out,
                                                            const float* __restrict__ params, int rows, int C, int H, int W) {
  const long long total = (long long)rows * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const int y = (int)((i / W) % H);
    const int n = (int)(i / ((long long)W * H));
    const TfRow r = load_row(params, n);
    const Taps t = make_taps(r, x, y, H, W);
    const float* src = in + (long long)n * C * H * W;
    float* dst = out + (long long)n * C * H * W + y * W + x;
    for (int c = 0; c < C; ++c) {
      const float* pl = src + (long long)c * H * W;
      float v = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (t.off[k] >= 0) v = fmaf(t.w[k], __ldg(pl + t.off[k]), v);
      dst[(long long)c * H * W] = v;
    }
  }
}

// adjoint: din(src) += w * dout(dst) for every tap of every output pixel (din zero-filled by the entry point unless
// `accumulate`); float atomics -- at most a handful of contributions per input pixel
__global__ void __launch_bounds__(256) post_transform_bwd_k(const float* __restrict__ dout, float* __restrict__ din,
                                                            const float* __restrict__ params, int rows, int C, int H, int W) {
  const long long total = (long long)rows * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const int y = (int)((i / W) % H);
    const int n = (int)(i / ((long long)W * H));
    const TfRow r = load_row(params, n);
    const Taps t = make_taps(r, x, y, H, W);
    const float* g = dout + (long long)n * C * H * W + y * W + x;
    float* dst = din + (long long)n * C * H * W;
    for (int c = 0; c < C; ++c) {
      const float gv = __ldg(g + (long long)c * H * W);
      float* pl = dst + (long long)c * H * W;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (t.off[k] >= 0 && t.w[k] != 0.f) atomicAdd(pl + t.off[k], t.w[k] * gv);
    }
  }
}

}  // namespace

extern "C" int combat_post_transform_fwd(const float* in, float* out, const float* params, int rows, int C, int H, int W,
                                         void* stream) {
  COMBAT_ARG(in && out && params && in != out, 0);
  COMBAT_ARG(rows >= 0 && C > 0 && H > 0 && W > 0, 1);
  if (rows == 0) return 0;
  const long long total = (long long)rows * H * W;
  int grid = cdiv(total, 256);
  const int cap = resident_ctas(post_transform_fwd_k, 256) * 4;
  if (cap > 0 && grid > cap) grid = cap;
  post_transform_fwd_k<<<grid, 256, 0, (cudaStream_t)stream>>>(in, out, params, rows, C, H, W);
  COMBAT_RETURN_LAUNCH("post_transform_fwd");
}

extern "C" int combat_post_transform_bwd(const float* dout, float* din, const float* params, int rows, int C, int H, int W,
                                         int accumulate, void* stream) {
  COMBAT_ARG(dout && din && params && dout != din, 0);
  COMBAT_ARG(rows >= 0 && C > 0 && H > 0 && W > 0, 1);
  if (rows == 0) return 0;
  const long long total = (long long)rows * H * W;
  if (!accumulate) cudaMemsetAsync(din, 0, (size_t)total * C * sizeof(float), (cudaStream_t)stream);
  int grid = cdiv(total, 256);
  const int cap = resident_ctas(post_transform_bwd_k, 256) * 4;
  if (cap > 0 && grid > cap) grid = cap;
  post_transform_bwd_k<<<grid, 256, 0, (cudaStream_t)stream>>>(dout, din, params, rows, C, H, W);
  COMBAT_RETURN_LAUNCH("post_transform_bwd");
}
