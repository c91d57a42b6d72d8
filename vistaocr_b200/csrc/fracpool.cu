// FractionalMaxPool2d(kernel 2, output_ratio (0.5, 0.7)) over NHWC fp32 (reference cnnlstm.py:127,130 -> ATen
// fractional_max_pool2d).  The pooling windows are pseudo-random PER (sample, channel), in training and in eval:
//   alpha = (in - 2) / (out - 1)  (fp32),  start_i = int((i + u) * alpha) - int(u * alpha)  for i < out-1,
//   start_{out-1} = in - 2,       u = samples[n, c, 0] for W and samples[n, c, 1] for H.
// The max scans the 2x2 window in (h, w) order with ATen's `val > max || isnan(val)` rule; the flat input index
// (h*W + w) of the winner is kept for the backward scatter (deterministic, see fracpool_bwd_kernel).
#include "common.cuh"

namespace vocr {

__device__ __forceinline__ int fmp_start(float u, int i, int in_size, int out_size) {
  if (i == out_size - 1) return in_size - 2;
  const float alpha = (float)(in_size - 2) / (float)(out_size - 1);
  // written as two separately rounded fp32 operations, exactly like ATen (no fma contraction is possible here)
  const float a = __fmul_rn(__fadd_rn((float)i, u), alpha);
  const float b = __fmul_rn(u, alpha);
  return (int)a - (int)b;
}

// grid = (chunks of Wo*C, Ho, B): the only division left per element is the 32-bit q / C
__global__ void __launch_bounds__(256)
fracpool_fwd_kernel(const float* __restrict__ x, const float* __restrict__ samples, float* __restrict__ y,
                    int32_t* __restrict__ idx, int B, int H, int W, int C, int Ho, int Wo) {
  const int ho = blockIdx.y, b = blockIdx.z;
  const unsigned row_elems = (unsigned)Wo * (unsigned)C;
  const float* xb = x + (size_t)b * H * W * C;
  const size_t orow = ((size_t)b * Ho + ho) * row_elems;
  for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < row_elems; q += gridDim.x * blockDim.x) {
    const unsigned wo = q / (unsigned)C;
    const int c = (int)(q - wo * (unsigned)C);
    const float* up = samples + ((size_t)b * C + c) * 2;  // (u_w, u_h)
    const int w0 = fmp_start(__ldg(up), (int)wo, W, Wo);
    const int h0 = fmp_start(__ldg(up + 1), ho, H, Ho);
    const int p00 = h0 * W + w0;
    const float* xp = xb + (size_t)p00 * C + c;
    const float v00 = __ldg(xp), v01 = __ldg(xp + C), v10 = __ldg(xp + (size_t)W * C), v11 = __ldg(xp + (size_t)(W + 1) * C);
    float best = -INFINITY;
    int bi = p00;
    // (h, w) scan order with ATen's `val > max || isnan(val)` rule
    if (v00 > best || v00 != v00) best = v00, bi = p00;
    if (v01 > best || v01 != v01) best = v01, bi = p00 + 1;
    if (v10 > best || v10 != v10) best = v10, bi = p00 + W;
    if (v11 > best || v11 != v11) best = v11, bi = p00 + W + 1;
    y[orow + q] = best;
    if (idx) idx[orow + q] = bi;
  }
}

// Backward: windows overlap (along W by one column whenever the start advances by 1), so the gradient is a scatter with
// collisions.  ATen resolves them with atomicAdd (run-to-run order, non-deterministic sums once three or more windows
// meet); here the scatter runs as FOUR passes over the (ho parity, wo parity) classes of the output: windows i and i+2
// of one axis never overlap (starts strictly increase), so inside a pass every input element is touched by at most
// one window and a plain read-modify-write is race free; the passes are stream ordered, i.e. the summation order is
// fixed: deterministic gradients, no atomics.
__global__ void __launch_bounds__(256)
fracpool_bwd_kernel(const float* __restrict__ dy, const int32_t* __restrict__ idx, float* __restrict__ dx, int HW, int C,
                    int Ho, int Wo, int ph, int pw) {
  // grid = (chunks of a row's selected windows x C, selected rows, B): 32-bit index arithmetic inside a sample
  const int b = blockIdx.z;
  const int ho = 2 * (int)blockIdx.y + ph;
  const unsigned nw = (unsigned)((Wo - pw + 1) >> 1);  // windows of this row in the pass: wo = pw, pw+2, ...
  const unsigned per = nw * (unsigned)C;
  const size_t o0 = ((size_t)b * Ho + ho) * (size_t)Wo * C;
  float* dxb = dx + (size_t)b * HW * C;
  for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < per; q += gridDim.x * blockDim.x) {
    const unsigned j = q / (unsigned)C;
    const unsigned c = q - j * (unsigned)C;
    const size_t o = o0 + (size_t)(2 * j + pw) * C + c;
    float* d = dxb + (size_t)__ldg(idx + o) * C + c;
    *d += __ldg(dy + o);
  }
}

}  // namespace vocr

using namespace vocr;

extern "C" int vocr_fracpool_fwd_f32(const float* x, const float* samples, float* y, int32_t* idx, int B, int H,
                                     int W, int C, int Ho, int Wo, vocr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VOCR_REQUIRE(B >= 0 && C > 0 && H >= 2 && W >= 2 && Ho >= 2 && Wo >= 2 && Ho <= H - 1 && Wo <= W - 1);
  const long long total = (long long)B * Ho * Wo * C;
  if (total == 0) return VOCR_OK;
  VOCR_REQUIRE(x && samples && y);
  VOCR_REQUIRE((long long)Wo * C < (1ll << 31) && Ho <= 65535 && B <= 65535);
  dim3 grid((unsigned)min(64ll, ceil_div64((long long)Wo * C, 256)), (unsigned)Ho, (unsigned)B);
  fracpool_fwd_kernel<<<grid, 256, 0, stream>>>(x, samples, y, idx, B, H, W, C, Ho, Wo);
  VOCR_CHECK_LAUNCH();
  return VOCR_OK;
}

extern "C" int vocr_fracpool_bwd_f32(const float* dy, const int32_t* idx, float* dx, int B, int H, int W, int C,
                                     int Ho, int Wo, vocr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VOCR_REQUIRE(B >= 0 && C > 0 && H >= 2 && W >= 2 && dx);
  if (cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)B * H * W * C, stream) != cudaSuccess)
    return VOCR_MEMOPS_FAILED;
  const long long total = (long long)B * Ho * Wo * C;
  if (total == 0) return VOCR_OK;
  VOCR_REQUIRE(dy && idx);
  VOCR_REQUIRE((long long)Ho * Wo * C < (1ll << 31) && B <= 65535 && Ho <= 2 * 65535);
  for (int ph = 0; ph < 2; ++ph)
    for (int pw = 0; pw < 2; ++pw) {
      const int rows = (Ho - ph + 1) / 2, nw = (Wo - pw + 1) / 2;
      if (rows <= 0 || nw <= 0) continue;
      dim3 grid((unsigned)min(16ll, ceil_div64((long long)nw * C, 256)), (unsigned)rows, (unsigned)B);
      fracpool_bwd_kernel<<<grid, 256, 0, stream>>>(dy, idx, dx, H * W, C, Ho, Wo, ph, pw);
      VOCR_CHECK_LAUNCH();
    }
  return VOCR_OK;
}
