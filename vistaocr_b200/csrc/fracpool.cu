// FractionalMaxPool2d(kernel 2, output_ratio (0.5, 0.7)) over NHWC fp32 (reference cnnlstm.py:127,130 -> ATen
// fractional_max_pool2d).  The pooling windows are pseudo-random PER (sample, channel), in training and in eval:
//   alpha = (in - 2) / (out - 1)  (fp32),  start_i = int((i + u) * alpha) - int(u * alpha)  for i < out-1,
//   start_{out-1} = in - 2,       u = samples[n, c, 0] for W and samples[n, c, 1] for H.
// The max scans the 2x2 window in (h, w) order with ATen's `val > max || isnan(val)` rule; the flat input index
// (h*W + w) of the winner is kept for the backward scatter.  Windows overlap along W, so backward is an atomic
// scatter-add.
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

__global__ void __launch_bounds__(256)
fracpool_fwd_kernel(const float* __restrict__ x, const float* __restrict__ samples, float* __restrict__ y,
                    int32_t* __restrict__ idx, int B, int H, int W, int C, int Ho, int Wo) {
  const long long total = (long long)B * Ho * Wo * C;
  for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total;
       o += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(o % C);
    const int wo = (int)((o / C) % Wo);
    const int ho = (int)((o / ((long long)C * Wo)) % Ho);
    const int b = (int)(o / ((long long)C * Wo * Ho));
    const float uw = __ldg(samples + ((size_t)b * C + c) * 2 + 0);
    const float uh = __ldg(samples + ((size_t)b * C + c) * 2 + 1);
    const int w0 = fmp_start(uw, wo, W, Wo);
    const int h0 = fmp_start(uh, ho, H, Ho);
    const float* xb = x + (size_t)b * H * W * C + c;
    float best = -INFINITY;
    int bi = h0 * W + w0;
#pragma unroll
    for (int dh = 0; dh < 2; ++dh)
#pragma unroll
      for (int dw = 0; dw < 2; ++dw) {
        const int p = (h0 + dh) * W + (w0 + dw);
        const float v = __ldg(xb + (size_t)p * C);
        if (v > best || v != v) {
          best = v;
          bi = p;
        }
      }
    y[o] = best;
    if (idx) idx[o] = bi;
  }
}

__global__ void __launch_bounds__(256)
fracpool_bwd_kernel(const float* __restrict__ dy, const int32_t* __restrict__ idx, float* __restrict__ dx, int B,
                    int HW, int C, int HoWo) {
  const long long total = (long long)B * HoWo * C;
  for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total;
       o += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(o % C);
    const int b = (int)(o / ((long long)C * HoWo));
    atomicAdd(dx + ((size_t)b * HW + idx[o]) * C + c, dy[o]);
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
  const int grid = (int)min((long long)kNumSMs * 16, ceil_div64(total, 256));
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
  const int grid = (int)min((long long)kNumSMs * 16, ceil_div64(total, 256));
  fracpool_bwd_kernel<<<grid, 256, 0, stream>>>(dy, idx, dx, B, H * W, C, Ho * Wo);
  VOCR_CHECK_LAUNCH();
  return VOCR_OK;
}
