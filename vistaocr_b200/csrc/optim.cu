// Fused gradient clamp + Adam step over a flat fp32 parameter buffer (reference train_cnn_lstm.py:143-149,363:
// per-tensor grad.clamp_(-5,5) then torch.optim.Adam -> ~120 small launches; here one launch over all 17.3 M
// parameters).  Update rule = torch.optim.Adam (no amsgrad): L2 weight decay added to the gradient, bias-corrected
// first/second moments, eps added to sqrt(v_hat).  HBM-bound: 4 reads + 3 writes of 4 B per parameter.
#include "common.cuh"

namespace vocr {

__global__ void __launch_bounds__(256)
clamp_adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                  long long n, float lr, float beta1, float beta2, float eps, float weight_decay, float clamp,
                  float bias_c1, float bias_c2_sqrt, float grad_scale) {
  const float step_size = lr / bias_c1;
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* P = &pp.x;
    const float* G = &gg.x;
    float* M = &mm.x;
    float* V = &vv.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float gr = G[j] * grad_scale;
      if (clamp > 0.f) gr = fminf(fmaxf(gr, -clamp), clamp);
      if (weight_decay != 0.f) gr = fmaf(weight_decay, P[j], gr);
      M[j] = M[j] + (gr - M[j]) * (1.f - beta1);
      V[j] = fmaf(V[j], beta2, (1.f - beta2) * gr * gr);
      const float denom = sqrtf(V[j]) / bias_c2_sqrt + eps;
      P[j] = P[j] - step_size * (M[j] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  // tail
  for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float gr = g[i] * grad_scale;
    if (clamp > 0.f) gr = fminf(fmaxf(gr, -clamp), clamp);
    if (weight_decay != 0.f) gr = fmaf(weight_decay, p[i], gr);
    const float mi = m[i] + (gr - m[i]) * (1.f - beta1);
    const float vi = fmaf(v[i], beta2, (1.f - beta2) * gr * gr);
    m[i] = mi;
    v[i] = vi;
    p[i] = p[i] - step_size * (mi / (sqrtf(vi) / bias_c2_sqrt + eps));
  }
}

}  // namespace vocr

using namespace vocr;

// step >= 1.  clamp <= 0 disables the clamp.  grad_scale multiplies the gradient first (1/world_size for averaging).
extern "C" int vocr_clamp_adam_f32(float* p, const float* g, float* m, float* v, long long n, int step, float lr,
                                   float beta1, float beta2, float eps, float weight_decay, float clamp,
                                   float grad_scale, vocr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VOCR_REQUIRE(n >= 0 && step >= 1);
  if (n == 0) return VOCR_OK;
  VOCR_REQUIRE(p && g && m && v);
  VOCR_REQUIRE(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                 reinterpret_cast<uintptr_t>(v)) & 15) == 0);
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const int grid = (int)min((long long)kNumSMs * 8, ceil_div64(max(1ll, n / 4), 256));
  clamp_adam_kernel<<<grid, 256, 0, stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, clamp, (float)bc1,
                                              (float)sqrt(bc2), grad_scale);
  VOCR_CHECK_LAUNCH();
  return VOCR_OK;
}
