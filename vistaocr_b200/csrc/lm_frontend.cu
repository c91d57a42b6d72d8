// LM-decode front end: fused log-softmax + model-alphabet -> LM-unit remap + per-line slicing,
// emitted in the float64 [len_b, |units|] layout the reference hands to the EESEN lattice decoder
// (reference src/decoder.py:61-101: torch log_softmax, np.full(log(1e-10)), fancy-index scatter per line on the host).
// One warp per valid frame (t < lens[b]): row log-sum-exp with warp shuffles, then every LM unit u gets
// logits[t,b,inv[u]] - lse (inv[u] = model index carrying that unit's string, -1 -> log(1e-10)).  Reads the logits once,
// writes sum(lens) * U * 8 bytes; padded frames are never touched.
#include "common.cuh"

namespace vocr {

__global__ void __launch_bounds__(256)
lm_frontend_kernel(const float* __restrict__ logits, int T, int B, int A, const int32_t* __restrict__ lens,
                   const long long* __restrict__ row_offsets, const int32_t* __restrict__ inv, int U, double fill,
                   double* __restrict__ out) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= (long long)T * B) return;
  const int t = (int)(row / B), b = (int)(row % B);
  if (t >= lens[b]) return;
  const int lane = threadIdx.x & 31;
  const float* x = logits + (size_t)row * A;
  float m = kNegInf;
  for (int a = lane; a < A; a += 32) m = fmaxf(m, __ldg(x + a));
  m = warp_max(m);
  float s = 0.f;
  for (int a = lane; a < A; a += 32) s += expf(__ldg(x + a) - m);
  s = warp_sum(s);
  const float lse = m + logf(s);
  double* o = out + (row_offsets[b] + t) * (long long)U;
  for (int u = lane; u < U; u += 32) {
    const int a = inv[u];
    o[u] = (a >= 0) ? (double)(__ldg(x + a) - lse) : fill;
  }
}

}  // namespace vocr

using namespace vocr;

// out: float64 [sum_b min(lens[b],T), U]; row_offsets[b] = first output row of line b (exclusive scan of the lengths).
extern "C" int vocr_lm_frontend_f32(const float* logits, int T, int B, int A, const int32_t* lens,
                                    const long long* row_offsets, const int32_t* inv, int U, double fill, double* out,
                                    vocr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VOCR_REQUIRE(T >= 0 && B >= 0 && A >= 1 && U >= 1);
  const long long rows = (long long)T * B;
  if (rows == 0) return VOCR_OK;
  VOCR_REQUIRE(logits && lens && row_offsets && inv && out);
  lm_frontend_kernel<<<(unsigned)ceil_div64(rows, 8), 256, 0, stream>>>(logits, T, B, A, lens, row_offsets, inv, U,
                                                                         fill, out);
  VOCR_CHECK_LAUNCH();
  return VOCR_OK;
}
