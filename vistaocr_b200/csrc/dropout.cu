// Inter-layer LSTM dropout (reference cnnlstm.py:148-149: nn.LSTM(..., dropout=p), p = 0.5 from train_cnn_lstm.py:331;
// applied in training to the output of every layer but the last) as ONE elementwise kernel whose keep-mask is either
//   * generated in the kernel from a counter-based Philox4x32-10 stream (nothing is stored: the backward pass
//     regenerates the same mask from the same (seed, offset)), or
//   * injected by the caller (uint8 keep mask) - how the parity tests feed the oracle's mask to both sides.
// y = x * keep * 1/(1-p), the same single fp32 multiply torch's dropout performs on kept elements.
// The generator state may live on the DEVICE (rng[0] = seed, rng[1] = offset) so that a CUDA graph replays the step with
// a fresh mask each time (vocr_rng_advance bumps the offset inside the graph).
//
// Element i uses word (i & 3) of Philox4x32-10(counter = {lo32(i>>2), hi32(i>>2), lo32(offset), hi32(offset)},
// key = {lo32(seed), hi32(seed)}) and is KEPT iff word >= floor(p * 2^32).  oracle/philox_ref.py restates this.
#include "common.cuh"

namespace vocr {

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&out)[4]) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// one thread per group of 4 consecutive elements (one Philox block)
__global__ void __launch_bounds__(256)
dropout_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, float scale, uint32_t thresh,
               const unsigned long long* __restrict__ rng, unsigned long long seed, unsigned long long offset,
               unsigned long long* __restrict__ rng_used, const uint8_t* __restrict__ mask_in,
               uint8_t* __restrict__ mask_out) {
  if (rng) {
    seed = rng[0];
    offset += rng[1];
  }
  if (rng_used && blockIdx.x == 0 && threadIdx.x == 0) {
    rng_used[0] = seed;
    rng_used[1] = offset;
  }
  const long long groups = (n + 3) >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const bool vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
  for (long long gidx = (long long)blockIdx.x * blockDim.x + threadIdx.x; gidx < groups; gidx += stride) {
    const long long i0 = gidx << 2;
    bool keep[4];
    if (mask_in) {
#pragma unroll
      for (int j = 0; j < 4; ++j) keep[j] = (i0 + j < n) && mask_in[i0 + j] != 0;
    } else {
      uint32_t r[4];
      philox4x32_10((uint32_t)gidx, (uint32_t)((unsigned long long)gidx >> 32), (uint32_t)offset,
                    (uint32_t)(offset >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), r);
#pragma unroll
      for (int j = 0; j < 4; ++j) keep[j] = r[j] >= thresh;
    }
    if (x) {
      if (vec && i0 + 4 <= n) {
        float4 v = __ldg(reinterpret_cast<const float4*>(x + i0));
        v.x = keep[0] ? v.x * scale : 0.f;
        v.y = keep[1] ? v.y * scale : 0.f;
        v.z = keep[2] ? v.z * scale : 0.f;
        v.w = keep[3] ? v.w * scale : 0.f;
        *reinterpret_cast<float4*>(y + i0) = v;
      } else {
        for (int j = 0; j < 4 && i0 + j < n; ++j) y[i0 + j] = keep[j] ? x[i0 + j] * scale : 0.f;
      }
    }
    if (mask_out)
      for (int j = 0; j < 4 && i0 + j < n; ++j) mask_out[i0 + j] = keep[j] ? 1 : 0;
  }
}

__global__ void rng_advance_kernel(unsigned long long* rng, unsigned long long inc) { rng[1] += inc; }

}  // namespace vocr

using namespace vocr;

extern "C" int vocr_dropout_f32(const float* x, float* y, long long n, float p, const unsigned long long* rng,
                                unsigned long long seed, unsigned long long offset, unsigned long long* rng_used,
                                const uint8_t* mask_in, uint8_t* mask_out, vocr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VOCR_REQUIRE(n >= 0 && p >= 0.f && p < 1.f);
  if (n == 0) return VOCR_OK;
  VOCR_REQUIRE((x && y) || (!x && !y && mask_out));
  const float scale = 1.f / (1.f - p);
  const double t = (double)p * 4294967296.0;
  const uint32_t thresh = t >= 4294967295.0 ? 0xffffffffu : (uint32_t)t;
  const int grid = (int)min((long long)kNumSMs * 8, ceil_div64((n + 3) / 4, 256));
  dropout_kernel<<<grid, 256, 0, stream>>>(x, y, n, scale, thresh, rng, seed, offset, rng_used, mask_in, mask_out);
  VOCR_CHECK_LAUNCH();
  return VOCR_OK;
}

extern "C" int vocr_rng_advance(unsigned long long* rng, unsigned long long inc, vocr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VOCR_REQUIRE(rng);
  rng_advance_kernel<<<1, 1, 0, stream>>>(rng, inc);
  VOCR_CHECK_LAUNCH();
  return VOCR_OK;
}
