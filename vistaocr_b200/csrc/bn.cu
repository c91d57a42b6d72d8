// BatchNorm2d (+ReLU) over NHWC fp32 activations [P = B*H*W pixels, C channels] (reference cnnlstm.py:263-266:
// nn.BatchNorm2d(eps 1e-5, momentum 0.1) + nn.ReLU).  The per-channel sum / sum of squares come out of the conv
// epilogue (conv.cu) in float64; here: finalize (batch or running statistics -> scale/shift, running-stat update),
// apply+ReLU (strided output so the last block can write the [T,B,h*C] sequence layout directly), and the backward
// pass (masked reductions, then dz and the conv-bias gradient).  HBM-bound elementwise / reduction kernels: float4
// along C, grids sized in multiples of the SM count.
#include "common.cuh"
#include "pair_f16.cuh"

namespace vocr {

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t t;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(x));
  return __uint_as_float(t);
}

__global__ void bn_finalize_kernel(const double* __restrict__ stats, double count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, float momentum, float eps, int training,
                                   float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ save_mean, float* __restrict__ save_invstd, int C,
                                   unsigned* __restrict__ aux, const float* __restrict__ zmax) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mean, invstd;
  if (training) {
    const double m = stats[c] / count;
    double var = stats[C + c] / count - m * m;  // biased batch variance (used for normalisation)
    if (var < 0.0) var = 0.0;
    mean = (float)m;
    invstd = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) {
      const double unbiased = (count > 1.0) ? var * count / (count - 1.0) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
  } else {
    mean = running_mean[c];
    invstd = 1.0f / sqrtf(running_var[c] + eps);
  }
  const float sc = gamma[c] * invstd;
  scale[c] = sc;
  shift[c] = beta[c] - mean * sc;
  if (save_mean) save_mean[c] = mean;
  if (save_invstd) save_invstd[c] = invstd;
  if (aux) {
    // aux[0]: an upper bound of the activations relu(gamma * xhat + beta): with batch statistics |xhat| <= sqrt(count)
    // (one value cannot exceed the whole sum of squares).  Loose by ~2^9, which the FP16 pair format absorbs
    // (pair_f16.cuh).  Not available with running statistics (0 -> the consumer computes an absmax).
    // aux[1]: max_c |scale_c| (bounds the backward pass's dz).  Non-negative floats order like their bit patterns.
    // With running statistics the bound comes from the convolution's measured max |z|: |a| <= |scale| zmax + |shift|.
    if (training) atomicMax(&aux[0], __float_as_uint(fabsf(gamma[c]) * sqrtf((float)count) + fabsf(beta[c])));
    else if (zmax) atomicMax(&aux[0], __float_as_uint(fmaf(fabsf(sc), __ldg(zmax), fabsf(beta[c] - mean * sc))));
    atomicMax(&aux[1], __float_as_uint(fabsf(sc)));
  }
}

// a[b,y,x,c] (strided) = relu(z[p,c]*scale[c] + shift[c]).  IDX = unsigned when P*C4 < 2^31: five 64-bit divisions per
// float4 made this streaming kernel instruction-bound (3 TB/s); 32-bit index arithmetic is ~5x cheaper.
template <typename IDX>
__global__ void __launch_bounds__(256)
bn_relu_apply_kernel(const float* __restrict__ z, const float* __restrict__ scale, const float* __restrict__ shift,
                     float* __restrict__ a, float* __restrict__ a_hi, float* __restrict__ a_lo, long long P, int H,
                     int W, int C4, long long sB, long long sH, long long sW, __half* __restrict__ a_hi16,
                     __half* __restrict__ a_lo16, const unsigned* __restrict__ bound_bits, int* __restrict__ exp_out) {
  const long long total = P * C4;
  float psc = 1.f;  // 2^e of the FP16 pair planes
  if (a_hi16) {
    const int e = pair_exponent(__ldg(bound_bits));
    psc = exp2i(e);
    if (blockIdx.x == 0 && threadIdx.x == 0) *exp_out = e;
  }
  for (IDX idx = (IDX)blockIdx.x * blockDim.x + threadIdx.x; idx < (IDX)total; idx += (IDX)gridDim.x * blockDim.x) {
    const IDX p = idx / (IDX)C4;
    const int c4 = (int)(idx - p * (IDX)C4);
    const IDX row = p / (IDX)W;
    const int x = (int)(p - row * (IDX)W);
    const long long b = (long long)(row / (IDX)H);
    const int y = (int)(row - (IDX)b * (IDX)H);
    const float4 v = __ldg(reinterpret_cast<const float4*>(z) + idx);
    const float4 sc = __ldg(reinterpret_cast<const float4*>(scale) + c4);
    const float4 sh = __ldg(reinterpret_cast<const float4*>(shift) + c4);
    float4 r;
    r.x = fmaxf(fmaf(v.x, sc.x, sh.x), 0.f);
    r.y = fmaxf(fmaf(v.y, sc.y, sh.y), 0.f);
    r.z = fmaxf(fmaf(v.z, sc.z, sh.z), 0.f);
    r.w = fmaxf(fmaf(v.w, sc.w, sh.w), 0.f);
    const long long o = b * sB + y * sH + x * sW + (long long)c4 * 4;
    if (a) *reinterpret_cast<float4*>(a + o) = r;
    if (a_hi) {  // TF32 (hi, lo) planes for the tensor-core conv that consumes this activation
      float4 h, l;
      h.x = tf32_rna(r.x); l.x = tf32_rna(r.x - h.x);
      h.y = tf32_rna(r.y); l.y = tf32_rna(r.y - h.y);
      h.z = tf32_rna(r.z); l.z = tf32_rna(r.z - h.z);
      h.w = tf32_rna(r.w); l.w = tf32_rna(r.w - h.w);
      *reinterpret_cast<float4*>(a_hi + o) = h;
      *reinterpret_cast<float4*>(a_lo + o) = l;
    }
    if (a_hi16) {  // FP16 pair planes (kind::f16 kernels)
      uint2 ph, pl;
      pair_pack4(r.x * psc, r.y * psc, r.z * psc, r.w * psc, ph, pl);
      *reinterpret_cast<uint2*>(a_hi16 + o) = ph;
      *reinterpret_cast<uint2*>(a_lo16 + o) = pl;
    }
  }
}

// Backward pass 1: per-channel  s1 = sum g,  s2 = sum g * xhat,  g = da * [z*scale+shift > 0],
// xhat = (z - mean) * invstd.  block = 256 threads = (256/C4) pixel rows x C4 float4 columns.
__global__ void __launch_bounds__(256)
bn_relu_bwd_reduce_kernel(const float* __restrict__ da, const float* __restrict__ z,
                          const float* __restrict__ scale, const float* __restrict__ shift,
                          const float* __restrict__ mean, const float* __restrict__ invstd, long long P, int H,
                          int W, int C4, long long sB, long long sH, long long sW, long long rows_per_cta,
                          double* __restrict__ red, unsigned* __restrict__ gmax_bits) {
  extern __shared__ float s_red[];  // [rows][C4*8]
  float gmax = 0.f;
  const int rows = 256 / C4;
  const int c4 = threadIdx.x % C4, r = threadIdx.x / C4;
  const float4 sc = __ldg(reinterpret_cast<const float4*>(scale) + c4);
  const float4 sh = __ldg(reinterpret_cast<const float4*>(shift) + c4);
  const float4 mu = __ldg(reinterpret_cast<const float4*>(mean) + c4);
  const float4 is = __ldg(reinterpret_cast<const float4*>(invstd) + c4);
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  const long long p0 = (long long)blockIdx.x * rows_per_cta, p1 = min(P, p0 + rows_per_cta);
  if (r < rows) {
    for (long long p = p0 + r; p < p1; p += rows) {
      const int x = (int)(p % W);
      const int y = (int)((p / W) % H);
      const long long b = p / ((long long)W * H);
      const float4 v = __ldg(reinterpret_cast<const float4*>(z) + p * C4 + c4);
      const float4 d = __ldg(reinterpret_cast<const float4*>(da + b * sB + y * sH + x * sW + (long long)c4 * 4));
      const float g0 = fmaf(v.x, sc.x, sh.x) > 0.f ? d.x : 0.f;
      const float g1 = fmaf(v.y, sc.y, sh.y) > 0.f ? d.y : 0.f;
      const float g2 = fmaf(v.z, sc.z, sh.z) > 0.f ? d.z : 0.f;
      const float g3 = fmaf(v.w, sc.w, sh.w) > 0.f ? d.w : 0.f;
      s1[0] += g0; s1[1] += g1; s1[2] += g2; s1[3] += g3;
      gmax = fmaxf(gmax, fmaxf(fmaxf(fabsf(g0), fabsf(g1)), fmaxf(fabsf(g2), fabsf(g3))));
      s2[0] = fmaf(g0, (v.x - mu.x) * is.x, s2[0]);
      s2[1] = fmaf(g1, (v.y - mu.y) * is.y, s2[1]);
      s2[2] = fmaf(g2, (v.z - mu.z) * is.z, s2[2]);
      s2[3] = fmaf(g3, (v.w - mu.w) * is.w, s2[3]);
    }
  }
  float* mine = s_red + (size_t)threadIdx.x * 8;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    mine[i] = s1[i];
    mine[4 + i] = s2[i];
  }
  __syncthreads();
  // first C4*8 threads... C4*8 may exceed 256 (C=256 -> 512): loop
  const int C = C4 * 4;
  for (int o = threadIdx.x; o < 2 * C; o += 256) {
    const int which = o / C, c = o % C;  // which: 0 -> s1, 1 -> s2
    double t = 0.0;
    for (int rr = 0; rr < rows; ++rr) t += (double)s_red[((size_t)rr * C4 + (c >> 2)) * 8 + which * 4 + (c & 3)];
    atomicAdd(&red[which * C + c], t);
  }
  if (gmax_bits) {  // max |g| over the tensor: bounds dz for the FP16 pair planes of pass 2
    const unsigned b = __reduce_max_sync(0xffffffffu, __float_as_uint(gmax));
    if ((threadIdx.x & 31) == 0 && b) atomicMax(gmax_bits, b);
  }
}

// Backward pass 2: dz = scale * (g - s1/N - xhat * s2/N)  (training)  or  scale * g  (eval);  also sums dz per channel
// (the gradient of the conv bias that precedes the BatchNorm) into red_bias (float64).
__global__ void __launch_bounds__(256)
bn_relu_bwd_apply_kernel(const float* __restrict__ da, const float* __restrict__ z, const float* __restrict__ scale,
                         const float* __restrict__ shift, const float* __restrict__ mean,
                         const float* __restrict__ invstd, const double* __restrict__ red, double inv_count,
                         int training, long long P, int H, int W, int C4, long long sB, long long sH, long long sW,
                         long long rows_per_cta, float* __restrict__ dz, float* __restrict__ dz_hi,
                         float* __restrict__ dz_lo, double* __restrict__ red_bias, __half* __restrict__ dz_hi16,
                         __half* __restrict__ dz_lo16, const unsigned* __restrict__ scmax_bits,
                         int* __restrict__ state) {
  extern __shared__ float s_red[];  // [256][4]
  float psc = 1.f;
  if (dz_hi16) {
    // |dz| <= max|scale| * max|g| * (2 + sqrt(N)) with batch statistics (|mean g| <= G, |mean g xhat| <= G,
    // |xhat| <= sqrt(N)), max|scale| * max|g| with running statistics
    const float G = __uint_as_float(*reinterpret_cast<const unsigned*>(state + 1));
    const float bound = __uint_as_float(__ldg(scmax_bits)) * G * (training ? 2.f + sqrtf((float)P) : 1.f);
    const int e = pair_exponent(__float_as_uint(fminf(bound, 3.0e38f)));
    psc = exp2i(e);
    if (blockIdx.x == 0 && threadIdx.x == 0) state[0] = e;
  }
  const int rows = 256 / C4;
  const int c4 = threadIdx.x % C4, r = threadIdx.x / C4;
  const int C = C4 * 4;
  const float4 sc = __ldg(reinterpret_cast<const float4*>(scale) + c4);
  const float4 sh = __ldg(reinterpret_cast<const float4*>(shift) + c4);
  const float4 mu = __ldg(reinterpret_cast<const float4*>(mean) + c4);
  const float4 is = __ldg(reinterpret_cast<const float4*>(invstd) + c4);
  float m1[4] = {0.f, 0.f, 0.f, 0.f}, m2[4] = {0.f, 0.f, 0.f, 0.f};
  if (training) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      m1[i] = (float)(red[c4 * 4 + i] * inv_count);
      m2[i] = (float)(red[C + c4 * 4 + i] * inv_count);
    }
  }
  float sb[4] = {0.f, 0.f, 0.f, 0.f};
  const long long p0 = (long long)blockIdx.x * rows_per_cta, p1 = min(P, p0 + rows_per_cta);
  if (r < rows) {
    for (long long p = p0 + r; p < p1; p += rows) {
      const int x = (int)(p % W);
      const int y = (int)((p / W) % H);
      const long long b = p / ((long long)W * H);
      const float4 v = __ldg(reinterpret_cast<const float4*>(z) + p * C4 + c4);
      const float4 d = __ldg(reinterpret_cast<const float4*>(da + b * sB + y * sH + x * sW + (long long)c4 * 4));
      const float g0 = fmaf(v.x, sc.x, sh.x) > 0.f ? d.x : 0.f;
      const float g1 = fmaf(v.y, sc.y, sh.y) > 0.f ? d.y : 0.f;
      const float g2 = fmaf(v.z, sc.z, sh.z) > 0.f ? d.z : 0.f;
      const float g3 = fmaf(v.w, sc.w, sh.w) > 0.f ? d.w : 0.f;
      float4 o;
      o.x = sc.x * (g0 - m1[0] - (v.x - mu.x) * is.x * m2[0]);
      o.y = sc.y * (g1 - m1[1] - (v.y - mu.y) * is.y * m2[1]);
      o.z = sc.z * (g2 - m1[2] - (v.z - mu.z) * is.z * m2[2]);
      o.w = sc.w * (g3 - m1[3] - (v.w - mu.w) * is.w * m2[3]);
      sb[0] += o.x; sb[1] += o.y; sb[2] += o.z; sb[3] += o.w;
      if (dz) *(reinterpret_cast<float4*>(dz) + p * C4 + c4) = o;  // (null: the consumers read the FP16 pair planes only)
      if (dz_hi) {
        float4 h, l;
        h.x = tf32_rna(o.x); l.x = tf32_rna(o.x - h.x);
        h.y = tf32_rna(o.y); l.y = tf32_rna(o.y - h.y);
        h.z = tf32_rna(o.z); l.z = tf32_rna(o.z - h.z);
        h.w = tf32_rna(o.w); l.w = tf32_rna(o.w - h.w);
        *(reinterpret_cast<float4*>(dz_hi) + p * C4 + c4) = h;
        *(reinterpret_cast<float4*>(dz_lo) + p * C4 + c4) = l;
      }
      if (dz_hi16) {
        uint2 ph, pl;
        pair_pack4(o.x * psc, o.y * psc, o.z * psc, o.w * psc, ph, pl);
        *(reinterpret_cast<uint2*>(dz_hi16) + p * C4 + c4) = ph;
        *(reinterpret_cast<uint2*>(dz_lo16) + p * C4 + c4) = pl;
      }
    }
  }
  if (red_bias) {
    float* mine = s_red + (size_t)threadIdx.x * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) mine[i] = sb[i];
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
      double t = 0.0;
      for (int rr = 0; rr < rows; ++rr) t += (double)s_red[((size_t)rr * C4 + (c >> 2)) * 4 + (c & 3)];
      atomicAdd(&red_bias[c], t);
    }
  }
}

__global__ void f64_to_f32_kernel(const double* __restrict__ in, float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)in[i];
}

}  // namespace vocr

using namespace vocr;

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int vocr_bn_finalize_f32(const double* stats, long long count, const float* gamma, const float* beta,
                                    float* running_mean, float* running_var, float momentum, float eps,
                                    int training, float* scale, float* shift, float* save_mean, float* save_invstd,
                                    int C, float* aux, const float* zmax, vocr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VOCR_REQUIRE(C > 0 && gamma && beta && scale && shift);
  VOCR_REQUIRE(training ? (stats != nullptr && count > 0) : (running_mean && running_var));
  if (aux && cudaMemsetAsync(aux, 0, 2 * sizeof(float), stream) != cudaSuccess) return VOCR_MEMOPS_FAILED;
  bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, stream>>>(stats, (double)count, gamma, beta, running_mean,
                                                           running_var, momentum, eps, training, scale, shift,
                                                           save_mean, save_invstd, C,
                                                           reinterpret_cast<unsigned*>(aux), zmax);
  VOCR_CHECK_LAUNCH();
  return VOCR_OK;
}

// Upper bound of a = relu((conv(x) + bias) * scale + shift) from the weights alone:
//   |a_c| <= |scale_c| (sum_k |w_ck| * xbound + |bias_c|) + |shift_c|,      aux[0] = max_c of that (atomicMax on the bits)
// One warp per output channel.  Loose by the usual sum-of-magnitudes factor, which the FP16 pair format absorbs
// (pair_f16.cuh); it lets the fused convolution epilogue emit the next layer's operand planes without seeing max |z|.
__global__ void __launch_bounds__(256)
bn_eval_bound_kernel(const float* __restrict__ w, const float* __restrict__ bias, const float* __restrict__ scale,
                     const float* __restrict__ shift, const float* __restrict__ xbound, int C, int K,
                     unsigned* __restrict__ aux) {
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (c >= C) return;
  float s = 0.f;
  for (int k = lane; k < K; k += 32) s += fabsf(__ldg(w + (size_t)c * K + k));
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) {
    const float zb = fmaf(s * 1.0001f, __ldg(xbound), bias ? fabsf(__ldg(bias + c)) : 0.f);  // (margin for the fp32 sum)
    atomicMax(&aux[0], __float_as_uint(fmaf(fabsf(__ldg(scale + c)), zb, fabsf(__ldg(shift + c))) * 1.0001f));
  }
}

// w [C][K] (K = 9 * Cin; any order inside a row), bias [C] or NULL, scale / shift [C], xbound: device scalar >= max |x|,
// aux[0] (zeroed by vocr_bn_finalize_f32) receives the bound.
extern "C" int vocr_bn_eval_bound_f32(const float* w, const float* bias, const float* scale, const float* shift,
                                      const float* xbound, int C, int K, float* aux, vocr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VOCR_REQUIRE(w && scale && shift && xbound && aux && C > 0 && K > 0);
  bn_eval_bound_kernel<<<ceil_div(C, 8), 256, 0, stream>>>(w, bias, scale, shift, xbound, C, K,
                                                          reinterpret_cast<unsigned*>(aux));
  VOCR_CHECK_LAUNCH();
  return VOCR_OK;
}

extern "C" int vocr_bn_relu_apply_f32(const float* z, const float* scale, const float* shift, float* a, float* a_hi,
                                      float* a_lo, int B, int H, int W, int C, long long sB, long long sH,
                                      long long sW, uint16_t* a_hi16, uint16_t* a_lo16, const float* bound,
                                      int32_t* pair_exp, vocr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const long long P = (long long)B * H * W;
  if (P == 0) return VOCR_OK;
  VOCR_REQUIRE(z && scale && shift && (a || a_hi16) && C > 0 && C % 4 == 0);  // a may be NULL: planes only
  VOCR_REQUIRE(aligned16(z) && (!a || aligned16(a)) && aligned16(scale) && aligned16(shift));
  VOCR_REQUIRE(sB % 4 == 0 && sH % 4 == 0 && sW % 4 == 0);
  const long long total = P * (C / 4);
  const int grid = (int)min((long long)kNumSMs * 16, ceil_div64(total, 256));
  VOCR_REQUIRE((a_hi == nullptr) == (a_lo == nullptr));
  VOCR_REQUIRE(a_hi16 ? (a_lo16 && bound && pair_exp) : !a_lo16);
  if (total + (long long)grid * 256 < (1ll << 31))
    bn_relu_apply_kernel<unsigned><<<grid, 256, 0, stream>>>(
        z, scale, shift, a, a_hi, a_lo, P, H, W, C / 4, sB, sH, sW, reinterpret_cast<__half*>(a_hi16),
        reinterpret_cast<__half*>(a_lo16), reinterpret_cast<const unsigned*>(bound), pair_exp);
  else
    bn_relu_apply_kernel<long long><<<grid, 256, 0, stream>>>(
        z, scale, shift, a, a_hi, a_lo, P, H, W, C / 4, sB, sH, sW, reinterpret_cast<__half*>(a_hi16),
        reinterpret_cast<__half*>(a_lo16), reinterpret_cast<const unsigned*>(bound), pair_exp);
  VOCR_CHECK_LAUNCH();
  return VOCR_OK;
}

// red: double[2*C] (zeroed here), filled with s1 (= dbeta) and s2 (= dgamma); dz [P,C]; dgamma/dbeta/dbias fp32 [C].
// red_ws: double[3*C] scratch.
extern "C" int vocr_bn_relu_bwd_f32(const float* da, const float* z, const float* scale, const float* shift,
                                    const float* save_mean, const float* save_invstd, int training, int B, int H,
                                    int W, int C, long long sB, long long sH, long long sW, float* dz, float* dz_hi,
                                    float* dz_lo, float* dgamma, float* dbeta, float* dbias, double* red_ws,
                                    uint16_t* dz_hi16, uint16_t* dz_lo16, const float* scale_max, int32_t* pair_state,
                                    vocr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const long long P = (long long)B * H * W;
  VOCR_REQUIRE(C > 0 && C % 4 == 0 && C <= 1024 && red_ws && (dz || dz_hi16) && dgamma && dbeta);
  if (cudaMemsetAsync(red_ws, 0, sizeof(double) * 3 * C, stream) != cudaSuccess) return VOCR_MEMOPS_FAILED;
  VOCR_REQUIRE(dz_hi16 ? (dz_lo16 && scale_max && pair_state) : !dz_lo16);
  if (dz_hi16 && cudaMemsetAsync(pair_state, 0, 2 * sizeof(int32_t), stream) != cudaSuccess) return VOCR_MEMOPS_FAILED;
  if (P > 0) {
    VOCR_REQUIRE(da && z && scale && shift && save_mean && save_invstd);
    VOCR_REQUIRE(aligned16(z) && aligned16(da) && (!dz || aligned16(dz)) && sB % 4 == 0 && sH % 4 == 0 && sW % 4 == 0);
    const int C4 = C / 4;
    const int rows = 256 / C4;
    long long rows_per_cta = ceil_div64(P, (long long)kNumSMs * 4);
    rows_per_cta = max((long long)rows * 8, ceil_div64(rows_per_cta, rows) * rows);
    const int grid = (int)ceil_div64(P, rows_per_cta);
    bn_relu_bwd_reduce_kernel<<<grid, 256, sizeof(float) * 256 * 8, stream>>>(
        da, z, scale, shift, save_mean, save_invstd, P, H, W, C4, sB, sH, sW, rows_per_cta, red_ws,
        dz_hi16 ? reinterpret_cast<unsigned*>(pair_state + 1) : nullptr);
    VOCR_CHECK_LAUNCH();
    bn_relu_bwd_apply_kernel<<<grid, 256, sizeof(float) * 256 * 4, stream>>>(
        da, z, scale, shift, save_mean, save_invstd, red_ws, 1.0 / (double)P, training, P, H, W, C4, sB, sH, sW,
        rows_per_cta, dz, dz_hi, dz_lo, dbias ? red_ws + 2 * C : nullptr, reinterpret_cast<__half*>(dz_hi16),
        reinterpret_cast<__half*>(dz_lo16), reinterpret_cast<const unsigned*>(scale_max), pair_state);
    VOCR_CHECK_LAUNCH();
  }
  f64_to_f32_kernel<<<ceil_div(C, 128), 128, 0, stream>>>(red_ws, dbeta, C);
  f64_to_f32_kernel<<<ceil_div(C, 128), 128, 0, stream>>>(red_ws + C, dgamma, C);
  if (dbias) f64_to_f32_kernel<<<ceil_div(C, 128), 128, 0, stream>>>(red_ws + 2 * C, dbias, C);
  VOCR_CHECK_LAUNCH();
  return VOCR_OK;
}
