// 3x3 / pad 1 / stride 1 convolutions of the CNN feature extractor as implicit GEMM over NHWC fp32 activations
// (reference cnnlstm.py:114-134,263-266 -> cuDNN): forward (+bias, + per-channel sum / sum-of-squares for the
// BatchNorm that follows), data gradient (same kernel on flipped weights) and weight gradient (split-K over pixels).
// The rapid-downsample stage (conv Cin->16 + ReLU + 2x2 max-pool, cnnlstm.py:114-121) is one fused direct kernel so
// the full-resolution 16-channel intermediate is never written.
//
// GEMM views (P = B*H*W pixels, K index = (ky*3+kx)*Cin + ci):
//   fwd   : Z[P,Cout]  = im2col(X)[P,9Cin]   * Wk[9Cin,Cout]
//   dgrad : dX[P,Cin]  = im2col(dZ)[P,9Cout] * Wd[9Cout,Cin]      Wd[(ky,kx,co),ci] = W[co,ci,2-ky,2-kx]
//   wgrad : dWk[9Cin,Cout] = im2col(X)^T[9Cin,P] * dZ[P,Cout]     split over P, reduced + re-laid-out afterwards
#include "gemm_core.cuh"

namespace vocr {

// ---- im2col loader, contiguous along K (forward / dgrad): A(m = pixel, k..k+3) --------------------------------
struct ConvALoad {
  static constexpr bool kContigK = true;
  const float* x;
  int H, W, C;
  long long P;
  bool vec;  // C % 4 == 0 and base 16-B aligned: 4 consecutive k share a tap
  struct State {
    const float* base;  // &x[pixel, 0]
    int y, xx, tap, ci;
    bool valid;
  };
  __device__ __forceinline__ void init(State& st, int m, int k_first) const {
    st.valid = m < P;
    const int mm = st.valid ? m : 0;
    st.xx = mm % W;
    st.y = (mm / W) % H;
    st.base = x + (size_t)mm * C;
    st.tap = k_first / C;
    st.ci = k_first - st.tap * C;
  }
  __device__ __forceinline__ float tap_elem(const State& st, int tap, int ci) const {
    const int ky = tap / 3, kx = tap - ky * 3;
    const int yy = st.y + ky - 1, xc = st.xx + kx - 1;
    if (yy < 0 || yy >= H || xc < 0 || xc >= W) return 0.f;
    return __ldg(st.base + ((ky - 1) * W + (kx - 1)) * C + ci);
  }
  __device__ __forceinline__ float4 fetch(State& st, int k, int k_end) const {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (st.valid && k < k_end) {
      if (vec) {
        const int ky = st.tap / 3, kx = st.tap - ky * 3;
        const int yy = st.y + ky - 1, xc = st.xx + kx - 1;
        if (yy >= 0 && yy < H && xc >= 0 && xc < W)
          v = __ldg(reinterpret_cast<const float4*>(st.base + ((ky - 1) * W + (kx - 1)) * C + st.ci));
      } else {
        int tap = st.tap, ci = st.ci;
        float r[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (k + j < k_end) r[j] = tap_elem(st, tap, ci);
          if (++ci == C) {
            ci = 0;
            ++tap;
          }
        }
        v = make_float4(r[0], r[1], r[2], r[3]);
      }
    }
    st.ci += kGemmBK;
    while (st.ci >= C) {
      st.ci -= C;
      ++st.tap;
    }
    return v;
  }
};

// ---- im2col loader, contiguous along M (wgrad): A(m..m+3 = (tap,ci..ci+3), k = pixel) --------------------------
struct ConvATLoad {
  static constexpr bool kContigK = false;
  const float* x;
  int H, W, C, M;  // M = 9*C
  bool vec;
  struct State {
    int m, tap, ci;  // first of the 4 rows
    int y, xx;       // coordinates of the current pixel
    long long p;
  };
  __device__ __forceinline__ void init(State& st, int m, int k_first) const {
    st.m = m;
    st.tap = m / C;
    st.ci = m - st.tap * C;
    st.p = k_first;
    st.xx = k_first % W;
    st.y = (k_first / W) % H;
  }
  __device__ __forceinline__ float elem(const State& st, int tap, int ci) const {
    const int ky = tap / 3, kx = tap - ky * 3;
    const int yy = st.y + ky - 1, xc = st.xx + kx - 1;
    if (yy < 0 || yy >= H || xc < 0 || xc >= W) return 0.f;
    return __ldg(x + (size_t)(st.p + (ky - 1) * W + (kx - 1)) * C + ci);
  }
  __device__ __forceinline__ float4 fetch(State& st, int k, int k_end) const {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k < k_end && st.m < M) {
      if (vec) {
        const int ky = st.tap / 3, kx = st.tap - ky * 3;
        const int yy = st.y + ky - 1, xc = st.xx + kx - 1;
        if (yy >= 0 && yy < H && xc >= 0 && xc < W)
          v = __ldg(reinterpret_cast<const float4*>(x + (size_t)(st.p + (ky - 1) * W + (kx - 1)) * C + st.ci));
      } else {
        int tap = st.tap, ci = st.ci;
        float r[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (st.m + j < M) r[j] = elem(st, tap, ci);
          if (++ci == C) {
            ci = 0;
            ++tap;
          }
        }
        v = make_float4(r[0], r[1], r[2], r[3]);
      }
    }
    st.p += kGemmBK;
    st.xx += kGemmBK;
    while (st.xx >= W) {
      st.xx -= W;
      if (++st.y == H) st.y = 0;
    }
    return v;
  }
};

// ---- forward epilogue: z = acc + bias, optional per-channel sum / sum of squares (for BatchNorm) -------------------
template <int BN>
struct ConvFwdEpilogue {
  float* z;
  long long P;
  int N;
  const float* bias;
  double* stats;  // [2*N] or nullptr
  unsigned* zmax; // optional: max |z| (bit pattern of the non-negative float, atomicMax)
  float vmax;
  float* s_sum;   // shared [BN], s_sq shared [BN] (zeroed by the kernel)
  float* s_sq;
  int n0;
  float cs[4 * (BN / 64)], cq[4 * (BN / 64)];
  __device__ __forceinline__ void begin() {
#pragma unroll
    for (int j = 0; j < 4 * (BN / 64); ++j) cs[j] = cq[j] = 0.f;
    vmax = 0.f;
  }
  __device__ __forceinline__ void operator()(int m, int n, float4 v, int j) {
    if (m >= P || n >= N) return;
    float r[4] = {v.x, v.y, v.z, v.w};
    float* q = z + (size_t)m * N + n;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (n + c < N) {
        if (bias) r[c] += __ldg(bias + n + c);
        cs[4 * j + c] += r[c];
        cq[4 * j + c] = fmaf(r[c], r[c], cq[4 * j + c]);
        vmax = fmaxf(vmax, fabsf(r[c]));
      }
    }
    if ((N & 3) == 0) {
      *reinterpret_cast<float4*>(q) = make_float4(r[0], r[1], r[2], r[3]);
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (n + c < N) q[c] = r[c];
    }
  }
  __device__ __forceinline__ void finish() {
    if (zmax) {
      const float m = warp_max(vmax);
      if ((threadIdx.x & 31) == 0) atomicMax(zmax, __float_as_uint(m));
    }
    if (!stats) return;
    const int tx = threadIdx.x & 15;
#pragma unroll
    for (int j = 0; j < BN / 64; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int col = j * 64 + tx * 4 + c;
        atomicAdd(&s_sum[col], cs[4 * j + c]);
        atomicAdd(&s_sq[col], cq[4 * j + c]);
      }
    __syncthreads();
    for (int col = threadIdx.x; col < BN; col += kGemmThreads) {
      const int n = n0 + col;
      if (n < N) {
        atomicAdd(&stats[n], (double)s_sum[col]);
        atomicAdd(&stats[N + n], (double)s_sq[col]);
      }
    }
  }
};

template <int BN>
__global__ void __launch_bounds__(kGemmThreads)
conv3x3_fwd_kernel(ConvALoad la, BLoadContigN lb, float* z, const float* bias, double* stats, unsigned* zmax, int N) {
  __shared__ float s_sum[BN], s_sq[BN];
  for (int i = threadIdx.x; i < BN; i += kGemmThreads) s_sum[i] = s_sq[i] = 0.f;
  ConvFwdEpilogue<BN> ep;
  ep.z = z; ep.P = la.P; ep.N = N; ep.bias = bias; ep.stats = stats; ep.zmax = zmax; ep.s_sum = s_sum; ep.s_sq = s_sq;
  // pixel tiles in grid.x (up to 2^31 - 1 of them: 512 lines x 30 x 1200 px is 144k tiles), channel tiles in grid.y
  ep.n0 = blockIdx.y * BN;
  ep.begin();
  // gemm_tile's first __syncthreads orders the zeroing above before any shared atomic
  gemm_tile<BN>(la, lb, ep, blockIdx.x * kGemmBM, blockIdx.y * BN, 0, 9 * la.C);
}

template <int BN>
__global__ void __launch_bounds__(kGemmThreads)
conv3x3_wgrad_kernel(ConvATLoad la, BLoadContigN lb, float* partial, int M, int N, long long P,
                     long long k_chunk) {
  const long long k0 = (long long)blockIdx.z * k_chunk;
  const long long k1 = min(P, k0 + k_chunk);
  DenseEpilogue ep{partial + (size_t)blockIdx.z * M * N, M, N, N, nullptr, false, false, (N & 3) == 0};
  gemm_tile<BN>(la, lb, ep, blockIdx.y * kGemmBM, blockIdx.x * BN, (int)k0, (int)k1);
}

// sum split-K partials [S][9*Cin][Cout] and write the PyTorch layout dW[Cout][Cin][3][3]
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ partial, int splits, int Cin, int Cout, float* __restrict__ dw) {
  const int total = 9 * Cin * Cout;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    // idx enumerates the OUTPUT (co, ci, tap) so the write is coalesced
    const int tap = idx % 9;
    const int ci = (idx / 9) % Cin;
    const int co = idx / (9 * Cin);
    const size_t src = (size_t)(tap * Cin + ci) * Cout + co;
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += partial[(size_t)z * total + src];
    dw[idx] = s;
  }
}

// W[Cout][Cin][3][3] -> Wk[(ky,kx,ci)][co]  and  Wd[(ky,kx,co)][ci] = W[co][ci][2-ky][2-kx]
__global__ void __launch_bounds__(256)
conv_weight_layout_kernel(const float* __restrict__ w, int Cin, int Cout, float* __restrict__ wk,
                          float* __restrict__ wd) {
  const int total = 9 * Cin * Cout;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int tap = idx % 9;
    const int ci = (idx / 9) % Cin;
    const int co = idx / (9 * Cin);
    const float v = w[idx];
    if (wk) wk[(size_t)(tap * Cin + ci) * Cout + co] = v;
    if (wd) wd[(size_t)((8 - tap) * Cout + co) * Cin + ci] = v;
  }
}

// ---- rapid-downsample stage: conv3x3(Cin->16)+bias -> ReLU -> maxpool 2x2/2, fused, direct ----------------------
// One thread = one pooled pixel x 4 output channels.  wk: [9*Cin][16].  arg (uint8, optional): which of the 4 window
// positions won (first max in (dy,dx) scan order, like ATen max_pool2d), for the backward pass.
constexpr int kRdsCout = 16;
template <bool VEC4>  // VEC4: Cin % 4 == 0 and a 16-byte aligned x (second stage); separate instantiations keep the
                      // one-channel first stage at its low register count
__global__ void __launch_bounds__(256, VEC4 ? 2 : 3)
rds_fwd_kernel(const float* __restrict__ x, const float* __restrict__ wk, const float* __restrict__ bias,
               float* __restrict__ y, uint8_t* __restrict__ arg, int B, int H, int W, int Cin) {
  extern __shared__ float s_w[];  // [9*Cin][16]
  for (int i = threadIdx.x; i < 9 * Cin * kRdsCout; i += blockDim.x) s_w[i] = wk[i];
  __syncthreads();
  const int Ho = H / 2, Wo = W / 2;
  const long long total = (long long)B * Ho * Wo * 4;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int cq = (int)(gid & 3);
  const long long pp = gid >> 2;
  const int xo = (int)(pp % Wo);
  const int yo = (int)((pp / Wo) % Ho);
  const int b = (int)(pp / ((long long)Wo * Ho));
  float acc[4][4];  // [window position][channel]
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[p][c] = bias[cq * 4 + c];
  const int y0 = yo * 2 - 1, x0 = xo * 2 - 1;  // top-left of the 4x4 input patch
  const float* xb = x + (size_t)b * H * W * Cin;
  if (VEC4) {
    // channels four at a time: 16 float4 loads of the patch feed 576 multiply-adds (second stage, Cin = 16)
    for (int c4 = 0; c4 < Cin; c4 += 4) {
      float4 patch[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int yy = y0 + r, xc = x0 + c;
          patch[r][c] = (yy >= 0 && yy < H && xc >= 0 && xc < W)
                            ? __ldg(reinterpret_cast<const float4*>(xb + ((size_t)yy * W + xc) * Cin + c4))
                            : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            const float4 w4 = *reinterpret_cast<const float4*>(&s_w[((ky * 3 + kx) * Cin + c4 + cc) * kRdsCout + cq * 4]);
#pragma unroll
            for (int p = 0; p < 4; ++p) {
              const float4 pv = patch[(p >> 1) + ky][(p & 1) + kx];
              const float v = cc == 0 ? pv.x : cc == 1 ? pv.y : cc == 2 ? pv.z : pv.w;
              acc[p][0] = fmaf(v, w4.x, acc[p][0]);
              acc[p][1] = fmaf(v, w4.y, acc[p][1]);
              acc[p][2] = fmaf(v, w4.z, acc[p][2]);
              acc[p][3] = fmaf(v, w4.w, acc[p][3]);
            }
          }
    }
  } else
  for (int ci = 0; ci < Cin; ++ci) {
    float patch[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int yy = y0 + r, xc = x0 + c;
        patch[r][c] = (yy >= 0 && yy < H && xc >= 0 && xc < W) ? __ldg(xb + ((size_t)yy * W + xc) * Cin + ci) : 0.f;
      }
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const float4 w4 = *reinterpret_cast<const float4*>(&s_w[((ky * 3 + kx) * Cin + ci) * kRdsCout + cq * 4]);
        const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const float v = patch[(p >> 1) + ky][(p & 1) + kx];
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[p][c] = fmaf(v, wv[c], acc[p][c]);
        }
      }
  }
  float out[4];
  uint32_t amask = 0;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float best = fmaxf(acc[0][c], 0.f);
    int bi = 0;
#pragma unroll
    for (int p = 1; p < 4; ++p) {
      const float v = fmaxf(acc[p][c], 0.f);
      if (v > best || v != v) {
        best = v;
        bi = p;
      }
    }
    out[c] = best;
    amask |= (uint32_t)bi << (8 * c);
  }
  const size_t o = (size_t)pp * kRdsCout + cq * 4;
  *reinterpret_cast<float4*>(y + o) = make_float4(out[0], out[1], out[2], out[3]);
  if (arg) *reinterpret_cast<uint32_t*>(arg + o) = amask;
}

// backward of ReLU + maxpool: scatter dy to the winning pre-activation position, dense dpre[B,H,W,16]
// (positions that lost, and winners whose pooled value is 0 (ReLU inactive), get 0).
__global__ void __launch_bounds__(256)
rds_unpool_kernel(const float* __restrict__ dy, const float* __restrict__ y, const uint8_t* __restrict__ arg,
                  float* __restrict__ dpre, int B, int H, int W) {
  const int Ho = H / 2, Wo = W / 2;
  const long long total = (long long)B * H * W * 4;  // one thread per (full-res pixel, channel quad)
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int cq = (int)(gid & 3);
  const long long pp = gid >> 2;
  const int xc = (int)(pp % W);
  const int yy = (int)((pp / W) % H);
  const int b = (int)(pp / ((long long)W * H));
  float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
  const int yo = yy >> 1, xo = xc >> 1;
  if (yo < Ho && xo < Wo) {
    const size_t o = (((size_t)b * Ho + yo) * Wo + xo) * kRdsCout + cq * 4;
    const uint32_t am = *reinterpret_cast<const uint32_t*>(arg + o);
    const float4 d = *reinterpret_cast<const float4*>(dy + o);
    const float4 v = *reinterpret_cast<const float4*>(y + o);
    const uint32_t mine = (uint32_t)((yy & 1) * 2 + (xc & 1));
    if (((am >> 0) & 0xff) == mine && v.x > 0.f) g.x = d.x;
    if (((am >> 8) & 0xff) == mine && v.y > 0.f) g.y = d.y;
    if (((am >> 16) & 0xff) == mine && v.z > 0.f) g.z = d.z;
    if (((am >> 24) & 0xff) == mine && v.w > 0.f) g.w = d.w;
  }
  *reinterpret_cast<float4*>(dpre + (size_t)pp * kRdsCout + cq * 4) = g;
}

// Weight / bias gradient of the FIRST rapid-downsample stage (Cin = 1, no data gradient needed), straight from the
// pooled gradient: only the window position that won the max-pool (and survived the ReLU) carries gradient, so the
// dense full-resolution dpre tensor is never built.  One thread = one pooled pixel x 4 output channels; per-thread
// partials (4 x 9 weights + 4 biases) are reduced by warp shuffles, then shared memory, then float64 atomics.
__global__ void __launch_bounds__(256)
rds_wgrad_c1_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ y,
                    const uint8_t* __restrict__ arg, int B, int H, int W, double* __restrict__ out /*[16*9 + 16]*/) {
  const int Ho = H / 2, Wo = W / 2;
  const long long total = (long long)B * Ho * Wo * 4;
  float acc[4][9], accb[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    accb[c] = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[c][t] = 0.f;
  }
  // grid-stride in units that keep tid & 3 == channel quad fixed per thread
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x; gid < total; gid += stride) {
    const int cq = (int)(gid & 3);
    const long long pp = gid >> 2;
    const int xo = (int)(pp % Wo);
    const int yo = (int)((pp / Wo) % Ho);
    const int b = (int)(pp / ((long long)Wo * Ho));
    const size_t o = (size_t)pp * kRdsCout + cq * 4;
    const float4 d = __ldg(reinterpret_cast<const float4*>(dy + o));
    const float4 v = __ldg(reinterpret_cast<const float4*>(y + o));
    const uint32_t am = __ldg(reinterpret_cast<const uint32_t*>(arg + o));
    const float g[4] = {v.x > 0.f ? d.x : 0.f, v.y > 0.f ? d.y : 0.f, v.z > 0.f ? d.z : 0.f, v.w > 0.f ? d.w : 0.f};
    float patch[4][4];
    const int y0 = yo * 2 - 1, x0 = xo * 2 - 1;
    const float* xb = x + (size_t)b * H * W;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int yy = y0 + r, xc = x0 + c;
        patch[r][c] = (yy >= 0 && yy < H && xc >= 0 && xc < W) ? __ldg(xb + (size_t)yy * W + xc) : 0.f;
      }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int pos = (int)((am >> (8 * c)) & 0xff);
      accb[c] += g[c];
#pragma unroll
      for (int pv = 0; pv < 4; ++pv) {
        const float gg = (pos == pv) ? g[c] : 0.f;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) acc[c][ky * 3 + kx] = fmaf(gg, patch[(pv >> 1) + ky][(pv & 1) + kx], acc[c][ky * 3 + kx]);
      }
    }
  }
  // lanes with equal (lane & 3) hold the same channel quad: reduce over lane bits 2..4
#pragma unroll
  for (int c = 0; c < 4; ++c) {
#pragma unroll
    for (int off = 4; off < 32; off <<= 1) accb[c] += __shfl_xor_sync(0xffffffffu, accb[c], off);
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int off = 4; off < 32; off <<= 1) acc[c][t] += __shfl_xor_sync(0xffffffffu, acc[c][t], off);
  }
  __shared__ float s_part[8][4][40];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane < 4) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
      for (int t = 0; t < 9; ++t) s_part[warp][lane][c * 9 + t] = acc[c][t];
      s_part[warp][lane][36 + c] = accb[c];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 160; i += blockDim.x) {
    const int cq = i / 40, k = i % 40;
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += (double)s_part[w][cq][k];
    // out layout: dw[co][tap] (144) then db[co] (16)
    if (k < 36) atomicAdd(&out[(cq * 4 + k / 9) * 9 + k % 9], t);
    else atomicAdd(&out[144 + cq * 4 + (k - 36)], t);
  }
}

// ---- 16 -> 16 channel 3x3 convolutions of the SECOND rapid-downsample stage (cfg3: 60 x 1000 pixels, batch 64) -------
// With 16 channels on both sides the implicit-GEMM kernels waste 3/4 (forward / data gradient, N = 16 of a 64-wide tile)
// to 99 % (weight gradient, a 144 x 16 output tile) of their tiles: 1.9 + 4.2 ms of a 29-ms cfg3 step.  These two direct
// kernels are FFMA-issue bound instead (8.8 G fused multiply-adds each at that size).

// z[p][co] = sum_{tap,ci} x[p + tap][ci] wk[(tap, ci)][co] (+ bias).  One thread = 4 consecutive pixels of a row x 4
// output channels: per (ky, input-channel quad) it loads 6 float4 of x and reads 12 float4 of weights from shared memory
// for 192 multiply-adds.
__global__ void __launch_bounds__(256)
conv16_fwd_kernel(const float* __restrict__ x, const float* __restrict__ wk, const float* __restrict__ bias,
                  float* __restrict__ z, int H, int W) {
  __shared__ float4 s_w[144 * 4];  // [(tap, ci)][co quad]
  for (int i = threadIdx.x; i < 144 * 4; i += 256) s_w[i] = __ldg(reinterpret_cast<const float4*>(wk) + i);
  __syncthreads();
  const int cq = threadIdx.x & 3, pg = threadIdx.x >> 2;
  const int x0 = blockIdx.x * 256 + pg * 4, y = blockIdx.y, b = blockIdx.z;
  if (x0 >= W) return;
  float acc[4][4];
  const float4 bv = bias ? __ldg(reinterpret_cast<const float4*>(bias) + cq) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    acc[p][0] = bv.x; acc[p][1] = bv.y; acc[p][2] = bv.z; acc[p][3] = bv.w;
  }
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int yy = y + ky - 1;
    if (yy < 0 || yy >= H) continue;
    const float4* row = reinterpret_cast<const float4*>(x + ((size_t)b * H + yy) * W * 16);
#pragma unroll
    for (int q = 0; q < 4; ++q) {  // input channels 4q .. 4q+3
      float4 xv[6];
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const int xc = x0 - 1 + j;
        xv[j] = (xc >= 0 && xc < W) ? __ldg(row + (size_t)xc * 4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 w4 = s_w[((ky * 3 + kx) * 16 + q * 4 + c) * 4 + cq];
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const float4 xx = xv[p + kx];
            const float v = c == 0 ? xx.x : c == 1 ? xx.y : c == 2 ? xx.z : xx.w;
            acc[p][0] = fmaf(v, w4.x, acc[p][0]);
            acc[p][1] = fmaf(v, w4.y, acc[p][1]);
            acc[p][2] = fmaf(v, w4.z, acc[p][2]);
            acc[p][3] = fmaf(v, w4.w, acc[p][3]);
          }
        }
      }
    }
  }
  float4* zrow = reinterpret_cast<float4*>(z + (((size_t)b * H + y) * W) * 16);
#pragma unroll
  for (int p = 0; p < 4; ++p)
    if (x0 + p < W) zrow[(size_t)(x0 + p) * 4 + cq] = make_float4(acc[p][0], acc[p][1], acc[p][2], acc[p][3]);
}

// dw[co][ci][tap] = sum_p x[p + tap][ci] dz[p][co],  db[co] = sum_p dz[p][co]  -> float64 out[2304 + 16] (atomics).
// A CTA walks over (image, 64-pixel column chunk) items and, inside one, over the rows; per row it stages the three x rows
// (with halo) and the dz row in shared memory.  256 threads = 4 pixel groups x (16 ci x 4 co quads): a thread keeps
// 9 taps x 4 output channels in registers and slides a 3-wide window of x along its 16 pixels - per pixel 3 + 1
// shared-memory loads for 36 multiply-adds.
constexpr int kW16Chunk = 64;
__global__ void __launch_bounds__(256)
conv16_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dz, int B, int H, int W, int n_items,
                    int chunks, double* __restrict__ out) {
  __shared__ __align__(16) float xs[3][kW16Chunk + 2][16];
  __shared__ __align__(16) float ds[kW16Chunk][16];
  __shared__ float red[64][37];
  const int tid = threadIdx.x, g = tid >> 6, t = tid & 63;
  const int ci = t & 15, coq = t >> 4;
  float acc[9][4], accb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 9; ++k)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[k][c] = 0.f;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int b = item / chunks, xc0 = (item - b * chunks) * kW16Chunk;
    for (int y = 0; y < H; ++y) {
      __syncthreads();  // the previous row's tiles are consumed
      // stage: 3 rows x 66 pixels x 4 float4 of x, 64 pixels x 4 float4 of dz
      for (int i = tid; i < 3 * (kW16Chunk + 2) * 4; i += 256) {
        const int r = i / ((kW16Chunk + 2) * 4), rem = i - r * (kW16Chunk + 2) * 4;
        const int c = rem >> 2, q = rem & 3;
        const int yy = y + r - 1, xx = xc0 - 1 + c;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (yy >= 0 && yy < H && xx >= 0 && xx < W)
          v = __ldg(reinterpret_cast<const float4*>(x + (((size_t)b * H + yy) * W + xx) * 16) + q);
        *reinterpret_cast<float4*>(&xs[r][c][q * 4]) = v;
      }
      for (int i = tid; i < kW16Chunk * 4; i += 256) {
        const int c = i >> 2, q = i & 3;
        const int xx = xc0 + c;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (xx < W) v = __ldg(reinterpret_cast<const float4*>(dz + (((size_t)b * H + y) * W + xx) * 16) + q);
        *reinterpret_cast<float4*>(&ds[c][q * 4]) = v;
      }
      __syncthreads();
      const int p0 = g * 16;
      float w0[3], w1[3];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        w0[r] = xs[r][p0][ci];
        w1[r] = xs[r][p0 + 1][ci];
      }
#pragma unroll
      for (int p = 0; p < 16; ++p) {
        const float4 d = *reinterpret_cast<const float4*>(&ds[p0 + p][coq * 4]);
        if (ci == 0) {
          accb[0] += d.x; accb[1] += d.y; accb[2] += d.z; accb[3] += d.w;
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const float w2 = xs[r][p0 + p + 2][ci];
          acc[r * 3 + 0][0] = fmaf(w0[r], d.x, acc[r * 3 + 0][0]);
          acc[r * 3 + 0][1] = fmaf(w0[r], d.y, acc[r * 3 + 0][1]);
          acc[r * 3 + 0][2] = fmaf(w0[r], d.z, acc[r * 3 + 0][2]);
          acc[r * 3 + 0][3] = fmaf(w0[r], d.w, acc[r * 3 + 0][3]);
          acc[r * 3 + 1][0] = fmaf(w1[r], d.x, acc[r * 3 + 1][0]);
          acc[r * 3 + 1][1] = fmaf(w1[r], d.y, acc[r * 3 + 1][1]);
          acc[r * 3 + 1][2] = fmaf(w1[r], d.z, acc[r * 3 + 1][2]);
          acc[r * 3 + 1][3] = fmaf(w1[r], d.w, acc[r * 3 + 1][3]);
          acc[r * 3 + 2][0] = fmaf(w2, d.x, acc[r * 3 + 2][0]);
          acc[r * 3 + 2][1] = fmaf(w2, d.y, acc[r * 3 + 2][1]);
          acc[r * 3 + 2][2] = fmaf(w2, d.z, acc[r * 3 + 2][2]);
          acc[r * 3 + 2][3] = fmaf(w2, d.w, acc[r * 3 + 2][3]);
          w0[r] = w1[r];
          w1[r] = w2;
        }
      }
    }
  }
  // the four pixel groups add their partials in turn (fixed order), then one float64 atomic per output and CTA
  for (int gg = 0; gg < 4; ++gg) {
    __syncthreads();
    if (g == gg) {
#pragma unroll
      for (int k = 0; k < 9; ++k)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float* r = &red[t][k * 4 + c];
          *r = (gg == 0 ? 0.f : *r) + acc[k][c];
        }
      if (ci == 0) {  // bias partials: 16 values per group, straight out
#pragma unroll
        for (int c = 0; c < 4; ++c) atomicAdd(&out[2304 + coq * 4 + c], (double)accb[c]);
      }
    }
  }
  __syncthreads();
  if (tid < 64) {
#pragma unroll
    for (int k = 0; k < 9; ++k)
#pragma unroll
      for (int c = 0; c < 4; ++c)
        atomicAdd(&out[((coq * 4 + c) * 16 + ci) * 9 + k], (double)red[t][k * 4 + c]);
  }
}

__global__ void f64_to_f32_conv_kernel(const double* __restrict__ in, float* __restrict__ a, int na,
                                       float* __restrict__ b, int nb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < na) a[i] = (float)in[i];
  else if (i < na + nb) b[i - na] = (float)in[i];
}

}  // namespace vocr

using namespace vocr;

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int vocr_conv_weight_layout_f32(const float* w, int Cin, int Cout, float* wk, float* wd,
                                           vocr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VOCR_REQUIRE(w && Cin > 0 && Cout > 0 && (wk || wd));
  const int total = 9 * Cin * Cout;
  conv_weight_layout_kernel<<<min(ceil_div(total, 256), 4 * kNumSMs), 256, 0, stream>>>(w, Cin, Cout, wk, wd);
  VOCR_CHECK_LAUNCH();
  return VOCR_OK;
}

// x [B,H,W,Cin] NHWC, wk [9*Cin,Cout], z [B,H,W,Cout]; stats (optional) double[2*Cout] is ACCUMULATED into.
extern "C" int vocr_conv3x3_fwd_f32(const float* x, const float* wk, const float* bias, float* z, int B, int H,
                                    int W, int Cin, int Cout, double* stats, float* zmax, vocr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VOCR_REQUIRE(B >= 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0);
  const long long P = (long long)B * H * W;
  if (P == 0) return VOCR_OK;
  VOCR_REQUIRE(x && wk && z && P < (1ll << 31) - 256);
  ConvALoad la;
  la.x = x; la.H = H; la.W = W; la.C = Cin; la.P = P; la.vec = (Cin % 4 == 0) && aligned16(x);
  BLoadContigN lb;
  lb.p = wk; lb.cols = Cout; lb.ld = Cout; lb.vec = (Cout % 4 == 0) && aligned16(wk);
  if (Cout <= 64) {
    dim3 grid((unsigned)ceil_div64(P, kGemmBM), ceil_div(Cout, 64));
    conv3x3_fwd_kernel<64><<<grid, kGemmThreads, 0, stream>>>(la, lb, z, bias, stats, reinterpret_cast<unsigned*>(zmax), Cout);
  } else {
    dim3 grid((unsigned)ceil_div64(P, kGemmBM), ceil_div(Cout, 128));
    conv3x3_fwd_kernel<128><<<grid, kGemmThreads, 0, stream>>>(la, lb, z, bias, stats, reinterpret_cast<unsigned*>(zmax), Cout);
  }
  VOCR_CHECK_LAUNCH();
  return VOCR_OK;
}

extern "C" size_t vocr_conv3x3_wgrad_workspace_size(int B, int H, int W, int Cin, int Cout) {
  // generous upper bound on split count (see below)
  return sizeof(float) * (size_t)9 * Cin * Cout * 64 + 256;
}

// dw [Cout,Cin,3,3] (PyTorch layout) = sum over pixels of im2col(x)^T dz.  x [B,H,W,Cin], dz [B,H,W,Cout].
extern "C" int vocr_conv3x3_wgrad_f32(const float* x, const float* dz, float* dw, int B, int H, int W, int Cin,
                                      int Cout, void* workspace, size_t workspace_bytes, vocr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VOCR_REQUIRE(B >= 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && dw && workspace);
  const long long P = (long long)B * H * W;
  VOCR_REQUIRE(P < (1ll << 31) - 256);
  const int M = 9 * Cin, N = Cout;
  if (P == 0) {
    if (cudaMemsetAsync(dw, 0, sizeof(float) * M * N, stream) != cudaSuccess) return VOCR_MEMOPS_FAILED;
    return VOCR_OK;
  }
  VOCR_REQUIRE(x && dz);
  const int bn = (N <= 64) ? 64 : 128;
  const int tiles = ceil_div(M, kGemmBM) * ceil_div(N, bn);
  int splits = max(1, min(64, (2 * kNumSMs + tiles - 1) / tiles));
  long long k_chunk = ceil_div64(P, splits);
  k_chunk = ceil_div64(k_chunk, kGemmBK) * kGemmBK;
  splits = (int)ceil_div64(P, k_chunk);
  VOCR_REQUIRE(sizeof(float) * (size_t)M * N * splits <= workspace_bytes);
  float* partial = static_cast<float*>(workspace);
  ConvATLoad la;
  la.x = x; la.H = H; la.W = W; la.C = Cin; la.M = M; la.vec = (Cin % 4 == 0) && aligned16(x);
  BLoadContigN lb;
  lb.p = dz; lb.cols = N; lb.ld = N; lb.vec = (N % 4 == 0) && aligned16(dz);
  dim3 grid(ceil_div(N, bn), ceil_div(M, kGemmBM), splits);
  if (bn == 64) conv3x3_wgrad_kernel<64><<<grid, kGemmThreads, 0, stream>>>(la, lb, partial, M, N, P, k_chunk);
  else conv3x3_wgrad_kernel<128><<<grid, kGemmThreads, 0, stream>>>(la, lb, partial, M, N, P, k_chunk);
  VOCR_CHECK_LAUNCH();
  wgrad_reduce_kernel<<<min(ceil_div(M * N, 256), 4 * kNumSMs), 256, 0, stream>>>(partial, splits, Cin, Cout, dw);
  VOCR_CHECK_LAUNCH();
  return VOCR_OK;
}

extern "C" int vocr_rds_fwd_f32(const float* x, const float* wk, const float* bias, float* y, uint8_t* arg, int B,
                                int H, int W, int Cin, vocr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VOCR_REQUIRE(B >= 0 && H >= 2 && W >= 2 && Cin > 0 && Cin <= 64);
  const long long total = (long long)B * (H / 2) * (W / 2) * 4;
  if (total == 0) return VOCR_OK;
  VOCR_REQUIRE(x && wk && bias && y && aligned16(y));
  const size_t smem = sizeof(float) * 9 * Cin * kRdsCout;
  if ((Cin & 3) == 0 && aligned16(x))
    rds_fwd_kernel<true><<<(unsigned)ceil_div64(total, 256), 256, smem, stream>>>(x, wk, bias, y, arg, B, H, W, Cin);
  else
    rds_fwd_kernel<false><<<(unsigned)ceil_div64(total, 256), 256, smem, stream>>>(x, wk, bias, y, arg, B, H, W, Cin);
  VOCR_CHECK_LAUNCH();
  return VOCR_OK;
}

extern "C" int vocr_rds_unpool_f32(const float* dy, const float* y, const uint8_t* arg, float* dpre, int B, int H,
                                   int W, vocr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VOCR_REQUIRE(B >= 0 && H >= 2 && W >= 2);
  const long long total = (long long)B * H * W * 4;
  if (total == 0) return VOCR_OK;
  VOCR_REQUIRE(dy && y && arg && dpre);
  rds_unpool_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, stream>>>(dy, y, arg, dpre, B, H, W);
  VOCR_CHECK_LAUNCH();
  return VOCR_OK;
}

// First rapid-downsample stage (Cin = 1): dw[16,1,3,3], db[16] from the pooled gradient.  ws: float64[160] scratch.
extern "C" int vocr_rds_wgrad_c1_f32(const float* x, const float* dy, const float* y, const uint8_t* arg, float* dw,
                                     float* db, int B, int H, int W, double* ws, vocr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VOCR_REQUIRE(B >= 0 && H >= 2 && W >= 2 && dw && db && ws);
  if (cudaMemsetAsync(ws, 0, sizeof(double) * 160, stream) != cudaSuccess) return VOCR_MEMOPS_FAILED;
  const long long total = (long long)B * (H / 2) * (W / 2) * 4;
  if (total > 0) {
    VOCR_REQUIRE(x && dy && y && arg);
    const int grid = (int)min((long long)kNumSMs * 8, ceil_div64(total, 256));
    rds_wgrad_c1_kernel<<<grid, 256, 0, stream>>>(x, dy, y, arg, B, H, W, ws);
    VOCR_CHECK_LAUNCH();
  }
  f64_to_f32_conv_kernel<<<1, 192, 0, stream>>>(ws, dw, 144, db, 16);
  VOCR_CHECK_LAUNCH();
  return VOCR_OK;
}

// 3x3 / pad 1 convolution with 16 input and 16 output channels (second rapid-downsample stage: its data gradient runs
// through here with the flipped weight matrix wd of vocr_conv_weight_layout_f32).  wk [144][16], bias [16] or NULL.
extern "C" int vocr_conv3x3_c16_fwd_f32(const float* x, const float* wk, const float* bias, float* z, int B, int H,
                                        int W, vocr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VOCR_REQUIRE(B >= 0 && H > 0 && W > 0 && B <= 65535 && H <= 65535);
  if (B == 0) return VOCR_OK;
  VOCR_REQUIRE(x && wk && z && aligned16(x) && aligned16(wk) && aligned16(z) && (!bias || aligned16(bias)));
  dim3 grid(ceil_div(W, 256), H, B);
  conv16_fwd_kernel<<<grid, 256, 0, stream>>>(x, wk, bias, z, H, W);
  VOCR_CHECK_LAUNCH();
  return VOCR_OK;
}

// Weight / bias gradient of that convolution: dw[16,16,3,3] (state_dict layout), db[16] from x and dz (both
// [B,H,W,16]).  ws: float64[2320] scratch.
extern "C" int vocr_conv3x3_c16_wgrad_f32(const float* x, const float* dz, float* dw, float* db, int B, int H, int W,
                                          double* ws, vocr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VOCR_REQUIRE(B >= 0 && H > 0 && W > 0 && dw && db && ws);
  if (cudaMemsetAsync(ws, 0, sizeof(double) * 2320, stream) != cudaSuccess) return VOCR_MEMOPS_FAILED;
  const int chunks = ceil_div(W, kW16Chunk);
  const long long items = (long long)B * chunks;
  VOCR_REQUIRE(items < (1ll << 31));
  if (items > 0) {
    VOCR_REQUIRE(x && dz && aligned16(x) && aligned16(dz));
    const int grid = (int)min((long long)kNumSMs * 4, items);
    conv16_wgrad_kernel<<<grid, 256, 0, stream>>>(x, dz, B, H, W, (int)items, chunks, ws);
    VOCR_CHECK_LAUNCH();
  }
  f64_to_f32_conv_kernel<<<ceil_div(2320, 256), 256, 0, stream>>>(ws, dw, 2304, db, 16);
  VOCR_CHECK_LAUNCH();
  return VOCR_OK;
}
