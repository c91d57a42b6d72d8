// fp32 SIMT tile GEMM engine shared by the dense GEMM and the implicit-GEMM convolutions.
//   C(m,n) = sum_k A(m,k) * B(k,n),   CTA tile 128 x BN x 16, 256 threads, 8 x (BN/16) micro-tiles,
//   register-prefetch double buffering (one __syncthreads per k-tile).
// Operands are described by loader functors so that the same main loop serves row-major / transposed matrices and
// the im2col gathers of conv3x3 forward, dgrad and wgrad.  A loader declares which logical dimension is contiguous
// in memory and returns 4 consecutive elements along it (zero-filled out of bounds):
//   ALoader::kContigK ? fetch(st, k) -> A(m, k..k+3)  :  fetch(st, k) -> A(m..m+3, k)
//   BLoader::kContigK ? fetch(st, k) -> B(k..k+3, n)  :  fetch(st, k) -> B(k, n..n+3)
// `st` is a per-thread State initialised once per tile with the thread's fixed coordinate (m or n) and its first k;
// successive fetches advance k by kGemmBK, which lets the conv loaders track (tap, channel) or (image, row, column)
// incrementally instead of dividing.
// fp32 FFMA is the exact-parity path (north_star tolerance 1e-5 relative); the tcgen05 path lives in tc_gemm.cu.
#pragma once
#include "common.cuh"

namespace vocr {

constexpr int kGemmBM = 128;
constexpr int kGemmBK = 16;
constexpr int kGemmThreads = 256;
constexpr int kGemmLd = kGemmBM + 4;  // padded leading dimension of the k-major smem tiles

template <int BN, class ALoader, class BLoader, class Epilogue>
__device__ __forceinline__ void gemm_tile(const ALoader& la, const BLoader& lb, Epilogue& ep, int m0, int n0,
                                          int k_begin, int k_end) {
  static_assert(BN == 64 || BN == 128, "BN must be 64 or 128");
  constexpr int NB = BN / 64;           // 4-wide column groups per thread
  constexpr int LDB = BN + 4;
  constexpr int BV = (BN * kGemmBK / 4) / kGemmThreads;  // float4 per thread for the B tile (1 or 2)
  __shared__ __align__(16) float As[2][kGemmBK][kGemmLd];
  __shared__ __align__(16) float Bs[2][kGemmBK][LDB];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;

  float acc[8][4 * NB];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4 * NB; ++j) acc[i][j] = 0.f;

  float4 ra[2], rb[BV];
  typename ALoader::State sa[2];
  typename BLoader::State sb[BV];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int v = tid + i * kGemmThreads;
    if (ALoader::kContigK) la.init(sa[i], m0 + (v >> 2), k_begin + (v & 3) * 4);
    else la.init(sa[i], m0 + (v & 31) * 4, k_begin + (v >> 5));
  }
#pragma unroll
  for (int i = 0; i < BV; ++i) {
    const int v = tid + i * kGemmThreads;
    if (BLoader::kContigK) lb.init(sb[i], n0 + (v >> 2), k_begin + (v & 3) * 4);
    else lb.init(sb[i], n0 + (v % (BN / 4)) * 4, k_begin + v / (BN / 4));
  }

  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int v = tid + i * kGemmThreads;
      ra[i] = la.fetch(sa[i], ALoader::kContigK ? k0 + (v & 3) * 4 : k0 + (v >> 5), k_end);
    }
#pragma unroll
    for (int i = 0; i < BV; ++i) {
      const int v = tid + i * kGemmThreads;
      rb[i] = lb.fetch(sb[i], BLoader::kContigK ? k0 + (v & 3) * 4 : k0 + v / (BN / 4), k_end);
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int v = tid + i * kGemmThreads;
      if (ALoader::kContigK) {
        const int row = v >> 2, kq = (v & 3) * 4;
        As[buf][kq + 0][row] = ra[i].x;
        As[buf][kq + 1][row] = ra[i].y;
        As[buf][kq + 2][row] = ra[i].z;
        As[buf][kq + 3][row] = ra[i].w;
      } else {
        const int k = v >> 5, mq = (v & 31) * 4;
        *reinterpret_cast<float4*>(&As[buf][k][mq]) = ra[i];
      }
    }
#pragma unroll
    for (int i = 0; i < BV; ++i) {
      const int v = tid + i * kGemmThreads;
      if (BLoader::kContigK) {
        const int col = v >> 2, kq = (v & 3) * 4;
        Bs[buf][kq + 0][col] = rb[i].x;
        Bs[buf][kq + 1][col] = rb[i].y;
        Bs[buf][kq + 2][col] = rb[i].z;
        Bs[buf][kq + 3][col] = rb[i].w;
      } else {
        const int k = v / (BN / 4), nq = (v % (BN / 4)) * 4;
        *reinterpret_cast<float4*>(&Bs[buf][k][nq]) = rb[i];
      }
    }
  };

  const int n_tiles = (k_end - k_begin + kGemmBK - 1) / kGemmBK;
  if (n_tiles > 0) {
    fetch(k_begin);
    stash(0);
  }
  __syncthreads();
  for (int t = 0; t < n_tiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < n_tiles) fetch(k_begin + (t + 1) * kGemmBK);
#pragma unroll
    for (int k = 0; k < kGemmBK; ++k) {
      float a[8], b[4 * NB];
      *reinterpret_cast<float4*>(&a[0]) = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      *reinterpret_cast<float4*>(&a[4]) = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
#pragma unroll
      for (int j = 0; j < NB; ++j)
        *reinterpret_cast<float4*>(&b[4 * j]) = *reinterpret_cast<const float4*>(&Bs[buf][k][j * 64 + tx * 4]);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4 * NB; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (t + 1 < n_tiles) stash(buf ^ 1);
    __syncthreads();
  }
  // epilogue: rows m0 + {ty*4+i, 64+ty*4+i}, column groups n0 + j*64 + tx*4
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const int n = n0 + j * 64 + tx * 4;
      ep(m, n, make_float4(acc[i][4 * j], acc[i][4 * j + 1], acc[i][4 * j + 2], acc[i][4 * j + 3]), j);
    }
  }
  ep.finish();
}

// ---- dense loaders ------------------------------------------------------------------------------------------
// Row-major matrix whose contiguous dimension is the GEMM's K (A given as [M,K], or B given as [N,K]).
struct DenseContigK {
  static constexpr bool kContigK = true;
  const float* p;
  long long rows;    // extent of the non-K dimension
  int ld;
  bool vec;          // base 16-B aligned and ld % 4 == 0
  // A(m, k..k+3)  or, used as a B loader, B(k..k+3, n) with (k, n) argument order
  __device__ __forceinline__ float4 get(int r, int k, int k_end) const {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r >= rows) return v;
    const float* q = p + (size_t)r * ld + k;
    if (vec && k + 3 < k_end) return __ldg(reinterpret_cast<const float4*>(q));
    if (k + 0 < k_end) v.x = __ldg(q + 0);
    if (k + 1 < k_end) v.y = __ldg(q + 1);
    if (k + 2 < k_end) v.z = __ldg(q + 2);
    if (k + 3 < k_end) v.w = __ldg(q + 3);
    return v;
  }
};
struct ALoadContigK : DenseContigK {
  typedef int State;
  __device__ __forceinline__ void init(State& st, int m, int) const { st = m; }
  __device__ __forceinline__ float4 fetch(State& st, int k, int k_end) const { return get(st, k, k_end); }
};
struct BLoadContigK : DenseContigK {
  typedef int State;
  __device__ __forceinline__ void init(State& st, int n, int) const { st = n; }
  __device__ __forceinline__ float4 fetch(State& st, int k, int k_end) const { return get(st, k, k_end); }
};
// Row-major matrix whose contiguous dimension is NOT K (A given as [K,M], or B given as [K,N]).
struct DenseContigMN {
  static constexpr bool kContigK = false;
  const float* p;
  int cols, ld;  // cols = extent of the contiguous (M or N) dimension
  bool vec;
  __device__ __forceinline__ float4 get(long long k, int c, long long k_end) const {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k >= k_end) return v;
    const float* q = p + (size_t)k * ld + c;
    if (vec && c + 3 < cols) return __ldg(reinterpret_cast<const float4*>(q));
    if (c + 0 < cols) v.x = __ldg(q + 0);
    if (c + 1 < cols) v.y = __ldg(q + 1);
    if (c + 2 < cols) v.z = __ldg(q + 2);
    if (c + 3 < cols) v.w = __ldg(q + 3);
    return v;
  }
};
struct ALoadContigM : DenseContigMN {
  typedef int State;
  __device__ __forceinline__ void init(State& st, int m, int) const { st = m; }
  __device__ __forceinline__ float4 fetch(State& st, int k, int k_end) const { return get(k, st, k_end); }
};
struct BLoadContigN : DenseContigMN {
  typedef int State;
  __device__ __forceinline__ void init(State& st, int n, int) const { st = n; }
  __device__ __forceinline__ float4 fetch(State& st, int k, int k_end) const { return get(k, st, k_end); }
};

// ---- dense epilogue: C = [C +] acc [+ bias[n]] [relu] ---------------------------------------------------------
struct DenseEpilogue {
  float* c;
  int M, N, ldc;
  const float* bias;  // [N] or nullptr
  bool relu, accumulate, vec;
  __device__ __forceinline__ void finish() const {}
  __device__ __forceinline__ void operator()(int m, int n, float4 v, int) const {
    if (m >= M || n >= N) return;
    float r[4] = {v.x, v.y, v.z, v.w};
    float* q = c + (size_t)m * ldc + n;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (n + j < N) {
        float x = r[j];
        if (bias) x += __ldg(bias + n + j);
        if (accumulate) x += q[j];
        if (relu) x = fmaxf(x, 0.f);
        r[j] = x;
      }
    }
    if (vec && n + 3 < N) {
      *reinterpret_cast<float4*>(q) = make_float4(r[0], r[1], r[2], r[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n + j < N) q[j] = r[j];
    }
  }
};

}  // namespace vocr
