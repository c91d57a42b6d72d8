// Dense fp32 GEMM with fused bias / ReLU / accumulate epilogue (bridge layer, LSTM input projections, prob layer and
// their backward GEMMs; reference cnnlstm.py:143-154,278,294 -> cuBLAS).  SIMT FFMA engine in gemm_core.cuh.
#include "gemm_core.cuh"

namespace vocr {

template <int BN, class AL, class BL>
__global__ void __launch_bounds__(kGemmThreads)
dense_gemm_kernel(AL la, BL lb, DenseEpilogue ep, int K) {
  gemm_tile<BN>(la, lb, ep, blockIdx.x * kGemmBM, blockIdx.y * BN, 0, K);  // row tiles in grid.x (no 65535 limit)
}

template <class AL, class BL>
static int launch_dense(const AL& la, const BL& lb, const DenseEpilogue& ep, int M, int N, int K,
                        cudaStream_t stream) {
  if (N <= 64) {
    dim3 grid(ceil_div(M, kGemmBM), ceil_div(N, 64));
    dense_gemm_kernel<64, AL, BL><<<grid, kGemmThreads, 0, stream>>>(la, lb, ep, K);
  } else {
    dim3 grid(ceil_div(M, kGemmBM), ceil_div(N, 128));
    dense_gemm_kernel<128, AL, BL><<<grid, kGemmThreads, 0, stream>>>(la, lb, ep, K);
  }
  VOCR_CHECK_LAUNCH();
  return VOCR_OK;
}

__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ x, long long rows, int cols, int ld, int rows_per_cta,
              float* __restrict__ out, int accumulate) {
  // grid.x over column groups of 32*? ; grid.y over row chunks; partial sums -> atomicAdd (fp32)
  __shared__ float part[8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int ry = threadIdx.x >> 5;
  const long long r0 = (long long)blockIdx.y * rows_per_cta;
  const long long r1 = min(rows, r0 + rows_per_cta);
  float s = 0.f;
  if (c < cols)
    for (long long r = r0 + ry; r < r1; r += 8) s += __ldg(x + r * ld + c);
  part[ry][threadIdx.x & 31] = s;
  __syncthreads();
  if (ry == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i][threadIdx.x & 31];
    atomicAdd(out + c, t);
  }
  (void)accumulate;
}

// vectorised variant (cols % 4 == 0, ld % 4 == 0, 16-B aligned base): a warp covers 128 columns per row with one
// 512-byte request, 8 warps x 4 rows in flight per CTA
__global__ void __launch_bounds__(256)
colsum_vec4_kernel(const float* __restrict__ x, long long rows, int cols4, int ld4, int rows_per_cta,
                   float* __restrict__ out) {
  __shared__ float4 part[8][32];
  const int lane = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c4 = blockIdx.x * 32 + lane;
  const long long r0 = (long long)blockIdx.y * rows_per_cta;
  const long long r1 = min(rows, r0 + rows_per_cta);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c4 < cols4) {
    const float4* p = reinterpret_cast<const float4*>(x) + c4;
    long long r = r0 + ry;
    for (; r + 24 < r1; r += 32) {
      const float4 a = __ldg(p + r * ld4), b = __ldg(p + (r + 8) * ld4), c = __ldg(p + (r + 16) * ld4),
                   d = __ldg(p + (r + 24) * ld4);
      s.x += (a.x + b.x) + (c.x + d.x);
      s.y += (a.y + b.y) + (c.y + d.y);
      s.z += (a.z + b.z) + (c.z + d.z);
      s.w += (a.w + b.w) + (c.w + d.w);
    }
    for (; r < r1; r += 8) {
      const float4 a = __ldg(p + r * ld4);
      s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
    }
  }
  part[ry][lane] = s;
  __syncthreads();
  if (ry == 0 && c4 < cols4) {
    float4 t = part[0][lane];
#pragma unroll
    for (int i = 1; i < 8; ++i) {
      t.x += part[i][lane].x; t.y += part[i][lane].y; t.z += part[i][lane].z; t.w += part[i][lane].w;
    }
    float* o = out + (size_t)c4 * 4;
    atomicAdd(o, t.x);
    atomicAdd(o + 1, t.y);
    atomicAdd(o + 2, t.z);
    atomicAdd(o + 3, t.w);
  }
}

}  // namespace vocr

using namespace vocr;

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int vocr_gemm_f32(int transa, int transb, int M, int N, int K, const float* A, int lda, const float* B,
                             int ldb, float* C, int ldc, const float* bias, int relu, int accumulate,
                             vocr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VOCR_REQUIRE(M >= 0 && N >= 0 && K >= 0);
  if (M == 0 || N == 0) return VOCR_OK;
  VOCR_REQUIRE(C && (K == 0 || (A && B)));
  DenseEpilogue ep{C, M, N, ldc, bias, relu != 0, accumulate != 0, aligned16(C) && (ldc % 4 == 0)};
  if (!transa && transb) {
    ALoadContigK la;
    la.p = A; la.rows = M; la.ld = lda; la.vec = aligned16(A) && (lda % 4 == 0);
    BLoadContigK lb;
    lb.p = B; lb.rows = N; lb.ld = ldb; lb.vec = aligned16(B) && (ldb % 4 == 0);
    return launch_dense(la, lb, ep, M, N, K, stream);
  } else if (!transa && !transb) {
    ALoadContigK la;
    la.p = A; la.rows = M; la.ld = lda; la.vec = aligned16(A) && (lda % 4 == 0);
    BLoadContigN lb;
    lb.p = B; lb.cols = N; lb.ld = ldb; lb.vec = aligned16(B) && (ldb % 4 == 0);
    return launch_dense(la, lb, ep, M, N, K, stream);
  } else if (transa && !transb) {
    ALoadContigM la;
    la.p = A; la.cols = M; la.ld = lda; la.vec = aligned16(A) && (lda % 4 == 0);
    BLoadContigN lb;
    lb.p = B; lb.cols = N; lb.ld = ldb; lb.vec = aligned16(B) && (ldb % 4 == 0);
    return launch_dense(la, lb, ep, M, N, K, stream);
  } else {
    ALoadContigM la;
    la.p = A; la.cols = M; la.ld = lda; la.vec = aligned16(A) && (lda % 4 == 0);
    BLoadContigK lb;
    lb.p = B; lb.rows = N; lb.ld = ldb; lb.vec = aligned16(B) && (ldb % 4 == 0);
    return launch_dense(la, lb, ep, M, N, K, stream);
  }
}

extern "C" int vocr_colsum_f32(const float* x, long long rows, int cols, int ld, float* out, int accumulate,
                               vocr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VOCR_REQUIRE(rows >= 0 && cols >= 0 && out);
  if (cols == 0) return VOCR_OK;
  if (!accumulate)
    if (cudaMemsetAsync(out, 0, sizeof(float) * cols, stream) != cudaSuccess) return VOCR_MEMOPS_FAILED;
  if (rows == 0) return VOCR_OK;
  VOCR_REQUIRE(x);
  if (cols % 4 == 0 && ld % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    const int gx4 = ceil_div(cols / 4, 32);
    const long long want = max(1, (4 * kNumSMs) / gx4);
    int rpc = (int)max(64ll, ceil_div64(rows, want));
    rpc = (rpc + 31) & ~31;
    dim3 grid4(gx4, (unsigned)ceil_div64(rows, rpc));
    colsum_vec4_kernel<<<grid4, 256, 0, stream>>>(x, rows, cols / 4, ld / 4, rpc, out);
    VOCR_CHECK_LAUNCH();
    return VOCR_OK;
  }
  const int gx = ceil_div(cols, 32);
  long long want_y = max(1, (2 * kNumSMs) / gx);
  int rows_per_cta = (int)max(64ll, ceil_div64(rows, want_y));
  rows_per_cta = (rows_per_cta + 7) & ~7;
  dim3 grid(gx, (unsigned)ceil_div64(rows, rows_per_cta));
  colsum_kernel<<<grid, 256, 0, stream>>>(x, rows, cols, ld, rows_per_cta, out, accumulate);
  VOCR_CHECK_LAUNCH();
  return VOCR_OK;
}
