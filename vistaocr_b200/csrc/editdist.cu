// Batched Levenshtein distance on int32 sequences: the O(n*m) dynamic programme behind the
// reference's CER / WER scoring (src/textutils.py:264-287 edit_distance, called per line from compute_cer_wer
// :326-351 inside test_on_val, src/train_cnn_lstm.py:61-79 - a NumPy double loop on the host today).
// One CTA per (hypothesis, reference) pair; cells on an anti-diagonal are independent and are computed in parallel,
// three diagonals live in shared memory.  Integer work: results are bit-exact.
#include "common.cuh"

namespace vocr {

__global__ void __launch_bounds__(128)
edit_distance_kernel(const int32_t* __restrict__ a_flat, const int32_t* __restrict__ a_off,
                     const int32_t* __restrict__ b_flat, const int32_t* __restrict__ b_off, int32_t* __restrict__ dist,
                     int max_n, int max_m) {
  extern __shared__ int32_t ed_smem[];
  const int p = blockIdx.x;
  const int n = a_off[p + 1] - a_off[p], m = b_off[p + 1] - b_off[p];
  if (n == 0 || m == 0) {
    if (threadIdx.x == 0) dist[p] = n + m;
    return;
  }
  int32_t* sa = ed_smem;                // [max_n]
  int32_t* sb = sa + max_n;             // [max_m]
  int32_t* d0 = sb + max_m;             // three diagonals, indexed by i in [0, n]
  int32_t* d1 = d0 + (max_n + 1);
  int32_t* d2 = d1 + (max_n + 1);
  for (int i = threadIdx.x; i < n; i += blockDim.x) sa[i] = a_flat[a_off[p] + i];
  for (int j = threadIdx.x; j < m; j += blockDim.x) sb[j] = b_flat[b_off[p] + j];
  // diagonal k holds D(i, k - i).  k = 0: D(0,0) = 0;  k = 1: D(0,1) = 1, D(1,0) = 1
  if (threadIdx.x == 0) {
    d0[0] = 0;
    d1[0] = 1;
    d1[1] = 1;
  }
  __syncthreads();
  int32_t *pp = d0, *pv = d1, *cur = d2;  // k-2, k-1, k
  for (int k = 2; k <= n + m; ++k) {
    const int ilo = max(0, k - m), ihi = min(n, k);
    for (int i = ilo + threadIdx.x; i <= ihi; i += blockDim.x) {
      const int j = k - i;
      int v;
      if (i == 0) v = j;
      else if (j == 0) v = i;
      else if (sa[i - 1] == sb[j - 1]) v = pp[i - 1];
      else v = 1 + min(pp[i - 1], min(pv[i - 1], pv[i]));  // D(i-1,j-1), D(i-1,j), D(i,j-1)
      cur[i] = v;
    }
    __syncthreads();
    int32_t* t = pp;
    pp = pv;
    pv = cur;
    cur = t;
  }
  if (threadIdx.x == 0) dist[p] = pv[n];  // diagonal n+m, i = n
}

}  // namespace vocr

using namespace vocr;

// a_off / b_off: [P+1] offsets into the flat sequences; max_n / max_m: upper bounds on the lengths (sizes shared mem).
extern "C" int vocr_edit_distance_i32(const int32_t* a_flat, const int32_t* a_off, const int32_t* b_flat,
                                      const int32_t* b_off, int P, int max_n, int max_m, int32_t* dist,
                                      vocr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VOCR_REQUIRE(P >= 0 && max_n >= 0 && max_m >= 0);
  if (P == 0) return VOCR_OK;
  VOCR_REQUIRE(a_off && b_off && dist && (max_n == 0 || a_flat) && (max_m == 0 || b_flat));
  const size_t smem = sizeof(int32_t) * ((size_t)max_n + max_m + 3 * ((size_t)max_n + 1));
  VOCR_REQUIRE(smem <= 200 * 1024);
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(edit_distance_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return VOCR_EXECUTION_FAILED;
  edit_distance_kernel<<<P, 128, smem, stream>>>(a_flat, a_off, b_flat, b_off, dist, max_n, max_m);
  VOCR_CHECK_LAUNCH();
  return VOCR_OK;
}
