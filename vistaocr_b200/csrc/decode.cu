// Greedy CTC decode on sm_100a: fused per-frame argmax + blank/low-confidence mapping (kernel 1, the HBM-bound
// one: reads T*B*A*4 bytes once) and collapse-repeats + drop-blanks with warp-level stream compaction (kernel 2,
// reads T*B*4 bytes).  Replaces the per-timestep D2H + NumPy/Python loops of ArgmaxDecoder.decode
// (reference src/decoder.py:116-185; twin src/models/cnnlstm.py:479-541).
//
// HBM layout: logits [T,B,A] fp32 is treated as a flat array of T*B rows of A floats.  A CTA owns a tile of
// kRowsPerTile consecutive rows and stages it in shared memory with one 1-D bulk async copy (TMA engine,
// mbarrier completion); 8 warps then reduce the rows out of shared memory, 8 lanes per row.
// Rows of A*4 bytes are in general not 16-B aligned, which is why the tile — not the row — is the copy unit:
// kRowsPerTile % 4 == 0 keeps every tile start 16-B aligned for any A.
#include "common.cuh"

namespace vocr {

constexpr int kDecThreads = 256;
constexpr int kDecWarps = kDecThreads / 32;

// numpy ordering: NaN is maximal, otherwise plain '>' (common.cuh).
__device__ __forceinline__ bool dec_gt(float a, float b) { return argmax_gt(a, b); }

// 8 lanes per row, 4 rows per warp pass: each lane scans A/8 elements serially, then 3 shuffle rounds finish the
// row.  (A full-warp-per-row reduction costs 5 rounds x 2 shuffles per 480-byte row and made the kernel
// issue-bound at ~30% of HBM peak; this form issues ~4x fewer instructions per row.)
constexpr int kLanesPerRow = 8;
constexpr int kRowsPerWarpPass = 32 / kLanesPerRow;

__device__ __forceinline__ void row_argmax8(const float* __restrict__ row, bool valid, int A, float& best_v,
                                            int& best_i) {
  const int l8 = lane_id() & (kLanesPerRow - 1);
  float v = 0.f;
  int i = 0x7fffffff;
  if (valid) {
    if (l8 < A) {
      v = row[l8];
      i = l8;
    }
    for (int a = l8 + kLanesPerRow; a < A; a += kLanesPerRow) {
      const float x = row[a];
      if (dec_gt(x, v)) {
        v = x;
        i = a;
      }
    }
  }
#pragma unroll
  for (int o = kLanesPerRow / 2; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, i, o);
    // an empty lane carries i = INT_MAX and loses every comparison below
    const bool take = (oi != 0x7fffffff) && (i == 0x7fffffff || dec_gt(ov, v) || (!dec_gt(v, ov) && oi < i));
    if (take) {
      v = ov;
      i = oi;
    }
  }
  best_v = v;
  best_i = i;
}

// Kernel 1.  grid = number of row tiles, block = 256.  dynamic smem = rows_per_tile*A*4 (+16 for the mbarrier).
template <bool kStaged>
__global__ void __launch_bounds__(kDecThreads)
argmax_path_kernel(const float* __restrict__ logits, long long n_rows, int T, int B, int A, int rows_per_tile,
                   const int32_t* __restrict__ lens, float thresh, int32_t* __restrict__ path) {
  extern __shared__ __align__(128) unsigned char dec_smem[];
  pdl_trigger();  // the collapse kernel may be scheduled behind this grid's last wave
  const long long row0 = (long long)blockIdx.x * rows_per_tile;
  const int rows_here = (int)min((long long)rows_per_tile, n_rows - row0);
  const float* gsrc = logits + row0 * A;
  const float* rows;
  if (kStaged) {
    uint64_t* bar = reinterpret_cast<uint64_t*>(dec_smem);
    float* tile = reinterpret_cast<float*>(dec_smem + 16);
    const uint32_t bytes = (uint32_t)rows_here * (uint32_t)A * 4u;
    const bool bulk_ok = (bytes % 16u) == 0u;  // only a partial last tile can fail this
    if (bulk_ok) {
      if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(bar, bytes);
        bulk_g2s(tile, gsrc, bytes, bar);
      }
      mbar_wait_or_trap(bar, 0);
    } else {
      const int n = rows_here * A;
      for (int i = threadIdx.x; i < n; i += kDecThreads) tile[i] = __ldg(gsrc + i);
      __syncthreads();
    }
    rows = tile;
  } else {
    rows = gsrc;
  }
  const int warp = threadIdx.x >> 5;
  const int sub = lane_id() / kLanesPerRow;
  // (t, b) of the tile's first row once per CTA; rows inside the tile then need only a 32-bit division (the 64-bit
  // division per row was a visible share of the instruction stream of this issue-bound kernel)
  const int t0 = (int)(row0 / B);
  const unsigned b0 = (unsigned)(row0 - (long long)t0 * B);
  for (int r0 = warp * kRowsPerWarpPass; r0 < rows_here; r0 += kDecWarps * kRowsPerWarpPass) {
    const int r = r0 + sub;
    const bool valid = r < rows_here;
    float mv;
    int mi;
    row_argmax8(rows + (size_t)r * A, valid, A, mv, mi);
    if (valid && (lane_id() & (kLanesPerRow - 1)) == 0) {
      const unsigned br = b0 + (unsigned)r, dt = br / (unsigned)B;
      const int t = t0 + (int)dt, b = (int)(br - dt * (unsigned)B);
      int label = -1;
      if (t < lens[b]) label = (mi == 0 || mv < thresh) ? 0 : mi;
      path[(size_t)b * T + t] = label;
    }
  }
}

// Kernel 2.  One warp per line: collapse repeats, drop blanks, compact with ballot/popc.
__global__ void __launch_bounds__(kDecThreads)
collapse_compact_kernel(const int32_t* __restrict__ path, int T, int B, const int32_t* __restrict__ lens,
                        const int32_t* __restrict__ canon, int32_t* __restrict__ labels,
                        int32_t* __restrict__ counts, int ld) {
  const int b = blockIdx.x * kDecWarps + (threadIdx.x >> 5);
  if (b >= B) return;
  const int lane = lane_id();
  const int len = max(0, min(lens[b], T));
  const int32_t* p = path + (size_t)b * T;
  int32_t* out = labels + (size_t)b * ld;
  pdl_wait();  // `path` comes from the arg-max kernel this launch depends on
  int base = 0;
  int carry = 0;  // canonical label of the frame before this chunk (0 = no previous character)
  // the loads of chunk k + 1 are issued before chunk k is processed: the loop runs at shuffle latency, not L2 latency
  int cur = (lane < len) ? p[lane] : 0;
  int cc = (cur > 0) ? (canon ? canon[cur] : cur) : 0;
  for (int t0 = 0; t0 < len; t0 += 32) {
    const int tn = t0 + 32 + lane;
    const int cur_n = (tn < len) ? p[tn] : 0;
    const int cc_n = (cur_n > 0) ? (canon ? canon[cur_n] : cur_n) : 0;
    int pc = __shfl_up_sync(0xffffffffu, cc, 1);
    if (lane == 0) pc = carry;
    const bool emit = (cur > 0) && (pc != cc);
    const unsigned m = __ballot_sync(0xffffffffu, emit);
    if (emit) {
      const int pos = base + __popc(m & ((1u << lane) - 1u));
      if (pos < ld) out[pos] = cur;
    }
    base += __popc(m);
    carry = __shfl_sync(0xffffffffu, cc, 31);
    cur = cur_n;
    cc = cc_n;
  }
  if (lane == 0) counts[b] = min(base, ld);
}

}  // namespace vocr

using namespace vocr;

extern "C" int vocr_greedy_decode_f32(const float* logits, int T, int B, int A, const int32_t* lens, float thresh,
                                      const int32_t* canon, int32_t* path, int32_t* labels, int32_t* counts,
                                      int ld, vocr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VOCR_REQUIRE(T >= 0 && B >= 0 && A >= 1 && ld >= 0);
  if (B == 0) return VOCR_OK;
  VOCR_REQUIRE(lens && labels && counts);
  if (T > 0) {
    VOCR_REQUIRE(logits && path);
    const long long n_rows = (long long)T * B;
    // tile = multiple of 4 rows (16-B aligned starts), at most ~32 KB of shared memory
    int rows_per_tile = 64;
    while (rows_per_tile > 4 && (size_t)rows_per_tile * A * 4 > 32768) rows_per_tile >>= 1;
    const size_t tile_bytes = (size_t)rows_per_tile * A * 4;
    const bool staged = (reinterpret_cast<uintptr_t>(logits) % 16 == 0) && tile_bytes <= 40960;
    const long long n_tiles = ceil_div64(n_rows, rows_per_tile);
    VOCR_REQUIRE(n_tiles < (1ll << 31));
    if (staged) {
      argmax_path_kernel<true><<<(unsigned)n_tiles, kDecThreads, tile_bytes + 16, stream>>>(
          logits, n_rows, T, B, A, rows_per_tile, lens, thresh, path);
    } else {
      argmax_path_kernel<false><<<(unsigned)n_tiles, kDecThreads, 0, stream>>>(
          logits, n_rows, T, B, A, rows_per_tile, lens, thresh, path);
    }
    VOCR_CHECK_LAUNCH();
  }
  // programmatic dependent launch: the collapse grid is scheduled while the arg-max grid drains
  if (launch_pdl(collapse_compact_kernel, dim3(ceil_div(B, kDecWarps)), dim3(kDecThreads), 0, stream, path, T, B, lens,
                 canon, labels, counts, ld) != cudaSuccess)
    return VOCR_EXECUTION_FAILED;
  return VOCR_OK;
}

// Second half alone: path [B,T] (frame labels: -1 beyond the line, 0 = blank / low confidence, else the arg-max) ->
// collapsed label strings.  For producers that already hold the frame path - the prob-layer GEMM with the arg-max in its
// epilogue (vocr_tc_gemm_f16x3_argmax) never writes the logits.
extern "C" int vocr_ctc_collapse_i32(const int32_t* path, int T, int B, const int32_t* lens, const int32_t* canon,
                                     int32_t* labels, int32_t* counts, int ld, vocr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VOCR_REQUIRE(T >= 0 && B >= 0 && ld >= 0);
  if (B == 0) return VOCR_OK;
  VOCR_REQUIRE(lens && labels && counts && (T == 0 || path));
  collapse_compact_kernel<<<ceil_div(B, kDecWarps), kDecThreads, 0, stream>>>(path, T, B, lens, canon, labels, counts, ld);
  VOCR_CHECK_LAUNCH();
  return VOCR_OK;
}
