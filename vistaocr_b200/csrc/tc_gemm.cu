// Tensor-core GEMM for sm_100a: TMA-fed tcgen05.mma (kind::tf32) with the accumulator in TMEM, error-compensated
// "3xTF32" so the result keeps fp32-level accuracy (north_star: 1e-5 relative in fp32):
//     C = A_hi B_hi + A_lo B_hi + A_hi B_lo,     x_hi = rna_tf32(x),  x_lo = x - x_hi   (split done by the producer)
// Replaces the FFMA engine (gemm_core.cuh) for the dense GEMMs of the path (LSTM input projections, bridge / prob
// layers and their backward GEMMs; reference cnnlstm.py:143-154 -> cuBLAS).
//
// One CTA = one 128 x 128 output tile, 192 threads, warp-specialised:
//   warp 0   TMA producer  : 4 operand tiles (A_hi, A_lo, B_hi, B_lo; 128 x 32 fp32 = 16 KB each, SWIZZLE_128B) per
//                            k-block into a 3-stage shared-memory ring, mbarrier expect_tx / complete_tx
//   warp 1   MMA issuer    : allocates 512 TMEM columns, one elected lane issues 12 tcgen05.mma per k-block
//                            (4 k-steps of 8 x 3 products), tcgen05.commit releases the stage / signals the epilogue
//   warps 2-5 epilogue     : tcgen05.ld (32 lanes x 32 columns per warp and pass) -> bias / ReLU / accumulate -> global
// Both operand layouts are supported without any transpose pass: K-major (the reduction dimension is contiguous in
// memory: activations [M,K], nn.Linear weights [N,K]) and MN-major (the reduction dimension is the row index: dY^T X
// weight gradients, dY W data gradients); they differ only in the TMA box shape and the UMMA descriptor strides.
// Every mbarrier wait is bounded and traps, so a descriptor bug surfaces as a launch error, not a hung GPU.
#include <cuda_fp16.h>

#include "tc_common.cuh"

namespace vocr {

constexpr int kTcBM = 128, kTcBN = 128, kTcBK = 32;  // BK fp32 = 128 bytes = one swizzle row
constexpr int kTcStages = 3;
constexpr int kTcThreads = 192;
constexpr int kTcTileBytes = kTcBM * kTcBK * 4;         // 16 KB
constexpr int kTcStageBytes = 4 * kTcTileBytes;         // A_hi, A_lo, B_hi, B_lo
constexpr int kTcSmemBytes = kTcStages * kTcStageBytes + 1024 /*align*/ + 256 /*barriers*/;
// persistent kernel: 8 epilogue warps (two per TMEM quadrant, 64 columns each) + a 32 x 32 fp32 transposition buffer per
// epilogue warp (16-byte chunks XOR-swizzled by the row), see its epilogue
constexpr int kTcPersistThreads = 64 + 8 * 32;
constexpr int kTcPersistSmemBytes = kTcSmemBytes + 8 * 32 * 32 * 4;
constexpr uint32_t kTmemCols = 512;  // 3 rotating hi*hi accumulators + 1 for the lo products, 128 columns each
constexpr int kTcHiAcc = 3;

struct TcGemmParams {
  float* c;        // output, or the split-K partial buffer [splits][M][N] when splits > 1
  const float* bias;
  int M, N, K, ldc;
  int relu, accumulate;
  int a_mn, b_mn;  // 1 = MN-major operand
  int kb_per_split;
  const int* exp_a;  // FP16 pair operands: planes hold A * 2^exp_a[0], B * 2^exp_b[0] (device scalars)
  const int* exp_b;
  int single;        // 1 = one product per k-step on the hi planes only (reduced-precision mode: fp16 operands, fp32
                     //     accumulation; the lo planes are neither loaded nor multiplied)
  int cm, cn;        // thread-block cluster = cm x cn neighbouring output tiles (1 or 2 each); see TcClusterPos
  // greedy-decode epilogue of the persistent kernel (vocr_tc_gemm_f16x3_argmax; N <= 128): row m = t * dec_B + b of the
  // prob-layer output is reduced to its frame label path[b * dec_T + t] (decode.cu semantics); c may then be null
  int32_t* path;
  const int32_t* lens;
  int dec_T, dec_B;
  float thresh;
};
constexpr int kTcChunk = 8;  // k-blocks accumulated in TMEM before the epilogue warps drain them into fp32 registers

// Cluster of cm x cn output tiles (CTA rank r -> tile row rm = r / cn, tile column rn = r % cn of the cluster's
// super-tile).  The cn CTAs of a tile row need the same A tile and the cm CTAs of a tile column the same B tile:
// CTA (rm, rn) fetches part rn of its A tile and part rm of its B tile and multicasts them (tc_common.cuh), so per
// k-block a CTA pulls 1/cn of A and 1/cm of B through its own L2 port instead of both tiles - at 2 x 2 half the bytes.
// A CTA may refill a stage once every CTA its parts land in has retired the MMAs that read the stage: the MMA warp's
// commit arrives on `empty` of every CTA in mask_a | mask_b (the senders into this CTA), and `empty` counts cm + cn - 1.
struct TcClusterPos {
  int size, rm, rn;
  uint16_t mask_a, mask_b;
};
__device__ __forceinline__ TcClusterPos tc_cluster_pos(const TcGemmParams& p) {
  TcClusterPos c;
  c.size = p.cm * p.cn;
  const int r = c.size > 1 ? (int)tc_cluster_rank() : 0;
  c.rm = r / p.cn;
  c.rn = r - c.rm * p.cn;
  c.mask_a = (uint16_t)(((1u << p.cn) - 1u) << (c.rm * p.cn));
  c.mask_b = 0;
  for (int i = 0; i < p.cm; ++i) c.mask_b |= (uint16_t)(1u << (i * p.cn + c.rn));
  return c;
}

// One operand tile (128 rows of the M or N dimension x one k-block) of both planes: part `part` of `parts`, multicast
// to `mask` when the tile is shared.  K-major: box {BK k, 128 / parts rows}; MN-major: boxes {32|64 m, BK k rows}.
template <bool F16>
__device__ __forceinline__ void tc_load_operand(unsigned char* hi, unsigned char* lo, const CUtensorMap* map_hi,
                                                const CUtensorMap* map_lo, uint64_t* full, int mn, int parts, int part,
                                                uint16_t mask, int k0, int r0, int single) {
  using E = TcElem<F16>;
  if (!mn) {
    if (parts == 1) {
      tma_load_2d(hi, map_hi, full, k0, r0);
      if (!single) tma_load_2d(lo, map_lo, full, k0, r0);
    } else {
      const int rows = kTcBM / parts;
      const uint32_t off = (uint32_t)(part * rows) * 128u;
      tma_load_2d_mc(hi + off, map_hi, full, k0, r0 + part * rows, mask);
      if (!single) tma_load_2d_mc(lo + off, map_lo, full, k0, r0 + part * rows, mask);
    }
  } else {
    const int per = (kTcBM / E::kMnBox) / parts;
    for (int j = part * per; j < (part + 1) * per; ++j) {
      if (parts == 1) {
        tma_load_2d(hi + j * E::kMnBoxBytes, map_hi, full, r0 + E::kMnBox * j, k0);
        if (!single) tma_load_2d(lo + j * E::kMnBoxBytes, map_lo, full, r0 + E::kMnBox * j, k0);
      } else {
        tma_load_2d_mc(hi + j * E::kMnBoxBytes, map_hi, full, r0 + E::kMnBox * j, k0, mask);
        if (!single) tma_load_2d_mc(lo + j * E::kMnBoxBytes, map_lo, full, r0 + E::kMnBox * j, k0, mask);
      }
    }
  }
}
// all four tiles of a stage (A_hi, A_lo, B_hi, B_lo); `full` expects the whole stage whoever delivers it
template <bool F16>
__device__ __forceinline__ void tc_load_stage(unsigned char* st, uint64_t* full, const CUtensorMap* ma_hi,
                                              const CUtensorMap* ma_lo, const CUtensorMap* mb_hi,
                                              const CUtensorMap* mb_lo, const TcGemmParams& p, const TcClusterPos& c,
                                              int k0, int m0, int n0) {
  mbar_arrive_expect_tx(full, p.single ? kTcStageBytes / 2 : kTcStageBytes);
  tc_load_operand<F16>(st, st + kTcTileBytes, ma_hi, ma_lo, full, p.a_mn, p.cn, c.rn, c.mask_a, k0, m0, p.single);
  tc_load_operand<F16>(st + 2 * kTcTileBytes, st + 3 * kTcTileBytes, mb_hi, mb_lo, full, p.b_mn, p.cm, c.rm, c.mask_b,
                       k0, n0, p.single);
}
// frees a stage: locally, or in every CTA whose loads land in this CTA's stage
__device__ __forceinline__ void tc_release_stage(uint64_t* empty, const TcClusterPos& c) {
  if (c.size > 1) umma_commit_mc(empty, (uint16_t)(c.mask_a | c.mask_b));
  else umma_commit(empty);
}

__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Short reductions (<= 96 k-blocks, no split-K): hi*hi rotates over the 3 accumulators per k-block and everything is
// drained once at the end (measured ~20 % faster than the chunked pipeline below at K = 1024).
template <bool F16>
__global__ void __launch_bounds__(kTcThreads, 1)
tc_gemm_x3_shortk_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                      const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                      TcGemmParams p) {
  extern __shared__ unsigned char tc_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) &
                                                         ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kTcStages * kTcStageBytes);
  uint64_t* full_bar = bars;                     // [stages]
  uint64_t* empty_bar = bars + kTcStages;        // [stages]
  uint64_t* tmem_full_bar = bars + 2 * kTcStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kTcStages + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const TcClusterPos cp = tc_cluster_pos(p);  // grid = (n tiles, m tiles) padded to the cluster shape (cn, cm)
  const int m0 = ((int)(blockIdx.y / p.cm) * p.cm + cp.rm) * kTcBM, n0 = ((int)(blockIdx.x / p.cn) * p.cn + cp.rn) * kTcBN;
  using E = TcElem<F16>;
  constexpr int BK = E::kBK;
  const int num_kb = (p.K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kTcStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], (uint32_t)(p.cm + p.cn - 1));
    }
    mbar_init(tmem_full_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (p.cm * p.cn > 1) tc_cluster_sync();  // every CTA's barriers exist before a peer can signal them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kTcStages;
        const uint32_t ph = (uint32_t)(kb / kTcStages) & 1u;
        mbar_wait_or_trap(&empty_bar[s], ph ^ 1u);
        tc_load_stage<F16>(smem + (size_t)s * kTcStageBytes, &full_bar[s], &map_a_hi, &map_a_lo, &map_b_hi, &map_b_lo, p,
                           cp, kb * BK, m0, n0);
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      // instruction descriptor: D = F32 (1 << 4), A/B = TF32 (2 << 7, 2 << 10), majors, N >> 3 at bit 17, M >> 4 at 24
      const uint32_t idesc = (1u << 4) | (E::kFmt << 7) | (E::kFmt << 10) | ((uint32_t)p.a_mn << 15) |
                             ((uint32_t)p.b_mn << 16) | ((uint32_t)(kTcBN >> 3) << 17) | ((uint32_t)(kTcBM >> 4) << 24);
      // per k-step (32 bytes of K): K-major advances 32 B inside the swizzle row, MN-major advances its k rows
      const uint32_t a_step = p.a_mn ? E::kMnStep : 32u, b_step = p.b_mn ? E::kMnStep : 32u;
      const uint32_t a_lbo = p.a_mn ? E::kMnBoxBytes : 16u, b_lbo = p.b_mn ? E::kMnBoxBytes : 16u;
      const uint32_t a_sbo = p.a_mn ? E::kMnSbo : 1024u, b_sbo = p.b_mn ? E::kMnSbo : 1024u;
      const uint32_t a_lt = p.a_mn ? E::kMnLayout : 2u, b_lt = p.b_mn ? E::kMnLayout : 2u;
      // The tensor core truncates (RZ) when it adds into the fp32 accumulator, once per instruction, by up to an ulp
      // of the RUNNING SUM - a bias that grows linearly with K.  Keep the running sums short and well scaled: the two
      // lo products go to their own accumulator (its sum is 2^-11 smaller), hi*hi rotates over 3 accumulators by
      // k-block; the epilogue adds the four in fp32 with round-to-nearest.
      const uint32_t tmem_lo = tmem_base + kTcHiAcc * kTcBN;
      uint32_t accum_lo = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kTcStages;
        const uint32_t ph = (uint32_t)(kb / kTcStages) & 1u;
        mbar_wait_or_trap(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t st = smem_u32(smem + (size_t)s * kTcStageBytes);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t a_hi = make_desc(st + ks * a_step, a_lbo, a_sbo, a_lt);
          const uint64_t a_lo = make_desc(st + kTcTileBytes + ks * a_step, a_lbo, a_sbo, a_lt);
          const uint64_t b_hi = make_desc(st + 2 * kTcTileBytes + ks * b_step, b_lbo, b_sbo, b_lt);
          const uint64_t b_lo = make_desc(st + 3 * kTcTileBytes + ks * b_step, b_lbo, b_sbo, b_lt);
          if (!p.single) {
            E::mma(tmem_lo, a_lo, b_hi, idesc, accum_lo);
            accum_lo = 1;
            E::mma(tmem_lo, a_hi, b_lo, idesc, 1);
          }
          E::mma(tmem_base + (uint32_t)(kb % kTcHiAcc) * kTcBN, a_hi, b_hi, idesc,
                 (kb >= kTcHiAcc || ks > 0) ? 1u : 0u);
        }
        tc_release_stage(&empty_bar[s], cp);  // frees the stage once the MMAs above have read it
      }
      umma_commit(tmem_full_bar);    // accumulator complete
    }
  } else {
    // ================================ epilogue (warps 2..5) ================================
    mbar_wait_or_trap(tmem_full_bar, 0);
    tc_fence_after();
    const int lane_grp = warp & 3;  // TMEM lanes [32*lane_grp, +32) are the ones this warp may read
    const int m = m0 + lane_grp * 32 + lane;
    float* crow = p.c + (size_t)m * p.ldc;
    const bool vec = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.c) & 15) == 0);
    const int out_shift = F16 ? -(__ldg(p.exp_a) + __ldg(p.exp_b)) : 0;
    const int n_hi = min(kTcHiAcc, num_kb);  // hi accumulators that were actually written
    for (int cb = 0; cb < kTcBN; cb += 32) {
      float r[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) r[j] = 0.f;
      for (int acc = 0; acc < n_hi + (p.single ? 0 : 1); ++acc) {  // acc == n_hi -> the lo accumulator
        const int which = (acc == n_hi) ? kTcHiAcc : acc;
        uint32_t t[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(which * kTcBN + cb);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]),
              "=r"(t[8]), "=r"(t[9]), "=r"(t[10]), "=r"(t[11]), "=r"(t[12]), "=r"(t[13]), "=r"(t[14]), "=r"(t[15]),
              "=r"(t[16]), "=r"(t[17]), "=r"(t[18]), "=r"(t[19]), "=r"(t[20]), "=r"(t[21]), "=r"(t[22]), "=r"(t[23]),
              "=r"(t[24]), "=r"(t[25]), "=r"(t[26]), "=r"(t[27]), "=r"(t[28]), "=r"(t[29]), "=r"(t[30]), "=r"(t[31])
            : "r"(taddr)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const float w = (F16 && which == kTcHiAcc) ? 1.f / kPairLoScale : 1.f;  // lo products carry the 2^11
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = fmaf(__uint_as_float(t[j]), w, r[j]);
      }
      if (F16) {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = scale_pow2(r[j], out_shift);
      }
      if (m < p.M) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const int n = n0 + cb + j;
          if (n >= p.N) break;
          float v[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float x = r[j + q];
            if (n + q < p.N) {
              if (p.bias) x += __ldg(p.bias + n + q);
              if (p.accumulate) x += crow[n + q];
              if (p.relu) x = fmaxf(x, 0.f);
            }
            v[q] = x;
          }
          if (vec && n + 3 < p.N) {
            *reinterpret_cast<float4*>(crow + n) = make_float4(v[0], v[1], v[2], v[3]);
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (n + q < p.N) crow[n + q] = v[q];
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (p.cm * p.cn > 1) tc_cluster_sync();  // no CTA leaves while a peer's commit could still arrive here
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}


// Short reductions, PERSISTENT: one CTA per SM walks over output tiles; the accumulators are double-buffered in TMEM
// (set = {hi*hi, lo products} x 128 columns, two sets = 512 columns), so the epilogue warps drain tile j while the MMA
// warp already runs tile j+1 and the TMA warp prefetches across tile boundaries.  With K <= 24 k-blocks (<= 96
// accumulations into the hi accumulator) the round-toward-zero bias stays below 6e-6 relative; longer reductions use
// the rotating / chunked kernels.  (The xproj GEMM of the LSTM layers, K = 1024, spent ~half its time in un-overlapped
// prologues and epilogues before this.)
constexpr int kTcPersistMaxKb = 24;

template <bool F16>
__global__ void __launch_bounds__(kTcPersistThreads, 1)
tc_gemm_x3_persist_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                          const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                          TcGemmParams p, int super_n, int num_super) {
  extern __shared__ unsigned char tc_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) &
                                                         ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kTcStages * kTcStageBytes);
  uint64_t* full_bar = bars;                     // [stages]
  uint64_t* empty_bar = bars + kTcStages;        // [stages]
  uint64_t* acc_full = bars + 2 * kTcStages;     // [2] accumulator set complete
  uint64_t* acc_empty = acc_full + 2;            // [2] accumulator set drained (128 epilogue threads)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  using E = TcElem<F16>;
  constexpr int BK = E::kBK;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = (p.K + BK - 1) / BK;
  // a cluster walks over super-tiles of cm x cn output tiles; its CTAs run the same trip counts
  const TcClusterPos cp = tc_cluster_pos(p);
  const int cluster_id = blockIdx.x / cp.size, n_clusters = gridDim.x / cp.size;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kTcStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], (uint32_t)(p.cm + p.cn - 1));
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], kTcPersistThreads - 64);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  if (p.cm * p.cn > 1) tc_cluster_sync();  // every CTA's barriers exist before a peer can signal them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int it = 0;  // k-blocks issued so far (the stage ring runs across tiles)
      for (int st = cluster_id; st < num_super; st += n_clusters) {
        const int m0 = ((st / super_n) * p.cm + cp.rm) * kTcBM, n0 = ((st % super_n) * p.cn + cp.rn) * kTcBN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % kTcStages;
          const uint32_t ph = (uint32_t)(it / kTcStages) & 1u;
          mbar_wait_or_trap(&empty_bar[s], ph ^ 1u);
          tc_load_stage<F16>(smem + (size_t)s * kTcStageBytes, &full_bar[s], &map_a_hi, &map_a_lo, &map_b_hi, &map_b_lo, p,
                             cp, kb * BK, m0, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (E::kFmt << 7) | (E::kFmt << 10) | ((uint32_t)p.a_mn << 15) |
                             ((uint32_t)p.b_mn << 16) | ((uint32_t)(kTcBN >> 3) << 17) | ((uint32_t)(kTcBM >> 4) << 24);
      const uint32_t a_step = p.a_mn ? E::kMnStep : 32u, b_step = p.b_mn ? E::kMnStep : 32u;
      const uint32_t a_lbo = p.a_mn ? E::kMnBoxBytes : 16u, b_lbo = p.b_mn ? E::kMnBoxBytes : 16u;
      const uint32_t a_sbo = p.a_mn ? E::kMnSbo : 1024u, b_sbo = p.b_mn ? E::kMnSbo : 1024u;
      const uint32_t a_lt = p.a_mn ? E::kMnLayout : 2u, b_lt = p.b_mn ? E::kMnLayout : 2u;
      int it = 0, j = 0;
      for (int st = cluster_id; st < num_super; st += n_clusters, ++j) {
        const int set = j & 1;
        mbar_wait_or_trap(&acc_empty[set], ((uint32_t)(j >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t tmem_hi = tmem_base + (uint32_t)set * 2 * kTcBN, tmem_lo = tmem_hi + kTcBN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % kTcStages;
          const uint32_t ph = (uint32_t)(it / kTcStages) & 1u;
          mbar_wait_or_trap(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t st = smem_u32(smem + (size_t)s * kTcStageBytes);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t a_hi = make_desc(st + ks * a_step, a_lbo, a_sbo, a_lt);
            const uint64_t a_lo = make_desc(st + kTcTileBytes + ks * a_step, a_lbo, a_sbo, a_lt);
            const uint64_t b_hi = make_desc(st + 2 * kTcTileBytes + ks * b_step, b_lbo, b_sbo, b_lt);
            const uint64_t b_lo = make_desc(st + 3 * kTcTileBytes + ks * b_step, b_lbo, b_sbo, b_lt);
            const uint32_t acc = (kb > 0 || ks > 0) ? 1u : 0u;
            if (!p.single) {
              E::mma(tmem_lo, a_lo, b_hi, idesc, acc);
              E::mma(tmem_lo, a_hi, b_lo, idesc, 1);
            }
            E::mma(tmem_hi, a_hi, b_hi, idesc, acc);
          }
          tc_release_stage(&empty_bar[s], cp);
        }
        umma_commit(&acc_full[set]);
      }
    }
  } else {
    // ================================ epilogue (warps 2..9) ================================
    // warp w reads TMEM lanes [32 (w % 4), +32) - two warps per quadrant, each draining half of the tile's columns
    const int lane_grp = warp & 3, chalf = (warp - 2) >> 2;
    const bool vec = p.c && ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.c) & 15) == 0);
    const int out_shift = F16 ? -(__ldg(p.exp_a) + __ldg(p.exp_b)) : 0;
    // wide outputs go through the per-warp transposition buffer (behind the barriers, 16-B aligned)
    float* stg = reinterpret_cast<float*>(smem + kTcStages * kTcStageBytes + 256) + (warp - 2) * 32 * 32;
    const bool coalesce = vec && !p.accumulate && !p.path && p.N >= 256 && (p.N & 3) == 0 &&
                          (!p.bias || (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0);
    int j = 0;
    for (int st = cluster_id; st < num_super; st += n_clusters, ++j) {
      const int set = j & 1;
      const int m0 = ((st / super_n) * p.cm + cp.rm) * kTcBM, n0 = ((st % super_n) * p.cn + cp.rn) * kTcBN;
      mbar_wait_or_trap(&acc_full[set], (uint32_t)(j >> 1) & 1u);
      tc_fence_after();
      const int m = m0 + lane_grp * 32 + lane;
      float* crow = p.c ? p.c + (size_t)m * p.ldc : nullptr;
      const uint32_t lane_addr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)set * 2 * kTcBN;
      float best_v = 0.f;  // arg-max of the row (decode epilogue): columns arrive in order, the first maximum wins
      int best_i = -1;
      // (arg-max epilogue: a row must stay in one thread - the first warp of the quadrant takes all columns)
      const int cb0 = p.path ? 0 : chalf * (kTcBN / 2), cb1 = p.path ? (chalf == 0 ? kTcBN : 0) : (chalf + 1) * (kTcBN / 2);
#pragma unroll 1
      for (int cb = cb0; cb < cb1; cb += 32) {
        if (n0 + cb >= p.N) break;  // warp-uniform
        uint32_t th[32], tl[32];
        tmem_ld32_nowait(lane_addr + (uint32_t)cb, th);
        tmem_ld32_nowait(lane_addr + (uint32_t)(kTcBN + cb), tl);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float r[32];
#pragma unroll
        for (int q = 0; q < 32; ++q) {
          const float v = p.single ? __uint_as_float(th[q])
                                   : fmaf(__uint_as_float(tl[q]), F16 ? 1.f / kPairLoScale : 1.f, __uint_as_float(th[q]));
          r[q] = F16 ? scale_pow2(v, out_shift) : v;
        }
        if (coalesce) {
          // An epilogue thread owns a ROW of the tile: storing it directly makes every warp store touch 32 rows x 16 B
          // (half-filled sectors, 16 KB apart at N = 4096) - measured 0.6 TB/s, which bounds the K = 128 projection
          // outright (1.22 ms for an 814-MB output).  Transpose the 32 x 32 block through shared memory instead so that
          // a store instruction writes four full 128-byte lines.
#pragma unroll
          for (int q0 = 0; q0 < 32; q0 += 4) {
            const int n = n0 + cb + q0;
            float4 v = make_float4(r[q0], r[q0 + 1], r[q0 + 2], r[q0 + 3]);
            if (p.bias && n + 3 < p.N) {  // (N % 4 == 0 here: ldc % 4 == 0 and the tail is masked on the way out)
              const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + n));
              v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
            }
            if (p.relu) {
              v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
            }
            *reinterpret_cast<float4*>(stg + lane * 32 + (((q0 >> 2) ^ (lane & 7)) << 2)) = v;
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = 4 * i + (lane >> 3), g = lane & 7;
            const float4 v = *reinterpret_cast<const float4*>(stg + row * 32 + ((g ^ (row & 7)) << 2));
            const int mr = m0 + lane_grp * 32 + row, n = n0 + cb + 4 * g;
            if (mr < p.M && n + 3 < p.N) *reinterpret_cast<float4*>(p.c + (size_t)mr * p.ldc + n) = v;
          }
          __syncwarp();  // the block is read before the next one overwrites the buffer
        } else if (m < p.M) {
#pragma unroll
          for (int q0 = 0; q0 < 32; q0 += 4) {
            const int n = n0 + cb + q0;
            if (n >= p.N) break;
            float v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float x = r[q0 + q];
              if (n + q < p.N) {
                if (p.bias) x += __ldg(p.bias + n + q);
                if (p.accumulate) x += crow[n + q];
                if (p.relu) x = fmaxf(x, 0.f);
                if (p.path && (best_i < 0 || argmax_gt(x, best_v))) {
                  best_v = x;
                  best_i = n + q;
                }
              }
              v[q] = x;
            }
            if (!crow) continue;
            if (vec && n + 3 < p.N) {
              *reinterpret_cast<float4*>(crow + n) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
              for (int q = 0; q < 4; ++q)
                if (n + q < p.N) crow[n + q] = v[q];
            }
          }
        }
      }
      if (p.path && chalf == 0 && m < p.M) {  // frame label as in decode.cu: -1 beyond the line, 0 = blank or below the threshold
        const int t = m / p.dec_B, b = m - t * p.dec_B;
        int label = -1;
        if (t < __ldg(p.lens + b)) label = (best_i == 0 || best_v < p.thresh) ? 0 : best_i;
        p.path[(size_t)b * p.dec_T + t] = label;
      }
      tc_fence_before();
      mbar_arrive_cta(&acc_empty[set]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (p.cm * p.cn > 1) tc_cluster_sync();  // no CTA leaves while a peer's commit could still arrive here
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}


// Accumulation scheme (all K): the tensor core adds into its fp32 accumulator with round-toward-zero, once per
// instruction, by up to an ulp of the RUNNING SUM - a bias that grows linearly with the number of accumulations.
// So (1) the two lo products go to their own accumulator (its running sum is 2^-11 smaller), and (2) hi*hi is cut
// into chunks of kTcChunk k-blocks (32 accumulations) that rotate over 3 TMEM accumulators; the epilogue warps drain a
// finished chunk into fp32 registers with round-to-nearest adds while the next chunk runs on another accumulator.
template <bool F16>
__global__ void __launch_bounds__(kTcThreads, 1)
tc_gemm_x3_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                      const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                      TcGemmParams p) {
  extern __shared__ unsigned char tc_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) &
                                                         ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kTcStages * kTcStageBytes);
  uint64_t* full_bar = bars;                     // [stages]
  uint64_t* empty_bar = bars + kTcStages;        // [stages]
  uint64_t* acc_full = bars + 2 * kTcStages;     // [3]
  uint64_t* acc_empty = acc_full + kTcHiAcc;     // [3]
  uint64_t* lo_full = acc_empty + kTcHiAcc;      // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(lo_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const TcClusterPos cp = tc_cluster_pos(p);  // grid = (n tiles, m tiles) padded to the cluster shape (cn, cm)
  const int m0 = ((int)(blockIdx.y / p.cm) * p.cm + cp.rm) * kTcBM, n0 = ((int)(blockIdx.x / p.cn) * p.cn + cp.rn) * kTcBN;
  using E = TcElem<F16>;
  constexpr int BK = E::kBK;
  const int total_kb = (p.K + BK - 1) / BK;
  const int kb_begin = blockIdx.z * p.kb_per_split;
  const int num_kb = max(0, min(total_kb, kb_begin + p.kb_per_split) - kb_begin);
  // short reductions (<= 96 k-blocks): one chunk per accumulator, nothing is drained before the end
  const int chunk_len = (num_kb <= 96) ? max(1, (num_kb + kTcHiAcc - 1) / kTcHiAcc) : kTcChunk;
  const int num_chunks = (num_kb + chunk_len - 1) / chunk_len;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kTcStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], (uint32_t)(p.cm + p.cn - 1));
    }
    for (int a = 0; a < kTcHiAcc; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], 128);
    }
    mbar_init(lo_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  if (p.cm * p.cn > 1) tc_cluster_sync();  // every CTA's barriers exist before a peer can signal them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % kTcStages;
        const uint32_t ph = (uint32_t)(i / kTcStages) & 1u;
        mbar_wait_or_trap(&empty_bar[s], ph ^ 1u);
        tc_load_stage<F16>(smem + (size_t)s * kTcStageBytes, &full_bar[s], &map_a_hi, &map_a_lo, &map_b_hi, &map_b_lo, p,
                           cp, (kb_begin + i) * BK, m0, n0);
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      // instruction descriptor: D = F32 (1 << 4), A/B = TF32 (2 << 7, 2 << 10), majors, N >> 3 at bit 17, M >> 4 at 24
      const uint32_t idesc = (1u << 4) | (E::kFmt << 7) | (E::kFmt << 10) | ((uint32_t)p.a_mn << 15) |
                             ((uint32_t)p.b_mn << 16) | ((uint32_t)(kTcBN >> 3) << 17) | ((uint32_t)(kTcBM >> 4) << 24);
      // per k-step (32 bytes of K): K-major advances 32 B inside the swizzle row, MN-major advances its k rows
      const uint32_t a_step = p.a_mn ? E::kMnStep : 32u, b_step = p.b_mn ? E::kMnStep : 32u;
      const uint32_t a_lbo = p.a_mn ? E::kMnBoxBytes : 16u, b_lbo = p.b_mn ? E::kMnBoxBytes : 16u;
      const uint32_t a_sbo = p.a_mn ? E::kMnSbo : 1024u, b_sbo = p.b_mn ? E::kMnSbo : 1024u;
      const uint32_t a_lt = p.a_mn ? E::kMnLayout : 2u, b_lt = p.b_mn ? E::kMnLayout : 2u;
      const uint32_t tmem_lo = tmem_base + kTcHiAcc * kTcBN;
      uint32_t accum_lo = 0;
      for (int c = 0; c < num_chunks; ++c) {
        const int a = c % kTcHiAcc;
        mbar_wait_or_trap(&acc_empty[a], ((uint32_t)(c / kTcHiAcc) & 1u) ^ 1u);
        tc_fence_after();
        const int i1 = min(num_kb, (c + 1) * chunk_len);
        for (int i = c * chunk_len; i < i1; ++i) {
          const int s = i % kTcStages;
          const uint32_t ph = (uint32_t)(i / kTcStages) & 1u;
          mbar_wait_or_trap(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t st = smem_u32(smem + (size_t)s * kTcStageBytes);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t a_hi = make_desc(st + ks * a_step, a_lbo, a_sbo, a_lt);
            const uint64_t a_lo = make_desc(st + kTcTileBytes + ks * a_step, a_lbo, a_sbo, a_lt);
            const uint64_t b_hi = make_desc(st + 2 * kTcTileBytes + ks * b_step, b_lbo, b_sbo, b_lt);
            const uint64_t b_lo = make_desc(st + 3 * kTcTileBytes + ks * b_step, b_lbo, b_sbo, b_lt);
            if (!p.single) {
              E::mma(tmem_lo, a_lo, b_hi, idesc, accum_lo);
              accum_lo = 1;
              E::mma(tmem_lo, a_hi, b_lo, idesc, 1);
            }
            E::mma(tmem_base + (uint32_t)a * kTcBN, a_hi, b_hi, idesc, (i > c * chunk_len || ks > 0) ? 1u : 0u);
          }
          tc_release_stage(&empty_bar[s], cp);  // frees the stage once the MMAs above have read it
        }
        umma_commit(&acc_full[a]);
      }
      umma_commit(lo_full);
    }
  } else {
    // ================================ epilogue (warps 2..5) ================================
    const int lane_grp = warp & 3;  // TMEM lanes [32*lane_grp, +32) are the ones this warp may read
    const uint32_t lane_addr = tmem_base + ((uint32_t)(lane_grp * 32) << 16);
    float sum[kTcBN];
#pragma unroll
    for (int j = 0; j < kTcBN; ++j) sum[j] = 0.f;
    for (int c = 0; c < num_chunks; ++c) {
      const int a = c % kTcHiAcc;
      mbar_wait_or_trap(&acc_full[a], (uint32_t)(c / kTcHiAcc) & 1u);
      tc_fence_after();
#pragma unroll
      for (int cb = 0; cb < kTcBN; cb += 32) {
        uint32_t t[32];
        tmem_ld32(lane_addr + (uint32_t)(a * kTcBN + cb), t);
#pragma unroll
        for (int j = 0; j < 32; ++j) sum[cb + j] += __uint_as_float(t[j]);
      }
      tc_fence_before();
      mbar_arrive_cta(&acc_empty[a]);
    }
    if (num_kb > 0 && !p.single) {
      mbar_wait_or_trap(lo_full, 0);
      tc_fence_after();
#pragma unroll
      for (int cb = 0; cb < kTcBN; cb += 32) {
        uint32_t t[32];
        tmem_ld32(lane_addr + (uint32_t)(kTcHiAcc * kTcBN + cb), t);
#pragma unroll
        for (int j = 0; j < 32; ++j)
          sum[cb + j] = fmaf(__uint_as_float(t[j]), F16 ? 1.f / kPairLoScale : 1.f, sum[cb + j]);
      }
    }
    if (F16) {
      const int out_shift = -(__ldg(p.exp_a) + __ldg(p.exp_b));
#pragma unroll
      for (int j = 0; j < kTcBN; ++j) sum[j] = scale_pow2(sum[j], out_shift);
    }
    const int m = m0 + lane_grp * 32 + lane;
    if (m < p.M) {
      const bool partial = gridDim.z > 1;  // split-K: raw partial sums, the reduce kernel applies the epilogue
      float* crow = partial ? p.c + ((size_t)blockIdx.z * p.M + m) * p.N : p.c + (size_t)m * p.ldc;
      const int ld_eff = partial ? p.N : p.ldc;
      const bool vec = ((ld_eff & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.c) & 15) == 0);
#pragma unroll
      for (int j = 0; j < kTcBN; j += 4) {
        const int n = n0 + j;
        if (n < p.N) {
          float v[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float x = sum[j + q];
            if (!partial && n + q < p.N) {
              if (p.bias) x += __ldg(p.bias + n + q);
              if (p.accumulate) x += crow[n + q];
              if (p.relu) x = fmaxf(x, 0.f);
            }
            v[q] = x;
          }
          if (vec && n + 3 < p.N) {
            *reinterpret_cast<float4*>(crow + n) = make_float4(v[0], v[1], v[2], v[3]);
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (n + q < p.N) crow[n + q] = v[q];
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (p.cm * p.cn > 1) tc_cluster_sync();  // no CTA leaves while a peer's commit could still arrive here
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// split-K epilogue: C = sum_z partial[z] (+bias) (+C) (relu)
__global__ void __launch_bounds__(256)
tc_gemm_reduce_kernel(const float* __restrict__ partial, int splits, int M, int N, float* __restrict__ c, int ldc,
                      const float* __restrict__ bias, int relu, int accumulate) {
  const long long total = (long long)M * N;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(idx / N), n = (int)(idx - (long long)m * N);
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += partial[(size_t)z * total + idx];
    if (bias) s += __ldg(bias + n);
    float* q = c + (size_t)m * ldc + n;
    if (accumulate) s += *q;
    if (relu) s = fmaxf(s, 0.f);
    *q = s;
  }
}

__device__ __forceinline__ float rna_tf32(float x) {
  uint32_t t;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(x));
  return __uint_as_float(t);
}
// x -> hi = rna_tf32(x), lo = rna_tf32(x - hi): both exactly representable in TF32, so the tensor core (which
// truncates its inputs) reads them unchanged and the only input error left is the unbiased 2^-23 |x| of rounding lo
__global__ void __launch_bounds__(256)
split_tf32_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, long long n4,
                  long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    float4 h, l;
    h.x = rna_tf32(v.x); l.x = rna_tf32(v.x - h.x);
    h.y = rna_tf32(v.y); l.y = rna_tf32(v.y - h.y);
    h.z = rna_tf32(v.z); l.z = rna_tf32(v.z - h.z);
    h.w = rna_tf32(v.w); l.w = rna_tf32(v.w - h.w);
    reinterpret_cast<float4*>(hi)[i] = h;
    reinterpret_cast<float4*>(lo)[i] = l;
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float h = rna_tf32(x[i]);
    hi[i] = h;
    lo[i] = rna_tf32(x[i] - h);
  }
}


// ---- FP16 pair planes ------------------------------------------------------------------------------------------------
// x * 2^e -> hi = fp16(x 2^e), lo = fp16((x 2^e - hi) * 2^11): 22 significant bits where hi is a normal FP16 number
// and an absolute error floor of bound * 2^-50 below that (lo keeps 11 bits down to 2^-25 after scaling), i.e. a
// 22-bit format with ~40 binades of dynamic range under the tensor's bound.  e puts the bound in [2^14, 2^15).
__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ x, long long n4, long long n,
                                                     unsigned* __restrict__ out_bits) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  float m = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    m = fmaxf(m, fabsf(x[i]));
  // fmaxf drops NaNs; that is fine for a scale (a NaN element stays NaN in the planes)
  unsigned bits = __float_as_uint(m);
  bits = __reduce_max_sync(0xffffffffu, bits);  // non-negative floats order like their bit patterns
  if ((threadIdx.x & 31) == 0 && bits) atomicMax(out_bits, bits);
}

__global__ void __launch_bounds__(256)
split_f16_kernel(const float* __restrict__ x, __half* __restrict__ hi, __half* __restrict__ lo, long long n4,
                 long long n, const unsigned* __restrict__ bound_bits, int* __restrict__ exp_out) {
  const int e = pair_exponent(__ldg(bound_bits));
  if (blockIdx.x == 0 && threadIdx.x == 0) *exp_out = e;
  const float sc = exp2i(e);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    uint2 ph, pl;
    pair_pack4(v.x * sc, v.y * sc, v.z * sc, v.w * sc, ph, pl);
    reinterpret_cast<uint2*>(hi)[i] = ph;
    reinterpret_cast<uint2*>(lo)[i] = pl;
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float a = x[i] * sc;
    const __half h = __float2half_rn(a);
    hi[i] = h;
    lo[i] = __float2half_rn((a - __half2float(h)) * kPairLoScale);
  }
}


// 3x3 / pad 1 patches of a narrow NHWC activation (C < 64, e.g. the 16 channels behind the rapid-downsample block) as
// FP16 pair planes cols[p][tap*C + c] = x[p + tap][c] * 2^e (zero outside the image).  With so few channels the
// implicit-GEMM conv kernels would move 128-byte swizzle rows that are mostly padding; a K = 9C GEMM over these planes
// (forward: cols W^T, weight gradient: cols^T dz) runs on the kind::f16 GEMM kernels instead.
template <typename IDX>  // unsigned when the plane has < 2^31 4-element groups (64-bit divisions made it instruction-bound)
__global__ void __launch_bounds__(256)
im2col3x3_f16_kernel(const float* __restrict__ x, int B, int H, int W, int C4, __half* __restrict__ hi,
                     __half* __restrict__ lo, const unsigned* __restrict__ bound_bits, int* __restrict__ exp_out) {
  const int e = pair_exponent(__ldg(bound_bits));
  if (blockIdx.x == 0 && threadIdx.x == 0) *exp_out = e;
  const float sc = exp2i(e);
  const long long total = (long long)B * H * W * 9 * C4;
  for (IDX idx = (IDX)blockIdx.x * blockDim.x + threadIdx.x; idx < (IDX)total; idx += (IDX)gridDim.x * blockDim.x) {
    const IDX r9 = idx / (IDX)C4;
    const int c4 = (int)(idx - r9 * (IDX)C4);
    const IDX r = r9 / 9;  // pixel
    const int tap = (int)(r9 - r * 9);
    const IDX row = r / (IDX)W;
    const int xx = (int)(r - row * (IDX)W);
    const long long b = (long long)(row / (IDX)H);
    const int yy = (int)(row - (IDX)b * (IDX)H);
    const int sy = yy + tap / 3 - 1, sx = xx + tap % 3 - 1;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (sy >= 0 && sy < H && sx >= 0 && sx < W)
      v = __ldg(reinterpret_cast<const float4*>(x) + ((b * H + sy) * W + sx) * C4 + c4);
    uint2 ph, pl;
    pair_pack4(v.x * sc, v.y * sc, v.z * sc, v.w * sc, ph, pl);
    reinterpret_cast<uint2*>(hi)[idx] = ph;  // idx = (pixel * 9 + tap) * C4 + c4: the [P][9C] row-major plane
    reinterpret_cast<uint2*>(lo)[idx] = pl;
  }
}

}  // namespace vocr

using namespace vocr;

extern "C" int vocr_split_tf32_f32(const float* x, float* hi, float* lo, long long n, vocr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VOCR_REQUIRE(n >= 0);
  if (n == 0) return VOCR_OK;
  VOCR_REQUIRE(x && hi && lo);
  const bool al = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(hi) | reinterpret_cast<uintptr_t>(lo)) & 15) == 0;
  const long long n4 = al ? n / 4 : 0;
  const int grid = (int)min((long long)kNumSMs * 8, ceil_div64(max(1ll, n / 4), 256));
  split_tf32_kernel<<<grid, 256, 0, stream>>>(x, hi, lo, n4, n);
  VOCR_CHECK_LAUNCH();
  return VOCR_OK;
}

// x[n] -> FP16 pair planes of x * 2^e.  state (device, 2 x int32): [0] receives e, [1] is scratch for the absmax pass.
// bound (device float, optional): any upper bound of max|x| known to the caller (skips the absmax pass).
extern "C" int vocr_split_f16_f32(const float* x, long long n, const float* bound, int32_t* state, uint16_t* hi,
                                  uint16_t* lo, vocr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VOCR_REQUIRE(n >= 0 && state);
  const unsigned* bits = reinterpret_cast<const unsigned*>(bound);
  const bool al = ((reinterpret_cast<uintptr_t>(x) & 15) | (reinterpret_cast<uintptr_t>(hi) & 7) |
                   (reinterpret_cast<uintptr_t>(lo) & 7)) == 0;
  const long long n4 = al ? n / 4 : 0;
  const int grid = (int)min((long long)kNumSMs * 8, ceil_div64(max(1ll, n / 4), 256));
  if (!bound) {
    if (cudaMemsetAsync(state + 1, 0, sizeof(int32_t), stream) != cudaSuccess) return VOCR_MEMOPS_FAILED;
    if (n > 0) {
      VOCR_REQUIRE(x);
      absmax_kernel<<<grid, 256, 0, stream>>>(x, n4, n, reinterpret_cast<unsigned*>(state + 1));
      VOCR_CHECK_LAUNCH();
    }
    bits = reinterpret_cast<const unsigned*>(state + 1);
  }
  VOCR_REQUIRE(n == 0 || (x && hi && lo));
  split_f16_kernel<<<max(1, grid), 256, 0, stream>>>(x, reinterpret_cast<__half*>(hi), reinterpret_cast<__half*>(lo), n4,
                                                     n, bits, state);
  VOCR_CHECK_LAUNCH();
  return VOCR_OK;
}

// x [B,H,W,C] (C % 4 == 0) -> FP16 pair planes cols [B*H*W][9*C] of the 3x3 / pad 1 patches, cols[p][tap*C + c].
// state / bound as in vocr_split_f16_f32 (the planes share x's scale).
extern "C" int vocr_im2col3x3_f16(const float* x, int B, int H, int W, int C, const float* bound, int32_t* state,
                                  uint16_t* hi, uint16_t* lo, vocr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VOCR_REQUIRE(B >= 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0 && state);
  const long long n = (long long)B * H * W * C;
  const unsigned* bits = reinterpret_cast<const unsigned*>(bound);
  if (!bound) {
    if (cudaMemsetAsync(state + 1, 0, sizeof(int32_t), stream) != cudaSuccess) return VOCR_MEMOPS_FAILED;
    if (n > 0) {
      VOCR_REQUIRE(x && (reinterpret_cast<uintptr_t>(x) & 15) == 0);
      const int grid = (int)min((long long)kNumSMs * 8, ceil_div64(n / 4, 256));
      absmax_kernel<<<grid, 256, 0, stream>>>(x, n / 4, n, reinterpret_cast<unsigned*>(state + 1));
      VOCR_CHECK_LAUNCH();
    }
    bits = reinterpret_cast<const unsigned*>(state + 1);
  }
  if (n == 0) return VOCR_OK;
  VOCR_REQUIRE(x && hi && lo && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
               ((reinterpret_cast<uintptr_t>(hi) | reinterpret_cast<uintptr_t>(lo)) & 7) == 0);
  const long long total = n * 9 / 4;
  const int grid = (int)min((long long)kNumSMs * 16, ceil_div64(total, 256));
  if (total + (long long)grid * 256 < (1ll << 31))
    im2col3x3_f16_kernel<unsigned><<<grid, 256, 0, stream>>>(x, B, H, W, C / 4, reinterpret_cast<__half*>(hi),
                                                             reinterpret_cast<__half*>(lo), bits, state);
  else
    im2col3x3_f16_kernel<long long><<<grid, 256, 0, stream>>>(x, B, H, W, C / 4, reinterpret_cast<__half*>(hi),
                                                              reinterpret_cast<__half*>(lo), bits, state);
  VOCR_CHECK_LAUNCH();
  return VOCR_OK;
}

// C[M,N] = op(A) op(B) (+bias) (+C) (relu), operands pre-split into (hi, lo) planes with identical layout.
//   a_mn = 0: A planes are [M,K] row-major (lda)      a_mn = 1: A planes are [K,M] row-major (lda)
//   b_mn = 0: B planes are [N,K] row-major (ldb)      b_mn = 1: B planes are [K,N] row-major (ldb)
// TF32 planes: lda, ldb multiples of 4; FP16 pair planes: multiples of 8; plane bases 16-B aligned.
// workspace (optional, 16-B aligned): enables split-K for long reductions with few output tiles; any size works (the
// split count adapts), M*N*16 floats is ample.
template <bool F16>
static int tc_gemm_launch(int a_mn, int b_mn, int M, int N, int K, const void* a_hi, const void* a_lo, int lda,
                          const int* exp_a, const void* b_hi, const void* b_lo, int ldb, const int* exp_b, float* C,
                          int ldc, const float* bias, int relu, int accumulate, void* workspace,
                          size_t workspace_bytes, int products, cudaStream_t stream,
                          const TcGemmParams* decode = nullptr) {
  VOCR_REQUIRE(products == 0 || products == 1 || products == 3);
  using E = TcElem<F16>;
  constexpr int BK = E::kBK, ALIGN = F16 ? 8 : 4;
  VOCR_REQUIRE(M >= 0 && N >= 0 && K >= 1);
  if (M == 0 || N == 0) return VOCR_OK;
  VOCR_REQUIRE(a_hi && a_lo && b_hi && b_lo && (C || decode) && (!F16 || (exp_a && exp_b)));
  VOCR_REQUIRE(lda % ALIGN == 0 && ldb % ALIGN == 0);
  VOCR_REQUIRE(((reinterpret_cast<uintptr_t>(a_hi) | reinterpret_cast<uintptr_t>(a_lo) |
                 reinterpret_cast<uintptr_t>(b_hi) | reinterpret_cast<uintptr_t>(b_lo)) & 15) == 0);
  const int tiles_m = ceil_div(M, kTcBM), tiles_n = ceil_div(N, kTcBN);
  const int tiles = tiles_n * tiles_m;
  const int total_kb = ceil_div(K, BK);
  const bool single = F16 && resolve_tc_products(products) == 1;
  // split-K for long reductions that would otherwise leave most SMs idle (weight-gradient GEMMs)
  int splits = 1;
  if (workspace && tiles * 2 <= kNumSMs && total_kb >= 32) {
    splits = min(min(64, (2 * kNumSMs) / tiles), total_kb / 16);
    while (splits > 1 && sizeof(float) * (size_t)M * N * splits > workspace_bytes) --splits;
    if (splits < 1) splits = 1;
  }
  int kb_per_split = ceil_div(total_kb, splits);
  splits = ceil_div(total_kb, kb_per_split);
  const int kind = (splits == 1 && (total_kb <= kTcPersistMaxKb || single)) ? 0 : (splits == 1 && total_kb <= 96) ? 1 : 2;
  // Cluster shape: 2 x 2 output tiles (2 x 1 / 1 x 2 for one-tile-wide problems) for the long reductions of the weight
  // gradients (kind 2; measured on cfg2 shapes: dW_ih 531 -> 438 us, dW_hh 161 -> 138 us).  The short-K kernels already
  // run at ~80 % of the sustained tensor rate on their own and lose 5 - 10 % to the lock step of a cluster.
  const bool clusters = tc_clusters_enabled() && kind == 2 && !single;
  const int cm = (clusters && tiles_m >= 2) ? 2 : 1, cn = (clusters && tiles_n >= 2) ? 2 : 1;
  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  auto map2d = [](CUtensorMap* m, const void* base, long long rows, long long cols, long long ld, int bc, int br,
                  bool mn) {
    if (F16) return make_map_2d_f16(m, base, rows, cols, ld, bc, br);
    return make_map_2d(m, static_cast<const float*>(base), rows, cols, ld, bc, br, mn);
  };
  bool ok = true;
  if (!a_mn)  // K-major: a CTA fetches 1 / cn of the tile's rows
    ok = ok && map2d(&ma_hi, a_hi, M, K, lda, BK, kTcBM / cn, false) && map2d(&ma_lo, a_lo, M, K, lda, BK, kTcBM / cn, false);
  else
    ok = ok && map2d(&ma_hi, a_hi, K, M, lda, E::kMnBox, BK, true) && map2d(&ma_lo, a_lo, K, M, lda, E::kMnBox, BK, true);
  if (!b_mn)
    ok = ok && map2d(&mb_hi, b_hi, N, K, ldb, BK, kTcBN / cm, false) && map2d(&mb_lo, b_lo, N, K, ldb, BK, kTcBN / cm, false);
  else
    ok = ok && map2d(&mb_hi, b_hi, K, N, ldb, E::kMnBox, BK, true) && map2d(&mb_lo, b_lo, K, N, ldb, E::kMnBox, BK, true);
  if (!ok) return VOCR_EXECUTION_FAILED;
  static DeviceLatch attr_latch;
  if (attr_latch.need()) {
    if (cudaFuncSetAttribute(tc_gemm_x3_kernel<F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes) !=
            cudaSuccess ||
        cudaFuncSetAttribute(tc_gemm_x3_shortk_kernel<F16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             kTcSmemBytes) != cudaSuccess ||
        cudaFuncSetAttribute(tc_gemm_x3_persist_kernel<F16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             kTcPersistSmemBytes) != cudaSuccess)
      return VOCR_EXECUTION_FAILED;
    attr_latch.set();
  }
  TcGemmParams p{splits > 1 ? static_cast<float*>(workspace) : C, bias, M, N, K, ldc, relu, accumulate, a_mn ? 1 : 0,
                 b_mn ? 1 : 0, kb_per_split, exp_a, exp_b, single ? 1 : 0, cm, cn, nullptr, nullptr, 0, 0, 0.f};
  if (decode) {  // arg-max epilogue: the persistent kernel with the whole row in one tile
    VOCR_REQUIRE(kind == 0 && tiles_n == 1 && decode->path && decode->lens && (long long)decode->dec_T * decode->dec_B == M);
    p.path = decode->path; p.lens = decode->lens; p.dec_T = decode->dec_T; p.dec_B = decode->dec_B; p.thresh = decode->thresh;
  }
  cudaError_t err = cudaSuccess;
  if (kind == 0) {  // persistent, one CTA per SM
    tc_gemm_x3_persist_kernel<F16><<<min(tiles, kNumSMs), kTcPersistThreads, kTcPersistSmemBytes, stream>>>(ma_hi, ma_lo, mb_hi, mb_lo, p,
                                                                                             tiles_n, tiles);
  } else if (cm * cn == 1) {
    dim3 grid(tiles_n, tiles_m, splits);
    if (kind == 1) tc_gemm_x3_shortk_kernel<F16><<<grid, kTcThreads, kTcSmemBytes, stream>>>(ma_hi, ma_lo, mb_hi, mb_lo, p);
    else tc_gemm_x3_kernel<F16><<<grid, kTcThreads, kTcSmemBytes, stream>>>(ma_hi, ma_lo, mb_hi, mb_lo, p);
  } else {
    // (a cluster launch also changes the CTA -> SM placement, so launches without a cluster keep the plain path)
    cudaLaunchConfig_t cfg{};
    cfg.blockDim = dim3(kTcThreads);
    cfg.dynamicSmemBytes = kTcSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim = {(unsigned)cn, (unsigned)cm, 1};
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cfg.gridDim = dim3(ceil_div(tiles_n, cn) * cn, ceil_div(tiles_m, cm) * cm, splits);
    err = cudaLaunchKernelEx(&cfg, tc_gemm_x3_kernel<F16>, ma_hi, ma_lo, mb_hi, mb_lo, p);
  }
  if (err != cudaSuccess) return VOCR_EXECUTION_FAILED;
  VOCR_CHECK_LAUNCH();
  if (splits > 1) {
    const long long total = (long long)M * N;
    tc_gemm_reduce_kernel<<<(int)min((long long)4 * kNumSMs, ceil_div64(total, 256)), 256, 0, stream>>>(
        static_cast<const float*>(workspace), splits, M, N, C, ldc, bias, relu, accumulate);
  }
  VOCR_CHECK_LAUNCH();
  return VOCR_OK;
}

extern "C" int vocr_tc_gemm_tf32x3(int a_mn, int b_mn, int M, int N, int K, const float* a_hi, const float* a_lo,
                                   int lda, const float* b_hi, const float* b_lo, int ldb, float* C, int ldc,
                                   const float* bias, int relu, int accumulate, void* workspace,
                                   size_t workspace_bytes, vocr_stream_t stream_) {
  return tc_gemm_launch<false>(a_mn, b_mn, M, N, K, a_hi, a_lo, lda, nullptr, b_hi, b_lo, ldb, nullptr, C, ldc, bias,
                               relu, accumulate, workspace, workspace_bytes, 3, static_cast<cudaStream_t>(stream_));
}

// Same GEMM on FP16 pair planes (vocr_split_f16_f32): planes hold op * 2^exp[0]; the epilogue undoes both scales.
extern "C" int vocr_tc_gemm_f16x3(int a_mn, int b_mn, int M, int N, int K, const uint16_t* a_hi,
                                  const uint16_t* a_lo, int lda, const int32_t* exp_a, const uint16_t* b_hi,
                                  const uint16_t* b_lo, int ldb, const int32_t* exp_b, float* C, int ldc,
                                  const float* bias, int relu, int accumulate, void* workspace,
                                  size_t workspace_bytes, int products, vocr_stream_t stream_) {
  return tc_gemm_launch<true>(a_mn, b_mn, M, N, K, a_hi, a_lo, lda, exp_a, b_hi, b_lo, ldb, exp_b, C, ldc, bias, relu,
                              accumulate, workspace, workspace_bytes, products, static_cast<cudaStream_t>(stream_));
}

// Prob layer + first half of the greedy decode in one kernel: logits[m, :] = A[m, :] . W^T + bias for row m = t * B + b
// (A [T*B, K] activations, W [N, K], both K-major FP16 pair planes; N <= 128, K <= 1536) and, in the epilogue,
// path[b * T + t] = the frame label of decoder.py:116-185 (arg-max with numpy's NaN ordering; 0 when the arg-max is the
// blank or its value is below thresh; -1 for t >= lens[b]).  C may be NULL: the logits are then never written.
extern "C" int vocr_tc_gemm_f16x3_argmax(int M, int N, int K, const uint16_t* a_hi, const uint16_t* a_lo, int lda,
                                         const int32_t* exp_a, const uint16_t* b_hi, const uint16_t* b_lo, int ldb,
                                         const int32_t* exp_b, float* C, int ldc, const float* bias,
                                         const int32_t* lens, int T, int B, float thresh, int32_t* path, int products,
                                         vocr_stream_t stream_) {
  TcGemmParams d{};
  d.path = path; d.lens = lens; d.dec_T = T; d.dec_B = B; d.thresh = thresh;
  VOCR_REQUIRE(N >= 1 && N <= kTcBN && T >= 0 && B >= 0);
  return tc_gemm_launch<true>(0, 0, M, N, K, a_hi, a_lo, lda, exp_a, b_hi, b_lo, ldb, exp_b, C, ldc, bias, 0, 0, nullptr, 0,
                              products, static_cast<cudaStream_t>(stream_), &d);
}
