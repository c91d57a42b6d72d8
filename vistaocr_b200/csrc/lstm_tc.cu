// BiLSTM forward recurrence as a CLUSTER-RESIDENT tcgen05 kernel, sm_100a.
// (reference cnnlstm.py:148-149,285-290: nn.LSTM on a packed sequence -> cuDNN RNN; the input projections of all
// timesteps are one tensor-core GEMM done beforehand, see lstm.cu / ops.py.)
//
// One thread-block cluster of 16 CTAs serves one (direction, group of 32 samples) for the whole sequence; nothing ever
// leaves the cluster except the layer's outputs:
//   * CTA s of the cluster owns 32 hidden units = 128 gate rows of W_hh, kept ON CHIP for all timesteps as FP16 pairs
//     (hi, lo * 2^11; 22 significant bits): the hi plane lives in TENSOR MEMORY (128 lanes x 256 columns) and feeds
//     tcgen05.mma as the A operand straight from TMEM, the lo plane lives in shared memory (128 KB, K-major SW128).
//   * per step   D[128 gate rows x 32 samples] = W_slice[128 x H] . h_{t-1}^T   with kind::f16, M = 128, N = 32, K = 16 per
//     instruction: three error-compensated products in TWO instructions per k-step - W_hi (from TMEM) times the hi and lo
//     planes of h laid side by side as one N = 64 operand, and W_lo (from shared memory) times the hi plane - 64
//     instructions per step issued by one thread.
//   * eight epilogue warps read the accumulators (tcgen05.ld), regroup the four gates of a unit with warp shuffles, add
//     the input projection (staged by a loader warp with cp.async one step ahead), apply the gates and write the CTA's
//     piece of h_t (32 units x 32 samples, FP16 pair, already in the operand layout) to a 4-KB staging area;
//   * ONE multicast bulk copy per CTA and step (cp.async.bulk ... .multicast::cluster) drops that piece into the operand
//     tile of all 16 CTAs; the copies complete on each destination's mbarrier, so a CTA starts the next products the
//     moment its 16 pieces have landed - no flags, no polling, no grid or cluster barrier.  Flow control is one more
//     tcgen05.commit per step, multicast to the whole cluster: a CTA overwrites the operand tiles only after every CTA's
//     products of the step have retired.
// The operand tile of h uses the K-major SWIZZLE_NONE canonical layout (8 x 16-byte core matrices) so that a CTA's piece
// is one contiguous 4-KB block of every destination tile.
// Ragged lengths use packed-sequence semantics by masking: sample b is active at step k iff k < lens[b]; the reverse
// direction visits t = lens[b]-1-k; outputs beyond lens[b] stay zero; finished samples keep publishing their last state.
#include <cuda_fp16.h>
#include <cstdlib>
#ifdef VOCR_LSTM_PROF
#include <cstdio>
#endif

#include "tc_common.cuh"

namespace vocr {

constexpr int kClM = 128;             // gate rows per CTA = UMMA M (32 unit slots x 4 gates)
constexpr int kClSize = 16;           // CTAs per cluster
constexpr int kClMaxKB = 8;           // k-blocks of 64 units (H <= 512)
constexpr int kClThreads = 320;       // 8 epilogue warps, 1 MMA warp, 1 input-projection loader warp
constexpr uint32_t kClWloKb = kClM * 128;                  // W_lo bytes per k-block: 16 KB
constexpr uint32_t kColA = 0, kColAlo = 256;               // TMEM columns: W_hi operand | (NS = 64) first half of W_lo
constexpr float kClLoScale = 2048.f;

// Geometry of one cluster work item = (direction, NS samples).  NS = 32: four items of a 64-line batch run on four clusters
// in parallel (training).  NS = 64 (batches >= 128): every product instruction streams the same 4 KB of W whatever its N,
// so twice the samples per instruction halve the per-sample cost; the 128-KB operand tile then only fits because the
// first half of W_lo moves into TMEM next to W_hi (TS-mode products for k < 256, SS-mode above).
template <int NS>
struct ClGeom {
  static constexpr uint32_t kPlane = (NS / 8) * 128;          // one plane of a k-chunk of 8 units: NS rows x 16 B
  static constexpr uint32_t kChunk = 2 * kPlane;              // hi plane | lo plane
  static constexpr uint32_t kTileBytes = 64 * kChunk;         // operand tile of h: 64 / 128 KB
  static constexpr int kWloSmemKB = (NS == 32) ? 8 : 4;       // k-blocks of W_lo kept in shared memory (the upper ones)
  static constexpr int kWloTmemKB = kClMaxKB - kWloSmemKB;    // ... and in TMEM (the lower ones)
  static constexpr int kXsLd = (NS == 32) ? 136 : 132;        // row stride (floats) of the staged input projections
  static constexpr uint32_t kXsBytes = NS * kXsLd * 4;
  static constexpr uint32_t kColD = (NS == 32) ? 256 : 384;   // accumulator: 2 NS columns (hi.hi | cross terms)
  static constexpr uint32_t kStageBytes = 4 * kChunk;         // a CTA's piece: up to 4 k-chunks
  static constexpr size_t kSmem = (size_t)kWloSmemKB * kClWloKb + kTileBytes + kXsBytes + 64 + 1024;
};

struct LstmClArgs {
  const float* xproj;   // [T,B,2,4H]
  const float* whh;     // [2,4H,H]
  const int32_t* lens;  // [B]
  float* out;           // [T,B,2H]  (pre-zeroed)
  float* gates;         // [T,B,2,4H] activated i,f,g,o (may be null)
  float* cst;           // [T,B,2,H]  (may be null)
  unsigned char* stage; // [clusters][16 CTAs][2 parities][piece] staging of the published pieces
  int T, B, H, KB, US, NSL, Tmax, NG, n_items;
};

// Gate nonlinearities on the serial chain of the recurrence: ex2.approx-based (absolute error ~2e-7 on values in
// [-1, 1] - below the fp32 rounding of the pre-activation sums they are applied to).
__device__ __forceinline__ float cl_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));  // one MUFU, <= 1 ulp
  return r;
}
__device__ __forceinline__ float cl_sigmoidf(float x) { return cl_rcp(1.f + __expf(-x)); }
__device__ __forceinline__ float cl_tanhf(float x) { return fmaf(-2.f, cl_rcp(1.f + __expf(2.f * x)), 1.f); }
__device__ __forceinline__ void cl_split_f16(float x, __half& hi, __half& lo) {
  hi = __float2half_rn(x);
  lo = __float2half_rn((x - __half2float(hi)) * kClLoScale);
}
// byte offset of element (row r, k < 64) inside a K-major SWIZZLE_128B tile (rows of 128 B, 8-row atoms of 1024 B)
__device__ __forceinline__ uint32_t cl_sw128(int r, int k) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + (((k >> 3) ^ (r & 7)) << 4) + ((k & 7) << 1));
}
__device__ __forceinline__ void cl_umma_ts(uint32_t tmem_c, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_c),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives (when all prior tcgen05 operations of this thread have completed) on the mbarrier at the same shared-memory
// offset in every CTA of `mask`
__device__ __forceinline__ void cl_commit_multicast(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}
__device__ __forceinline__ void cl_tmem_ld16(uint32_t taddr, uint32_t (&t)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]), "=r"(t[8]),
        "=r"(t[9]), "=r"(t[10]), "=r"(t[11]), "=r"(t[12]), "=r"(t[13]), "=r"(t[14]), "=r"(t[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void cl_tmem_st32(uint32_t taddr, const uint32_t (&t)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(t[0]), "r"(t[1]), "r"(t[2]), "r"(t[3]), "r"(t[4]), "r"(t[5]), "r"(t[6]), "r"(t[7]), "r"(t[8]), "r"(t[9]),
      "r"(t[10]), "r"(t[11]), "r"(t[12]), "r"(t[13]), "r"(t[14]), "r"(t[15]), "r"(t[16]), "r"(t[17]), "r"(t[18]),
      "r"(t[19]), "r"(t[20]), "r"(t[21]), "r"(t[22]), "r"(t[23]), "r"(t[24]), "r"(t[25]), "r"(t[26]), "r"(t[27]),
      "r"(t[28]), "r"(t[29]), "r"(t[30]), "r"(t[31])
      : "memory");
}
__device__ __forceinline__ uint32_t cl_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cl_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// the 256 epilogue threads only
__device__ __forceinline__ void cl_epi_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

template <int NS>
__global__ void __launch_bounds__(kClThreads, 1) bilstm_fwd_cluster_kernel(LstmClArgs a) {
  using G = ClGeom<NS>;
  constexpr int NP = NS / 32;  // epilogue passes: a warp handles 16 samples per pass
  extern __shared__ unsigned char cl_smem_raw[];
  unsigned char* smem =
      reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(cl_smem_raw) + 1023) & ~uintptr_t(1023));
  const int KB = a.KB, H = a.H, US = a.US;
  unsigned char* Wlo = smem;                                         // [kWloSmemKB][128 rows x 128 B] SW128 K-major
  unsigned char* Ht = Wlo + (size_t)G::kWloSmemKB * kClWloKb;        // [64 k-chunks][hi plane | lo plane] SWIZZLE_NONE
  float* xs = reinterpret_cast<float*>(Ht + G::kTileBytes);          // [NS samples][kXsLd]: gate g, unit u at g*32 + u
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(xs) + G::kXsBytes);
  uint64_t* full = bars;           // h tile of the step landed (1 arrival + NSL pieces of tx bytes)
  uint64_t* mma_done = bars + 1;   // this CTA's products of the step retired (tcgen05.commit)
  uint64_t* tile_free = bars + 2;  // EVERY CTA's products of the step retired (NSL multicast commits)
  uint64_t* xs_full = bars + 3;    // input projections of the step staged (1 arrival)
  uint64_t* xs_free = bars + 4;    // ... and consumed (8 warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  const int slice = (int)cl_cluster_rank();
  const int cluster_id = blockIdx.x / kClSize, n_clusters = gridDim.x / kClSize;
  const int u0 = slice * US;
  const int nu = max(0, min(US, H - u0));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint16_t cta_mask = (uint16_t)((1u << a.NSL) - 1u);
  const uint32_t piece_bytes = (uint32_t)(US / 8) * G::kChunk;
  const bool active_cta = slice < a.NSL;

  if (tid == 0) {
    mbar_init(full, 1);
    mbar_init(mma_done, 1);
    mbar_init(tile_free, (uint32_t)a.NSL);
    mbar_init(xs_full, 1);
    mbar_init(xs_free, 8);
    mbar_fence_init();
  }
  if (warp == 8) tmem_alloc(tmem_slot, 512);
  // the operand tile starts as zeros: unit columns nobody publishes (H < 64 KB) and the state of step 0
  for (int i = tid; i < (int)(G::kTileBytes / 16); i += kClThreads) reinterpret_cast<uint4*>(Ht)[i] = make_uint4(0, 0, 0, 0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  cl_cluster_sync();  // every CTA's barriers are initialised before any peer can signal them
  if (!active_cta) {
    // H < 512: fewer than 16 slices.  Nobody addresses this CTA; it only keeps the cluster barriers balanced.
    for (int item = cluster_id; item + n_clusters < a.n_items; item += n_clusters) cl_cluster_sync();
    cl_cluster_sync();
    if (warp == 8) tmem_dealloc(tmem_base, 512);
    return;
  }

  unsigned n_full = 0, n_done = 0, n_free = 0, n_xsfull = 0, n_xsfree = 0;
  int loaded_dir = -1;
  unsigned char* stage = a.stage + ((size_t)(cluster_id * kClSize + slice) * 2) * G::kStageBytes;

  for (int item = cluster_id; item < a.n_items; item += n_clusters) {
    const int dir = item & 1, grp = item >> 1;
    const int b_base = grp * NS;
    int tm = 0;  // steps of this item: its longest sample
    for (int j = b_base; j < min(a.B, b_base + NS); ++j) tm = max(tm, min(a.lens[j], a.Tmax));

    if (loaded_dir != dir) {
      // ---- W slice: row m = 32q + 4 u8 + g  <->  gate g of unit slot U = 8q + u8 ------------------------------------
      const float* wd = a.whh + (size_t)dir * 4 * H * H;
      if (warp < 4) {  // TMEM planes: lane m, column k/2 (two halves per column): W_hi, and W_lo for the lower k-blocks
        const int g = lane & 3, U = 8 * warp + (lane >> 2);
        const float* wr = wd + ((size_t)g * H + u0 + U) * H;
        for (int plane = 0; plane < (G::kWloTmemKB > 0 ? 2 : 1); ++plane) {
          const int ncols = (plane == 0 ? KB : min(KB, G::kWloTmemKB)) * 32;
          for (int c0 = 0; c0 < ncols; c0 += 32) {
            uint32_t v[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const int k = 2 * (c0 + c);
              const float w0 = (U < nu && k < H) ? __ldg(wr + k) : 0.f, w1 = (U < nu && k + 1 < H) ? __ldg(wr + k + 1) : 0.f;
              __half h0, l0, h1, l1;
              cl_split_f16(w0, h0, l0);
              cl_split_f16(w1, h1, l1);
              v[c] = plane == 0 ? ((uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16))
                                : ((uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16));
            }
            cl_tmem_st32(tmem_base + ((uint32_t)(32 * warp) << 16) + (plane == 0 ? kColA : kColAlo) + c0, v);
          }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      const int k_lo0 = G::kWloTmemKB * 64;  // first k of W_lo that lives in shared memory
      for (int i = tid; i < kClM * (KB * 64 - min(KB * 64, k_lo0)); i += kClThreads) {
        const int kw = KB * 64 - k_lo0;
        const int m = i / kw, k = k_lo0 + (i - m * kw);
        const int q = m >> 5, g = m & 3, U = 8 * q + ((m & 31) >> 2);
        float v = 0.f;
        if (U < nu && k < H) v = __ldg(wd + ((size_t)g * H + u0 + U) * H + k);
        __half hi, lo;
        cl_split_f16(v, hi, lo);
        *reinterpret_cast<__half*>(Wlo + (size_t)((k >> 6) - G::kWloTmemKB) * kClWloKb + cl_sw128(m, k & 63)) = lo;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      loaded_dir = dir;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp < 8) {
      // ===================================== epilogue warps ======================================================
      const int q = warp & 3, hh = warp >> 2;       // TMEM quadrant, half of a 32-sample pass
      const int u8 = lane >> 2, g = lane & 3;
      const int U = 8 * q + u8;                     // unit slot of this thread
      const bool unit_ok = U < nu;
      int bs[NP][4], len[NP][4];
      float c_reg[NP][4], h_reg[NP][4];
#pragma unroll
      for (int ps = 0; ps < NP; ++ps)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int n = 32 * ps + 16 * hh + 4 * j + g;  // this thread's samples
          bs[ps][j] = b_base + n;
          len[ps][j] = (bs[ps][j] < a.B) ? min(a.lens[bs[ps][j]], a.Tmax) : 0;
          c_reg[ps][j] = h_reg[ps][j] = 0.f;
        }
      for (int k = 0; k < tm; ++k) {
        if (k > 0) {
          mbar_wait_or_trap(mma_done, n_done & 1u);
          ++n_done;
          tc_fence_after();
        }
        unsigned char* pc = stage + (size_t)(k & 1) * G::kStageBytes;
        float ig[NP][4], fg[NP][4], gv[NP][4], og[NP][4];
        bool act[NP][4];
#pragma unroll
        for (int ps = 0; ps < NP; ++ps) {
          const int sb = 32 * ps + 16 * hh;  // first of the 16 samples (accumulator columns) of this pass
          float pre[4][4];                   // [gate][j]
#pragma unroll
          for (int gg = 0; gg < 4; ++gg)
#pragma unroll
            for (int j = 0; j < 4; ++j) pre[gg][j] = 0.f;
          if (k > 0) {
            // D = W_hi . [h_hi | h_lo] (columns 0..NS-1 | NS..2NS-1); W_lo . h_hi is accumulated into the second half as
            // well (both cross terms carry the 2^-11 scale)
            const uint32_t tbase = tmem_base + ((uint32_t)(32 * q) << 16) + G::kColD + sb;
            uint32_t a_hh[16], a_x[16];
            float v[16];
            cl_tmem_ld16(tbase, a_hh);
            cl_tmem_ld16(tbase + NS, a_x);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int c = 0; c < 16; ++c) v[c] = fmaf(__uint_as_float(a_x[c]), 1.f / kClLoScale, __uint_as_float(a_hh[c]));
            // regroup: this lane holds gate g of unit U for 16 samples.  A 4 x 4 transpose over the four sibling lanes
            // (same u8; two xor-shuffle rounds, static register indices only) leaves it with all four gates of the
            // samples c = 4j + g.
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float x0 = v[4 * j], x1 = v[4 * j + 1], x2 = v[4 * j + 2], x3 = v[4 * j + 3];
              const bool o1 = (g & 1) != 0, o2 = (g & 2) != 0;
              float s0 = o1 ? x0 : x1, s1 = o1 ? x2 : x3;
              s0 = __shfl_xor_sync(0xffffffffu, s0, 1);
              s1 = __shfl_xor_sync(0xffffffffu, s1, 1);
              if (o1) { x0 = s0; x2 = s1; } else { x1 = s0; x3 = s1; }
              s0 = o2 ? x0 : x2;
              s1 = o2 ? x1 : x3;
              s0 = __shfl_xor_sync(0xffffffffu, s0, 2);
              s1 = __shfl_xor_sync(0xffffffffu, s1, 2);
              if (o2) { x0 = s0; x1 = s1; } else { x2 = s0; x3 = s1; }
              pre[0][j] = x0; pre[1][j] = x1; pre[2][j] = x2; pre[3][j] = x3;
            }
          }
          if (ps == 0) {  // input projections staged by the loader warp
            if (k > 0) tc_fence_before();
            mbar_wait_or_trap(xs_full, n_xsfull & 1u);
            ++n_xsfull;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            act[ps][j] = unit_ok && k < len[ps][j];
            if (act[ps][j]) {
              const float* xr = xs + (sb + 4 * j + g) * G::kXsLd + U;
#pragma unroll
              for (int gg = 0; gg < 4; ++gg) pre[gg][j] += xr[gg * 32];
            }
          }
          if (ps == NP - 1) {
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(xs_free)) : "memory");
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            ig[ps][j] = fg[ps][j] = gv[ps][j] = og[ps][j] = 0.f;
            if (act[ps][j]) {
              ig[ps][j] = cl_sigmoidf(pre[0][j]); fg[ps][j] = cl_sigmoidf(pre[1][j]);
              gv[ps][j] = cl_tanhf(pre[2][j]);    og[ps][j] = cl_sigmoidf(pre[3][j]);
              c_reg[ps][j] = fmaf(fg[ps][j], c_reg[ps][j], ig[ps][j] * gv[ps][j]);
              h_reg[ps][j] = og[ps][j] * cl_tanhf(c_reg[ps][j]);
            }
          }
          if (k + 1 < tm && U < US) {
            // publish h_k: this CTA's piece of the operand tile, laid out exactly like the tile (k-chunk q of the piece,
            // plane, row group n/8, row n%8, unit u8): finished / padding samples and missing units publish their state
            // (zeros), the consumers take the whole piece
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int n = sb + 4 * j + g;
              __half hi, lo;
              cl_split_f16(h_reg[ps][j], hi, lo);
              unsigned char* p = pc + (size_t)q * G::kChunk + (n >> 3) * 128 + (n & 7) * 16 + u8 * 2;
              *reinterpret_cast<__half*>(p) = hi;
              *reinterpret_cast<__half*>(p + G::kPlane) = lo;
            }
          }
        }
        if (k + 1 < tm) {
          cl_epi_sync();
          if (tid == 0) {
            // every CTA's products of step k have retired: the tiles may be overwritten (k = 0: nothing ran yet)
            if (k > 0) {
              mbar_wait_or_trap(tile_free, n_free & 1u);
              ++n_free;
            }
            asm volatile("fence.proxy.async.global;" ::: "memory");
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                ::"r"(smem_u32(Ht + (size_t)slice * piece_bytes)), "l"(pc), "r"(piece_bytes), "r"(smem_u32(full)), "h"(cta_mask)
                : "memory");
          }
        }
        // the layer's outputs (and what backward needs) are written after the exchange has been started: off the chain
#pragma unroll
        for (int ps = 0; ps < NP; ++ps)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (act[ps][j]) {
              const int tt = dir == 0 ? k : len[ps][j] - 1 - k;
              const size_t tb_ = (size_t)tt * a.B + bs[ps][j];
              a.out[(tb_ * 2 + dir) * H + u0 + U] = h_reg[ps][j];
              if (a.gates) {
                float* gp = a.gates + (tb_ * 2 + dir) * 4 * H + u0 + U;
                gp[0] = ig[ps][j];
                gp[(size_t)H] = fg[ps][j];
                gp[(size_t)2 * H] = gv[ps][j];
                gp[(size_t)3 * H] = og[ps][j];
              }
              if (a.cst) a.cst[(tb_ * 2 + dir) * H + u0 + U] = c_reg[ps][j];
            }
      }
      if (tid == 0 && tm > 1) {  // the last step's multicast commits (keeps the phase counter in step)
        mbar_wait_or_trap(tile_free, n_free & 1u);
        ++n_free;
      }
    } else if (warp == 8) {
      // ===================================== MMA issuer ==========================================================
      if (lane == 0) {
        // instruction descriptors: D = F32, A / B = F16 K-major, M = 128, N = 2 NS (hi | lo planes of h side by side) or NS
        const uint32_t idesc_w = (1u << 4) | ((uint32_t)(2 * NS >> 3) << 17) | ((uint32_t)(kClM >> 4) << 24);
        const uint32_t idesc_n = (1u << 4) | ((uint32_t)(NS >> 3) << 17) | ((uint32_t)(kClM >> 4) << 24);
        // descriptors advance by constants: W_lo 32 B per k-step inside a k-block of 16 KB, h two k-chunks per k-step; the
        // start-address field counts 16-byte units
        const uint64_t d_wlo0 = make_desc(smem_u32(Wlo), 16, 1024, 2);
        const uint64_t d_h0 = make_desc(smem_u32(Ht), G::kChunk, 128, 0);
        const uint32_t t_a = tmem_base + kColA, t_alo = tmem_base + kColAlo, t_d = tmem_base + G::kColD;
#ifdef VOCR_LSTM_PROF
        long long mf_wait = 0, mf_issue = 0;
#endif
        for (int k = 1; k < tm; ++k) {
#ifdef VOCR_LSTM_PROF
          const long long m0 = clock64();
#endif
          mbar_arrive_expect_tx(full, (uint32_t)a.NSL * piece_bytes);  // arm: the pieces of h_{k-1}
          mbar_wait_or_trap(full, n_full & 1u);
          ++n_full;
          tc_fence_after();
#ifdef VOCR_LSTM_PROF
          const long long m1 = clock64();
#endif
#pragma unroll
          for (int kb = 0; kb < kClMaxKB; ++kb) {
            if (kb < KB) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                const uint64_t d_h = d_h0 + (uint64_t)(((kb * 8 + ks * 2) * G::kChunk) >> 4);
                const uint32_t first = (kb == 0 && ks == 0) ? 0u : 1u;
                cl_umma_ts(t_d, t_a + (uint32_t)(kb * 32 + ks * 8), d_h, idesc_w, first);  // W_hi . [h_hi | h_lo]
                if (kb < G::kWloTmemKB) {  // += W_lo . h_hi into the cross-term columns: W_lo from TMEM ...
                  cl_umma_ts(t_d + NS, t_alo + (uint32_t)(kb * 32 + ks * 8), d_h, idesc_n, 1u);
                } else {                   // ... or from shared memory
                  const uint64_t d_wlo = d_wlo0 + (uint64_t)((((kb - G::kWloTmemKB) * kClWloKb) + ks * 32) >> 4);
                  umma_f16(t_d + NS, d_wlo, d_h, idesc_n, 1u);
                }
              }
            }
          }
          umma_commit(mma_done);
          cl_commit_multicast(tile_free, cta_mask);
#ifdef VOCR_LSTM_PROF
          const long long m2 = clock64();
          mf_wait += m1 - m0; mf_issue += m2 - m1;
#endif
        }
#ifdef VOCR_LSTM_PROF
        if (blockIdx.x == 3) printf("lstm cluster<%d> mma thread: steps %d  wait-full %lld  issue %lld\n", NS, tm - 1, mf_wait, mf_issue);
#endif
      }
      __syncwarp();
    } else {
      // ===================================== input-projection loader warp ========================================
      // lane = sample (and sample + 32): 4 gate rows x US units of xproj[t, b, dir] -> xs[sample][gate*32 + unit], one
      // step ahead
      const int nch = (nu * 4 + 15) / 16;  // 16-byte chunks per gate row (H % 4 == 0)
      int len[NP];
#pragma unroll
      for (int ps = 0; ps < NP; ++ps) {
        const int b = b_base + 32 * ps + lane;
        len[ps] = (b < a.B) ? min(a.lens[b], a.Tmax) : 0;
      }
      for (int k = 0; k < tm; ++k) {
        if (k > 0) {
          mbar_wait_or_trap(xs_free, n_xsfree & 1u);
          ++n_xsfree;
        }
#pragma unroll
        for (int ps = 0; ps < NP; ++ps)
          if (k < len[ps]) {
            const int n = 32 * ps + lane;
            const int tt = dir == 0 ? k : len[ps] - 1 - k;
            const float* src = a.xproj + (((size_t)tt * a.B + b_base + n) * 2 + dir) * 4 * H + u0;
            float* dst = xs + n * G::kXsLd;
            for (int gg = 0; gg < 4; ++gg)
              for (int c = 0; c < nch; ++c)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst + gg * 32 + c * 4)),
                             "l"(src + (size_t)gg * H + c * 4)
                             : "memory");
          }
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(xs_full)) : "memory");
      }
      if (tm > 0) {  // drain the last step's release
        mbar_wait_or_trap(xs_free, n_xsfree & 1u);
        ++n_xsfree;
      }
    }
    // the next item of this cluster starts from a zero state: clear the operand tile once every CTA of the cluster has
    // finished this item (a peer may still be reading its tile / a late piece may still be in flight otherwise)
    tc_fence_before();
    __syncthreads();
    if (item + n_clusters < a.n_items) {
      cl_cluster_sync();
      for (int i = tid; i < (int)(G::kTileBytes / 16); i += kClThreads) reinterpret_cast<uint4*>(Ht)[i] = make_uint4(0, 0, 0, 0);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
    }
  }
  tc_fence_before();
  __syncthreads();
  cl_cluster_sync();  // no CTA leaves while a peer could still multicast into it
  if (warp == 8) tmem_dealloc(tmem_base, 512);
}


// =====================================================================================================================
// Backward recurrence, cluster-resident (same cluster shape and W_hh residency as the forward kernel).
//
// Per step k (k = tm-1 ... 0) and sample:   da_k = gate gradients(dh_k, dc_k, saved gates)   [4H]
//                                            dh_{k-1} = dout_{k-1} + W_hh^T da_k              [H], reduction over 4H.
// CTA s of the cluster owns the same 32 units as in the forward pass, i.e. 128 of the 4H gate rows = a K-SLICE of the
// reduction.  It keeps W_slice^T on chip as the A operand (out unit x gate row; hi plane in TMEM, lo plane in shared
// memory) and multiplies it by ITS OWN gate gradients (the B operand, 128 gate rows x 32 samples, written to shared memory
// by its own epilogue warps): up to four M = 128 tiles of out units, 16 instructions each, three compensated products
// as in the forward kernel.  The result is a PARTIAL dh_{k-1} for all H units; the cluster reduce-scatters it:
//   * the epilogue warps drain a finished tile from TMEM while the next tile's products run, scale it back and store it to
//     a staging area in global memory, one 4-KB piece per consumer CTA (the owner of those 32 units);
//   * one thread pushes each piece into the consumer's shared memory with a bulk copy that completes on the consumer's
//     mbarrier (cp.async.bulk ... multicast::cluster with a one-CTA mask): no flags, no polling; only the last tile's
//     drain + copy + landing is exposed;
//   * the consumer sums its 16 pieces in a fixed order (deterministic), adds dout and forms the next gate gradients.
// Flow control: a consumer tells every producer (remote mbarrier arrive) when it has read the pieces of a step; a
// producer pushes the next ones only after that.  Gate gradients have no fixed range: each sample's 128 values of a CTA
// are scaled by a power of two that puts the largest near 2^14 before the FP16 pair split, and the product columns are
// scaled back in fp32 (as in the mma.sync kernel of lstm.cu).
// Thread mapping of the epilogue warps: warp w owns samples 4w .. 4w+3, lane = unit - every global access (saved gates,
// cell states, dout, dgates) is a 128-byte row, and the per-sample maximum is a warp reduction.
constexpr int kCbNS = 32;                          // samples per cluster work item
constexpr uint32_t kCbBtKb = 64 * 128;             // gate-gradient tile, one k-block of 64 gate rows: 64 rows (32 samples hi |
                                                   // 32 samples lo) x 128 B, K-major SWIZZLE_128B
constexpr uint32_t kCbBtBytes = 2 * kCbBtKb;       // K <= 128 gate rows
constexpr uint32_t kCbPieceFloats = 32 * kCbNS;    // a piece: 32 samples x <= 32 units (fp32), 4 KB slots
constexpr uint32_t kCbColD = 256;                  // accumulators: 4 tiles x 64 columns
constexpr int kCbThreads = 320;                    // 8 epilogue warps, 1 MMA warp, 1 push warp
constexpr size_t kCbSmem = (size_t)8 * kClWloKb + kCbBtBytes + 16 * kCbPieceFloats * 4 + 2 * kCbNS * 4 + 256 + 1024;

struct LstmCbArgs {
  const float* dout;    // [T,B,2H]
  const float* whh;     // [2,4H,H]
  const int32_t* lens;  // [B]
  const float* gates;   // [T,B,2,4H] activated i,f,g,o
  const float* cst;     // [T,B,2,H]
  float* dgates;        // [T,B,2,4H] (pre-zeroed)
  float* absmax;        // max |dgates| (bit pattern), may be null
  float* stage;         // [clusters][2 parities][16 consumers][16 producers][4 KB]
  int T, B, H, US, NSL, MT, UT, Tmax, NG, n_items;  // UT = out units per M tile (a multiple of US), MT tiles
};

// arrive on the mbarrier at the same offset as `bar` in CTA `rank` of the cluster
__device__ __forceinline__ void cb_remote_arrive(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)),
      "r"(rank)
      : "memory");
}
__device__ __forceinline__ void cb_tmem_ld32(uint32_t taddr, uint32_t (&t)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]), "=r"(t[8]),
        "=r"(t[9]), "=r"(t[10]), "=r"(t[11]), "=r"(t[12]), "=r"(t[13]), "=r"(t[14]), "=r"(t[15]), "=r"(t[16]),
        "=r"(t[17]), "=r"(t[18]), "=r"(t[19]), "=r"(t[20]), "=r"(t[21]), "=r"(t[22]), "=r"(t[23]), "=r"(t[24]),
        "=r"(t[25]), "=r"(t[26]), "=r"(t[27]), "=r"(t[28]), "=r"(t[29]), "=r"(t[30]), "=r"(t[31])
      : "r"(taddr)
      : "memory");
}

__global__ void __launch_bounds__(kCbThreads, 1) bilstm_bwd_cluster_kernel(LstmCbArgs a) {
  extern __shared__ unsigned char cl_smem_raw[];
  unsigned char* smem =
      reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(cl_smem_raw) + 1023) & ~uintptr_t(1023));
  const int H = a.H, US = a.US, NSL = a.NSL, MT = a.MT, UT = a.UT;
  const int KS = US / 4;   // k-steps of 16 gate rows (K = 4 US)
  const int CPT = UT / US; // consumer CTAs per M tile
  unsigned char* Wlo = smem;                                        // [MT][2 k-blocks][128 rows x 128 B] SW128 K-major
  unsigned char* Bt = Wlo + (size_t)8 * kClWloKb;                   // [2 k-blocks][64 rows x 128 B] SW128 K-major
  float* pieces = reinterpret_cast<float*>(Bt + kCbBtBytes);        // [16 producers][32 samples][US]
  float* rscale = pieces + 16 * kCbPieceFloats;                     // [2 parities][32 samples] 1 / scale of da
  uint64_t* bars = reinterpret_cast<uint64_t*>(rscale + 2 * kCbNS);
  uint64_t* full = bars;            // the 16 pieces of a step landed (1 arrival + tx bytes)
  uint64_t* b_ready = bars + 1;     // gate-gradient tile written (8 warp arrivals)
  uint64_t* tile_done = bars + 2;   // [4] products of an M tile retired (tcgen05.commit)
  uint64_t* free_bar = bars + 6;    // [16] consumer c has read the pieces of a step (8 remote warp arrivals)
  uint64_t* stored = bars + 22;     // [4] partial products of an M tile stored to the staging area (4 warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 26);

  const int slice = (int)cl_cluster_rank();
  const int cluster_id = blockIdx.x / kClSize, n_clusters = gridDim.x / kClSize;
  const int u0 = slice * US;
  const int nu = max(0, min(US, H - u0));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t piece_bytes = (uint32_t)US * kCbNS * 4u;
  const bool active_cta = slice < NSL;

  if (tid == 0) {
    mbar_init(full, 1);
    mbar_init(b_ready, 8);
    for (int m = 0; m < 4; ++m) mbar_init(&tile_done[m], 1);
    for (int c = 0; c < 16; ++c) mbar_init(&free_bar[c], 8);
    for (int m = 0; m < 4; ++m) mbar_init(&stored[m], 4);
    mbar_fence_init();
  }
  if (warp == 8) tmem_alloc(tmem_slot, 512);
  for (int i = tid; i < (int)(kCbBtBytes / 16); i += kCbThreads) reinterpret_cast<uint4*>(Bt)[i] = make_uint4(0, 0, 0, 0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  cl_cluster_sync();  // every CTA's barriers are initialised before any peer can signal them
  if (!active_cta) {
    for (int item = cluster_id; item + n_clusters < a.n_items; item += n_clusters) cl_cluster_sync();
    cl_cluster_sync();
    if (warp == 8) tmem_dealloc(tmem_base, 512);
    return;
  }

  unsigned n_full = 0, n_bready = 0, n_tile = 0, n_free = 0;
  int loaded_dir = -1;
  float da_max = 0.f;

  for (int item = cluster_id; item < a.n_items; item += n_clusters) {
    const int dir = item & 1, grp = item >> 1;
    const int b_base = grp * kCbNS;
    int tm = 0;
    for (int j = b_base; j < min(a.B, b_base + kCbNS); ++j) tm = max(tm, min(a.lens[j], a.Tmax));

    if (loaded_dir != dir) {
      // ---- A operand: A_m[i][k] = W[gate row (k & 3) H + u0 + (k >> 2)][out unit m UT + i],  k = 4 U + g ------------------
      const float* wd = a.whh + (size_t)dir * 4 * H * H;
      const int Kc = 4 * US;
      if (warp < 4) {  // hi plane -> TMEM: lane = tile row, two k per 32-bit column
        const int r = 32 * warp + lane;
        for (int m = 0; m < MT; ++m) {
          const int uo = m * UT + r;
          const bool row_ok = r < UT && uo < H;
          for (int c0 = 0; c0 < max(32, Kc / 2); c0 += 32) {
            uint32_t v[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const int k = 2 * (c0 + c);
              const int U = k >> 2;  // k and k + 1 belong to the same unit
              float w0 = 0.f, w1 = 0.f;
              if (row_ok && k < Kc && U < nu) {
                w0 = __ldg(wd + ((size_t)(k & 3) * H + u0 + U) * H + uo);
                w1 = __ldg(wd + ((size_t)((k + 1) & 3) * H + u0 + U) * H + uo);
              }
              __half h0, l0, h1, l1;
              cl_split_f16(w0, h0, l0);
              cl_split_f16(w1, h1, l1);
              v[c] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
            }
            cl_tmem_st32(tmem_base + ((uint32_t)(32 * warp) << 16) + (uint32_t)(m * 64 + c0), v);
          }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      for (int idx = tid; idx < MT * Kc * 128; idx += kCbThreads) {  // lo plane -> shared memory (row fastest: coalesced)
        const int i = idx & 127, rest = idx >> 7;
        const int k = rest % Kc, m = rest / Kc;
        const int U = k >> 2, uo = m * UT + i;
        float v = 0.f;
        if (i < UT && uo < H && U < nu) v = __ldg(wd + ((size_t)(k & 3) * H + u0 + U) * H + uo);
        __half hi, lo;
        cl_split_f16(v, hi, lo);
        *reinterpret_cast<__half*>(Wlo + (size_t)(m * 2 + (k >> 6)) * kClWloKb + cl_sw128(i, k & 63)) = lo;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      loaded_dir = dir;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp < 8) {
      // ===================================== epilogue warps ======================================================
      const int U = lane;
      const bool unit_ok = U < nu;
      const int q = warp & 3, half = warp >> 2;
      const int rot = (slice / CPT) % MT;  // this CTA's products start with the tile its own consumer group reads
      int bs[4], len[4];
      float dc[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        bs[j] = b_base + 4 * warp + j;
        len[j] = (bs[j] < a.B) ? min(a.lens[bs[j]], a.Tmax) : 0;
        dc[j] = 0.f;
      }
      // Saved activations of a step: the loads are issued right after the gate-gradient tile of the previous step has been
      // handed to the tensor core and consumed (into the per-(unit, sample) factors below) just before the next step
      // needs them - their latency hides behind the drains.
      float raw[4][7];   // i, f, g, o, c, c_prev, dout
      float fAo[4], fBc[4], fPi[4], fPf[4], fPg[4], fFg[4], fdy[4];
      size_t tbq[4], tbn[4];
      bool act[4], actn[4];
      auto issue_loads = [&](int k) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          actn[j] = unit_ok && k >= 0 && k < len[j];
          const int tq = actn[j] ? (dir == 0 ? k : len[j] - 1 - k) : 0;
          const int bj = actn[j] ? bs[j] : 0;
          const int uu = actn[j] ? u0 + U : 0;
          const size_t tb_ = (size_t)tq * a.B + bj;
          const int tp = (actn[j] && k > 0) ? (dir == 0 ? tq - 1 : tq + 1) : tq;
          const float* gp = a.gates + (tb_ * 2 + dir) * 4 * H + uu;   // inactive lanes read element 0 of valid rows
          raw[j][0] = __ldg(gp);
          raw[j][1] = __ldg(gp + (size_t)H);
          raw[j][2] = __ldg(gp + (size_t)2 * H);
          raw[j][3] = __ldg(gp + (size_t)3 * H);
          raw[j][4] = __ldg(a.cst + (tb_ * 2 + dir) * H + uu);
          raw[j][5] = __ldg(a.cst + (((size_t)tp * a.B + bj) * 2 + dir) * H + uu);
          raw[j][6] = __ldg(a.dout + (tb_ * 2 + dir) * H + uu);
          tbn[j] = tb_;
        }
      };
      auto make_factors = [&](int k) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          act[j] = actn[j];
          tbq[j] = tbn[j];
          const float ig = raw[j][0], fg = raw[j][1], gg = raw[j][2], og = raw[j][3];
          const float c_prev = k > 0 ? raw[j][5] : 0.f;
          const float tc = tanhf(raw[j][4]);
          fdy[j] = raw[j][6];
          fAo[j] = tc * og * (1.f - og);
          fBc[j] = og * (1.f - tc * tc);
          fPi[j] = gg * ig * (1.f - ig);
          fPf[j] = c_prev * fg * (1.f - fg);
          fPg[j] = ig * (1.f - gg * gg);
          fFg[j] = fg;
          // keep the factor arithmetic HERE (off the serial chain): without this the compiler sinks it to the first use
          asm volatile("" : "+f"(fAo[j]), "+f"(fBc[j]), "+f"(fPi[j]), "+f"(fPf[j]), "+f"(fPg[j]), "+f"(fdy[j]));
        }
      };
      if (tm > 0) {
        issue_loads(tm - 1);
        make_factors(tm - 1);
      }
#ifdef VOCR_LSTM_PROF
      long long pf[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#define CB_T(x) const long long x = clock64()
#define CB_ADD(i, a_, b_) pf[i] += (b_) - (a_)
#else
#define CB_T(x)
#define CB_ADD(i, a_, b_)
#endif
      for (int k = tm - 1; k >= 0; --k) {
        CB_T(c0);
        // 1. dh_k = dout + sum of the 16 partial products of step k+1 (fixed order).  A piece holds, for every group of
        //    four samples, [unit][4 samples]: one 16-byte load per producer
        float s[4] = {0.f, 0.f, 0.f, 0.f};
        if (k < tm - 1) {
          if (tid == 0) mbar_arrive_expect_tx(full, (uint32_t)NSL * piece_bytes);
          mbar_wait_or_trap(full, n_full & 1u);
          ++n_full;
          CB_T(c1);
          CB_ADD(0, c0, c1);
          if (U < US) {
            const float4* pp = reinterpret_cast<const float4*>(pieces) + (size_t)warp * US + U;
            const int pstride = US * kCbNS / 4;
            float4 v[16];
#pragma unroll
            for (int p = 0; p < 16; ++p) v[p] = (p < NSL) ? pp[(size_t)p * pstride] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int p = 0; p < 16; ++p) {
              s[0] += v[p].x; s[1] += v[p].y; s[2] += v[p].z; s[3] += v[p].w;
            }
          }
        }
        CB_T(c2);
        CB_ADD(1, c0, c2);
        // 2. gate gradients
        float da[4][4], mx[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float dh = fdy[j] + s[j];
          const float dct = fmaf(dh, fBc[j], dc[j]);
          da[j][0] = dct * fPi[j];
          da[j][1] = dct * fPf[j];
          da[j][2] = dct * fPg[j];
          da[j][3] = dh * fAo[j];
          dc[j] = dct * fFg[j];
          if (!act[j]) da[j][0] = da[j][1] = da[j][2] = da[j][3] = dc[j] = 0.f;
          mx[j] = fmaxf(fmaxf(fabsf(da[j][0]), fabsf(da[j][1])), fmaxf(fabsf(da[j][2]), fabsf(da[j][3])));
        }
        if (k > 0) {
          // 3. per-sample power-of-two scale over this CTA's gate rows (largest magnitude -> [2^14, 2^15)), FP16 pair
          //    split into the operand tile
#pragma unroll
          for (int off = 16; off >= 1; off >>= 1)
#pragma unroll
            for (int j = 0; j < 4; ++j) mx[j] = fmaxf(mx[j], __shfl_xor_sync(0xffffffffu, mx[j], off));
          float rs[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int e = min(240, max(16, 268 - (int)((__float_as_uint(mx[j]) >> 23) & 0xffu)));
            const float sc = __uint_as_float((unsigned)e << 23);
            rs[j] = __uint_as_float((unsigned)(254 - e) << 23);
            if (U < US) {
              const int n = 4 * warp + j;
              __half2 h01 = __floats2half2_rn(da[j][0] * sc, da[j][1] * sc), h23 = __floats2half2_rn(da[j][2] * sc, da[j][3] * sc);
              const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
              __half2 l01 = __floats2half2_rn((da[j][0] * sc - f01.x) * kClLoScale, (da[j][1] * sc - f01.y) * kClLoScale);
              __half2 l23 = __floats2half2_rn((da[j][2] * sc - f23.x) * kClLoScale, (da[j][3] * sc - f23.y) * kClLoScale);
              // gate rows k = 4 U .. 4 U + 3 of sample n: 8 bytes of row n (hi) and of row 32 + n (lo).  With the 128-byte
              // swizzle the 16 units of a k-block land on all 32 banks.
              unsigned char* pb = Bt + (size_t)(U >> 4) * kCbBtKb + cl_sw128(n, (4 * U) & 63);
              *reinterpret_cast<uint2*>(pb) = make_uint2(*reinterpret_cast<uint32_t*>(&h01), *reinterpret_cast<uint32_t*>(&h23));
              *reinterpret_cast<uint2*>(pb + 4096) = make_uint2(*reinterpret_cast<uint32_t*>(&l01), *reinterpret_cast<uint32_t*>(&l23));
            }
          }
          if (lane == 0) *reinterpret_cast<float4*>(rscale + (k & 1) * kCbNS + 4 * warp) = make_float4(rs[0], rs[1], rs[2], rs[3]);
          CB_T(g0);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          CB_T(g1);
          CB_ADD(8, g0, g1);
          __syncwarp();
          if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b_ready)) : "memory");
        }
        CB_T(c3);
        CB_ADD(2, c2, c3);
        // 4. off the chain: the pieces of this step are consumed (producers may push the next ones), loads of the next
        //    step, dgates of this one
        if (k < tm - 1) {
          __syncwarp();
          if (lane < NSL) cb_remote_arrive(&free_bar[slice], (uint32_t)lane);
        }
        if (k > 0) issue_loads(k - 1);  // (writes raw / actn / tbn only: act / tbq still describe step k)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (act[j]) {
            float* dg = a.dgates + (tbq[j] * 2 + dir) * 4 * H + u0 + U;
            dg[0] = da[j][0];
            dg[(size_t)H] = da[j][1];
            dg[(size_t)2 * H] = da[j][2];
            dg[(size_t)3 * H] = da[j][3];
            da_max = fmaxf(da_max, mx[j]);
          }
        if (k == 0) break;
        CB_T(c4);
        CB_ADD(3, c3, c4);
        cl_epi_sync();  // every warp's rscale entries of this step are visible
        CB_T(c5);
        CB_ADD(4, c4, c5);
        // 5. drain the M tiles as their products retire: warps 0-3 take the 1st and 3rd tile of the CTA's order, warps 4-7
        //    the 2nd and 4th
        const int par = k & 1;
        for (int i = half; i < MT; i += 2) {
          const int t = (i + rot) % MT;
          CB_T(d0);
          mbar_wait_or_trap(&tile_done[i], n_tile & 1u);
          tc_fence_after();
          CB_T(d1);
          CB_ADD(5, d0, d1);
          uint32_t xh[32], xx[32];
          const uint32_t tbase = tmem_base + ((uint32_t)(32 * q) << 16) + kCbColD + (uint32_t)(t * 64);
          cb_tmem_ld32(tbase, xh);
          cb_tmem_ld32(tbase + 32, xx);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          tc_fence_before();
          const int r = 32 * q + lane, uo = t * UT + r;
          if (r < UT && uo < H) {
            const int c = uo / US, ul = uo - c * US;
            float4* dst = reinterpret_cast<float4*>(
                              a.stage + ((((size_t)cluster_id * 2 + par) * 16 + c) * 16 + slice) * kCbPieceFloats) + ul;
            const float4* rsc = reinterpret_cast<const float4*>(rscale + par * kCbNS);
#pragma unroll
            for (int g4 = 0; g4 < 8; ++g4) {
              const float4 rr = rsc[g4];
              float4 o;
              o.x = fmaf(__uint_as_float(xx[4 * g4 + 0]), 1.f / kClLoScale, __uint_as_float(xh[4 * g4 + 0])) * rr.x;
              o.y = fmaf(__uint_as_float(xx[4 * g4 + 1]), 1.f / kClLoScale, __uint_as_float(xh[4 * g4 + 1])) * rr.y;
              o.z = fmaf(__uint_as_float(xx[4 * g4 + 2]), 1.f / kClLoScale, __uint_as_float(xh[4 * g4 + 2])) * rr.z;
              o.w = fmaf(__uint_as_float(xx[4 * g4 + 3]), 1.f / kClLoScale, __uint_as_float(xh[4 * g4 + 3])) * rr.w;
              dst[(size_t)g4 * US] = o;
            }
          }
          CB_T(d2);
          CB_ADD(6, d1, d2);
          // stored: the push warp sends the tile's pieces to their consumers
          __syncwarp();
          if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&stored[i])) : "memory");
          CB_T(d3);
          CB_ADD(7, d2, d3);
        }
        CB_T(c6);
        // the gate-gradient tile and the accumulators are rewritten by the next step: every product of this one has retired
        mbar_wait_or_trap(&tile_done[MT - 1], n_tile & 1u);
        ++n_tile;
        make_factors(k - 1);
        CB_T(c7);
        CB_ADD(9, c6, c7);
      }
#ifdef VOCR_LSTM_PROF
      if (cluster_id == 0 && (slice == 0 || slice == 15) && lane == 0 && (warp == 0 || warp == 4))
        printf("lstm bwd cluster: slice %d warp %d steps %d | wait-full %lld  sum(+wait) %lld  gates->b_ready %lld  loads+dgates %lld  epi-sync %lld | "
               "wait-tile %lld  ld+store %lld  stored-arrive %lld  proxy-fence(gates) %lld  wait-last+factors %lld (cycles/step)\n",
               slice, warp, tm, pf[0] / tm, pf[1] / tm, pf[2] / tm, pf[3] / tm, pf[4] / tm, pf[5] / tm, pf[6] / tm, pf[7] / tm,
               pf[8] / tm, pf[9] / tm);
#endif
    } else if (warp == 9) {
      // ===================================== push warp ===========================================================
      // lane j pushes the piece of the j-th consumer of a tile (the copies of a tile are issued side by side)
      {
        const int rot = (slice / CPT) % MT;
#ifdef VOCR_LSTM_PROF
        long long ps_wait = 0, ps_fence = 0, ps_push = 0;
#endif
        for (int k = tm - 1; k >= 1; --k) {
          const int par = k & 1;
          for (int i = 0; i < MT; ++i) {
            const int t = (i + rot) % MT;
#ifdef VOCR_LSTM_PROF
            const long long q0 = clock64();
#endif
            mbar_wait_or_trap(&stored[i], n_tile & 1u);
#ifdef VOCR_LSTM_PROF
            const long long q1 = clock64();
#endif
            asm volatile("fence.proxy.async.global;" ::: "memory");
#ifdef VOCR_LSTM_PROF
            const long long q2 = clock64();
#endif
            const int c = t * CPT + lane;
            if (lane < CPT && c < NSL) {
              // consumer c has read the pieces of step k (nothing to wait for on the first push of an item)
              if (k < tm - 1) mbar_wait_or_trap(&free_bar[c], (n_free + (unsigned)(tm - 2 - k)) & 1u);
              const float* src = a.stage + ((((size_t)cluster_id * 2 + par) * 16 + c) * 16 + slice) * kCbPieceFloats;
              asm volatile(
                  "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                  ::"r"(smem_u32(pieces + (size_t)slice * US * kCbNS)), "l"(src), "r"(piece_bytes), "r"(smem_u32(full)),
                  "h"((uint16_t)(1u << c))
                  : "memory");
            }
            __syncwarp();
#ifdef VOCR_LSTM_PROF
            const long long q3 = clock64();
            ps_wait += q1 - q0; ps_fence += q2 - q1; ps_push += q3 - q2;
#endif
          }
          ++n_tile;
        }
#ifdef VOCR_LSTM_PROF
        if (cluster_id == 0 && (slice == 0 || slice == 15) && lane == 0)
          printf("lstm bwd cluster: slice %d push warp: wait-stored %lld  fence %lld  free-wait+issue %lld (cycles/step)\n", slice,
                 ps_wait / tm, ps_fence / tm, ps_push / tm);
#endif
        if (tm > 1) n_free += (unsigned)(tm - 1);
      }
      __syncwarp();
    } else {
      // ===================================== MMA issuer ==========================================================
      if (lane == 0) {
        const uint32_t idesc_w = (1u << 4) | ((uint32_t)(2 * kCbNS >> 3) << 17) | ((uint32_t)(kClM >> 4) << 24);
        const uint32_t idesc_n = (1u << 4) | ((uint32_t)(kCbNS >> 3) << 17) | ((uint32_t)(kClM >> 4) << 24);
        const uint64_t d_wlo0 = make_desc(smem_u32(Wlo), 16, 1024, 2);
        const uint64_t d_b0 = make_desc(smem_u32(Bt), 16, 1024, 2);
        const uint32_t t_a = tmem_base + kColA, t_d = tmem_base + kCbColD;
        // tile order rotated per CTA (the pushes of a step spread over its whole length); per-position operand bases
        const int rot = (slice / CPT) % MT;
        uint32_t tile_d[4], tile_a[4];
        uint64_t tile_wlo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int m = (i + rot) % MT;
          tile_d[i] = t_d + (uint32_t)(m * 64);
          tile_a[i] = t_a + (uint32_t)(m * 64);
          tile_wlo[i] = d_wlo0 + (uint64_t)((m * 2 * kClWloKb) >> 4);
        }
#ifdef VOCR_LSTM_PROF
        long long mf_wait = 0, mf_issue = 0;
#endif
        for (int k = tm - 1; k >= 1; --k) {
#ifdef VOCR_LSTM_PROF
          const long long m0 = clock64();
#endif
          mbar_wait_or_trap(b_ready, n_bready & 1u);
          ++n_bready;
          tc_fence_after();
#ifdef VOCR_LSTM_PROF
          const long long m1 = clock64();
#endif
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (i < MT) {
#pragma unroll
              for (int ks = 0; ks < 8; ++ks) {
                if (ks < KS) {
                  const uint64_t d_b = d_b0 + (uint64_t)(((ks >> 2) * kCbBtKb + (ks & 3) * 32) >> 4);
                  // W_hi . [da_hi | da_lo] -> columns 0..31 | 32..63;  += W_lo . da_hi into the cross-term columns
                  cl_umma_ts(tile_d[i], tile_a[i] + (uint32_t)(ks * 8), d_b, idesc_w, ks > 0 ? 1u : 0u);
                  umma_f16(tile_d[i] + kCbNS, tile_wlo[i] + (uint64_t)((((ks >> 2) * kClWloKb) + (ks & 3) * 32) >> 4), d_b,
                           idesc_n, 1u);
                }
              }
              umma_commit(&tile_done[i]);
            }
          }
#ifdef VOCR_LSTM_PROF
          const long long m2 = clock64();
          mf_wait += m1 - m0;
          mf_issue += m2 - m1;
#endif
        }
#ifdef VOCR_LSTM_PROF
        if (cluster_id == 0 && (slice == 0 || slice == 15))
          printf("lstm bwd cluster: slice %d mma thread steps %d  wait-b_ready %lld  issue %lld (cycles/step)\n", slice, tm,
                 mf_wait / tm, mf_issue / tm);
#endif
      }
      __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    // the next item reuses the piece buffer and the barriers' phases: every CTA has finished this one first
    if (item + n_clusters < a.n_items) cl_cluster_sync();
  }
  if (a.absmax && warp < 8) {  // non-negative floats order like their bit patterns
    da_max = warp_max(da_max);
    if (lane == 0) atomicMax(reinterpret_cast<unsigned*>(a.absmax), __float_as_uint(da_max));
  }
  tc_fence_before();
  __syncthreads();
  cl_cluster_sync();  // no CTA leaves while a peer could still push into it or arrive on its barriers
  if (warp == 8) tmem_dealloc(tmem_base, 512);
}

}  // namespace vocr

using namespace vocr;

static int lstm_cl_samples(int B) { return B >= 128 ? 64 : 32; }  // samples per cluster work item

static int lstm_cl_geometry(int B, int H, LstmClArgs* a, size_t* smem) {
  if (H < 1 || H > 64 * kClMaxKB || (H % 4) != 0) return VOCR_INVALID_VALUE;
  const int ns = lstm_cl_samples(B);
  a->KB = ceil_div(H, 64);
  a->US = 8 * ceil_div(H, 8 * kClSize);  // unit slots per CTA: a multiple of 8 (one 16-byte chunk), at most 32
  a->NSL = ceil_div(H, a->US);
  a->NG = ceil_div(B, ns);
  a->n_items = 2 * a->NG;
  *smem = ns == 64 ? ClGeom<64>::kSmem : ClGeom<32>::kSmem;
  return VOCR_OK;
}

bool lstm_tc_enabled() {
  const char* e = getenv("VOCR_LSTM_TC");
  return !(e && e[0] == '0');
}

static int lstm_cl_clusters(int n_items) {
  // at most 7 clusters of 16 CTAs are co-resident on a B200 (cudaOccupancyMaxActiveClusters at this footprint); with an
  // odd count a cluster alternates between the directions and re-loads its W_hh slice per item (~1 % of an item)
  const int nc = n_items < 7 ? n_items : 7;
  return nc < 1 ? 1 : nc;
}

size_t lstm_tc_fwd_workspace_bytes(int T, int B, int H) {
  (void)T;
  LstmClArgs a;
  size_t smem;
  if (!lstm_tc_enabled() || lstm_cl_geometry(B, H, &a, &smem) != VOCR_OK) return 0;
  return 1024 + (size_t)lstm_cl_clusters(a.n_items) * kClSize * 2 * ClGeom<64>::kStageBytes;
}

template <int NS>
static int lstm_cl_launch(const LstmClArgs& a, int nc, size_t smem, cudaStream_t stream) {
  static DeviceLatch latch;
  if (latch.need()) {
    if (cudaFuncSetAttribute(bilstm_fwd_cluster_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) !=
            cudaSuccess ||
        cudaFuncSetAttribute(bilstm_fwd_cluster_kernel<NS>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess)
      return VOCR_EXECUTION_FAILED;
    latch.set();
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(nc * kClSize);
  cfg.blockDim = dim3(kClThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = kClSize;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, bilstm_fwd_cluster_kernel<NS>, a) != cudaSuccess) return VOCR_EXECUTION_FAILED;
  return VOCR_OK;
}

int lstm_tc_fwd_launch(const float* xproj, const float* whh, const int32_t* lens, float* out, float* gates, float* cst,
                       int T, int B, int H, int Tmax, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  LstmClArgs a{};
  size_t smem;
  int st = lstm_cl_geometry(B, H, &a, &smem);
  if (st != VOCR_OK) return -1;  // shape not covered by this kernel: the caller uses the mma.sync kernel
  a.xproj = xproj; a.whh = whh; a.lens = lens; a.out = out; a.gates = gates; a.cst = cst;
  a.T = T; a.B = B; a.H = H; a.Tmax = Tmax;
  if ((reinterpret_cast<uintptr_t>(xproj) & 15) != 0) return -1;
  const int nc = lstm_cl_clusters(a.n_items);
  const uintptr_t w = (reinterpret_cast<uintptr_t>(workspace) + 1023) & ~uintptr_t(1023);
  const size_t need = (size_t)nc * kClSize * 2 * ClGeom<64>::kStageBytes;
  if ((w - reinterpret_cast<uintptr_t>(workspace)) + need > workspace_bytes) return VOCR_INVALID_VALUE;
  a.stage = reinterpret_cast<unsigned char*>(w);
  return lstm_cl_samples(B) == 64 ? lstm_cl_launch<64>(a, nc, smem, stream) : lstm_cl_launch<32>(a, nc, smem, stream);
}

// ---- backward -----------------------------------------------------------------------------------------------------------
static int lstm_cb_geometry(int B, int H, LstmCbArgs* a) {
  if (H < 1 || H > 64 * kClMaxKB || (H % 4) != 0) return VOCR_INVALID_VALUE;
  a->US = 8 * ceil_div(H, 8 * kClSize);
  if (128 % a->US != 0) return VOCR_INVALID_VALUE;  // 24 unit slots per CTA (256 < H <= 384): the mma.sync kernel serves it
  a->NSL = ceil_div(H, a->US);
  a->UT = 128;
  a->MT = ceil_div(H, 128);
  a->NG = ceil_div(B, kCbNS);
  a->n_items = 2 * a->NG;
  return VOCR_OK;
}

// The cluster backward kernel is correct (tests/test_gpu_ops.py runs it) but measured SLOWER than the mma.sync kernel of
// lstm.cu at the benchmark shape (T = 294, B = 64, H = 512: 7.2 vs 6.3 us per step; DESIGN.md 4.4 has the phase table), so
// it is opt-in: VOCR_LSTM_TC_BWD=1.
static bool lstm_tc_bwd_enabled() {
  const char* e = getenv("VOCR_LSTM_TC_BWD");
  return lstm_tc_enabled() && e && e[0] == '1';
}

size_t lstm_tc_bwd_workspace_bytes(int T, int B, int H) {
  (void)T;
  LstmCbArgs a;
  if (!lstm_tc_bwd_enabled() || lstm_cb_geometry(B, H, &a) != VOCR_OK) return 0;
  return 1024 + (size_t)lstm_cl_clusters(a.n_items) * 2 * 16 * 16 * kCbPieceFloats * sizeof(float);
}

// -1 = shape not covered (the caller uses the mma.sync kernel)
int lstm_tc_bwd_launch(const float* dout, const float* whh, const int32_t* lens, const float* gates, const float* cst,
                       float* dgates, float* absmax, int T, int B, int H, int Tmax, void* workspace,
                       size_t workspace_bytes, cudaStream_t stream) {
  LstmCbArgs a{};
  if (!lstm_tc_bwd_enabled() || lstm_cb_geometry(B, H, &a) != VOCR_OK) return -1;
  a.dout = dout; a.whh = whh; a.lens = lens; a.gates = gates; a.cst = cst; a.dgates = dgates; a.absmax = absmax;
  a.T = T; a.B = B; a.H = H; a.Tmax = Tmax;
  const int nc = lstm_cl_clusters(a.n_items);
  const uintptr_t w = (reinterpret_cast<uintptr_t>(workspace) + 1023) & ~uintptr_t(1023);
  const size_t need = (size_t)nc * 2 * 16 * 16 * kCbPieceFloats * sizeof(float);
  if ((w - reinterpret_cast<uintptr_t>(workspace)) + need > workspace_bytes) return VOCR_INVALID_VALUE;
  a.stage = reinterpret_cast<float*>(w);
  static DeviceLatch latch;
  if (latch.need()) {
    if (cudaFuncSetAttribute(bilstm_bwd_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) !=
            cudaSuccess ||
        cudaFuncSetAttribute(bilstm_bwd_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess)
      return VOCR_EXECUTION_FAILED;
    latch.set();
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(nc * kClSize);
  cfg.blockDim = dim3(kCbThreads);
  cfg.dynamicSmemBytes = kCbSmem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = kClSize;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, bilstm_bwd_cluster_kernel, a) != cudaSuccess) return VOCR_EXECUTION_FAILED;
  return VOCR_OK;
}
