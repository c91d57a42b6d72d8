// Bidirectional LSTM layer recurrence on sm_100a, forward and backward, as PERSISTENT cooperative kernels
// (reference cnnlstm.py:148-149,285-290: nn.LSTM on a packed sequence -> cuDNN RNN).
//
// The input projections x_t W_ih^T + b_ih + b_hh of ALL timesteps and both directions are one tensor-core GEMM done
// beforehand (xproj [T,B,2,4H]); what is left is the strictly sequential part  h_{t-1} W_hh^T  + gate nonlinearities.
//
// Work decomposition.  An "instance" = (direction, tile of 16 samples); it is served by NSL CTAs ("slices"), each
// owning US hidden units = 4*US rows of W_hh, which stay RESIDENT IN SHARED MEMORY for all timesteps as FP16 pairs
// (hi, lo * 2^11): the same 4 bytes per weight as fp32 (H=512: 32 slices x 64 rows x 512 = 130 KB each).  Per step a
// slice multiplies the 16 x H tile of h_{t-1} by its 64 rows on the tensor cores (mma.sync m16n8k16 FP16, three
// error-compensated products hi*hi + hi*lo + lo*hi accumulated in fp32: 22 significant bits, the accuracy of the
// 3xTF32 GEMMs at twice the legacy tensor rate and with no split arithmetic in the loop; K split over 4 warp pairs
// and reduced through shared memory), applies the gates for its units, and publishes its 16 x US piece of h_t -
// already split into the FP16 pair - through a ping-pong buffer in global memory (L2); the slices of an instance
// meet at a monotonically increasing flag (release/acquire at gpu scope) - there is no grid-wide barrier.
// Latency hiding: every CTA serves TWO instances (two sample tiles of one direction) in alternation, so the L2 round
// trip of one tile's h exchange (flag + 32 KB load) overlaps the other tile's product and gate math.
// B=64, H=512: 2 directions x 2 tile pairs x 32 slices = 128 CTAs, one per SM, both directions concurrent.
// Ragged lengths use packed-sequence semantics by masking: sample b is active at step k iff k < lens[b]; the reverse
// direction visits t = lens[b]-1-k, i.e. starts at the sample's own last frame; outputs beyond lens[b] stay zero.
//
// Backward keeps the same residency and interleaving: a slice turns dh into gate gradients for its own units,
// multiplies them by its W_hh rows (16 x 64 by 64 x H, tensor cores) into a partial dh_{t-1} for ALL units, and the
// instance reduce-scatters the partials through L2 in a fixed order (deterministic).  dW_hh / dW_ih / db / dx are
// tensor-core GEMMs over the saved gate gradients afterwards (host side, vistaocr_b200/ops.py).
#include <cuda_fp16.h>
#ifdef VOCR_LSTM_PROF
#include <cstdio>
#endif

#include <algorithm>

#include "common.cuh"

namespace vocr {

constexpr int kLstmThreads = 320;     // 8 compute warps + one communication warp per interleaved instance
constexpr int kLstmBT = 16;        // samples per instance = MMA M
constexpr int kLstmNI = 2;         // instances interleaved per CTA
constexpr int kLstmMaxUS = 16;     // hidden units per slice
constexpr int kLstmRows = 64;      // 4 * kLstmMaxUS gate rows per slice (zero padded)
constexpr int kLstmDaLd = kLstmRows + 16;  // gate-gradient rows: 128-bit fragment loads of 2 rows x 4 lanes conflict-free

struct LstmArgs {
  const float* xproj;   // [T,B,2,4H]  (fwd)            | dout [T,B,2H] (bwd)
  const float* whh;     // [2,4H,H]
  const int32_t* lens;  // [B]
  float* out;           // [T,B,2H]    (fwd, pre-zeroed) | dgates [T,B,2,4H] (bwd, pre-zeroed)
  float* gates;         // [T,B,2,4H] activated i,f,g,o (fwd: written if non-null; bwd: read)
  float* cst;           // [T,B,2,H]  cell state        (fwd: written if non-null; bwd: read)
  float* xchg;          // fwd: [n_inst][2][16][Hp hi | Hp lo | pad] fp16 | bwd: [n_inst][2][NSL cons][NSL prod][16][US]
  unsigned* flags;      // [n_inst] arrival counters, zeroed before launch
  float* absmax;        // bwd: receives max |dgates| (bit pattern, atomicMax on the non-negative float), may be null
  int T, B, H, Hp, US, NSL, Tmax, NBT, gpd, n_groups;  // gpd = instance pairs per direction
};

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release(unsigned* p) {
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// ---- forward operands: FP16 pairs (hi, lo * 2^11), error-compensated like 3xTF32 but at twice the tensor rate --------
// x = hi + lo' * 2^-11 with hi = fp16(x), lo' = fp16((x - hi) * 2^11): 22 significant bits, and the three products
// hi*hi + (hi*lo' + lo'*hi) * 2^-11 accumulate in fp32.  Both recurrence operands are range-safe in FP16: |h| < 1 and
// the lo parts are pre-scaled out of the denormal range; W_hh must stay below 65504 in magnitude.  Values smaller than
// 2^-14 fall on FP16's denormal grid: an ABSOLUTE error below 2^-25 per element, under the fp32 rounding of the sum.
constexpr float kLoScale = 2048.f;
__device__ __forceinline__ void split_f16(float x, __half& hi, __half& lo) {
  hi = __float2half_rn(x);
  lo = __float2half_rn((x - __half2float(hi)) * kLoScale);
}
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// forward copy of the slice's W_hh rows: row r = g*US + u holds [Hp halves hi | Hp halves lo'] (+ padding to ld words)
template <int ROWS, int NTHREADS>
__device__ __forceinline__ void load_w_slice_f16(float* Ws, const float* __restrict__ whh_dir, int H, int Hp, int US,
                                                 int u0, int nu, int ld) {
  __half* Wh = reinterpret_cast<__half*>(Ws);
  for (int i = threadIdx.x; i < ROWS * Hp; i += NTHREADS) {
    const int r = i / Hp, k = i - r * Hp;
    const int g = r / US, u = r - g * US;
    float v = 0.f;
    if (g < 4 && u < nu && k < H) v = __ldg(whh_dir + ((size_t)g * H + u0 + u) * H + k);
    __half hi, lo;
    split_f16(v, hi, lo);
    Wh[(size_t)r * 2 * ld + k] = hi;
    Wh[(size_t)r * 2 * ld + Hp + k] = lo;
  }
}

// A CTA group serves NI consecutive entries of the folded tile order: batches arrive sorted by width
// (SortByWidthCollater), so long tiles are mixed with short ones and every group gets about the same number of steps.
// folded tile order 0, NBT-1, 1, NBT-2, ...: consecutive entries pair a long tile with a short one
__device__ __forceinline__ int lstm_fold(int idx, int NBT) { return (idx & 1) ? NBT - 1 - (idx >> 1) : (idx >> 1); }
// steps an instance really needs: the longest sample of its tile (the rest of [0, Tmax) would be all-masked work)
__device__ __forceinline__ int lstm_tile_tmax(const int32_t* lens, int b0, int B, int Tmax) {
  int m = 0;
  for (int j = b0; j < min(B, b0 + kLstmBT); ++j) m = max(m, min(lens[j], Tmax));
  return m;
}

// named barrier among the 256 compute threads only (the producer warps never join it)
__device__ __forceinline__ void compute_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive1(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// one arrival per warp, after the warp's lanes have synchronised (orders every lane's prior writes before the arrive)
__device__ __forceinline__ void warp_arrive(uint64_t* bar) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive1(bar);
}
__device__ __forceinline__ void poll_flag(const unsigned* flag, unsigned target) {
  unsigned spins = 0;
  while (ld_acquire(flag) < target) {
    if (++spins > (1u << 26)) asm volatile("trap;");  // ~seconds: a lost peer must not hang the box
  }
}

// Warp roles: warps 0-7 compute; warp 8+i is the COMMUNICATION warp of instance i.  It waits for the instance's flag,
// pulls the 16 x H tile of h_{t-1} into shared memory with bulk async copies (mbarrier `full`), and - once the compute
// warps have published their piece of h_t (mbarrier `written`) - makes it visible (fence) and bumps the flag.  The
// compute warps therefore never sit on an L2 round trip: while instance 0's exchange is in flight they work on
// instance 1.

// ================================================ forward =====================================================
// ROWS = gate rows per slice (4 * units: 64 -> 32 slices at H=512, 32 -> 64 slices), NI = sample tiles interleaved per
// CTA.  <64,2> serves small batches; <32,4> puts all four 16-sample tiles of a 64-line batch on every CTA of a
// direction, so three other tiles' worth of tensor work hides each tile's L2 exchange.
template <int ROWS, int NI>
__global__ void __launch_bounds__(256 + 32 * NI, 1) bilstm_fwd_kernel(LstmArgs a) {
  constexpr int NT = ROWS / 16;          // n-tiles (8 gate rows each) per warp
  constexpr int PLD = ROWS + 8;          // row stride of the K-split partial sums
  constexpr int NTHREADS = 256 + 32 * NI;
  extern __shared__ __align__(16) float lstm_smem[];
  const int H = a.H, Hp = a.Hp, US = a.US, ld = Hp + 8;  // ld in 32-bit words; +8 keeps the 64-bit fragment loads
                                                         // of 8 rows x 4 lanes on distinct banks
  float* Ws = lstm_smem;                              // [ROWS][ld]   fp16 pairs: hi plane | lo plane
  float* hbuf = Ws + ROWS * ld;                       // [NI][16][ld] same row format
  float* part = hbuf + (size_t)NI * kLstmBT * ld;     // [4][16][72]
  uint64_t* bars = reinterpret_cast<uint64_t*>(part + 4 * kLstmBT * PLD);
  uint64_t* full = bars;                  // [NI] h tile landed           (tx bytes)
  uint64_t* empty = bars + NI;       // [NI] product done with hs[i]  (8 warp arrivals)
  uint64_t* written = bars + 2 * NI; // [NI] h_t piece stored         (8 warp arrivals)
  const int slice = blockIdx.x;
  const int u0 = slice * US;
  const int nu = min(US, H - u0);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < NI; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 8);
      mbar_init(&written[i], 8);
    }
    mbar_fence_init();
  }
  // running phase counters (barriers are used across groups without re-initialisation)
  unsigned n_full[NI], n_empty[NI], n_written[NI];
#pragma unroll
  for (int i = 0; i < NI; ++i) n_full[i] = n_empty[i] = n_written[i] = 0;

  for (int grp = blockIdx.y; grp < a.n_groups; grp += gridDim.y) {
    const int dir = grp / a.gpd, pair = grp - dir * a.gpd;
    const int ni = min(NI, a.NBT - NI * pair);
    __syncthreads();
    load_w_slice_f16<ROWS, NTHREADS>(Ws, a.whh + (size_t)dir * 4 * H * H, H, Hp, US, u0, nu, ld);
    for (int i = tid; i < NI * kLstmBT * ld; i += NTHREADS) hbuf[i] = 0.f;  // padding columns stay zero
    __syncthreads();

    if (warp >= 8) {
      // ------------------------------ communication warp of instance i ------------------------------
      const int i = warp - 8;
      if (i < ni) {
        const int tile = lstm_fold(NI * pair + i, a.NBT);
        const int inst = dir * a.NBT + tile;
        const int tm = lstm_tile_tmax(a.lens, tile * kLstmBT, a.B, a.Tmax);
        float* hx = a.xchg + (size_t)inst * 2 * kLstmBT * ld;  // rows padded like the shared-memory tile
        unsigned* flag = a.flags + inst;
        float* hs = hbuf + (size_t)i * kLstmBT * ld;
#ifdef VOCR_LSTM_PROF
        long long cw_empty = 0, cw_poll = 0, cw_copy = 0, cw_written = 0, cw_red = 0;
        unsigned cw_nfull = 0;
#endif
        for (int k = 0; k < tm; ++k) {
          if (k > 0) {
#ifdef VOCR_LSTM_PROF
            const long long q0 = clock64();
#endif
            if (k > 1) {  // hs[i] is free once the product of step k-1 has consumed it
              mbar_wait_or_trap(&empty[i], (n_empty[i] & 1u));
              ++n_empty[i];
            }
#ifdef VOCR_LSTM_PROF
            const long long q1 = clock64();
#endif
            if (lane == 0) {
              poll_flag(flag, (unsigned)(a.NSL * k));
              asm volatile("fence.proxy.async;" ::: "memory");
              mbar_arrive_expect_tx(&full[i], (uint32_t)(kLstmBT * ld * 4));
              // ONE bulk copy for the whole tile (a bulk copy costs ~120 cycles of serial issue whatever its size)
              bulk_g2s(hs, hx + (size_t)((k - 1) & 1) * kLstmBT * ld, (uint32_t)(kLstmBT * ld * 4), &full[i]);
            }
            __syncwarp();
#ifdef VOCR_LSTM_PROF
            const long long q2 = clock64();
#endif
#ifdef VOCR_LSTM_PROF
            mbar_wait_or_trap(&full[i], cw_nfull & 1u);  // probe only: how long the 32 KB take to land
            ++cw_nfull;
            const long long q3 = clock64();
            cw_empty += q1 - q0; cw_poll += q2 - q1; cw_copy += q3 - q2;
#endif
          }
          if (k + 1 < tm) {
#ifdef VOCR_LSTM_PROF
            const long long q4 = clock64();
#endif
            mbar_wait_or_trap(&written[i], (n_written[i] & 1u));
            ++n_written[i];
#ifdef VOCR_LSTM_PROF
            const long long q5 = clock64();
#endif
            if (lane == 0) red_release(flag);  // release: cumulative over the stores ordered by `written`
#ifdef VOCR_LSTM_PROF
            __syncwarp();
            const long long q6 = clock64();
            cw_written += q5 - q4; cw_red += q6 - q5;
#endif
          }
        }
#ifdef VOCR_LSTM_PROF
        if (blockIdx.x == 0 && blockIdx.y == 0 && lane == 0)
          printf("lstm fwd comm warp %d: steps %d  wait-empty %lld  poll %lld  copy-land %lld  wait-written %lld  red.release %lld (cycles)\n",
                 i, tm, cw_empty, cw_poll, cw_copy, cw_written, cw_red);
#endif
        if (tm > 1) {  // drain: the last product's release of hs[i] (keeps the phase counters in step)
          mbar_wait_or_trap(&empty[i], (n_empty[i] & 1u));
          ++n_empty[i];
        }
      }
    } else {
      // ------------------------------------- compute warps -------------------------------------
      const int g = lane >> 2, t = lane & 3;          // mma fragment coordinates
      const int kgrp = warp >> 1, nh = warp & 1;      // K quarter, half of the 64 gate rows
      const int kqw = Hp / 8;                         // words of a K quarter (Hp % 64 == 0 -> whole k16 steps)
      const int lo_off = Hp / 2;                      // word offset of the lo plane inside a row
      const int pb = tid / US, pu = tid - pb * US;    // gate epilogue: one (sample, unit) pair per thread, instance
      int b0[NI], plen[NI], tmx[NI];
      float* hx[NI];
      float c_reg[NI], h_reg[NI];
      bool pok[NI];
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int tile = lstm_fold(NI * pair + i, a.NBT);
        const int inst = dir * a.NBT + tile;
        b0[i] = tile * kLstmBT;
        tmx[i] = (i < ni) ? lstm_tile_tmax(a.lens, b0[i], a.B, a.Tmax) : 0;
        const int nb = min(kLstmBT, a.B - b0[i]);
        pok[i] = (i < ni) && pb < kLstmBT && pb < nb && pu < nu;
        plen[i] = pok[i] ? min(a.lens[b0[i] + pb], a.Tmax) : 0;
        hx[i] = a.xchg + (size_t)inst * 2 * kLstmBT * ld;
        c_reg[i] = 0.f;
        h_reg[i] = 0.f;
      }
#ifdef VOCR_LSTM_PROF
      long long pf_wait = 0, pf_prod = 0, pf_sync = 0, pf_epi = 0, pf_pub = 0, pf_t0 = clock64();
#define PF_T(v) const long long v = clock64()
#define PF_ADD(acc, x, y) acc += (y) - (x)
#else
#define PF_T(v)
#define PF_ADD(acc, x, y)
#endif
      for (int k = 0; k < a.Tmax; ++k) {
#pragma unroll
        for (int i = 0; i < NI; ++i) {
          if (i >= ni || k >= tmx[i]) continue;
          const float* hs = hbuf + (size_t)i * kLstmBT * ld;
          const bool act = pok[i] && k < plen[i];
          const int tt = act ? (dir == 0 ? k : plen[i] - 1 - k) : -1;
          float pre[4] = {0.f, 0.f, 0.f, 0.f};
          if (act) {  // input projections: independent of the recurrence, loads issued before any wait
            const float* xr = a.xproj + (((size_t)tt * a.B + b0[i] + pb) * 2 + dir) * 4 * H + u0 + pu;
#pragma unroll
            for (int q = 0; q < 4; ++q) pre[q] = __ldg(xr + (size_t)q * H);
          }
          PF_T(c0);
          if (k > 0) {
            mbar_wait_or_trap(&full[i], (n_full[i] & 1u));
            ++n_full[i];
            PF_T(c1);
            PF_ADD(pf_wait, c0, c1);
            // D[16 x ROWS/2 rows of this warp] += h[16 x K quarter] * W^T.  One 64-bit load per row fetches the two
            // k-pairs a lane owns in a k16 step (A and B use the same assignment of the 16 k to fragment slots).
            float acc[NT][4], acl[NT][4];
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
              for (int q = 0; q < 4; ++q) acc[j][q] = acl[j][q] = 0.f;
            const uint32_t* hA = reinterpret_cast<const uint32_t*>(hs) + (size_t)g * ld + kgrp * kqw + 2 * t;
            const uint32_t* wB =
                reinterpret_cast<const uint32_t*>(Ws) + (size_t)(nh * (ROWS / 2) + g) * ld + kgrp * kqw + 2 * t;
#pragma unroll 4
            for (int w0 = 0; w0 < kqw; w0 += 8) {
              const uint2 h0 = *reinterpret_cast<const uint2*>(hA + w0);
              const uint2 h1 = *reinterpret_cast<const uint2*>(hA + (size_t)8 * ld + w0);
              const uint2 l0 = *reinterpret_cast<const uint2*>(hA + lo_off + w0);
              const uint2 l1 = *reinterpret_cast<const uint2*>(hA + (size_t)8 * ld + lo_off + w0);
              const uint32_t ah[4] = {h0.x, h1.x, h0.y, h1.y};
              const uint32_t al[4] = {l0.x, l1.x, l0.y, l1.y};
#pragma unroll
              for (int j = 0; j < NT; ++j) {
                const uint2 wh = *reinterpret_cast<const uint2*>(wB + (size_t)(j * 8) * ld + w0);
                const uint2 wl = *reinterpret_cast<const uint2*>(wB + (size_t)(j * 8) * ld + lo_off + w0);
                const uint32_t bh[2] = {wh.x, wh.y}, bl[2] = {wl.x, wl.y};
                mma_f16(acl[j], al, bh);
                mma_f16(acl[j], ah, bl);
                mma_f16(acc[j], ah, bh);
              }
            }
            PF_T(c2);
            PF_ADD(pf_prod, c1, c2);
            warp_arrive(&empty[i]);  // hs[i] may be refilled
            compute_sync();          // the previous instance-step's readers of `part` are done
#pragma unroll
            for (int j = 0; j < NT; ++j) {
              const int col = nh * (ROWS / 2) + j * 8 + 2 * t;
              *reinterpret_cast<float2*>(part + (size_t)(kgrp * kLstmBT + g) * PLD + col) =
                  make_float2(fmaf(acl[j][0], 1.f / kLoScale, acc[j][0]), fmaf(acl[j][1], 1.f / kLoScale, acc[j][1]));
              *reinterpret_cast<float2*>(part + (size_t)(kgrp * kLstmBT + g + 8) * PLD + col) =
                  make_float2(fmaf(acl[j][2], 1.f / kLoScale, acc[j][2]), fmaf(acl[j][3], 1.f / kLoScale, acc[j][3]));
            }
            compute_sync();
            PF_T(c3);
            PF_ADD(pf_sync, c2, c3);
            if (act) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int r = q * US + pu;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) pre[q] += part[(size_t)(kk * kLstmBT + pb) * PLD + r];
              }
            }
            PF_T(c3b);
            PF_ADD(pf_epi, c3, c3b);
          }
          PF_T(c4);
          float ig = 0.f, fg = 0.f, gg = 0.f, og = 0.f;
          if (act) {
            ig = sigmoidf_(pre[0]); fg = sigmoidf_(pre[1]); gg = tanhf(pre[2]); og = sigmoidf_(pre[3]);
            c_reg[i] = fmaf(fg, c_reg[i], ig * gg);
            h_reg[i] = og * tanhf(c_reg[i]);
          }
          PF_T(c5);
          PF_ADD(pf_epi, c4, c5);
          // publish first (finished samples keep publishing their last state): the release that follows `written`
          // then does not have to wait for the bulkier stores of this step's outputs below
          if (pok[i]) {
            __half* row = reinterpret_cast<__half*>(hx[i] + (size_t)(k & 1) * kLstmBT * ld + (size_t)pb * ld);
            __half hi, lo;
            split_f16(h_reg[i], hi, lo);
            row[u0 + pu] = hi;
            row[Hp + u0 + pu] = lo;
          }
          if (k + 1 < tmx[i]) warp_arrive(&written[i]);
          if (act) {
            const size_t tb_ = (size_t)tt * a.B + b0[i] + pb;
            a.out[(tb_ * 2 + dir) * H + u0 + pu] = h_reg[i];
            if (a.gates) {
              float* gp = a.gates + (tb_ * 2 + dir) * 4 * H + u0 + pu;
              gp[0] = ig;
              gp[(size_t)H] = fg;
              gp[(size_t)2 * H] = gg;
              gp[(size_t)3 * H] = og;
            }
            if (a.cst) a.cst[(tb_ * 2 + dir) * H + u0 + pu] = c_reg[i];
          }
          PF_T(c6);
          PF_ADD(pf_pub, c5, c6);
        }
      }
#ifdef VOCR_LSTM_PROF
      if (blockIdx.x == 0 && blockIdx.y == 0 && (tid == 0 || tid == 255))
        printf("lstm fwd prof tid %d: total %lld  wait %lld  prod %lld  sync+part %lld  epi %lld  pub %lld (cycles), Tmax %d\n",
               tid, clock64() - pf_t0, pf_wait, pf_prod, pf_sync, pf_epi, pf_pub, a.Tmax);
#endif
    }
  }
}

// ================================================ backward ====================================================
// Per (instance, step): gate gradients of the slice's units (fp32) -> partial dh_{t-1}[16, all units] = da[16 x 64]
// . W_slice[64 x H] on the tensor cores -> reduce-scatter through L2.  The partial is stored consumer-major
// (px[parity][consumer slice][producer slice][16][US]) so that what a slice has to sum is ONE contiguous block: its
// communication warp pulls that block into shared memory with a bulk async copy as soon as the flag says every
// producer has published, and the compute warps sum it from shared memory in a fixed order (deterministic).
//
// Operands are FP16 (hi, lo * 2^11) pairs like the forward pass.  Gate gradients have no fixed range, so each sample
// row of da is scaled by a power of two (exact) that puts its largest magnitude near 2^14, and the product row is
// scaled back in fp32: the error is relative to the largest term of the row's dot products, as in fp32 itself.
constexpr uint32_t kBulkChunk = 32768u;  // bytes per bulk copy of the backward gather
constexpr int kBwdWtLd = 72;   // words per W^T row: 32 hi pairs | 32 lo pairs | pad (64-bit fragment loads conflict-free)

template <int NTHREADS>
__device__ __forceinline__ void load_wt_slice_f16(uint32_t* Wt, const float* __restrict__ whh_dir, int H, int Hp,
                                                  int US, int u0, int nu) {
  __half* Wh = reinterpret_cast<__half*>(Wt);
  for (int i = threadIdx.x; i < kLstmRows * Hp; i += NTHREADS) {
    const int r = i / Hp, n = i - r * Hp;
    const int g = r / US, u = r - g * US;
    float v = 0.f;
    if (g < 4 && u < nu && n < H) v = __ldg(whh_dir + ((size_t)g * H + u0 + u) * H + n);
    __half hi, lo;
    split_f16(v, hi, lo);
    Wh[(size_t)n * 2 * kBwdWtLd + r] = hi;
    Wh[(size_t)n * 2 * kBwdWtLd + kLstmRows + r] = lo;
  }
}

__global__ void __launch_bounds__(kLstmThreads, 1) bilstm_bwd_kernel(LstmArgs a) {
  extern __shared__ __align__(16) float lstm_smem[];
  const int H = a.H, Hp = a.Hp, US = a.US;
  const int blk = a.NSL * kLstmBT * US;     // floats one slice sums per (instance, step)
  uint32_t* Wt = reinterpret_cast<uint32_t*>(lstm_smem);               // [Hp][72]  W_slice^T, fp16 pairs
  float* pbuf = lstm_smem + (size_t)Hp * kBwdWtLd;                     // [NI][NSL][16][US] gathered partials
  float* dabuf = pbuf + (size_t)kLstmNI * blk;                         // [NI][16][80] gate gradients, row = sample
  uint64_t* bars = reinterpret_cast<uint64_t*>(dabuf + kLstmNI * kLstmBT * kLstmDaLd);
  uint64_t* ready = bars;                   // [NI] the gathered partials landed in pbuf[i]   (tx bytes)
  uint64_t* written = bars + kLstmNI;       // [NI] this slice's partial stored, pbuf[i] free (8 warp arrivals)
  const int slice = blockIdx.x;
  const int u0 = slice * US;
  const int nu = min(US, H - u0);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* dout = a.xproj;  // [T,B,2H]
  float* dgates = a.out;        // [T,B,2,4H]
  if (tid == 0) {
    for (int i = 0; i < kLstmNI; ++i) {
      mbar_init(&ready[i], 1);
      mbar_init(&written[i], 8);
    }
    mbar_fence_init();
  }
  unsigned n_ready[kLstmNI] = {0, 0}, n_written[kLstmNI] = {0, 0};

  for (int grp = blockIdx.y; grp < a.n_groups; grp += gridDim.y) {
    const int dir = grp / a.gpd, pair = grp - dir * a.gpd;
    const int ni = min(kLstmNI, a.NBT - kLstmNI * pair);
    __syncthreads();
    load_wt_slice_f16<kLstmThreads>(Wt, a.whh + (size_t)dir * 4 * H * H, H, Hp, US, u0, nu);
    for (int i = tid; i < kLstmNI * kLstmBT * kLstmDaLd; i += kLstmThreads) dabuf[i] = 0.f;
    __syncthreads();

    if (warp >= 8) {
      // ------------------------------ communication warp of instance i ------------------------------
      // (Per-producer flags with one 1-KB bulk copy per producer as it arrives were tried and are much slower,
      // 9.0 vs 6.1 us/step: a bulk copy costs ~120 cycles of serial issue in the TMA unit whatever its size, so the
      // exchange uses as few, as large copies as it can.)
      const int i = warp - 8;
      if (i < ni && lane == 0) {
        const int tile = lstm_fold(kLstmNI * pair + i, a.NBT);
        const int inst = dir * a.NBT + tile;
        unsigned* flag = a.flags + inst;
        const int tm = lstm_tile_tmax(a.lens, tile * kLstmBT, a.B, a.Tmax);
        const float* px = a.xchg + (size_t)inst * 2 * a.NSL * blk;
        float* dst = pbuf + (size_t)i * blk;
        unsigned round = 0;
        for (int k = tm - 1; k >= 1; --k) {
          mbar_wait_or_trap(&written[i], (n_written[i] & 1u));
          ++n_written[i];
          red_release(flag);  // release: cumulative over the compute warps' stores ordered by `written`
          const float* src = px + ((size_t)(round & 1) * a.NSL + slice) * blk;
          ++round;
          poll_flag(flag, (unsigned)a.NSL * round);
          asm volatile("fence.proxy.async;" ::: "memory");
          const uint32_t bytes = (uint32_t)blk * 4u;  // multiple of 64
          mbar_arrive_expect_tx(&ready[i], bytes);
          for (uint32_t off = 0; off < bytes; off += kBulkChunk)
            bulk_g2s(reinterpret_cast<char*>(dst) + off, reinterpret_cast<const char*>(src) + off,
                     min(kBulkChunk, bytes - off), &ready[i]);
        }
      }
      __syncwarp();  // reconverge before the block-wide barrier at the top of the next group
    } else {
      // ------------------------------------- compute warps -------------------------------------
      const int g = lane >> 2, t = lane & 3;
      const int pb = tid / US, pu = tid - pb * US;
      const int ntiles = Hp / 8;  // 8-column tiles of the partial product; warp w owns tiles w, w+8, ...
      const float inv_us = 1.f / (float)US;
      int b0[kLstmNI], plen[kLstmNI], tmx[kLstmNI];
      float* px[kLstmNI];
      float dc_reg[kLstmNI], dh_reg[kLstmNI];
      float da_max = 0.f;  // largest gate-gradient magnitude this thread wrote (operand bound of the GEMMs that follow)
      bool pok[kLstmNI];
      unsigned round[kLstmNI];
#pragma unroll
      for (int i = 0; i < kLstmNI; ++i) {
        const int tile = lstm_fold(kLstmNI * pair + i, a.NBT);
        const int inst = dir * a.NBT + tile;
        b0[i] = tile * kLstmBT;
        tmx[i] = (i < ni) ? lstm_tile_tmax(a.lens, b0[i], a.B, a.Tmax) : 0;
        const int nb = min(kLstmBT, a.B - b0[i]);
        pok[i] = (i < ni) && pb < kLstmBT && pb < nb && pu < nu;
        plen[i] = pok[i] ? min(a.lens[b0[i] + pb], a.Tmax) : 0;
        px[i] = a.xchg + (size_t)inst * 2 * a.NSL * blk;
        dc_reg[i] = 0.f;
        dh_reg[i] = 0.f;
        round[i] = 0;
      }
#ifdef VOCR_LSTM_PROF
      long long pf_wait = 0, pf_red = 0, pf_grad = 0, pf_frag = 0, pf_prod = 0, pf_store = 0, pf_t0 = clock64();
#endif
      for (int k = a.Tmax - 1; k >= 0; --k) {
#pragma unroll
        for (int i = 0; i < kLstmNI; ++i) {
          if (i >= ni || k >= tmx[i]) continue;
          float* das = dabuf + (size_t)i * kLstmBT * kLstmDaLd;
          // saved activations of this step: independent of the recurrence, loads issued before the wait
          const bool act = pok[i] && k < plen[i];
          float ig = 0.f, fg = 0.f, gg = 0.f, og = 0.f, c = 0.f, c_prev = 0.f, dy = 0.f;
          size_t tb_ = 0;
          if (act) {
            const int tq = (dir == 0) ? k : plen[i] - 1 - k;
            tb_ = (size_t)tq * a.B + b0[i] + pb;
            const float* gp = a.gates + (tb_ * 2 + dir) * 4 * H + u0 + pu;
            ig = gp[0]; fg = gp[(size_t)H]; gg = gp[(size_t)2 * H]; og = gp[(size_t)3 * H];
            c = a.cst[(tb_ * 2 + dir) * H + u0 + pu];
            if (k > 0) {
              const int tp = (dir == 0) ? tq - 1 : tq + 1;
              c_prev = a.cst[(((size_t)tp * a.B + b0[i] + pb) * 2 + dir) * H + u0 + pu];
            }
            dy = dout[(tb_ * 2 + dir) * H + u0 + pu];
          }
          // 0. fold in the partial dh produced by the previous (later-in-time) step of this instance
          PF_T(c0);
          if (k < tmx[i] - 1) {
            mbar_wait_or_trap(&ready[i], (n_ready[i] & 1u));
            ++n_ready[i];
            PF_T(c1);
            PF_ADD(pf_wait, c0, c1);
            if (pok[i]) {
              const float* q = pbuf + (size_t)i * blk + (size_t)pb * US + pu;
              const int sstride = kLstmBT * US;
              float s = 0.f;  // fixed summation order; loads batched by 8 ahead of the dependent adds
              int sl = 0;
              for (; sl + 8 <= a.NSL; sl += 8) {
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = q[(size_t)(sl + j) * sstride];
#pragma unroll
                for (int j = 0; j < 8; ++j) s += v[j];
              }
              for (; sl < a.NSL; ++sl) s += q[(size_t)sl * sstride];
              dh_reg[i] += s;
            }
            PF_T(c2);
            PF_ADD(pf_red, c1, c2);
          }
          PF_T(c3);
          // 1. gate gradients of this slice's units at step k
          float da[4] = {0.f, 0.f, 0.f, 0.f};
          if (act) {
            const float dh = dy + dh_reg[i];
            const float tc = tanhf(c);
            const float dc = fmaf(dh * og, 1.f - tc * tc, dc_reg[i]);
            da[0] = dc * gg * ig * (1.f - ig);
            da[1] = dc * c_prev * fg * (1.f - fg);
            da[2] = dc * ig * (1.f - gg * gg);
            da[3] = dh * tc * og * (1.f - og);
            dc_reg[i] = dc * fg;
            float* dg = dgates + (tb_ * 2 + dir) * 4 * H + u0 + pu;
            dg[0] = da[0];
            dg[(size_t)H] = da[1];
            dg[(size_t)2 * H] = da[2];
            dg[(size_t)3 * H] = da[3];
            da_max = fmaxf(da_max, fmaxf(fmaxf(fabsf(da[0]), fabsf(da[1])), fmaxf(fabsf(da[2]), fabsf(da[3]))));
            dh_reg[i] = 0.f;  // consumed; the next reduce-scatter refills it
          }
          if (k == 0) continue;
          if (pok[i]) {
#pragma unroll
            for (int q = 0; q < 4; ++q) das[(size_t)pb * kLstmDaLd + q * US + pu] = da[q];
          }
          PF_T(c4);
          PF_ADD(pf_grad, c3, c4);
          compute_sync();
          // 2. partial dh_{k-1}[16, :] = das[16, 0:64] . Ws[0:64, :]  (tensor cores) -> px[round&1][consumer][slice]
          //    A fragments: rows g and g+8, lane t owns k = 16*ks + 4t .. +3 of every k16 step (B uses the same map)
          float4 x0[4], x1[4];
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            x0[ks] = *reinterpret_cast<const float4*>(das + (size_t)g * kLstmDaLd + ks * 16 + 4 * t);
            x1[ks] = *reinterpret_cast<const float4*>(das + (size_t)(g + 8) * kLstmDaLd + ks * 16 + 4 * t);
          }
          // (das[i] is rewritten only after ready[i] of the next step, i.e. after every warp's `written` arrival below)
          float m0 = 0.f, m1 = 0.f;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            m0 = fmaxf(m0, fmaxf(fmaxf(fabsf(x0[ks].x), fabsf(x0[ks].y)), fmaxf(fabsf(x0[ks].z), fabsf(x0[ks].w))));
            m1 = fmaxf(m1, fmaxf(fmaxf(fabsf(x1[ks].x), fabsf(x1[ks].y)), fmaxf(fabsf(x1[ks].z), fabsf(x1[ks].w))));
          }
          m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
          m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
          m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
          m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
          // power-of-two row scales: largest magnitude -> [2^14, 2^15); exponent clamped so 1/scale stays normal
          const int e0 = min(240, max(16, 268 - (int)((__float_as_uint(m0) >> 23) & 0xffu)));
          const int e1 = min(240, max(16, 268 - (int)((__float_as_uint(m1) >> 23) & 0xffu)));
          const float s0 = __uint_as_float((unsigned)e0 << 23), s1 = __uint_as_float((unsigned)e1 << 23);
          const float r0 = __uint_as_float((unsigned)(254 - e0) << 23), r1 = __uint_as_float((unsigned)(254 - e1) << 23);
          uint32_t ah[4][4], al[4][4];
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            auto pack = [](float p, float q, float sc, uint32_t& hi, uint32_t& lo) {
              const float ps = p * sc, qs = q * sc;
              const __half2 h = __floats2half2_rn(ps, qs);
              const float2 hf = __half22float2(h);
              const __half2 l = __floats2half2_rn((ps - hf.x) * kLoScale, (qs - hf.y) * kLoScale);
              hi = *reinterpret_cast<const uint32_t*>(&h);
              lo = *reinterpret_cast<const uint32_t*>(&l);
            };
            pack(x0[ks].x, x0[ks].y, s0, ah[ks][0], al[ks][0]);
            pack(x1[ks].x, x1[ks].y, s1, ah[ks][1], al[ks][1]);
            pack(x0[ks].z, x0[ks].w, s0, ah[ks][2], al[ks][2]);
            pack(x1[ks].z, x1[ks].w, s1, ah[ks][3], al[ks][3]);
          }
          PF_T(c5);
          PF_ADD(pf_frag, c4, c5);
          float* pdst = px[i] + (size_t)(round[i] & 1) * a.NSL * blk + (size_t)slice * kLstmBT * US;
          // four column tiles at a time: 12 independent accumulator chains keep the tensor pipe busy.
          // (A "fragment order" block layout with one 16-byte store per lane measured SLOWER than these 8-byte
          // stores: 7.25 vs 6.09 us/step at H=512, B=64.)
          for (int base = 0; base < ntiles; base += 32) {
            int ntj[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
              ntj[j] = base + warp + 8 * j;
            // branch-free: a tile index past the end (small H) is redirected to tile 0 and its result dropped below
            const uint32_t* wB[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
              wB[j] = Wt + (size_t)((ntj[j] < ntiles ? ntj[j] : 0) * 8 + g) * kBwdWtLd + 2 * t;
            float acc[4][4], acl[4][4], acm[4][4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
              for (int q = 0; q < 4; ++q) acc[j][q] = acl[j][q] = acm[j][q] = 0.f;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint2 wh = *reinterpret_cast<const uint2*>(wB[j] + ks * 8);
                const uint2 wl = *reinterpret_cast<const uint2*>(wB[j] + kLstmRows / 2 + ks * 8);
                const uint32_t bh[2] = {wh.x, wh.y}, bl[2] = {wl.x, wl.y};
                mma_f16(acl[j], al[ks], bh);
                mma_f16(acm[j], ah[ks], bl);
                mma_f16(acc[j], ah[ks], bh);
              }
            }
            PF_T(c5b);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int nt = ntj[j];
              if (nt >= ntiles) continue;
              const float v00 = fmaf(acl[j][0] + acm[j][0], 1.f / kLoScale, acc[j][0]) * r0;
              const float v01 = fmaf(acl[j][1] + acm[j][1], 1.f / kLoScale, acc[j][1]) * r0;
              const float v10 = fmaf(acl[j][2] + acm[j][2], 1.f / kLoScale, acc[j][2]) * r1;
              const float v11 = fmaf(acl[j][3] + acm[j][3], 1.f / kLoScale, acc[j][3]) * r1;
              // columns col, col+1 (units of the layer) -> consumer slice cs, unit offset cu inside it
              const int col = nt * 8 + 2 * t;
              const int cs = (int)(((float)col + 0.5f) * inv_us), cu = col - cs * US;
              if (cs < a.NSL) {
                float* d0 = pdst + (size_t)cs * blk + (size_t)g * US + cu;
                float* d1 = d0 + (size_t)8 * US;
                if ((US & 1) == 0) {
                  *reinterpret_cast<float2*>(d0) = make_float2(v00, v01);
                  *reinterpret_cast<float2*>(d1) = make_float2(v10, v11);
                } else {
                  d0[0] = v00;
                  d1[0] = v10;
                  // the odd column may belong to the next consumer slice
                  const int cs1 = (cu + 1 == US) ? cs + 1 : cs, cu1 = (cu + 1 == US) ? 0 : cu + 1;
                  if (cs1 < a.NSL) {
                    pdst[(size_t)cs1 * blk + (size_t)g * US + cu1] = v01;
                    pdst[(size_t)cs1 * blk + (size_t)(g + 8) * US + cu1] = v11;
                  }
                }
              }
            }
            PF_T(c5c);
            PF_ADD(pf_store, c5b, c5c);
          }
          // 3. publish through the communication warp; the reduce-scatter happens at the top of the next step
          warp_arrive(&written[i]);
          ++round[i];
          PF_T(c6);
          PF_ADD(pf_prod, c5, c6);
        }
      }
      if (a.absmax) {  // non-negative floats order like their bit patterns; NaN (0x7fc00000) wins, so it is not hidden
        da_max = warp_max(da_max);
        if (lane == 0) atomicMax(reinterpret_cast<unsigned*>(a.absmax), __float_as_uint(da_max));
      }
#ifdef VOCR_LSTM_PROF
      if (blockIdx.x == 0 && blockIdx.y == 0 && (tid == 0 || tid == 255))
        printf("lstm bwd prof tid %d: total %lld  wait %lld  reduce %lld  grad %lld  sync+frag %lld  prod+store %lld (store %lld) (cycles), Tmax %d\n",
               tid, clock64() - pf_t0, pf_wait, pf_red, pf_grad, pf_frag, pf_prod, pf_store, a.Tmax);
#endif
    }
  }
}

}  // namespace vocr

using namespace vocr;

// variant 0: 64 gate rows per slice (32 slices at H=512), 2 interleaved tiles  - backward, and forward of small batches
// variant 1: 32 gate rows per slice (64 slices at H=512), 4 interleaved tiles  - forward when the batch has >= 3 tiles
static int lstm_geometry(int B, int H, LstmArgs* a, size_t* smem, int* grid_y, bool bwd, int* variant) {
  if (H < 1 || H > 32 * kLstmMaxUS) return VOCR_INVALID_VALUE;
  a->NBT = ceil_div(B, kLstmBT);
  *variant = 0;
  if (const char* e = getenv("VOCR_LSTM_VARIANT")) *variant = (!bwd && atoi(e) == 1) ? 1 : 0;
  const int slices = *variant ? 64 : 32, ni = *variant ? 4 : kLstmNI, rows = *variant ? 32 : kLstmRows;
  a->US = ceil_div(H, slices);
  a->NSL = ceil_div(H, a->US);
  a->Hp = ceil_div(H, 64) * 64;
  a->gpd = ceil_div(a->NBT, ni);
  a->n_groups = 2 * a->gpd;
  const size_t ld = a->Hp + 8;
  const size_t blk = (size_t)a->NSL * kLstmBT * a->US;
  *smem = sizeof(float) * (bwd ? (size_t)a->Hp * kBwdWtLd + kLstmNI * blk + (size_t)kLstmNI * kLstmBT * kLstmDaLd
                               : rows * ld + (size_t)ni * kLstmBT * ld + 4 * kLstmBT * (rows + 8)) + 128;
  *grid_y = max(1, min(a->n_groups, kNumSMs / a->NSL));
  return VOCR_OK;
}

static size_t lstm_flag_bytes(const LstmArgs& a, bool bwd) {
  (void)bwd;
  return (sizeof(unsigned) * 2 * a.NBT + 255) & ~size_t(255);
}
static size_t lstm_xchg_bytes(const LstmArgs& a, bool bwd) {
  if (bwd) return sizeof(float) * (size_t)(2 * a.NBT) * 2 * a.NSL * ((size_t)a.NSL * kLstmBT * a.US);
  return sizeof(float) * (size_t)(2 * a.NBT) * 2 * kLstmBT * (a.Hp + 8);
}

// lstm_tc.cu: the tcgen05 forward kernel (all samples of a 64-sample group per CTA, sentinel-polled exchange)
size_t lstm_tc_fwd_workspace_bytes(int T, int B, int H);
int lstm_tc_fwd_launch(const float* xproj, const float* whh, const int32_t* lens, float* out, float* gates, float* cst,
                       int T, int B, int H, int Tmax, void* workspace, size_t workspace_bytes, cudaStream_t stream);
bool lstm_tc_enabled();
size_t lstm_tc_bwd_workspace_bytes(int T, int B, int H);
int lstm_tc_bwd_launch(const float* dout, const float* whh, const int32_t* lens, const float* gates, const float* cst,
                       float* dgates, float* absmax, int T, int B, int H, int Tmax, void* workspace,
                       size_t workspace_bytes, cudaStream_t stream);

extern "C" size_t vocr_bilstm_workspace_size(int T, int B, int H, int backward) {
  LstmArgs a;
  size_t smem;
  int gy, variant;
  if (T < 0 || lstm_geometry(B, H, &a, &smem, &gy, backward != 0, &variant) != VOCR_OK) return 0;
  size_t need = 256 + lstm_flag_bytes(a, backward != 0) + lstm_xchg_bytes(a, backward != 0);
  need = std::max(need, backward ? lstm_tc_bwd_workspace_bytes(T, B, H) : lstm_tc_fwd_workspace_bytes(T, B, H));
  return need;
}

static int lstm_launch(bool bwd, LstmArgs a, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  size_t smem;
  int gy, variant;
  int st = lstm_geometry(a.B, a.H, &a, &smem, &gy, bwd, &variant);
  if (st != VOCR_OK) return st;
  uintptr_t w = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255);
  const size_t flag_bytes = lstm_flag_bytes(a, bwd);
  const size_t xchg = lstm_xchg_bytes(a, bwd);
  if ((w - reinterpret_cast<uintptr_t>(workspace)) + flag_bytes + xchg > workspace_bytes) return VOCR_INVALID_VALUE;
  a.flags = reinterpret_cast<unsigned*>(w);
  a.xchg = reinterpret_cast<float*>(w + flag_bytes);
  // flags start at 0; the exchange buffer is zeroed so padded rows / columns never inject NaNs into the products
  if (cudaMemsetAsync(a.flags, 0, flag_bytes + xchg, stream) != cudaSuccess) return VOCR_MEMOPS_FAILED;
  const void* fn = bwd ? (const void*)bilstm_bwd_kernel
                       : (variant ? (const void*)bilstm_fwd_kernel<32, 4> : (const void*)bilstm_fwd_kernel<64, 2>);
  const int threads = bwd ? kLstmThreads : (variant ? 256 + 32 * 4 : 256 + 32 * 2);
  static DeviceLatch latch[3];  // cudaFuncSetAttribute is per device
  DeviceLatch& l = latch[bwd ? 0 : 1 + variant];
  if (l.need()) {
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return VOCR_EXECUTION_FAILED;
    l.set();
  }
  dim3 grid(a.NSL, gy);
  void* params[] = {&a};
  // cooperative launch: the runtime refuses the launch unless every CTA can be co-resident, which the flag
  // protocol relies on
  if (cudaLaunchCooperativeKernel(fn, grid, dim3(threads), params, smem, stream) != cudaSuccess)
    return VOCR_EXECUTION_FAILED;
  return VOCR_OK;
}

// xproj [T,B,2,4H] (= x W_ih^T + b_ih + b_hh for both directions), whh [2,4H,H], lens [B] (device), out [T,B,2H].
// gates / cst may be NULL (inference).  Tmax = max(lens) (host knows it: lens are computed on the host).
extern "C" int vocr_bilstm_fwd_f32(const float* xproj, const float* whh, const int32_t* lens, float* out,
                                   float* gates, float* cst, int T, int B, int H, int Tmax, void* workspace,
                                   size_t workspace_bytes, vocr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VOCR_REQUIRE(T >= 0 && B >= 0 && H >= 1 && Tmax >= 0 && Tmax <= T);
  if (T == 0 || B == 0) return VOCR_OK;
  VOCR_REQUIRE(xproj && whh && lens && out && workspace);
  if (cudaMemsetAsync(out, 0, sizeof(float) * (size_t)T * B * 2 * H, stream) != cudaSuccess)
    return VOCR_MEMOPS_FAILED;
  if (Tmax == 0) return VOCR_OK;
  LstmArgs a{};
  a.xproj = xproj; a.whh = whh; a.lens = lens; a.out = out; a.gates = gates; a.cst = cst;
  a.T = T; a.B = B; a.H = H; a.Tmax = Tmax;
  if (lstm_tc_enabled()) {  // cluster-resident tcgen05 kernel (lstm_tc.cu); -1 = shape it does not cover
    const int st = lstm_tc_fwd_launch(xproj, whh, lens, out, gates, cst, T, B, H, Tmax, workspace, workspace_bytes, stream);
    if (st != -1) return st;
  }
  return lstm_launch(false, a, workspace, workspace_bytes, stream);
}

// dout [T,B,2H], gates/cst from the forward pass -> dgates [T,B,2,4H] (gradient w.r.t. xproj; zero beyond lens).
extern "C" int vocr_bilstm_bwd_f32(const float* dout, const float* whh, const int32_t* lens, const float* gates,
                                   const float* cst, float* dgates, float* dgates_absmax, int T, int B, int H, int Tmax,
                                   void* workspace, size_t workspace_bytes, vocr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VOCR_REQUIRE(T >= 0 && B >= 0 && H >= 1 && Tmax >= 0 && Tmax <= T);
  if (dgates_absmax && cudaMemsetAsync(dgates_absmax, 0, sizeof(float), stream) != cudaSuccess) return VOCR_MEMOPS_FAILED;
  if (T == 0 || B == 0) return VOCR_OK;
  VOCR_REQUIRE(dout && whh && lens && gates && cst && dgates && workspace);
  if (cudaMemsetAsync(dgates, 0, sizeof(float) * (size_t)T * B * 8 * H, stream) != cudaSuccess)
    return VOCR_MEMOPS_FAILED;
  if (Tmax == 0) return VOCR_OK;
  LstmArgs a{};
  a.xproj = dout; a.whh = whh; a.lens = lens; a.out = dgates;
  a.gates = const_cast<float*>(gates); a.cst = const_cast<float*>(cst);
  a.absmax = dgates_absmax;
  a.T = T; a.B = B; a.H = H; a.Tmax = Tmax;
  if (lstm_tc_enabled()) {  // cluster-resident tcgen05 kernel (lstm_tc.cu); -1 = shape it does not cover
    const int st = lstm_tc_bwd_launch(dout, whh, lens, gates, cst, dgates, dgates_absmax, T, B, H, Tmax, workspace,
                                      workspace_bytes, stream);
    if (st != -1) return st;
  }
  return lstm_launch(true, a, workspace, workspace_bytes, stream);
}
