// Bidirectional LSTM layer recurrence on sm_100a, forward and backward, as PERSISTENT cooperative kernels
// (reference cnnlstm.py:148-149,285-290: nn.LSTM on a packed sequence -> cuDNN RNN).
//
// The input projections x_t W_ih^T + b_ih + b_hh of ALL timesteps and both directions are one tensor-core GEMM done
// beforehand (xproj [T,B,2,4H]); what is left is the strictly sequential part  h_{t-1} W_hh^T  + gate nonlinearities.
//
// Work decomposition.  An "instance" = (direction, tile of 16 samples); it is served by NSL CTAs ("slices"), each
// owning US hidden units = 4*US rows of W_hh, which stay RESIDENT IN SHARED MEMORY (fp32) for all timesteps
// (H=512: 32 slices x 64 rows x 512 = 129 KB each).  Per step a slice multiplies the 16 x H tile of h_{t-1} by its
// 64 rows on the tensor cores (mma.sync m16n8k8 TF32 with the 3xTF32 hi/lo split done in registers, so the product
// keeps fp32-level accuracy without doubling the resident weights; K split over 4 warp pairs and reduced through
// shared memory), applies the gates for its units, and publishes its 16 x US piece of h_t through a ping-pong buffer
// in global memory (L2); the slices of an instance meet at a monotonically increasing flag (release/acquire at gpu
// scope) - there is no grid-wide barrier.
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
#include "common.cuh"

namespace vocr {

constexpr int kLstmThreads = 320;     // 8 compute warps + one communication warp per interleaved instance
constexpr int kLstmBT = 16;        // samples per instance = MMA M
constexpr int kLstmNI = 2;         // instances interleaved per CTA
constexpr int kLstmMaxUS = 16;     // hidden units per slice
constexpr int kLstmRows = 64;      // 4 * kLstmMaxUS gate rows per slice (zero padded)
constexpr int kLstmPartLd = 72;    // row stride of the K-split partial sums [4][16][72]
constexpr int kLstmDaLd = kLstmRows + 4;

struct LstmArgs {
  const float* xproj;   // [T,B,2,4H]  (fwd)            | dout [T,B,2H] (bwd)
  const float* whh;     // [2,4H,H]
  const int32_t* lens;  // [B]
  float* out;           // [T,B,2H]    (fwd, pre-zeroed) | dgates [T,B,2,4H] (bwd, pre-zeroed)
  float* gates;         // [T,B,2,4H] activated i,f,g,o (fwd: written if non-null; bwd: read)
  float* cst;           // [T,B,2,H]  cell state        (fwd: written if non-null; bwd: read)
  float* xchg;          // fwd: [n_inst][2][16][Hp]      | bwd: [n_inst][2][NSL][16][Hp]
  unsigned* flags;      // [n_inst], zeroed before launch
  int T, B, H, Hp, US, NSL, Tmax, NBT, gpd, n_groups;  // gpd = instance pairs per direction
};

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release(unsigned* p) {
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// ---- tensor-core helpers: m16n8k8 TF32, operands split hi/lo in registers (3xTF32) ---------------------------------
// hi = x with the 13 low mantissa bits cleared (exactly a TF32 value; one LOP3), lo = x - hi (exact; one FADD).
// cvt.rna.tf32 would halve |lo| but expands to ~6 instructions here, and this split runs 12x per k-step on the
// critical path; with truncation the dropped lo*lo term is still only 2^-20 relative.
__device__ __forceinline__ void split2(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));  // the tensor core ignores the 13 low mantissa bits of lo
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// load this slice's W_hh rows into shared memory: Ws[r = g*US + u][k], zero padded to 64 rows x (Hp+4)
__device__ __forceinline__ void load_w_slice(float* Ws, const float* __restrict__ whh_dir, int H, int Hp, int US,
                                             int u0, int nu) {
  const int ld = Hp + 4;
  for (int i = threadIdx.x; i < kLstmRows * ld; i += kLstmThreads) {
    const int r = i / ld, k = i - r * ld;
    const int g = r / US, u = r - g * US;
    float v = 0.f;
    if (g < 4 && u < nu && k < H) v = __ldg(whh_dir + ((size_t)g * H + u0 + u) * H + k);
    Ws[i] = v;
  }
}

// A CTA serves the sample tiles `pair` and NBT-1-pair: batches arrive sorted by width (SortByWidthCollater), so this
// pairs the longest tile with the shortest one and every CTA group gets about the same number of instance-steps.
__device__ __forceinline__ int lstm_tile(int pair, int i, int NBT) { return i == 0 ? pair : NBT - 1 - pair; }
__device__ __forceinline__ int lstm_ni(int pair, int NBT) { return (NBT - 1 - pair > pair) ? 2 : 1; }
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
__global__ void __launch_bounds__(kLstmThreads, 1) bilstm_fwd_kernel(LstmArgs a) {
  extern __shared__ __align__(16) float lstm_smem[];
  const int H = a.H, Hp = a.Hp, US = a.US, ld = Hp + 4;
  float* Ws = lstm_smem;                                   // [64][ld]
  float* hbuf = Ws + kLstmRows * ld;                       // [NI][16][ld]
  float* part = hbuf + (size_t)kLstmNI * kLstmBT * ld;     // [4][16][72]
  uint64_t* bars = reinterpret_cast<uint64_t*>(part + 4 * kLstmBT * kLstmPartLd);
  uint64_t* full = bars;                  // [NI] h tile landed           (tx bytes)
  uint64_t* empty = bars + kLstmNI;       // [NI] product done with hs[i]  (8 warp arrivals)
  uint64_t* written = bars + 2 * kLstmNI; // [NI] h_t piece stored         (8 warp arrivals)
  const int slice = blockIdx.x;
  const int u0 = slice * US;
  const int nu = min(US, H - u0);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < kLstmNI; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 8);
      mbar_init(&written[i], 8);
    }
    mbar_fence_init();
  }
  // running phase counters (barriers are used across groups without re-initialisation)
  unsigned n_full[kLstmNI] = {0, 0}, n_empty[kLstmNI] = {0, 0}, n_written[kLstmNI] = {0, 0};

  for (int grp = blockIdx.y; grp < a.n_groups; grp += gridDim.y) {
    const int dir = grp / a.gpd, pair = grp - dir * a.gpd;
    const int ni = lstm_ni(pair, a.NBT);
    __syncthreads();
    load_w_slice(Ws, a.whh + (size_t)dir * 4 * H * H, H, Hp, US, u0, nu);
    for (int i = tid; i < kLstmNI * kLstmBT * ld; i += kLstmThreads) hbuf[i] = 0.f;  // padding columns stay zero
    __syncthreads();

    if (warp >= 8) {
      // ------------------------------ communication warp of instance i ------------------------------
      const int i = warp - 8;
      if (i < ni) {
        const int tile = lstm_tile(pair, i, a.NBT);
        const int inst = dir * a.NBT + tile;
        const int tm = lstm_tile_tmax(a.lens, tile * kLstmBT, a.B, a.Tmax);
        float* hx = a.xchg + (size_t)inst * 2 * kLstmBT * Hp;
        unsigned* flag = a.flags + inst;
        float* hs = hbuf + (size_t)i * kLstmBT * ld;
        for (int k = 0; k < tm; ++k) {
          if (k > 0) {
            if (k > 1) {  // hs[i] is free once the product of step k-1 has consumed it
              mbar_wait_or_trap(&empty[i], (n_empty[i] & 1u));
              ++n_empty[i];
            }
            if (lane == 0) {
              poll_flag(flag, (unsigned)(a.NSL * k));
              asm volatile("fence.proxy.async;" ::: "memory");
              mbar_arrive_expect_tx(&full[i], (uint32_t)(kLstmBT * Hp * 4));
            }
            __syncwarp();
            if (lane < kLstmBT) {
              asm volatile("fence.proxy.async;" ::: "memory");
              bulk_g2s(hs + (size_t)lane * ld, hx + (size_t)((k - 1) & 1) * kLstmBT * Hp + (size_t)lane * Hp,
                       (uint32_t)(Hp * 4), &full[i]);
            }
          }
          if (k + 1 < tm) {
            mbar_wait_or_trap(&written[i], (n_written[i] & 1u));
            ++n_written[i];
            if (lane == 0) {
              __threadfence();
              red_release(flag);
            }
          }
        }
        if (tm > 1) {  // drain: the last product's release of hs[i] (keeps the phase counters in step)
          mbar_wait_or_trap(&empty[i], (n_empty[i] & 1u));
          ++n_empty[i];
        }
      }
    } else {
      // ------------------------------------- compute warps -------------------------------------
      const int g = lane >> 2, t = lane & 3;          // mma fragment coordinates
      const int kgrp = warp >> 1, nh = warp & 1;      // K quarter, half of the 64 gate rows
      const int kq = Hp / 4;                          // K range of a quarter (Hp % 32 == 0 -> multiple of 8)
      const int pb = tid / US, pu = tid - pb * US;    // gate epilogue: one (sample, unit) pair per thread, instance
      int b0[kLstmNI], plen[kLstmNI], tmx[kLstmNI];
      float* hx[kLstmNI];
      float c_reg[kLstmNI], h_reg[kLstmNI];
      bool pok[kLstmNI];
#pragma unroll
      for (int i = 0; i < kLstmNI; ++i) {
        const int tile = lstm_tile(pair, i, a.NBT);
        const int inst = dir * a.NBT + tile;
        b0[i] = tile * kLstmBT;
        tmx[i] = (i < ni) ? lstm_tile_tmax(a.lens, b0[i], a.B, a.Tmax) : 0;
        const int nb = min(kLstmBT, a.B - b0[i]);
        pok[i] = (i < ni) && pb < kLstmBT && pb < nb && pu < nu;
        plen[i] = pok[i] ? min(a.lens[b0[i] + pb], a.Tmax) : 0;
        hx[i] = a.xchg + (size_t)inst * 2 * kLstmBT * Hp;
        c_reg[i] = 0.f;
        h_reg[i] = 0.f;
      }
      for (int k = 0; k < a.Tmax; ++k) {
#pragma unroll
        for (int i = 0; i < kLstmNI; ++i) {
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
          if (k > 0) {
            mbar_wait_or_trap(&full[i], (n_full[i] & 1u));
            ++n_full[i];
            // D[16 x 32 rows of this warp] += h[16 x kq] * W^T, 3xTF32
            float acc[4][4], acl[4][4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
              for (int q = 0; q < 4; ++q) acc[j][q] = acl[j][q] = 0.f;
            const float* hA = hs + kgrp * kq;
            const float* wB = Ws + (size_t)(nh * 32) * ld + kgrp * kq;
#pragma unroll 4
            for (int k0 = 0; k0 < kq; k0 += 8) {
              uint32_t ah[4], al[4];
              split2(hA[(size_t)g * ld + k0 + t], ah[0], al[0]);
              split2(hA[(size_t)(g + 8) * ld + k0 + t], ah[1], al[1]);
              split2(hA[(size_t)g * ld + k0 + t + 4], ah[2], al[2]);
              split2(hA[(size_t)(g + 8) * ld + k0 + t + 4], ah[3], al[3]);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint32_t bh[2], bl[2];
                split2(wB[(size_t)(j * 8 + g) * ld + k0 + t], bh[0], bl[0]);
                split2(wB[(size_t)(j * 8 + g) * ld + k0 + t + 4], bh[1], bl[1]);
                mma_tf32(acl[j], al, bh);
                mma_tf32(acl[j], ah, bl);
                mma_tf32(acc[j], ah, bh);
              }
            }
            warp_arrive(&empty[i]);  // hs[i] may be refilled
            compute_sync();          // the previous instance-step's readers of `part` are done
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int col = nh * 32 + j * 8 + 2 * t;
              *reinterpret_cast<float2*>(part + (size_t)(kgrp * kLstmBT + g) * kLstmPartLd + col) =
                  make_float2(acc[j][0] + acl[j][0], acc[j][1] + acl[j][1]);
              *reinterpret_cast<float2*>(part + (size_t)(kgrp * kLstmBT + g + 8) * kLstmPartLd + col) =
                  make_float2(acc[j][2] + acl[j][2], acc[j][3] + acl[j][3]);
            }
            compute_sync();
            if (act) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int r = q * US + pu;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) pre[q] += part[(size_t)(kk * kLstmBT + pb) * kLstmPartLd + r];
              }
            }
          }
          if (act) {
            const float ig = sigmoidf_(pre[0]), fg = sigmoidf_(pre[1]), gg = tanhf(pre[2]), og = sigmoidf_(pre[3]);
            const float c = fmaf(fg, c_reg[i], ig * gg);
            const float h = og * tanhf(c);
            c_reg[i] = c;
            h_reg[i] = h;
            const size_t tb_ = (size_t)tt * a.B + b0[i] + pb;
            a.out[(tb_ * 2 + dir) * H + u0 + pu] = h;
            if (a.gates) {
              float* gp = a.gates + (tb_ * 2 + dir) * 4 * H + u0 + pu;
              gp[0] = ig;
              gp[(size_t)H] = fg;
              gp[(size_t)2 * H] = gg;
              gp[(size_t)3 * H] = og;
            }
            if (a.cst) a.cst[(tb_ * 2 + dir) * H + u0 + pu] = c;
          }
          // finished samples keep publishing their last state
          if (pok[i]) hx[i][(size_t)(k & 1) * kLstmBT * Hp + (size_t)pb * Hp + u0 + pu] = h_reg[i];
          if (k + 1 < tmx[i]) warp_arrive(&written[i]);
        }
      }
    }
  }
}

// ================================================ backward ====================================================
__global__ void __launch_bounds__(kLstmThreads, 1) bilstm_bwd_kernel(LstmArgs a) {
  extern __shared__ __align__(16) float lstm_smem[];
  const int H = a.H, Hp = a.Hp, US = a.US, ld = Hp + 4;
  float* Ws = lstm_smem;                    // [64][ld]
  float* dabuf = Ws + kLstmRows * ld;       // [NI][16][68] gate gradients of this slice's units, row = sample
  uint64_t* bars = reinterpret_cast<uint64_t*>(dabuf + kLstmNI * kLstmBT * kLstmDaLd);
  uint64_t* ready = bars;                   // [NI] all slices published their partial dh  (1 arrival, comm warp)
  uint64_t* written = bars + kLstmNI;       // [NI] this slice's partial stored            (8 warp arrivals)
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
    const int ni = lstm_ni(pair, a.NBT);
    __syncthreads();
    load_w_slice(Ws, a.whh + (size_t)dir * 4 * H * H, H, Hp, US, u0, nu);
    for (int i = tid; i < kLstmNI * kLstmBT * kLstmDaLd; i += kLstmThreads) dabuf[i] = 0.f;
    __syncthreads();

    if (warp >= 8) {
      // ------------------------------ communication warp of instance i ------------------------------
      const int i = warp - 8;
      if (i < ni) {
        const int tile = lstm_tile(pair, i, a.NBT);
        unsigned* flag = a.flags + dir * a.NBT + tile;
        const int tm = lstm_tile_tmax(a.lens, tile * kLstmBT, a.B, a.Tmax);
        unsigned round = 0;
        for (int k = tm - 1; k >= 1; --k) {
          mbar_wait_or_trap(&written[i], (n_written[i] & 1u));
          ++n_written[i];
          if (lane == 0) {
            __threadfence();
            red_release(flag);
            ++round;
            poll_flag(flag, (unsigned)a.NSL * round);
            mbar_arrive1(&ready[i]);
          }
          round = __shfl_sync(0xffffffffu, round, 0);
        }
      }
    } else {
      // ------------------------------------- compute warps -------------------------------------
      const int g = lane >> 2, t = lane & 3;
      const int pb = tid / US, pu = tid - pb * US;
      const int ntiles = Hp / 8;  // 8-column tiles of the partial product; warp w owns tiles w, w+8, ...
      int b0[kLstmNI], plen[kLstmNI], tmx[kLstmNI];
      float* px[kLstmNI];
      float dc_reg[kLstmNI], dh_reg[kLstmNI];
      bool pok[kLstmNI];
      unsigned round[kLstmNI];
#pragma unroll
      for (int i = 0; i < kLstmNI; ++i) {
        const int tile = lstm_tile(pair, i, a.NBT);
        const int inst = dir * a.NBT + tile;
        b0[i] = tile * kLstmBT;
        tmx[i] = (i < ni) ? lstm_tile_tmax(a.lens, b0[i], a.B, a.Tmax) : 0;
        const int nb = min(kLstmBT, a.B - b0[i]);
        pok[i] = (i < ni) && pb < kLstmBT && pb < nb && pu < nu;
        plen[i] = pok[i] ? min(a.lens[b0[i] + pb], a.Tmax) : 0;
        px[i] = a.xchg + (size_t)inst * 2 * a.NSL * kLstmBT * Hp;
        dc_reg[i] = 0.f;
        dh_reg[i] = 0.f;
        round[i] = 0;
      }
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
          if (k < tmx[i] - 1) {
            mbar_wait_or_trap(&ready[i], (n_ready[i] & 1u));
            ++n_ready[i];
            if (pok[i]) {
              const float* psrc = px[i] + (size_t)((round[i] - 1) & 1) * a.NSL * kLstmBT * Hp;
              // NSL independent L2 reads: issue them in batches of 8 before any add (fixed summation order)
              const float* q = psrc + (size_t)pb * Hp + u0 + pu;
              const size_t sstride = (size_t)kLstmBT * Hp;
              float s = 0.f;
              int sl = 0;
              for (; sl + 8 <= a.NSL; sl += 8) {
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = __ldcg(q + (size_t)(sl + j) * sstride);
#pragma unroll
                for (int j = 0; j < 8; ++j) s += v[j];
              }
              for (; sl < a.NSL; ++sl) s += __ldcg(q + (size_t)sl * sstride);
              dh_reg[i] += s;
            }
          }
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
            dh_reg[i] = 0.f;  // consumed; the next reduce-scatter refills it
          }
          if (k == 0) continue;
          if (pok[i]) {
#pragma unroll
            for (int q = 0; q < 4; ++q) das[(size_t)pb * kLstmDaLd + q * US + pu] = da[q];
          }
          compute_sync();
          // 2. partial dh_{k-1}[16, :] = das[16, 0:64] . Ws[0:64, :]  (tensor cores) -> px[round&1][slice]
          float* pdst = px[i] + ((size_t)(round[i] & 1) * a.NSL + slice) * kLstmBT * Hp;
          uint32_t ah[8][4], al[8][4];
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            split2(das[(size_t)g * kLstmDaLd + ks * 8 + t], ah[ks][0], al[ks][0]);
            split2(das[(size_t)(g + 8) * kLstmDaLd + ks * 8 + t], ah[ks][1], al[ks][1]);
            split2(das[(size_t)g * kLstmDaLd + ks * 8 + t + 4], ah[ks][2], al[ks][2]);
            split2(das[(size_t)(g + 8) * kLstmDaLd + ks * 8 + t + 4], ah[ks][3], al[ks][3]);
          }
          compute_sync();  // das[i] may be rewritten by the next step of this instance
          for (int nt = warp; nt < ntiles; nt += 8) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f}, acl[4] = {0.f, 0.f, 0.f, 0.f};
            const float* wB = Ws + nt * 8 + g;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              uint32_t bh[2], bl[2];
              split2(wB[(size_t)(ks * 8 + t) * ld], bh[0], bl[0]);
              split2(wB[(size_t)(ks * 8 + t + 4) * ld], bh[1], bl[1]);
              mma_tf32(acl, al[ks], bh);
              mma_tf32(acl, ah[ks], bl);
              mma_tf32(acc, ah[ks], bh);
            }
            const int col = nt * 8 + 2 * t;
            *reinterpret_cast<float2*>(pdst + (size_t)g * Hp + col) = make_float2(acc[0] + acl[0], acc[1] + acl[1]);
            *reinterpret_cast<float2*>(pdst + (size_t)(g + 8) * Hp + col) =
                make_float2(acc[2] + acl[2], acc[3] + acl[3]);
          }
          // 3. publish through the communication warp; the reduce-scatter happens at the top of the next step
          warp_arrive(&written[i]);
          ++round[i];
        }
      }
    }
  }
}

}  // namespace vocr

using namespace vocr;

static int lstm_geometry(int B, int H, LstmArgs* a, size_t* smem, int* grid_y, bool bwd) {
  if (H < 1 || H > 32 * kLstmMaxUS) return VOCR_INVALID_VALUE;
  a->US = ceil_div(H, 32);
  a->NSL = ceil_div(H, a->US);
  a->Hp = ceil_div(H, 32) * 32;
  a->NBT = ceil_div(B, kLstmBT);
  a->gpd = ceil_div(a->NBT, kLstmNI);
  a->n_groups = 2 * a->gpd;
  const size_t ld = a->Hp + 4;
  *smem = sizeof(float) * (kLstmRows * ld + (bwd ? (size_t)kLstmNI * kLstmBT * kLstmDaLd
                                                  : (size_t)kLstmNI * kLstmBT * ld + 4 * kLstmBT * kLstmPartLd)) + 64;
  *grid_y = max(1, min(a->n_groups, kNumSMs / a->NSL));
  return VOCR_OK;
}

static size_t lstm_xchg_bytes(const LstmArgs& a, bool bwd) {
  return sizeof(float) * (size_t)(2 * a.NBT) * 2 * kLstmBT * a.Hp * (bwd ? a.NSL : 1);
}

extern "C" size_t vocr_bilstm_workspace_size(int B, int H, int backward) {
  LstmArgs a;
  size_t smem;
  int gy;
  if (lstm_geometry(B, H, &a, &smem, &gy, backward != 0) != VOCR_OK) return 0;
  return 256 + ((sizeof(unsigned) * 2 * a.NBT + 255) & ~size_t(255)) + lstm_xchg_bytes(a, backward != 0);
}

static int lstm_launch(bool bwd, LstmArgs a, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  size_t smem;
  int gy;
  int st = lstm_geometry(a.B, a.H, &a, &smem, &gy, bwd);
  if (st != VOCR_OK) return st;
  uintptr_t w = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255);
  const size_t flag_bytes = (sizeof(unsigned) * 2 * a.NBT + 255) & ~size_t(255);
  const size_t xchg = lstm_xchg_bytes(a, bwd);
  if ((w - reinterpret_cast<uintptr_t>(workspace)) + flag_bytes + xchg > workspace_bytes) return VOCR_INVALID_VALUE;
  a.flags = reinterpret_cast<unsigned*>(w);
  a.xchg = reinterpret_cast<float*>(w + flag_bytes);
  // flags start at 0; the exchange buffer is zeroed so padded rows / columns never inject NaNs into the products
  if (cudaMemsetAsync(a.flags, 0, flag_bytes + xchg, stream) != cudaSuccess) return VOCR_MEMOPS_FAILED;
  const void* fn = bwd ? (const void*)bilstm_bwd_kernel : (const void*)bilstm_fwd_kernel;
  if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return VOCR_EXECUTION_FAILED;
  dim3 grid(a.NSL, gy);
  void* params[] = {&a};
  // cooperative launch: the runtime refuses the launch unless every CTA can be co-resident, which the flag
  // protocol relies on
  if (cudaLaunchCooperativeKernel(fn, grid, dim3(kLstmThreads), params, smem, stream) != cudaSuccess)
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
  return lstm_launch(false, a, workspace, workspace_bytes, stream);
}

// dout [T,B,2H], gates/cst from the forward pass -> dgates [T,B,2,4H] (gradient w.r.t. xproj; zero beyond lens).
extern "C" int vocr_bilstm_bwd_f32(const float* dout, const float* whh, const int32_t* lens, const float* gates,
                                   const float* cst, float* dgates, int T, int B, int H, int Tmax, void* workspace,
                                   size_t workspace_bytes, vocr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VOCR_REQUIRE(T >= 0 && B >= 0 && H >= 1 && Tmax >= 0 && Tmax <= T);
  if (T == 0 || B == 0) return VOCR_OK;
  VOCR_REQUIRE(dout && whh && lens && gates && cst && dgates && workspace);
  if (cudaMemsetAsync(dgates, 0, sizeof(float) * (size_t)T * B * 8 * H, stream) != cudaSuccess)
    return VOCR_MEMOPS_FAILED;
  if (Tmax == 0) return VOCR_OK;
  LstmArgs a{};
  a.xproj = dout; a.whh = whh; a.lens = lens; a.out = dgates;
  a.gates = const_cast<float*>(gates); a.cst = const_cast<float*>(cst);
  a.T = T; a.B = B; a.H = H; a.Tmax = Tmax;
  return lstm_launch(true, a, workspace, workspace_bytes, stream);
}
