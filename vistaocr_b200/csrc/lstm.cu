// Bidirectional LSTM layer recurrence on sm_100a, forward and backward, as PERSISTENT cooperative kernels
// (reference cnnlstm.py:148-149,285-290: nn.LSTM on a packed sequence -> cuDNN RNN).
//
// The input projections x_t W_ih^T + b_ih + b_hh of ALL timesteps and both directions are one GEMM done beforehand
// (xproj [T,B,2,4H]); what is left is the strictly sequential part  h_{t-1} W_hh^T  + gate nonlinearities.
// Work decomposition: an "instance" = (direction, tile of 32 samples); it is served by NSL CTAs ("slices"), each
// owning US hidden units = 4*US rows of W_hh, which stay RESIDENT IN SHARED MEMORY for all timesteps
// (H=512: 32 slices x 64 rows x 512 fp32 = 129 KB each).  Per step a slice multiplies the 32 x H tile of h_{t-1}
// by its 64 rows (K split 4 ways across warps, reduced through shared memory), applies the gates for its units,
// and publishes its 32 x US piece of h_t through a ping-pong buffer in global memory (L2); the slices of an
// instance meet at a monotonically increasing flag (release/acquire at gpu scope).  Directions and batch tiles never
// synchronise with each other, so both directions run concurrently (2 dirs x 2 tiles x 32 slices = 128 CTAs for
// B=64, H=512, one per SM).  Ragged lengths use packed-sequence semantics by masking: sample b is active at step k
// iff k < lens[b]; the reverse direction visits t = lens[b]-1-k, i.e. starts at the sample's own last frame;
// outputs beyond lens[b] stay zero.
//
// Backward keeps the same residency: a slice turns dh into gate gradients for its own units, multiplies them by its
// W_hh rows (32 x 64 by 64 x H) into a partial dh_{t-1} for ALL units, and the instance reduce-scatters the partials
// through L2 in a fixed order (deterministic).  dW_hh / dW_ih / db / dx are plain GEMMs over the saved gate
// gradients afterwards (host side).
#include <cooperative_groups.h>
#include "common.cuh"

namespace vocr {

constexpr int kLstmThreads = 256;
constexpr int kLstmBT = 32;        // samples per instance
constexpr int kLstmMaxUS = 16;     // hidden units per slice
constexpr int kLstmRows = 64;      // 4 * kLstmMaxUS gate rows per slice (padded)
constexpr int kLstmPairs = 2;      // (sample, unit) pairs per thread: 32*16/256
constexpr int kLstmPartLd = kLstmRows + 4;  // padded row of the K-split partial sums (conflict-free stores)

struct LstmArgs {
  const float* xproj;   // [T,B,2,4H]  (fwd)            | dout [T,B,2H] (bwd)
  const float* whh;     // [2,4H,H]
  const int32_t* lens;  // [B]
  float* out;           // [T,B,2H]    (fwd, pre-zeroed) | dgates [T,B,2,4H] (bwd, pre-zeroed)
  float* gates;         // [T,B,2,4H] activated i,f,g,o (fwd: written if non-null; bwd: read)
  float* cst;           // [T,B,2,H]  cell state        (fwd: written if non-null; bwd: read)
  float* xchg;          // fwd: [n_inst][2][32][Hp]      | bwd: [n_inst][2][NSL][32][Hp]
  unsigned* flags;      // [n_inst], zeroed before launch
  int T, B, H, Hp, US, NSL, Tmax, NBT, n_inst;
};

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release(unsigned* p) {
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}
// all slices of an instance have published step data `target/NSL` times
__device__ __forceinline__ void wait_flag(const unsigned* flag, unsigned target) {
  if (threadIdx.x == 0) {
    unsigned spins = 0;
    while (ld_acquire(flag) < target) {
      if (++spins > (1u << 26)) asm volatile("trap;");  // ~seconds: a lost peer must not hang the box
    }
  }
  __syncthreads();
}
__device__ __forceinline__ void signal_flag(unsigned* flag) {
  __syncthreads();  // every thread's global writes of this step are issued
  if (threadIdx.x == 0) {
    __threadfence();
    red_release(flag);
  }
}
__device__ __forceinline__ void cp_async16_cg(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

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

// ================================================ forward =====================================================
__global__ void __launch_bounds__(kLstmThreads, 1) bilstm_fwd_kernel(LstmArgs a) {
  extern __shared__ __align__(16) float lstm_smem[];
  const int H = a.H, Hp = a.Hp, US = a.US, ld = Hp + 4;
  float* Ws = lstm_smem;                 // [64][ld]
  float* hs = Ws + kLstmRows * ld;       // [32][ld]   (aliased by part[4][32][68] after the product)
  float* part = hs;
  const int slice = blockIdx.x;
  const int u0 = slice * US;
  const int nu = min(US, H - u0);
  const int tid = threadIdx.x;
  // product mapping: 4 K-groups x 64 threads; thread -> samples tb+8i (i<4), rows tr+8j (j<8)
  const int kg = tid >> 6, t64 = tid & 63, tb = t64 & 7, tr = t64 >> 3;
  const int kq = Hp / 4;  // K range of a group (Hp % 16 == 0)

  for (int inst = blockIdx.y; inst < a.n_inst; inst += gridDim.y) {
    const int dir = inst / a.NBT, bt = inst - dir * a.NBT;
    const int b0 = bt * kLstmBT;
    const int nb = min(kLstmBT, a.B - b0);
    __syncthreads();
    load_w_slice(Ws, a.whh + (size_t)dir * 4 * H * H, H, Hp, US, u0, nu);
    float* hx = a.xchg + (size_t)inst * 2 * kLstmBT * Hp;
    unsigned* flag = a.flags + inst;

    // (sample, unit) pairs owned by this thread for the gate epilogue
    int pb[kLstmPairs], pu[kLstmPairs], plen[kLstmPairs];
    float c_reg[kLstmPairs], h_reg[kLstmPairs];
#pragma unroll
    for (int i = 0; i < kLstmPairs; ++i) {
      const int p = tid + i * kLstmThreads;
      pb[i] = p / US;
      pu[i] = p - pb[i] * US;
      const bool ok = pb[i] < nb && pu[i] < nu && pb[i] < kLstmBT;
      plen[i] = ok ? min(a.lens[b0 + pb[i]], a.Tmax) : 0;
      if (!ok) pb[i] = -1;
      c_reg[i] = 0.f;
      h_reg[i] = 0.f;
    }
    __syncthreads();

    for (int k = 0; k < a.Tmax; ++k) {
      // prefetch this step's input projections (independent of the recurrence)
      float xp[kLstmPairs][4];
      int tt[kLstmPairs];
#pragma unroll
      for (int i = 0; i < kLstmPairs; ++i) {
        const bool act = pb[i] >= 0 && k < plen[i];
        tt[i] = act ? (dir == 0 ? k : plen[i] - 1 - k) : -1;
        if (act) {
          const float* xr = a.xproj + (((size_t)tt[i] * a.B + b0 + pb[i]) * 2 + dir) * 4 * H + u0 + pu[i];
#pragma unroll
          for (int g = 0; g < 4; ++g) xp[i][g] = __ldg(xr + (size_t)g * H);
        } else {
#pragma unroll
          for (int g = 0; g < 4; ++g) xp[i][g] = 0.f;
        }
      }
      float acc[4][8];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
      if (k > 0) {
        wait_flag(flag, (unsigned)(a.NSL * k));
        const float* src = hx + (size_t)((k - 1) & 1) * kLstmBT * Hp;
        const int chunks = kLstmBT * (Hp / 4);
        for (int i = tid; i < chunks; i += kLstmThreads) {
          const int r = i / (Hp / 4), c4 = i - r * (Hp / 4);
          cp_async16_cg(hs + (size_t)r * ld + c4 * 4, src + (size_t)r * Hp + c4 * 4);
        }
        cp_async_wait_all();
        __syncthreads();
        const float* hrow = hs + (size_t)tb * ld + kg * kq;
        const float* wrow = Ws + (size_t)tr * ld + kg * kq;
        for (int kk = 0; kk < kq; kk += 4) {
          float4 hv[4], wv[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) hv[i] = *reinterpret_cast<const float4*>(hrow + (size_t)(8 * i) * ld + kk);
#pragma unroll
          for (int j = 0; j < 8; ++j) wv[j] = *reinterpret_cast<const float4*>(wrow + (size_t)(8 * j) * ld + kk);
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              acc[i][j] = fmaf(hv[i].x, wv[j].x, acc[i][j]);
              acc[i][j] = fmaf(hv[i].y, wv[j].y, acc[i][j]);
              acc[i][j] = fmaf(hv[i].z, wv[j].z, acc[i][j]);
              acc[i][j] = fmaf(hv[i].w, wv[j].w, acc[i][j]);
            }
        }
        __syncthreads();  // everyone is done reading hs before it is reused as `part`
      }
      // partial sums -> part[kg][b][r]
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) part[((size_t)kg * kLstmBT + tb + 8 * i) * kLstmPartLd + tr + 8 * j] = acc[i][j];
      __syncthreads();
      float* hdst = hx + (size_t)(k & 1) * kLstmBT * Hp;
#pragma unroll
      for (int i = 0; i < kLstmPairs; ++i) {
        if (pb[i] < 0) continue;
        if (tt[i] >= 0) {
          float pre[4];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int r = g * US + pu[i];
            float s = xp[i][g];
#pragma unroll
            for (int q = 0; q < 4; ++q) s += part[((size_t)q * kLstmBT + pb[i]) * kLstmPartLd + r];
            pre[g] = s;
          }
          const float ig = sigmoidf_(pre[0]), fg = sigmoidf_(pre[1]), gg = tanhf(pre[2]), og = sigmoidf_(pre[3]);
          const float c = fmaf(fg, c_reg[i], ig * gg);
          const float h = og * tanhf(c);
          c_reg[i] = c;
          h_reg[i] = h;
          const size_t tb_ = (size_t)tt[i] * a.B + b0 + pb[i];
          a.out[(tb_ * 2 + dir) * H + u0 + pu[i]] = h;
          if (a.gates) {
            float* gp = a.gates + (tb_ * 2 + dir) * 4 * H + u0 + pu[i];
            gp[0] = ig;
            gp[(size_t)H] = fg;
            gp[(size_t)2 * H] = gg;
            gp[(size_t)3 * H] = og;
          }
          if (a.cst) a.cst[(tb_ * 2 + dir) * H + u0 + pu[i]] = c;
        }
        hdst[(size_t)pb[i] * Hp + u0 + pu[i]] = h_reg[i];  // finished samples keep publishing their last state
      }
      if (k + 1 < a.Tmax) signal_flag(flag);
      else __syncthreads();
    }
  }
}

// ================================================ backward ====================================================
__global__ void __launch_bounds__(kLstmThreads, 1) bilstm_bwd_kernel(LstmArgs a) {
  extern __shared__ __align__(16) float lstm_smem[];
  const int H = a.H, Hp = a.Hp, US = a.US, ld = Hp + 4;
  float* Ws = lstm_smem;                   // [64][ld]
  float* das = Ws + kLstmRows * ld;        // [32][68] gate gradients of this slice's units, row = sample
  constexpr int ldd = kLstmRows + 4;
  const int slice = blockIdx.x;
  const int u0 = slice * US;
  const int nu = min(US, H - u0);
  const int tid = threadIdx.x;
  const float* dout = a.xproj;  // [T,B,2H]
  float* dgates = a.out;        // [T,B,2,4H]
  // partial-product mapping: thread -> samples tb+4i (i<8), columns tk*4 + 256*j .. +3
  const int tb = tid & 3, tk = tid >> 2;
  const int ncol4 = Hp / 4;  // float4 columns

  for (int inst = blockIdx.y; inst < a.n_inst; inst += gridDim.y) {
    const int dir = inst / a.NBT, bt = inst - dir * a.NBT;
    const int b0 = bt * kLstmBT;
    const int nb = min(kLstmBT, a.B - b0);
    __syncthreads();
    load_w_slice(Ws, a.whh + (size_t)dir * 4 * H * H, H, Hp, US, u0, nu);
    float* px = a.xchg + (size_t)inst * 2 * a.NSL * kLstmBT * Hp;
    unsigned* flag = a.flags + inst;

    int pb[kLstmPairs], pu[kLstmPairs], plen[kLstmPairs];
    float dc_reg[kLstmPairs], dh_reg[kLstmPairs];
#pragma unroll
    for (int i = 0; i < kLstmPairs; ++i) {
      const int p = tid + i * kLstmThreads;
      pb[i] = p / US;
      pu[i] = p - pb[i] * US;
      const bool ok = pb[i] < nb && pu[i] < nu && pb[i] < kLstmBT;
      plen[i] = ok ? min(a.lens[b0 + pb[i]], a.Tmax) : 0;
      if (!ok) pb[i] = -1;
      dc_reg[i] = 0.f;
      dh_reg[i] = 0.f;
    }
    for (int i = tid; i < kLstmBT * ldd; i += kLstmThreads) das[i] = 0.f;
    __syncthreads();

    unsigned round = 0;
    for (int k = a.Tmax - 1; k >= 0; --k) {
      // 1. gate gradients of this slice's units at step k
#pragma unroll
      for (int i = 0; i < kLstmPairs; ++i) {
        if (pb[i] < 0) continue;
        float da[4] = {0.f, 0.f, 0.f, 0.f};
        if (k < plen[i]) {
          const int t = (dir == 0) ? k : plen[i] - 1 - k;
          const size_t tb_ = (size_t)t * a.B + b0 + pb[i];
          const float* gp = a.gates + (tb_ * 2 + dir) * 4 * H + u0 + pu[i];
          const float ig = gp[0], fg = gp[(size_t)H], gg = gp[(size_t)2 * H], og = gp[(size_t)3 * H];
          const float c = a.cst[(tb_ * 2 + dir) * H + u0 + pu[i]];
          float c_prev = 0.f;
          if (k > 0) {
            const int tp = (dir == 0) ? t - 1 : t + 1;
            c_prev = a.cst[(((size_t)tp * a.B + b0 + pb[i]) * 2 + dir) * H + u0 + pu[i]];
          }
          const float dh = dout[(tb_ * 2 + dir) * H + u0 + pu[i]] + dh_reg[i];
          const float tc = tanhf(c);
          const float dc = fmaf(dh * og, 1.f - tc * tc, dc_reg[i]);
          da[0] = dc * gg * ig * (1.f - ig);
          da[1] = dc * c_prev * fg * (1.f - fg);
          da[2] = dc * ig * (1.f - gg * gg);
          da[3] = dh * tc * og * (1.f - og);
          dc_reg[i] = dc * fg;
          float* dg = dgates + (tb_ * 2 + dir) * 4 * H + u0 + pu[i];
          dg[0] = da[0];
          dg[(size_t)H] = da[1];
          dg[(size_t)2 * H] = da[2];
          dg[(size_t)3 * H] = da[3];
          dh_reg[i] = 0.f;  // replaced by the reduced partials below (stays 0 at k == 0)
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) das[(size_t)pb[i] * ldd + g * US + pu[i]] = da[g];
      }
      if (k == 0) break;
      __syncthreads();
      // 2. partial dh_{k-1}[b, :] = das[b, 0:64] . Ws[0:64, :]   ->  px[round&1][slice][b][:]
      float* pdst = px + ((size_t)(round & 1) * a.NSL + slice) * kLstmBT * Hp;
      for (int cb = 0; cb < ncol4; cb += 128) {  // 128 float4 columns (= 2 per thread) per pass
        float4 acc[8][2];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
        const int c0 = cb + tk, c1 = cb + tk + 64;
        const bool v0 = c0 < ncol4, v1 = c1 < ncol4;
        for (int r = 0; r < kLstmRows; r += 4) {
          float4 dv[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) dv[i] = *reinterpret_cast<const float4*>(das + (size_t)(tb + 4 * i) * ldd + r);
#pragma unroll
          for (int rr = 0; rr < 4; ++rr) {
            const float4 w0 = v0 ? *reinterpret_cast<const float4*>(Ws + (size_t)(r + rr) * ld + c0 * 4)
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 w1 = v1 ? *reinterpret_cast<const float4*>(Ws + (size_t)(r + rr) * ld + c1 * 4)
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float d = (rr == 0) ? dv[i].x : (rr == 1) ? dv[i].y : (rr == 2) ? dv[i].z : dv[i].w;
              acc[i][0].x = fmaf(d, w0.x, acc[i][0].x);
              acc[i][0].y = fmaf(d, w0.y, acc[i][0].y);
              acc[i][0].z = fmaf(d, w0.z, acc[i][0].z);
              acc[i][0].w = fmaf(d, w0.w, acc[i][0].w);
              acc[i][1].x = fmaf(d, w1.x, acc[i][1].x);
              acc[i][1].y = fmaf(d, w1.y, acc[i][1].y);
              acc[i][1].z = fmaf(d, w1.z, acc[i][1].z);
              acc[i][1].w = fmaf(d, w1.w, acc[i][1].w);
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float* row = pdst + (size_t)(tb + 4 * i) * Hp;
          if (v0) *reinterpret_cast<float4*>(row + c0 * 4) = acc[i][0];
          if (v1) *reinterpret_cast<float4*>(row + c1 * 4) = acc[i][1];
        }
      }
      // 3. meet the other slices, 4. reduce-scatter: own units <- sum over slices, fixed order
      signal_flag(flag);
      ++round;
      wait_flag(flag, (unsigned)a.NSL * round);
      const float* psrc = px + (size_t)((round - 1) & 1) * a.NSL * kLstmBT * Hp;
#pragma unroll
      for (int i = 0; i < kLstmPairs; ++i) {
        if (pb[i] < 0) continue;
        float s = 0.f;
        for (int sl = 0; sl < a.NSL; ++sl)
          s += __ldcg(psrc + ((size_t)sl * kLstmBT + pb[i]) * Hp + u0 + pu[i]);
        dh_reg[i] += s;
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
  a->Hp = ceil_div(H, 16) * 16;
  a->NBT = ceil_div(B, kLstmBT);
  a->n_inst = 2 * a->NBT;
  const size_t ld = a->Hp + 4;
  *smem = sizeof(float) * (kLstmRows * ld + (bwd ? kLstmBT * (kLstmRows + 4) : max((size_t)kLstmBT * ld, (size_t)4 * kLstmBT * kLstmPartLd)));
  *grid_y = max(1, min(a->n_inst, kNumSMs / a->NSL));
  return VOCR_OK;
}

extern "C" size_t vocr_bilstm_workspace_size(int B, int H, int backward) {
  LstmArgs a;
  size_t smem;
  int gy;
  if (lstm_geometry(B, H, &a, &smem, &gy, backward != 0) != VOCR_OK) return 0;
  const size_t xchg = sizeof(float) * (size_t)a.n_inst * 2 * kLstmBT * a.Hp * (backward ? a.NSL : 1);
  return 256 + ((sizeof(unsigned) * a.n_inst + 255) & ~size_t(255)) + xchg;
}

static int lstm_launch(bool bwd, LstmArgs a, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  size_t smem;
  int gy;
  int st = lstm_geometry(a.B, a.H, &a, &smem, &gy, bwd);
  if (st != VOCR_OK) return st;
  uintptr_t w = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255);
  const size_t flag_bytes = (sizeof(unsigned) * a.n_inst + 255) & ~size_t(255);
  const size_t xchg = sizeof(float) * (size_t)a.n_inst * 2 * kLstmBT * a.Hp * (bwd ? a.NSL : 1);
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
