// CTC loss forward + gradient on sm_100a, log-space fp32, blank = 0, softmax inside.
// Replaces warpctc_pytorch.CTCLoss / warp-ctc's compute_ctc_loss (reference call sites
// src/train_cnn_lstm.py:12,52,138,358).  Four launches per batch:
//   1. ctc_scan_kernel        label offsets (exclusive scan of label_lens)
//   2. ctc_lse_lattice_kernel streaming pass #1 over acts: row log-sum-exp (warp-shuffle max/sum) and the compact
//                             log-softmax lattice  lat[b][t][0]=blank, lat[b][t][1+j]=label j   (L+1 floats/frame)
//   3. ctc_ab_warp_kernel     (L <= 191; ctc_alpha_beta_kernel, a multi-warp shared-memory form, beyond that)
//                             one CTA per utterance; an alpha warp walks t forward while a beta warp
//                             walks t backward CONCURRENTLY over the blank-extended lattice (S=2L+1 states),
//                             one __syncthreads per time step for both; the lattice is streamed through shared
//                             memory by bulk async copies; alpha/beta are renormalised every 8 steps (float64
//                             offsets) so fp32 error does not grow with |log p|; lattices go to the workspace
//   4. ctc_grad_kernel        streaming pass #2 over acts: softmax - occupancy, written once; rows beyond
//                             act_len and infeasible utterances are written as exact zeros
// Passes 2 and 4 stage flat row tiles of acts in shared memory with a 1-D bulk async copy (see decode.cu).
// Algorithmic HBM bytes: 2*T*B*A*4 (acts read once if pass 4 hits L2, grads written once); the lattices add
// 4*S/A of that.
#include "common.cuh"

namespace vocr {

constexpr int kCtcThreads = 256;
constexpr int kCtcWarps = kCtcThreads / 32;

// The recursions run in the log2 domain: one MUFU.EX2 per exp, one MUFU.LG2 per log, no range-reduction code on
// the T-long dependent chain.  Natural-log quantities are converted once per frame (lattice) / utterance (cost).
constexpr float kLog2e = 1.4426950408889634f;
constexpr double kLn2 = 0.6931471805599453;

__device__ __forceinline__ float lse3_2(float a, float b, float c) {
  const float m = fmaxf(a, fmaxf(b, c));
  if (m == kNegInf) return kNegInf;
  return m + __log2f(exp2f(a - m) + exp2f(b - m) + exp2f(c - m));
}
__device__ __forceinline__ float lse2_2(float a, float b) {
  const float m = fmaxf(a, b);
  if (m == kNegInf) return kNegInf;
  return m + __log2f(exp2f(a - m) + exp2f(b - m));
}

constexpr int kCtcChunk = 16;  // lattice frames staged per bulk copy in the recursion kernel
constexpr int kCtcRenorm = 8;  // renormalise alpha/beta every this many steps

struct CtcWorkspace {
  int32_t* offsets;  // [B+1]
  int32_t* nxt;      // [sum L] next position with the same symbol, -1 if none
  int32_t* first;    // [sum L] 1 if no earlier position has the same symbol
  double* ll;        // [B] log2-likelihood (-inf = infeasible)
  double* offa;      // [B*T] cumulative renormalisation offset of alpha at frame t
  double* offb;      // [B*T] same for beta
  float* lse;        // [B*T]
  float* lat;        // [B*T*Lp] compact log2-softmax lattice, Lp = round_up(Lmax+1, 4)
  float* alpha;      // [B*T*Smax] renormalised alpha
  float* beta;       // [B*T*Smax] renormalised beta
};

__host__ __device__ inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }
__host__ __device__ inline int lat_stride(int Lmax) { return (Lmax + 1 + 3) & ~3; }

__host__ inline size_t ctc_carve(CtcWorkspace* ws, void* base, int T, int B, int max_label_len) {
  const size_t Lmax = (size_t)max_label_len, Smax = 2 * Lmax + 1, Lp = (size_t)lat_stride(max_label_len);
  size_t off = 0;
  unsigned char* p = static_cast<unsigned char*>(base);
  auto take = [&](size_t bytes) {
    unsigned char* r = p ? p + off : nullptr;
    off += align256(bytes);
    return r;
  };
  CtcWorkspace w;
  w.offsets = (int32_t*)take(sizeof(int32_t) * ((size_t)B + 1));
  w.nxt = (int32_t*)take(sizeof(int32_t) * (size_t)B * Lmax + 4);
  w.first = (int32_t*)take(sizeof(int32_t) * (size_t)B * Lmax + 4);
  w.ll = (double*)take(sizeof(double) * (size_t)B);
  w.offa = (double*)take(sizeof(double) * (size_t)B * T);
  w.offb = (double*)take(sizeof(double) * (size_t)B * T);
  w.lse = (float*)take(sizeof(float) * (size_t)B * T);
  w.lat = (float*)take(sizeof(float) * (size_t)B * T * Lp);
  w.alpha = (float*)take(sizeof(float) * (size_t)B * T * Smax);
  w.beta = (float*)take(sizeof(float) * (size_t)B * T * Smax);
  if (ws) *ws = w;
  return off;
}

// ---- 1. label offsets: exclusive scan of label_lens, one CTA, parallel --------------------------------------
__global__ void __launch_bounds__(1024)
ctc_scan_kernel(const int32_t* __restrict__ label_lens, int B, int Lmax, CtcWorkspace ws) {
  __shared__ int s_part[1024];
  const int per = ceil_div(B, (int)blockDim.x);
  const int lo = min(B, (int)threadIdx.x * per), hi = min(B, lo + per);
  int sum = 0;
  for (int i = lo; i < hi; ++i) sum += max(0, min(label_lens[i], Lmax));
  s_part[threadIdx.x] = sum;
  __syncthreads();
  // Hillis-Steele inclusive scan over the 1024 partials
  for (int d = 1; d < (int)blockDim.x; d <<= 1) {
    const int v = (threadIdx.x >= (unsigned)d) ? s_part[threadIdx.x - d] : 0;
    __syncthreads();
    s_part[threadIdx.x] += v;
    __syncthreads();
  }
  int run = s_part[threadIdx.x] - sum;  // exclusive
  for (int i = lo; i < hi; ++i) {
    ws.offsets[i] = run;
    run += max(0, min(label_lens[i], Lmax));
  }
  if (threadIdx.x == blockDim.x - 1) ws.offsets[B] = s_part[threadIdx.x];
}

// ---- 2. row LSE + compact lattice --------------------------------------------------------------------------
__global__ void __launch_bounds__(kCtcThreads)
ctc_lse_lattice_kernel(const float* __restrict__ acts, long long n_rows, int T, int B, int A, int rows_per_tile,
                       bool base_aligned, const int32_t* __restrict__ labels,
                       const int32_t* __restrict__ label_lens, const int32_t* __restrict__ act_lens, int Lmax,
                       CtcWorkspace ws) {
  extern __shared__ __align__(128) unsigned char ctc_smem[];
  const long long row0 = (long long)blockIdx.x * rows_per_tile;
  const int rows_here = (int)min((long long)rows_per_tile, n_rows - row0);
  const float* tile = stage_row_tile(ctc_smem, acts + row0 * A, rows_here * A, base_aligned);
  const int warp = threadIdx.x >> 5, lane = lane_id();
  const int Lp = lat_stride(Lmax);
  for (int r = warp; r < rows_here; r += kCtcWarps) {
    const long long gr = row0 + r;
    const int t = (int)(gr / B), b = (int)(gr % B);
    if (t >= act_lens[b]) continue;
    const float* row = tile + (size_t)r * A;
    float m = kNegInf;
    for (int a = lane; a < A; a += 32) m = fmaxf(m, row[a]);
    m = warp_max(m);
    float s = 0.f;
    for (int a = lane; a < A; a += 32) s += expf(row[a] - m);
    s = warp_sum(s);
    const float lse = m + logf(s);
    const size_t bt = (size_t)b * T + t;
    if (lane == 0) ws.lse[bt] = lse;
    const int L = max(0, min(label_lens[b], Lmax));
    const int off = ws.offsets[b];
    float* lat = ws.lat + bt * (size_t)Lp;
    for (int j = lane; j <= L; j += 32) {
      int sym = (j == 0) ? 0 : labels[off + j - 1];
      sym = max(0, min(sym, A - 1));
      lat[j] = (row[sym] - lse) * kLog2e;
    }
  }
}

// ---- 3. alpha / beta recursions ----------------------------------------------------------------------------
// block = 2*G threads: threads [0,G) own alpha (t ascending), [G,2G) own beta (t descending), in lock step with
// one __syncthreads per time step.  The compact lattice is streamed through shared memory kCtcChunk frames at a
// time with 1-D bulk async copies (double buffered per direction).  Every kCtcRenorm steps the state vector is
// renormalised by its maximum and the subtracted amount accumulated in float64, so the stored alpha/beta stay
// O(1)-O(10) in magnitude: fp32 log-space error no longer grows with |log p| (warp-ctc's does).
// dynamic smem: [mbarriers 4x8][lab Smax][bufs 2x2xSmax][wmax 2x32][lattice 2x2xkCtcChunk*Lp]
__global__ void ctc_alpha_beta_kernel(const int32_t* __restrict__ labels, const int32_t* __restrict__ label_lens,
                                      const int32_t* __restrict__ act_lens, int T, int B, int A, int Lmax, int G,
                                      CtcWorkspace ws, float* __restrict__ costs) {
  extern __shared__ __align__(128) unsigned char ab_smem[];
  const int Smax = 2 * Lmax + 1;
  const int Lp = lat_stride(Lmax);
  uint64_t* bars = reinterpret_cast<uint64_t*>(ab_smem);                  // [2 grp][2 buf]
  float* latbuf = reinterpret_cast<float*>(ab_smem + 32);                 // [2][2][kCtcChunk*Lp] (16-B aligned)
  float* bufs = latbuf + (size_t)4 * kCtcChunk * Lp;                      // [2][2][Smax]
  float* wmax = bufs + (size_t)4 * Smax;                                  // [2][32]
  int* lab = reinterpret_cast<int*>(wmax + 64);                           // [Smax]
  const int b = blockIdx.x;
  const int L = max(0, min(label_lens[b], Lmax));
  const int S = 2 * L + 1;
  const int Tb = max(0, min(act_lens[b], T));
  const int grp = threadIdx.x / G;  // 0 alpha, 1 beta
  const int tid = threadIdx.x - grp * G;
  const int off = ws.offsets[b];
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    int sym = 0;
    if (s & 1) sym = max(0, min(labels[off + (s >> 1)], A - 1));
    lab[s] = sym;
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    mbar_fence_init();
  }
  __syncthreads();
  // duplicate-symbol chains for the gradient kernel (label positions j = 0..L-1, symbol lab[2j+1])
  for (int j = threadIdx.x; j < L; j += blockDim.x) {
    const int sym = lab[2 * j + 1];
    int nx = -1, fi = 1;
    for (int q = j + 1; q < L; ++q)
      if (lab[2 * q + 1] == sym) {
        nx = q;
        break;
      }
    for (int q = 0; q < j; ++q)
      if (lab[2 * q + 1] == sym) {
        fi = 0;
        break;
      }
    ws.nxt[off + j] = nx;
    ws.first[off + j] = fi;
  }
  if (Tb == 0) {
    if (threadIdx.x == 0) {
      ws.ll[b] = (L == 0) ? 0.0 : -INFINITY;  // zero frames: only the empty labelling is feasible; cost 0 either way
      costs[b] = 0.f;
    }
    return;
  }
  float* my = bufs + (size_t)grp * 2 * Smax;  // [2][Smax] ping-pong
  float* mylat = latbuf + (size_t)grp * 2 * kCtcChunk * Lp;
  float* out_lat = (grp == 0 ? ws.alpha : ws.beta) + (size_t)b * T * Smax;
  double* out_off = (grp == 0 ? ws.offa : ws.offb) + (size_t)b * T;
  const float* lat_g = ws.lat + (size_t)b * T * Lp;
  const int nchunks = ceil_div(Tb, kCtcChunk);
  const int nwarps_g = G >> 5;
  double offset = 0.0;  // meaningful in tid == 0 of each group

  auto issue = [&](int c) {  // called by tid == 0 of each group
    const int k0 = c * kCtcChunk, k1 = min(Tb, k0 + kCtcChunk);
    const int t_lo = (grp == 0) ? k0 : Tb - k1;  // beta frames Tb-1-k for k in [k0,k1)  ->  [Tb-k1, Tb-1-k0]
    const uint32_t bytes = (uint32_t)(k1 - k0) * (uint32_t)Lp * 4u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_arrive_expect_tx(&bars[grp * 2 + (c & 1)], bytes);
    bulk_g2s(mylat + (size_t)(c & 1) * kCtcChunk * Lp, lat_g + (size_t)t_lo * Lp, bytes, &bars[grp * 2 + (c & 1)]);
  };
  if (tid == 0) issue(0);
  // per-thread constants of state s = tid
  const int dir = (grp == 0) ? -1 : 1;
  const int li0 = (tid & 1) ? (tid >> 1) + 1 : 0;
  bool n1 = false, n2 = false, init0 = false;
  if (tid < S) {
    if (grp == 0) {
      n1 = tid >= 1;
      n2 = (tid & 1) && tid >= 3 && lab[tid] != lab[tid - 2];
      init0 = tid <= 1;
    } else {
      n1 = tid + 1 < S;
      n2 = (tid & 1) && tid + 2 < S && lab[tid] != lab[tid + 2];
      init0 = tid >= S - 2;
    }
  }
  float own = kNegInf;

  for (int c = 0; c < nchunks; ++c) {
    if (tid == 0 && c + 1 < nchunks) issue(c + 1);
    mbar_wait_or_trap(&bars[grp * 2 + (c & 1)], (uint32_t)((c >> 1) & 1));
    const int k0 = c * kCtcChunk, k1 = min(Tb, k0 + kCtcChunk);
    const int t_lo = (grp == 0) ? k0 : Tb - k1;
    const float* chunk = mylat + (size_t)(c & 1) * kCtcChunk * Lp;
    for (int k = k0; k < k1; ++k) {
      const int t = (grp == 0) ? k : Tb - 1 - k;
      const float* lrow = chunk + (size_t)(t - t_lo) * Lp;
      const float* prev = my + (size_t)((k + 1) & 1) * Smax;
      float* cur = my + (size_t)(k & 1) * Smax;
      const bool renorm = ((k % kCtcRenorm) == kCtcRenorm - 1);
      float vmax = kNegInf;
      // state s = tid: neighbours, lattice slot and skip rule are per-thread constants, own value lives in a register
      if (tid < S) {
        const float lp = lrow[li0];
        float v;
        if (k == 0) {
          v = init0 ? lp : kNegInf;
        } else {
          const float a1 = n1 ? prev[tid + dir] : kNegInf;
          const float a2 = n2 ? prev[tid + 2 * dir] : kNegInf;
          v = lse3_2(own, a1, a2) + lp;
        }
        own = v;
        cur[tid] = v;
        vmax = v;
        if (!renorm) out_lat[(size_t)t * Smax + tid] = v;
      }
      for (int s = tid + G; s < S; s += G) {  // only when S > G (label length > 255)
        const float lp = lrow[(s & 1) ? (s >> 1) + 1 : 0];
        float v;
        if (k == 0) {
          v = ((grp == 0) ? (s <= 1) : (s >= S - 2)) ? lp : kNegInf;
        } else {
          float a0 = prev[s], a1 = kNegInf, a2 = kNegInf;
          if (grp == 0) {
            if (s >= 1) a1 = prev[s - 1];
            if ((s & 1) && s >= 3 && lab[s] != lab[s - 2]) a2 = prev[s - 2];
          } else {
            if (s + 1 < S) a1 = prev[s + 1];
            if ((s & 1) && s + 2 < S && lab[s] != lab[s + 2]) a2 = prev[s + 2];
          }
          v = lse3_2(a0, a1, a2) + lp;
        }
        cur[s] = v;
        vmax = fmaxf(vmax, v);
        if (!renorm) out_lat[(size_t)t * Smax + s] = v;
      }
      if (renorm) {
        vmax = warp_max(vmax);
        if ((tid & 31) == 0) wmax[grp * 32 + (tid >> 5)] = vmax;
        __syncthreads();
        float m = kNegInf;
        for (int w = 0; w < nwarps_g; ++w) m = fmaxf(m, wmax[grp * 32 + w]);
        if (m == kNegInf) m = 0.f;  // every state impossible: nothing to rescale
        own -= m;
        for (int s = tid; s < S; s += G) {
          const float v = cur[s] - m;
          cur[s] = v;
          out_lat[(size_t)t * Smax + s] = v;
        }
        offset += (double)m;
      }
      if (tid == 0) out_off[t] = offset;
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) {
    const float* fin = bufs + (size_t)((Tb - 1) & 1) * Smax;  // alpha at t = Tb-1
    const float l = lse2_2(fin[S - 1], (S >= 2) ? fin[S - 2] : kNegInf);
    const double ll2 = (l == kNegInf) ? -INFINITY : offset + (double)l;  // log2 p(labels | acts)
    ws.ll[b] = ll2;
    costs[b] = (l == kNegInf) ? 0.f : (float)(-ll2 * kLn2);
  }
}

// ---- 3b. alpha / beta recursions, one WARP per direction, state in registers, warp shuffles ------------------------
// The workhorse for label lengths up to 32*Q-1 (Q <= 6 -> L <= 191).  Lane i owns Q consecutive label positions
// j = i*Q .. i*Q+Q-1, i.e. the state pairs (blank 2j, label 2j+1) of the blank-extended labelling, in registers.
// A step needs exactly one remote value per lane - the previous lane's last label state - fetched with one
// __shfl_up: no shared-memory exchange and no block barrier on the T-long dependent chain.  The beta recursion is
// the alpha recursion of the REVERSED labelling walked backwards in time (same code, mirrored output index), so
// warp 0 (alpha) and warp 1 (beta) never synchronise with each other.  The compact lattice is streamed through
// shared memory with 1-D bulk async copies, double buffered per warp.
__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_fast(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// log2(2^a + 2^b) and log2(2^a + 2^b + 2^c); -inf safe (ex2(-inf) = 0, and an all--inf input returns -inf)
__device__ __forceinline__ float lse2_w(float a, float b) {
  const float m = fmaxf(a, b);
  const float r = m + lg2_fast(ex2_fast(a - m) + ex2_fast(b - m));
  return (m == kNegInf) ? kNegInf : r;
}
__device__ __forceinline__ float lse3_w(float a, float b, float c) {
  const float m = fmaxf(a, fmaxf(b, c));
  const float r = m + lg2_fast(ex2_fast(a - m) + ex2_fast(b - m) + ex2_fast(c - m));
  return (m == kNegInf) ? kNegInf : r;
}

template <int Q>
__global__ void __launch_bounds__(64)
ctc_ab_warp_kernel(const int32_t* __restrict__ labels, const int32_t* __restrict__ label_lens,
                   const int32_t* __restrict__ act_lens, int T, int B, int A, int Lmax, CtcWorkspace ws,
                   float* __restrict__ costs) {
  extern __shared__ __align__(128) unsigned char abw_smem[];
  const int Smax = 2 * Lmax + 1;
  const int Lp = lat_stride(Lmax);
  uint64_t* bars = reinterpret_cast<uint64_t*>(abw_smem);   // [2 dir][2 buf]
  float* latbuf = reinterpret_cast<float*>(abw_smem + 32);  // [2][2][kCtcChunk*Lp]
  int* lab = reinterpret_cast<int*>(latbuf + (size_t)4 * kCtcChunk * Lp);  // [Lmax]
  const int b = blockIdx.x;
  const int L = max(0, min(label_lens[b], Lmax));
  const int S = 2 * L + 1;
  const int Tb = max(0, min(act_lens[b], T));
  const int dirn = threadIdx.x >> 5;  // 0 alpha, 1 beta
  const int lane = threadIdx.x & 31;
  const int off = ws.offsets[b];
  for (int j = threadIdx.x; j < L; j += blockDim.x) lab[j] = max(0, min(labels[off + j], A - 1));
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    mbar_fence_init();
  }
  __syncthreads();
  for (int j = threadIdx.x; j < L; j += blockDim.x) {  // duplicate-symbol chains for the gradient kernel
    const int sym = lab[j];
    int nx = -1, fi = 1;
    for (int q = j + 1; q < L; ++q)
      if (lab[q] == sym) {
        nx = q;
        break;
      }
    for (int q = 0; q < j; ++q)
      if (lab[q] == sym) {
        fi = 0;
        break;
      }
    ws.nxt[off + j] = nx;
    ws.first[off + j] = fi;
  }
  if (Tb == 0) {
    if (threadIdx.x == 0) {
      ws.ll[b] = (L == 0) ? 0.0 : -INFINITY;
      costs[b] = 0.f;
    }
    return;
  }
  // per-lane constants.  Position j (in THIS direction's order) carries a blank state and, if j < L, a label state.
  bool has_blank[Q], has_label[Q], skip[Q];
  int slot[Q], oblank[Q], olabel[Q];
#pragma unroll
  for (int i = 0; i < Q; ++i) {
    const int j = lane * Q + i;
    has_blank[i] = j <= L;
    has_label[i] = j < L;
    const int jo = (dirn == 0) ? j : L - 1 - j;  // position in the original labelling
    slot[i] = has_label[i] ? 1 + jo : 0;
    int sym = 0, symp = -1;
    if (has_label[i]) {
      sym = lab[jo];
      if (j >= 1) symp = lab[(dirn == 0) ? jo - 1 : jo + 1];
    }
    skip[i] = has_label[i] && j >= 1 && sym != symp;
    oblank[i] = (dirn == 0) ? 2 * j : S - 1 - 2 * j;
    olabel[i] = (dirn == 0) ? 2 * j + 1 : S - 2 - 2 * j;
  }
  float* mylat = latbuf + (size_t)dirn * 2 * kCtcChunk * Lp;
  float* out_lat = (dirn == 0 ? ws.alpha : ws.beta) + (size_t)b * T * Smax;
  double* out_off = (dirn == 0 ? ws.offa : ws.offb) + (size_t)b * T;
  const float* lat_g = ws.lat + (size_t)b * T * Lp;
  const int nchunks = ceil_div(Tb, kCtcChunk);
  double offset = 0.0;
  float bl[Q], lb[Q];
#pragma unroll
  for (int i = 0; i < Q; ++i) bl[i] = lb[i] = kNegInf;

  auto issue = [&](int c) {  // lane 0 of each warp
    const int k0 = c * kCtcChunk, k1 = min(Tb, k0 + kCtcChunk);
    const int t_lo = (dirn == 0) ? k0 : Tb - k1;
    const uint32_t bytes = (uint32_t)(k1 - k0) * (uint32_t)Lp * 4u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_arrive_expect_tx(&bars[dirn * 2 + (c & 1)], bytes);
    bulk_g2s(mylat + (size_t)(c & 1) * kCtcChunk * Lp, lat_g + (size_t)t_lo * Lp, bytes, &bars[dirn * 2 + (c & 1)]);
  };
  if (lane == 0) issue(0);
  for (int c = 0; c < nchunks; ++c) {
    __syncwarp();  // every lane is done reading buffer (c+1)&1 (chunk c-1) before it is refilled
    if (lane == 0 && c + 1 < nchunks) issue(c + 1);
    mbar_wait_or_trap(&bars[dirn * 2 + (c & 1)], (uint32_t)((c >> 1) & 1));
    const int k0 = c * kCtcChunk, k1 = min(Tb, k0 + kCtcChunk);
    const int t_lo = (dirn == 0) ? k0 : Tb - k1;
    const float* chunk = mylat + (size_t)(c & 1) * kCtcChunk * Lp;
    for (int k = k0; k < k1; ++k) {
      const int t = (dirn == 0) ? k : Tb - 1 - k;
      const float* lrow = chunk + (size_t)(t - t_lo) * Lp;
      const float lpb = lrow[0];
      float lpl[Q];
#pragma unroll
      for (int i = 0; i < Q; ++i) lpl[i] = lrow[slot[i]];
      if (k == 0) {
#pragma unroll
        for (int i = 0; i < Q; ++i) {
          const int j = lane * Q + i;
          bl[i] = (j == 0) ? lpb : kNegInf;
          lb[i] = (j == 0 && has_label[i]) ? lpl[i] : kNegInf;
        }
      } else {
        float pl = __shfl_up_sync(0xffffffffu, lb[Q - 1], 1);  // previous lane's last label state
        if (lane == 0) pl = kNegInf;
#pragma unroll
        for (int i = 0; i < Q; ++i) {
          const float nb = lse2_w(bl[i], pl) + lpb;
          const float nl = lse3_w(lb[i], bl[i], skip[i] ? pl : kNegInf) + lpl[i];
          pl = lb[i];  // old label state of this position feeds the next position
          bl[i] = has_blank[i] ? nb : kNegInf;
          lb[i] = has_label[i] ? nl : kNegInf;
        }
      }
      if ((k % kCtcRenorm) == kCtcRenorm - 1) {
        float m = kNegInf;
#pragma unroll
        for (int i = 0; i < Q; ++i) m = fmaxf(m, fmaxf(bl[i], lb[i]));
        m = warp_max(m);
        if (m != kNegInf) {
#pragma unroll
          for (int i = 0; i < Q; ++i) {
            bl[i] -= m;
            lb[i] -= m;
          }
          offset += (double)m;
        }
      }
      float* orow = out_lat + (size_t)t * Smax;
#pragma unroll
      for (int i = 0; i < Q; ++i) {
        if (has_blank[i]) orow[oblank[i]] = bl[i];
        if (has_label[i]) orow[olabel[i]] = lb[i];
      }
      if (lane == 0) out_off[t] = offset;
    }
  }
  if (dirn == 0) {
    // alpha_{Tb-1}(S-1) is the blank of position L, alpha_{Tb-1}(S-2) the label of position L-1
    float fb = kNegInf, fl = kNegInf;
#pragma unroll
    for (int i = 0; i < Q; ++i) {
      const int j = lane * Q + i;
      if (j == L) fb = bl[i];
      if (j == L - 1) fl = lb[i];
    }
    fb = warp_max(fb);
    fl = warp_max(fl);
    if (lane == 0) {
      const float l = lse2_w(fb, fl);
      const double ll2 = (l == kNegInf) ? -INFINITY : offset + (double)l;
      ws.ll[b] = ll2;
      costs[b] = (l == kNegInf) ? 0.f : (float)(-ll2 * kLn2);
    }
  }
}

// ---- 4. gradient ------------------------------------------------------------------------------------------
// dynamic smem: 16 B mbarrier + tile + per-warp occupancy scratch kCtcWarps*Smax floats
__global__ void __launch_bounds__(kCtcThreads)
ctc_grad_kernel(const float* __restrict__ acts, float* __restrict__ grads, long long n_rows, int T, int B, int A,
                int rows_per_tile, bool base_aligned, const int32_t* __restrict__ labels,
                const int32_t* __restrict__ label_lens, const int32_t* __restrict__ act_lens, int Lmax,
                CtcWorkspace ws) {
  extern __shared__ __align__(128) unsigned char ctc_smem[];
  const int Smax = 2 * Lmax + 1;
  // walk tiles from the end: the rows touched last by pass 2 are the most likely to still sit in L2
  const long long tile_idx = (long long)gridDim.x - 1 - blockIdx.x;
  const long long row0 = tile_idx * rows_per_tile;
  const int rows_here = (int)min((long long)rows_per_tile, n_rows - row0);
  float* tile = stage_row_tile(ctc_smem, acts + row0 * A, rows_here * A, base_aligned);
  const size_t tile_floats = (size_t)rows_per_tile * A;
  float* scratch = tile + ((tile_floats + 3) & ~size_t(3));
  const int warp = threadIdx.x >> 5, lane = lane_id();
  float* gs = scratch + (size_t)warp * Smax;
  for (int r = warp; r < rows_here; r += kCtcWarps) {
    const long long gr = row0 + r;
    const int t = (int)(gr / B), b = (int)(gr % B);
    float* row = tile + (size_t)r * A;
    const double ll = ws.ll[b];
    if (t >= act_lens[b] || ll == -INFINITY) {
      for (int a = lane; a < A; a += 32) row[a] = 0.f;
      continue;
    }
    const size_t bt = (size_t)b * T + t;
    const float lse = ws.lse[bt];
    for (int a = lane; a < A; a += 32) row[a] = expf(row[a] - lse);
    const int L = max(0, min(label_lens[b], Lmax));
    const int S = 2 * L + 1;
    const int off = ws.offsets[b];
    const float* al = ws.alpha + bt * (size_t)Smax;
    const float* be = ws.beta + bt * (size_t)Smax;
    const float* lat = ws.lat + bt * (size_t)lat_stride(Lmax);
    // alpha and beta are stored renormalised, in log2 units: fold both float64 offsets and the log-likelihood into one O(1) term
    const float shift = (float)(ws.offa[bt] + ws.offb[bt] - ll);
    float blank_occ = 0.f;
    for (int s = lane; s < S; s += 32) {
      const float lp = lat[(s & 1) ? (s >> 1) + 1 : 0];
      const float ab = al[s] + be[s];
      // beta carries the emission at t as well as alpha: remove one copy; -inf states contribute 0
      const float g = (ab == kNegInf) ? 0.f : exp2f(ab - lp + shift);
      if (s & 1) gs[s >> 1] = g;
      else blank_occ += g;
    }
    blank_occ = warp_sum(blank_occ);
    __syncwarp();
    if (lane == 0) row[0] -= blank_occ;
    for (int j = lane; j < L; j += 32) {
      if (ws.first[off + j]) {
        float tot = 0.f;
        for (int q = j; q >= 0; q = ws.nxt[off + q]) tot += gs[q];
        const int sym = max(0, min(labels[off + j], A - 1));
        row[sym] -= tot;
      }
    }
    __syncwarp();
  }
  __syncthreads();
  // write the whole tile back, coalesced
  float* gdst = grads + row0 * A;
  const int n = rows_here * A;
  if (base_aligned && (n % 4) == 0) {
    const float4* src4 = reinterpret_cast<const float4*>(tile);
    float4* dst4 = reinterpret_cast<float4*>(gdst);
    for (int i = threadIdx.x; i < n / 4; i += kCtcThreads) dst4[i] = src4[i];
  } else {
    for (int i = threadIdx.x; i < n; i += kCtcThreads) gdst[i] = tile[i];
  }
}

}  // namespace vocr

using namespace vocr;

extern "C" size_t vocr_ctc_workspace_size(int T, int B, int A, int max_label_len) {
  (void)A;
  if (T < 0 || B < 0 || max_label_len < 0) return 0;
  return ctc_carve(nullptr, nullptr, T, B, max_label_len) + 256;
}

extern "C" int vocr_ctc_loss_f32(const float* acts, float* grads, const int32_t* labels, const int32_t* label_lens,
                                 const int32_t* act_lens, int T, int B, int A, int max_label_len, float* costs,
                                 void* workspace, size_t workspace_bytes, vocr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VOCR_REQUIRE(T >= 0 && B >= 0 && A >= 1 && max_label_len >= 0);
  if (B == 0) return VOCR_OK;
  VOCR_REQUIRE(label_lens && act_lens && costs && workspace);
  VOCR_REQUIRE(labels || max_label_len == 0);
  VOCR_REQUIRE(T == 0 || acts);
  uintptr_t wbase = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255);
  CtcWorkspace ws;
  const size_t need = ctc_carve(&ws, reinterpret_cast<void*>(wbase), T, B, max_label_len);
  VOCR_REQUIRE(need + (wbase - reinterpret_cast<uintptr_t>(workspace)) <= workspace_bytes);
  const int Lmax = max_label_len, Smax = 2 * Lmax + 1;
  static const int32_t kDummy = 0;
  const int32_t* labels_safe = labels ? labels : &kDummy;  // never dereferenced when Lmax == 0

  ctc_scan_kernel<<<1, 1024, 0, stream>>>(label_lens, B, Lmax, ws);
  VOCR_CHECK_LAUNCH();

  const long long n_rows = (long long)T * B;
  int rows_per_tile = 32;
  while (rows_per_tile > 4 && (size_t)rows_per_tile * A * 4 > 32768) rows_per_tile >>= 1;
  const size_t tile_bytes = ((size_t)rows_per_tile * A * 4 + 15) & ~size_t(15);
  const bool base_aligned = (reinterpret_cast<uintptr_t>(acts) % 16 == 0) &&
                            (grads == nullptr || reinterpret_cast<uintptr_t>(grads) % 16 == 0);
  const long long n_tiles = ceil_div64(n_rows, rows_per_tile);
  VOCR_REQUIRE(n_tiles < (1ll << 31));
  if (n_rows > 0) {
    const size_t smem2 = 16 + tile_bytes;
    if (smem2 > 48 * 1024) {
      VOCR_REQUIRE(smem2 <= 200 * 1024);
      if (cudaFuncSetAttribute(ctc_lse_lattice_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2) !=
          cudaSuccess)
        return VOCR_EXECUTION_FAILED;
    }
    ctc_lse_lattice_kernel<<<(unsigned)n_tiles, kCtcThreads, smem2, stream>>>(
        acts, n_rows, T, B, A, rows_per_tile, base_aligned, labels_safe, label_lens, act_lens, Lmax, ws);
    VOCR_CHECK_LAUNCH();
  }
  if (Lmax + 1 <= 32 * 6) {
    const int Q = ceil_div(Lmax + 1, 32);
    const size_t smem3 = 32 + sizeof(float) * (size_t)4 * kCtcChunk * lat_stride(Lmax) + sizeof(int) * (size_t)(Lmax + 1);
#define VOCR_LAUNCH_ABW(q)                                                                                         \
  case q:                                                                                                          \
    if (smem3 > 48 * 1024 && cudaFuncSetAttribute(ctc_ab_warp_kernel<q>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                  (int)smem3) != cudaSuccess)                                      \
      return VOCR_EXECUTION_FAILED;                                                                                \
    ctc_ab_warp_kernel<q><<<B, 64, smem3, stream>>>(labels_safe, label_lens, act_lens, T, B, A, Lmax, ws, costs);    \
    break;
    switch (Q) {
      VOCR_LAUNCH_ABW(1)
      VOCR_LAUNCH_ABW(2)
      VOCR_LAUNCH_ABW(3)
      VOCR_LAUNCH_ABW(4)
      VOCR_LAUNCH_ABW(5)
      VOCR_LAUNCH_ABW(6)
      default: return VOCR_INVALID_VALUE;
    }
#undef VOCR_LAUNCH_ABW
    VOCR_CHECK_LAUNCH();
  } else {
    int G = ((Smax + 31) / 32) * 32;
    if (G > 512) G = 512;
    const size_t smem3 = 32 + sizeof(float) * ((size_t)4 * kCtcChunk * lat_stride(Lmax) + 4 * (size_t)Smax + 64) +
                         sizeof(int) * (size_t)Smax;
    if (smem3 > 48 * 1024) {
      VOCR_REQUIRE(smem3 <= 200 * 1024);
      if (cudaFuncSetAttribute(ctc_alpha_beta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3) !=
          cudaSuccess)
        return VOCR_EXECUTION_FAILED;
    }
    ctc_alpha_beta_kernel<<<B, 2 * G, smem3, stream>>>(labels_safe, label_lens, act_lens, T, B, A, Lmax, G, ws,
                                                       costs);
    VOCR_CHECK_LAUNCH();
  }
  if (grads && n_rows > 0) {
    const size_t smem4 = 16 + tile_bytes + sizeof(float) * (size_t)kCtcWarps * Smax + 16;
    if (smem4 > 48 * 1024) {
      VOCR_REQUIRE(smem4 <= 200 * 1024);
      if (cudaFuncSetAttribute(ctc_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4) !=
          cudaSuccess)
        return VOCR_EXECUTION_FAILED;
    }
    ctc_grad_kernel<<<(unsigned)n_tiles, kCtcThreads, smem4, stream>>>(acts, grads, n_rows, T, B, A, rows_per_tile,
                                                                       base_aligned, labels_safe, label_lens,
                                                                       act_lens, Lmax, ws);
    VOCR_CHECK_LAUNCH();
  }
  return VOCR_OK;
}
