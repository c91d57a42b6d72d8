// CTC loss forward + gradient on sm_100a, log-space fp32, blank = 0, softmax inside.
// Replaces warpctc_pytorch.CTCLoss / warp-ctc's compute_ctc_loss (reference call sites
// src/train_cnn_lstm.py:12,52,138,358).  Four launches per batch:
//   1. ctc_scan_kernel     label offsets (exclusive scan of label_lens; warp-shuffle scan)
//   2. ctc_lattice_kernel  streaming pass #1 over acts, 8 lanes per row: row log-sum-exp and the compact log2-softmax
//                          lattice  lat[b][t][0] = blank, lat[b][t][1+j] = label j   (L+1 floats per frame)
//   3. ctc_mitm_kernel     one CTA (2 warps) per utterance: the alpha warp walks t forward while the beta warp walks
//                          t backward CONCURRENTLY; state in registers (lane = Q consecutive label positions), one
//                          __shfl_up per step, log2 domain.  MEET IN THE MIDDLE: up to frame Tb/2 each warp stores its
//                          vectors (alpha_t for t < mid, beta_t for t >= mid); past it, each warp reads the vector the
//                          OTHER warp stored for the frame it is on and emits the posterior occupancy of that frame
//                          directly (labels per position + the sum over blank states).  Half of the lattices are
//                          never stored, the other half is read back while still in L2, and pass 4 reads
//                          L+1 occupancies per frame instead of 2*(2L+1) alpha/beta values.
//   4. ctc_grad_kernel     streaming pass #2 over acts, 8 lanes per row: softmax - occupancy, written once; rows beyond
//                          act_len and infeasible utterances are written as exact zeros
// Label lengths beyond 191 use the multi-warp shared-memory recursion (ctc_alpha_beta_kernel) with full lattices and
// ctc_grad_legacy_kernel.
// Algorithmic HBM bytes: 2*T*B*A*4 (acts read once if pass 4 hits L2, grads written once).
#include <type_traits>
#include "common.cuh"

namespace vocr {

constexpr int kCtcThreads = 256;
constexpr int kCtcWarps = kCtcThreads / 32;

// The recursions run in the log2 domain: one MUFU.EX2 per exp, one MUFU.LG2 per log, no range-reduction code on
// the T-long dependent chain.  Natural-log quantities are converted once per frame (lattice) / utterance (cost).
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2f = 0.6931471805599453f;
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

constexpr int kCtcChunk = 16;  // lattice frames staged per bulk copy (multi-warp recursion kernel)
constexpr int kCtcRenorm = 8;  // ... which renormalises alpha/beta every this many steps
constexpr int kMitmChunk = 8;  // frames per bulk copy in the meet-in-the-middle kernel
// ring buffers per stream: NB - 1 copies in flight cover the bulk-copy latency (~2 us); fewer where a step is long
// anyway and the rings would otherwise limit the CTAs per SM
__host__ __device__ constexpr int mitm_depth(int Q) { return Q <= 2 ? 4 : (Q <= 4 ? 3 : 2); }
constexpr int kMitmMaxQ = 6;   // label positions per lane: L + 1 <= 32 * 6

struct CtcWorkspace {
  int32_t* offsets;  // [B+1]
  int32_t* nxt;      // [sum L] next position with the same symbol, -1 if none
  int32_t* first;    // [sum L] 1 if no earlier position has the same symbol
  double* ll;        // [B] log2-likelihood (-inf = infeasible)
  float* lse;        // [B*T] natural-log row log-sum-exp
  float* lat;        // [B*T*Lp] compact log2-softmax lattice, Lp = round_up(Lmax+1, 4)
  // meet-in-the-middle path (Lmax <= 191)
  float* half;       // [B*T*HS] alpha_t (t < mid_b) or beta_t without the emission (t >= mid_b) in the writer's lane
                     //          layout [2Q][32], + the writer's integer log2 offset; HS = 64Q + 4
  float* occ;        // [B*T*OS] label occupancies in the writer's layout [Q][32], blank-state sum at [32Q]; OS = 32Q+4
  // multi-warp path
  double* offa;      // [B*T] cumulative renormalisation offset of alpha at frame t
  double* offb;      // [B*T] same for beta
  float* alpha;      // [B*T*Smax] renormalised alpha
  float* beta;       // [B*T*Smax] renormalised beta
};

__host__ __device__ inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }
__host__ __device__ inline int lat_stride(int Lmax) { return (Lmax + 1 + 3) & ~3; }
__host__ __device__ inline int mitm_q(int Lmax) { return (Lmax + 1 + 31) / 32; }

__host__ inline size_t ctc_carve(CtcWorkspace* ws, void* base, int T, int B, int max_label_len) {
  const size_t Lmax = (size_t)max_label_len, Smax = 2 * Lmax + 1, Lp = (size_t)lat_stride(max_label_len);
  size_t off = 0;
  unsigned char* p = static_cast<unsigned char*>(base);
  auto take = [&](size_t bytes) {
    unsigned char* r = p ? p + off : nullptr;
    off += align256(bytes);
    return r;
  };
  CtcWorkspace w = {};
  w.offsets = (int32_t*)take(sizeof(int32_t) * ((size_t)B + 1));
  w.nxt = (int32_t*)take(sizeof(int32_t) * (size_t)B * Lmax + 4);
  w.first = (int32_t*)take(sizeof(int32_t) * (size_t)B * Lmax + 4);
  w.ll = (double*)take(sizeof(double) * (size_t)B);
  w.lse = (float*)take(sizeof(float) * (size_t)B * T);
  w.lat = (float*)take(sizeof(float) * (size_t)B * T * Lp);
  const int Q = mitm_q(max_label_len);
  if (Q <= kMitmMaxQ) {
    w.half = (float*)take(sizeof(float) * (size_t)B * T * (64 * Q + 4));
    w.occ = (float*)take(sizeof(float) * (size_t)B * T * (32 * Q + 4));
  } else {
    w.offa = (double*)take(sizeof(double) * (size_t)B * T);
    w.offb = (double*)take(sizeof(double) * (size_t)B * T);
    w.alpha = (float*)take(sizeof(float) * (size_t)B * T * Smax);
    w.beta = (float*)take(sizeof(float) * (size_t)B * T * Smax);
  }
  if (ws) *ws = w;
  return off;
}

// ---- 1. label offsets: exclusive scan of label_lens, one CTA, warp-shuffle scan ----------------------------------------
__global__ void __launch_bounds__(1024)
ctc_scan_kernel(const int32_t* __restrict__ label_lens, int B, int Lmax, CtcWorkspace ws) {
  __shared__ int s_warp[32];
  const int per = ceil_div(B, (int)blockDim.x);
  const int lo = min(B, (int)threadIdx.x * per), hi = min(B, lo + per);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  int sum = 0;
  for (int i = lo; i < hi; ++i) sum += max(0, min(__ldg(label_lens + i), Lmax));
  int v = sum;  // inclusive scan inside the warp
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += n;
  }
  if (lane == 31) s_warp[warp] = v;
  __syncthreads();
  if (warp == 0) {
    int w = (lane < nwarps) ? s_warp[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += n;
    }
    s_warp[lane] = w;
  }
  __syncthreads();
  int run = (warp > 0 ? s_warp[warp - 1] : 0) + v - sum;  // exclusive
  for (int i = lo; i < hi; ++i) {
    ws.offsets[i] = run;
    run += max(0, min(__ldg(label_lens + i), Lmax));
  }
  if (threadIdx.x == blockDim.x - 1) ws.offsets[B] = s_warp[nwarps - 1];
}

// ---- 2. row LSE + compact lattice: 8 lanes per row, 32 rows in flight per CTA ------------------------------------------
// A row (A floats, contiguous) is read with up to four independent 16-byte loads per lane, parked in a shared-memory row
// buffer, reduced with three shuffles per reduction inside the 8-lane group, and the L+1 lattice entries of its utterance
// are gathered from the buffer.  VEC = rows are 16-byte aligned (A % 4 == 0 and an aligned base).
__device__ __forceinline__ float group8_max(float v, unsigned mask) {
  v = fmaxf(v, __shfl_xor_sync(mask, v, 4));
  v = fmaxf(v, __shfl_xor_sync(mask, v, 2));
  return fmaxf(v, __shfl_xor_sync(mask, v, 1));
}
__device__ __forceinline__ float group8_sum(float v, unsigned mask) {
  v += __shfl_xor_sync(mask, v, 4);
  v += __shfl_xor_sync(mask, v, 2);
  return v + __shfl_xor_sync(mask, v, 1);
}
// stage one row into the group's buffer; returns this lane's maximum.  f(x) is applied to every element first.
template <bool VEC, typename F>
__device__ __forceinline__ float group8_stage_row(const float* __restrict__ row, float* rb, int A, int l8, F f) {
  float m = -INFINITY;
  if (VEC) {
    const float4* row4 = reinterpret_cast<const float4*>(row);
    float4* rb4 = reinterpret_cast<float4*>(rb);
    const int n4 = A >> 2;
    for (int base = 0; base < n4; base += 32) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = base + u * 8 + l8;
        if (i < n4) v[u] = ldg_stream_f4(row4 + i);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = base + u * 8 + l8;
        if (i < n4) {
          float4 w = v[u];
          w.x = f(w.x); w.y = f(w.y); w.z = f(w.z); w.w = f(w.w);
          rb4[i] = w;
          m = fmaxf(m, fmaxf(fmaxf(w.x, w.y), fmaxf(w.z, w.w)));
        }
      }
    }
  } else {
    for (int base = 0; base < A; base += 32) {
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = base + u * 8 + l8;
        if (i < A) v[u] = __ldg(row + i);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = base + u * 8 + l8;
        if (i < A) {
          const float w = f(v[u]);
          rb[i] = w;
          m = fmaxf(m, w);
        }
      }
    }
  }
  return m;
}

// One CTA = one utterance x a range of frames: the utterance's symbols are staged in shared memory once, so the row loop
// has no dependent global loads besides the row itself.
template <bool VEC>
__global__ void __launch_bounds__(kCtcThreads)
ctc_lattice_kernel(const float* __restrict__ acts, int T, int B, int A, int frames_per_cta,
                   const int32_t* __restrict__ labels, const int32_t* __restrict__ label_lens,
                   const int32_t* __restrict__ act_lens, int Lmax, CtcWorkspace ws) {
  extern __shared__ __align__(16) float ctc_rowbuf[];  // [G groups][Ap] | int sym[Lmax + 1]
  const int Ap = (A + 3) & ~3;
  const int G = blockDim.x >> 3;
  const int grp = threadIdx.x >> 3, l8 = threadIdx.x & 7;
  const unsigned gmask = 0xffu << (threadIdx.x & 24);
  float* rb = ctc_rowbuf + (size_t)grp * Ap;
  int* sym = reinterpret_cast<int*>(ctc_rowbuf + (size_t)G * Ap);
  const int b = blockIdx.x, t0 = blockIdx.y * frames_per_cta;
  const int Tb = min(__ldg(act_lens + b), T);
  if (t0 >= Tb) return;
  const int L = max(0, min(__ldg(label_lens + b), Lmax));
  const int off = ws.offsets[b];
  for (int j = threadIdx.x; j <= L; j += blockDim.x) sym[j] = (j == 0) ? 0 : max(0, min(__ldg(labels + off + j - 1), A - 1));
  __syncthreads();
  const int t1 = min(Tb, t0 + frames_per_cta);
  const int Lp = lat_stride(Lmax);
  for (int t = t0 + grp; t < t1; t += G) {
    float m = group8_stage_row<VEC>(acts + ((size_t)t * B + b) * A, rb, A, l8, [](float x) { return x; });
    m = group8_max(m, gmask);
    const float mk = m * kLog2e;
    float s = 0.f;
    if (VEC) {
      const float4* rb4 = reinterpret_cast<const float4*>(rb);
      for (int i = l8; i < (A >> 2); i += 8) {  // this lane's own elements: no synchronisation needed yet
        const float4 w = rb4[i];
        s += ex2_fast(fmaf(w.x, kLog2e, -mk)) + ex2_fast(fmaf(w.y, kLog2e, -mk)) + ex2_fast(fmaf(w.z, kLog2e, -mk)) +
             ex2_fast(fmaf(w.w, kLog2e, -mk));
      }
    } else {
      for (int i = l8; i < A; i += 8) s += ex2_fast(fmaf(rb[i], kLog2e, -mk));
    }
    s = group8_sum(s, gmask);
    const float lse = m + kLn2f * __log2f(s);
    const size_t bt = (size_t)b * T + t;
    if (l8 == 0) ws.lse[bt] = lse;
    float* lat = ws.lat + bt * (size_t)Lp;
    __syncwarp(gmask);  // the whole row is in the buffer
    for (int j = l8; j <= L; j += 8) lat[j] = (rb[sym[j]] - lse) * kLog2e;
    __syncwarp(gmask);  // gathers done before the next row overwrites the buffer
  }
}

// ---- 3L. alpha / beta recursions, multi-warp form (label lengths > 191) ----------------------------------------------------------------------------
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

// ---- 3. alpha / beta recursions, one WARP per direction, meeting in the middle --------------------------------------------
// Lane i owns Q consecutive label positions j = i*Q .. i*Q+Q-1 of ITS direction's labelling, i.e. the state pairs
// (blank 2j, label 2j+1) of the blank-extended labelling, in registers.  A step needs exactly one remote value per lane
// - the previous lane's last label state - fetched with one __shfl_up: no shared-memory exchange and no block barrier on
// the T-long dependent chain.  The beta recursion is the alpha recursion of the REVERSED labelling walked backwards in
// time (same code), so the two warps synchronise exactly once, at the middle frame.
// A lone warp per scheduler issues in order: every instruction of the step is on the clock, not only the data-dependent
// chain (measured: 396 cycles per step with address arithmetic, chunk bookkeeping and conversions in the loop body).
// The step is therefore written for instruction count:
//   * "-inf" is the finite sentinel kNeg (absorbing under +, ex2 -> 0): no NaN guards or selects on the chain.
//   * log2(2^a + 2^b [+ 2^c]) = max + lg2(1 + ex2(. - max) [+ ex2(. - max)]): 2 / 3 MUFU operations.
//   * Renormalisation: every second step the warp maximum is taken with one REDUX on order-preserving integer keys and
//     its floor is subtracted two steps LATER, folded into the emission term - off the dependent chain, and the
//     cumulative offset is an exact int32.
//   * Lattice rows (and, past the middle, the other warp's stored vectors) stream through shared-memory rings filled by
//     1-D bulk async copies, NB - 1 chunks ahead; the 8 steps of a chunk are unrolled, so every shared-memory load has a
//     compile-time offset from one running row index and the compiler hoists them across the steps of the chunk.
// dynamic smem (floats): [mbarriers: 32 floats][lattice ring 2xNBxCHxLp][partner ring 2xNBxCHxHS][lab Lmax]
constexpr float kNeg = -1.0e30f;
constexpr float kNegTest = -1.0e29f;  // anything below is "impossible"

__device__ __forceinline__ float lse2n(float a, float b) {
  const float m = fmaxf(a, b), n = fminf(a, b);
  return m + lg2_fast(1.f + ex2_fast(n - m));
}
__device__ __forceinline__ float lse3n(float a, float b, float c) {
  const float hi = fmaxf(a, b), lo = fminf(a, b);
  const float m = fmaxf(hi, c), x = fminf(hi, c);
  return m + lg2_fast(1.f + ex2_fast(x - m) + ex2_fast(lo - m));
}
__device__ __forceinline__ int float_key(float f) {  // monotone float -> int
  const int b = __float_as_int(f);
  return b >= 0 ? b : b ^ 0x7fffffff;
}
__device__ __forceinline__ float key_float(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }

template <int Q>
__global__ void __launch_bounds__(64)
ctc_mitm_kernel(const int32_t* __restrict__ labels, const int32_t* __restrict__ label_lens,
                const int32_t* __restrict__ act_lens, int T, int B, int A, int Lmax, CtcWorkspace ws,
                float* __restrict__ costs) {
  extern __shared__ __align__(128) float smf[];
  constexpr int CH = kMitmChunk, NB = mitm_depth(Q);
  constexpr int HS = 64 * Q + 4, OS = 32 * Q + 4;
  constexpr unsigned kFull = 0xffffffffu;
  const int Lp = lat_stride(Lmax);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smf);  // [2 dir][lat 0..NB-1, par 0..NB-1]
  const int kLatBase = 32;                            // float index of the lattice ring [2][NB][CH*Lp]
  const int kParBase = kLatBase + 2 * NB * CH * Lp;   // ... of the partner ring [2][NB][CH*HS]
  int* lab = reinterpret_cast<int*>(smf + kParBase + 2 * NB * CH * HS);  // [Lmax]
  const int b = blockIdx.x;
  const int L = max(0, min(label_lens[b], Lmax));
  const int Tb = max(0, min(act_lens[b], T));
  const int dirn = threadIdx.x >> 5;  // 0 alpha, 1 beta
  const int lane = threadIdx.x & 31;
  const int off = ws.offsets[b];
  for (int j = threadIdx.x; j < L; j += blockDim.x) lab[j] = max(0, min(labels[off + j], A - 1));
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4 * NB; ++i) mbar_init(&bars[i], 1);
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
      ws.ll[b] = (L == 0) ? 0.0 : -INFINITY;  // zero frames: only the empty labelling is feasible; cost 0 either way
      costs[b] = 0.f;
    }
    return;
  }
  // per-lane constants.  Position j (in THIS direction's order) carries a blank state and, if j < L, a label state.
  bool has_blank[Q], has_label[Q], skip[Q];
  int slot[Q], pidx_b[Q], pidx_l[Q];
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
    // the same states in the OTHER warp's order: blank L - j, label L - 1 - j; its row layout is [2Q][32]
    const int jb = has_blank[i] ? L - j : 0, jl = has_label[i] ? L - 1 - j : 0;
    pidx_b[i] = (jb % Q) * 32 + jb / Q;
    pidx_l[i] = (Q + jl % Q) * 32 + jl / Q;
  }
  const bool skip0 = skip[0] && lane != 0;
  const int mid = Tb >> 1;
  const int n1 = (dirn == 0) ? mid : Tb - mid;  // steps before the meeting point (vectors stored)
  const int n2 = Tb - n1;                       // steps past it (occupancies emitted)
  const int G1 = ceil_div(n1, CH), G2 = ceil_div(n2, CH);  // chunks of the two phases (each phase starts a chunk)
  const int mylat = kLatBase + dirn * NB * CH * Lp;
  const int mypar = kParBase + dirn * NB * CH * HS;
  uint64_t* latbars = bars + dirn * 2 * NB;
  uint64_t* parbars = latbars + NB;
  const float* lat_g = ws.lat + (size_t)b * T * Lp;
  float* half_g = ws.half + (size_t)b * T * HS;
  float* occ_g = ws.occ + (size_t)b * T * OS;

  // Ring discipline (both streams): chunk g lives in buffer g % NB; when the warp starts on chunk g it refills the
  // buffer of chunk g - 1 with chunk g + NB - 1.  The refill is ordered after the reads of chunk g - 1 by program
  // order + __syncwarp (no proxy fence: a fence here drains the lane's outstanding global stores, ~1 us per chunk).
  auto chunk_steps = [&](int g, int& k0, int& cnt) {  // lattice chunk g (phase 1 chunks, then phase 2 chunks)
    if (g < G1) {
      k0 = g * CH;
      cnt = min(CH, n1 - k0);
    } else {
      k0 = n1 + (g - G1) * CH;
      cnt = min(CH, Tb - k0);
    }
  };
  auto issue_lat = [&](int g) {  // lane 0
    int k0, cnt;
    chunk_steps(g, k0, cnt);
    const int t_lo = (dirn == 0) ? k0 : Tb - k0 - cnt;
    const uint32_t bytes = (uint32_t)cnt * (uint32_t)Lp * 4u;
    mbar_arrive_expect_tx(&latbars[g % NB], bytes);
    bulk_g2s(smf + mylat + (g % NB) * CH * Lp, lat_g + (size_t)t_lo * Lp, bytes, &latbars[g % NB]);
  };
  auto issue_par = [&](int c) {  // lane 0: the other warp's rows of second-half steps [c*CH, (c+1)*CH)
    const int q0 = c * CH, cnt = min(CH, n2 - q0);
    const int t_lo = (dirn == 0) ? mid + q0 : mid - q0 - cnt;
    const uint32_t bytes = (uint32_t)cnt * (uint32_t)HS * 4u;
    mbar_arrive_expect_tx(&parbars[c % NB], bytes);
    bulk_g2s(smf + mypar + (c % NB) * CH * HS, half_g + (size_t)t_lo * HS, bytes, &parbars[c % NB]);
  };

  float bl[Q], lb[Q];  // virtual state before the first frame: all mass on the first blank
#pragma unroll
  for (int i = 0; i < Q; ++i) {
    bl[i] = (lane == 0 && i == 0) ? 0.f : kNeg;
    lb[i] = kNeg;
  }
  int E = 0;        // log2 alpha = state + E  (exact integer offset)
  int rpend = 0;    // measured at the previous even step, applied at the next one
  int Eref = 0;     // occupancy normalisation: log2 p(labels) = Eref + cref as seen from the first combined frame
  float cref = 0.f;
  if (lane == 0)
    for (int g = 0; g < min(NB - 1, G1 + G2); ++g) issue_lat(g);

  // one phase = the chunks [gbeg, gend) of the lattice stream
  auto run = [&](auto second_tag, int gbeg, int gend) {
    constexpr bool SECOND = decltype(second_tag)::value;
#pragma unroll 1
    for (int g = gbeg; g < gend; ++g) {
      int k0, cnt;
      chunk_steps(g, k0, cnt);
      __syncwarp();  // every lane is done with the previous chunk
      if (lane == 0 && g + NB - 1 < G1 + G2) issue_lat(g + NB - 1);
      mbar_wait_or_trap(&latbars[g % NB], (uint32_t)((g / NB) & 1));
      // rows of a chunk are stored in ascending t: alpha walks them upwards, beta downwards
      const int rstep = (dirn == 0) ? Lp : -Lp;
      int lrow = mylat + (g % NB) * CH * Lp + ((dirn == 0) ? 0 : (cnt - 1) * Lp);
      const int t0 = (dirn == 0) ? k0 : Tb - 1 - k0;
      float* hrow = half_g + (size_t)t0 * HS + lane;   // this lane's column of the stored vectors
      float* orow = occ_g + (size_t)t0 * OS + lane;    // ... and of the occupancies
      const int hstep = (dirn == 0) ? HS : -HS, ostep = (dirn == 0) ? OS : -OS;
      int prow = 0;
      if constexpr (SECOND) {
        const int c = g - G1;
        if (lane == 0 && c + NB - 1 < G2) issue_par(c + NB - 1);
        mbar_wait_or_trap(&parbars[c % NB], (uint32_t)((c / NB) & 1));
        prow = mypar + (c % NB) * CH * HS + ((dirn == 0) ? 0 : (cnt - 1) * HS);
      }
      // lattice values of the first step; inside the chunk the values of step u + 1 are loaded during step u
      float lpb = smf[lrow], lpl[Q];
#pragma unroll
      for (int i = 0; i < Q; ++i) lpl[i] = smf[lrow + slot[i]];

      auto step = [&](auto u_tag) {
        constexpr int u = decltype(u_tag)::value;
        float nlpb = 0.f, nlpl[Q];
#pragma unroll
        for (int i = 0; i < Q; ++i) nlpl[i] = 0.f;
        if constexpr (u + 1 < CH) {  // (a row past a short chunk is still inside the ring: loaded, never used)
          nlpb = smf[lrow + rstep];
#pragma unroll
          for (int i = 0; i < Q; ++i) nlpl[i] = smf[lrow + rstep + slot[i]];
        }
        float thb[Q], thl[Q];
        int Eth = 0;
        if constexpr (SECOND) {
#pragma unroll
          for (int i = 0; i < Q; ++i) {
            thb[i] = smf[prow + pidx_b[i]];
            thl[i] = smf[prow + pidx_l[i]];
          }
          Eth = __float_as_int(smf[prow + 64 * Q]);
        }
        // emission terms with the pending renormalisation folded in
        const int Ebefore = E;
        float rf = 0.f;
        if constexpr ((u & 1) == 0) {
          rf = (float)rpend;
          E += rpend;
        }
        float cb[Q], cl[Q];
#pragma unroll
        for (int i = 0; i < Q; ++i) {
          cb[i] = has_blank[i] ? lpb - rf : kNeg;
          cl[i] = has_label[i] ? lpl[i] - rf : kNeg;
        }
        // the recursion
        const float sh = __shfl_up_sync(kFull, lb[Q - 1], 1);  // previous lane's last label state
        float pl = (lane == 0) ? kNeg : sh;
        float pb[Q], pq[Q];  // sums over the predecessors, without this frame's emission
#pragma unroll
        for (int i = 0; i < Q; ++i) {
          const float ps = (i == 0) ? (skip0 ? sh : kNeg) : (skip[i] ? pl : kNeg);
          pb[i] = lse2n(bl[i], pl);
          pq[i] = lse3n(lb[i], bl[i], ps);
          pl = lb[i];  // old label state of this position feeds the next position
        }
#pragma unroll
        for (int i = 0; i < Q; ++i) {
          bl[i] = pb[i] + cb[i];
          lb[i] = pq[i] + cl[i];
        }
        if constexpr ((u & 1) == 0) {  // measure now, apply two steps later
          int key = float_key(bl[0]);
#pragma unroll
          for (int i = 0; i < Q; ++i) key = max(key, max(float_key(bl[i]), float_key(lb[i])));
          const float m = key_float(__reduce_max_sync(kFull, key));
          rpend = (m > kNegTest) ? __float2int_rd(m) : 0;
        }
        // alpha keeps the emission of its frame, beta does not: alpha_t(s) * beta_t(s) is then the path mass through (t, s)
        float vb[Q], vl[Q];
#pragma unroll
        for (int i = 0; i < Q; ++i) {
          vb[i] = (dirn == 0) ? bl[i] : pb[i];
          vl[i] = (dirn == 0) ? lb[i] : pq[i];
        }
        const int Emine = (dirn == 0) ? E : Ebefore;
        if constexpr (!SECOND) {
#pragma unroll
          for (int i = 0; i < Q; ++i) {
            hrow[i * 32] = vb[i];
            hrow[(Q + i) * 32] = vl[i];
          }
          if (lane == 0) hrow[64 * Q] = __int_as_float(Emine);
          hrow += hstep;
        } else {
          float gb[Q], gl[Q];
#pragma unroll
          for (int i = 0; i < Q; ++i) {
            gb[i] = has_blank[i] ? vb[i] + thb[i] : kNeg;
            gl[i] = has_label[i] ? vl[i] + thl[i] : kNeg;
          }
          const int Esum = Emine + Eth;
          if (u == 0 && g == G1) {  // first combined frame: log2 p(labels) = Esum + log2 sum_s 2^(alpha + beta)
            float mx = kNeg;
#pragma unroll
            for (int i = 0; i < Q; ++i) mx = fmaxf(mx, fmaxf(gb[i], gl[i]));
            mx = warp_max(mx);
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i < Q; ++i) sum += ex2_fast(gb[i] - mx) + ex2_fast(gl[i] - mx);
            sum = warp_sum(sum);
            Eref = Esum;
            cref = (mx > kNegTest) ? mx + lg2_fast(sum) : 0.f;
          }
          const float D = (float)(Esum - Eref) - cref;
          unsigned fx = 0;
#pragma unroll
          for (int i = 0; i < Q; ++i) {
            gb[i] = ex2_fast(gb[i] + D);
            gl[i] = ex2_fast(gl[i] + D);
            fx += __float2uint_rn(fminf(gb[i], 2.f) * 1073741824.f);
          }
          fx = __reduce_add_sync(kFull, fx);  // sum over the blank states in 2^-30 fixed point: one REDUX, deterministic
#pragma unroll
          for (int i = 0; i < Q; ++i) orow[i * 32] = gl[i];
          if (lane == 0) orow[32 * Q] = (float)fx * (1.f / 1073741824.f);
          orow += ostep;
          prow += (dirn == 0) ? HS : -HS;
        }
        lrow += rstep;
        lpb = nlpb;
#pragma unroll
        for (int i = 0; i < Q; ++i) lpl[i] = nlpl[i];
      };
      if (cnt == CH) {  // full chunk: eight steps, no exit tests
        step(std::integral_constant<int, 0>{});
        step(std::integral_constant<int, 1>{});
        step(std::integral_constant<int, 2>{});
        step(std::integral_constant<int, 3>{});
        step(std::integral_constant<int, 4>{});
        step(std::integral_constant<int, 5>{});
        step(std::integral_constant<int, 6>{});
        step(std::integral_constant<int, 7>{});
      } else {  // the last chunk of a phase
        step(std::integral_constant<int, 0>{});
        if (cnt > 1) step(std::integral_constant<int, 1>{});
        if (cnt > 2) step(std::integral_constant<int, 2>{});
        if (cnt > 3) step(std::integral_constant<int, 3>{});
        if (cnt > 4) step(std::integral_constant<int, 4>{});
        if (cnt > 5) step(std::integral_constant<int, 5>{});
        if (cnt > 6) step(std::integral_constant<int, 6>{});
      }
    }
  };

#ifdef VOCR_CTC_PROF
  long long prof_t1 = clock64();
#endif
  run(std::false_type{}, 0, G1);
#ifdef VOCR_CTC_PROF
  long long prof_t2 = clock64();
#endif
  // the other warp reads these rows through the async proxy (bulk copies)
  __threadfence_block();
  asm volatile("fence.proxy.async.global;" ::: "memory");
  __syncthreads();
  if (lane == 0)
    for (int c = 0; c < min(NB - 1, G2); ++c) issue_par(c);
#ifdef VOCR_CTC_PROF
  long long prof_t3 = clock64();
#endif
  run(std::true_type{}, G1, G1 + G2);
#ifdef VOCR_CTC_PROF
  if (lane == 0 && b < 4) {  // cycles: first half, barrier, second half, steps (read back from the unused tail of ws.lse)
    float* dbg = ws.lse + (size_t)B * T - 64 + (b * 2 + dirn) * 4;
    dbg[0] = (float)(prof_t2 - prof_t1);
    dbg[1] = (float)(prof_t3 - prof_t2);
    dbg[2] = (float)(clock64() - prof_t3);
    dbg[3] = (float)n1;
  }
#endif

  if (dirn == 0) {
    // alpha_{Tb-1}(S-1) is the blank of position L, alpha_{Tb-1}(S-2) the label of position L-1
    float fb = kNeg, fl = kNeg;
#pragma unroll
    for (int i = 0; i < Q; ++i) {
      const int j = lane * Q + i;
      if (j == L) fb = bl[i];
      if (j == L - 1) fl = lb[i];
    }
    fb = warp_max(fb);
    fl = warp_max(fl);
    if (lane == 0) {
      const float l = lse2n(fb, fl);
      const bool feasible = l > kNegTest;
      const double ll2 = feasible ? (double)E + (double)l : -INFINITY;
      ws.ll[b] = ll2;
      costs[b] = feasible ? (float)(-ll2 * kLn2) : 0.f;
    }
  }
}

// ---- 4. gradient from the occupancies: 8 lanes per row, one CTA = one utterance x a range of frames ------------------------
// The utterance's symbols, duplicate chains and occupancy slots are staged in shared memory once; per row the acts and the
// occupancy row are requested together, and the duplicate-symbol chains are walked in shared memory.
template <bool VEC>
__global__ void __launch_bounds__(kCtcThreads, 4)
ctc_grad_kernel(const float* __restrict__ acts, float* __restrict__ grads, int T, int B, int A, int frames_per_cta,
                const int32_t* __restrict__ labels, const int32_t* __restrict__ label_lens,
                const int32_t* __restrict__ act_lens, int Lmax, int Q, CtcWorkspace ws) {
  extern __shared__ __align__(16) float ctc_rowbuf[];  // [G][Ap] | occupancy rows [G][OS] | int sym, first, nxt, slotA, slotB [Lmax]
  const int Ap = (A + 3) & ~3;
  const int OS = 32 * Q + 4;
  const int G = blockDim.x >> 3;
  const int grp = threadIdx.x >> 3, l8 = threadIdx.x & 7;
  const unsigned gmask = 0xffu << (threadIdx.x & 24);
  float* rb = ctc_rowbuf + (size_t)grp * Ap;
  float* oc = ctc_rowbuf + (size_t)G * Ap + (size_t)grp * OS;
  int* sym = reinterpret_cast<int*>(ctc_rowbuf + (size_t)G * (Ap + OS));
  int* first = sym + Lmax;
  int* nxt = first + Lmax;
  int* slotA = nxt + Lmax;
  int* slotB = slotA + Lmax;
  const int b = blockIdx.x, t0 = blockIdx.y * frames_per_cta;
  const int t_end = min(T, t0 + frames_per_cta);
  const int Tb = min(__ldg(act_lens + b), T);
  const bool feasible = ws.ll[b] != -INFINITY;
  const int t1 = feasible ? min(Tb, t_end) : t0;  // frames [t1, t_end) are written as zeros
  const int L = max(0, min(__ldg(label_lens + b), Lmax));
  if (t0 < t1) {
    const int off = ws.offsets[b];
    for (int j = threadIdx.x; j < L; j += blockDim.x) {
      sym[j] = max(0, min(__ldg(labels + off + j), A - 1));
      first[j] = ws.first[off + j];
      nxt[j] = ws.nxt[off + j];
      const int jr = L - 1 - j;
      slotA[j] = (j % Q) * 32 + j / Q;     // position j as written by the alpha warp (frames >= Tb/2)
      slotB[j] = (jr % Q) * 32 + jr / Q;   // ... by the beta warp, whose position order is reversed
    }
    __syncthreads();
  }
  const int mid = Tb >> 1;
  for (int t = t0 + grp; t < t_end; t += G) {
    float* grow = grads + ((size_t)t * B + b) * A;
    if (t >= t1) {  // uniform over the 8-lane group
      if (VEC) {
        for (int i = l8; i < (A >> 2); i += 8) stg_stream_f4(reinterpret_cast<float4*>(grow) + i, make_float4(0.f, 0.f, 0.f, 0.f));
      } else {
        for (int i = l8; i < A; i += 8) grow[i] = 0.f;
      }
      continue;
    }
    const size_t bt = (size_t)b * T + t;
    // request the occupancy row (asynchronous 16-byte copies straight into shared memory) and the log-sum-exp before
    // touching the acts: all of this row's loads are in flight together
    {
      const float4* occ4 = reinterpret_cast<const float4*>(ws.occ + bt * (size_t)OS);
      const uint32_t dst = smem_u32(oc);
      for (int i = l8; i < (OS >> 2); i += 8)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)i * 16u), "l"(occ4 + i) : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    const float nl2 = -ws.lse[bt] * kLog2e;
    group8_stage_row<VEC>(acts + ((size_t)t * B + b) * A, rb, A, l8, [nl2](float x) { return ex2_fast(fmaf(x, kLog2e, nl2)); });
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp(gmask);
    const int* slot = (t >= mid) ? slotA : slotB;
    for (int jo = l8; jo < L; jo += 8) {
      if (first[jo]) {
        float tot = 0.f;
        for (int q = jo; q >= 0; q = nxt[q]) tot += oc[slot[q]];
        rb[sym[jo]] -= tot;
      }
    }
    __syncwarp(gmask);
    if (l8 == 0) rb[0] -= oc[32 * Q];
    __syncwarp(gmask);
    if (VEC) {
      const float4* rb4 = reinterpret_cast<const float4*>(rb);
      for (int i = l8; i < (A >> 2); i += 8) stg_stream_f4(reinterpret_cast<float4*>(grow) + i, rb4[i]);
    } else {
      for (int i = l8; i < A; i += 8) grow[i] = rb[i];
    }
    __syncwarp(gmask);
  }
}

// ---- 4L. gradient from full alpha / beta lattices (multi-warp path) ------------------------------------------------------------------------------------------
// dynamic smem: 16 B mbarrier + tile + per-warp occupancy scratch kCtcWarps*Smax floats
__global__ void __launch_bounds__(kCtcThreads)
ctc_grad_legacy_kernel(const float* __restrict__ acts, float* __restrict__ grads, long long n_rows, int T, int B, int A,
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

template <int Q>
static int launch_mitm(const int32_t* labels, const int32_t* label_lens, const int32_t* act_lens, int T, int B, int A,
                       int Lmax, const CtcWorkspace& ws, float* costs, cudaStream_t stream) {
  const size_t smem = 128 + sizeof(float) * (size_t)2 * mitm_depth(Q) * kMitmChunk * (lat_stride(Lmax) + 64 * Q + 4) +
                      sizeof(int) * (size_t)(Lmax + 1);
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(ctc_mitm_kernel<Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return VOCR_EXECUTION_FAILED;
  ctc_mitm_kernel<Q><<<B, 64, smem, stream>>>(labels, label_lens, act_lens, T, B, A, Lmax, ws, costs);
  VOCR_CHECK_LAUNCH();
  return VOCR_OK;
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

  ctc_scan_kernel<<<1, min(1024, ((B + 31) / 32) * 32), 0, stream>>>(label_lens, B, Lmax, ws);
  VOCR_CHECK_LAUNCH();

  // streaming passes: 8 lanes per row, G row groups per CTA (32 unless the alphabet is very large), one CTA = one
  // utterance x 2G frames
  const long long n_rows = (long long)T * B;
  const int Ap = (A + 3) & ~3;
  const int Q = mitm_q(Lmax);
  const int OSq = (Q <= kMitmMaxQ) ? 32 * Q + 4 : 0;
  int G = 32;
  while (G > 1 && (size_t)G * (Ap + OSq) * 4 > 64 * 1024) G >>= 1;
  const size_t smem_k1 = (size_t)G * Ap * 4 + sizeof(int) * (size_t)(Lmax + 1);
  const size_t smem_k3 = (size_t)G * (Ap + OSq) * 4 + sizeof(int) * (size_t)5 * Lmax + 16;
  VOCR_REQUIRE(smem_k1 <= 200 * 1024 && smem_k3 <= 200 * 1024);
  const int frames_per_cta = 2 * G;
  const bool vec = (A % 4 == 0) && (reinterpret_cast<uintptr_t>(acts) % 16 == 0) &&
                   (grads == nullptr || reinterpret_cast<uintptr_t>(grads) % 16 == 0);
  const dim3 grid_rows((unsigned)B, (unsigned)ceil_div(max(T, 1), frames_per_cta));
  VOCR_REQUIRE(grid_rows.y <= 65535u);
  static DeviceLatch attr_latch;
  if (attr_latch.need()) {
    if (cudaFuncSetAttribute(ctc_lattice_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(ctc_lattice_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(ctc_grad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(ctc_grad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
      return VOCR_EXECUTION_FAILED;
    attr_latch.set();
  }
  if (n_rows > 0) {
    if (vec)
      ctc_lattice_kernel<true><<<grid_rows, 8 * G, smem_k1, stream>>>(acts, T, B, A, frames_per_cta, labels_safe, label_lens,
                                                                      act_lens, Lmax, ws);
    else
      ctc_lattice_kernel<false><<<grid_rows, 8 * G, smem_k1, stream>>>(acts, T, B, A, frames_per_cta, labels_safe, label_lens,
                                                                       act_lens, Lmax, ws);
    VOCR_CHECK_LAUNCH();
  }
  if (Q <= kMitmMaxQ) {
    int st = VOCR_INVALID_VALUE;
    switch (Q) {
      case 1: st = launch_mitm<1>(labels_safe, label_lens, act_lens, T, B, A, Lmax, ws, costs, stream); break;
      case 2: st = launch_mitm<2>(labels_safe, label_lens, act_lens, T, B, A, Lmax, ws, costs, stream); break;
      case 3: st = launch_mitm<3>(labels_safe, label_lens, act_lens, T, B, A, Lmax, ws, costs, stream); break;
      case 4: st = launch_mitm<4>(labels_safe, label_lens, act_lens, T, B, A, Lmax, ws, costs, stream); break;
      case 5: st = launch_mitm<5>(labels_safe, label_lens, act_lens, T, B, A, Lmax, ws, costs, stream); break;
      case 6: st = launch_mitm<6>(labels_safe, label_lens, act_lens, T, B, A, Lmax, ws, costs, stream); break;
    }
    if (st != VOCR_OK) return st;
    if (grads && n_rows > 0) {
      if (vec)
        ctc_grad_kernel<true><<<grid_rows, 8 * G, smem_k3, stream>>>(acts, grads, T, B, A, frames_per_cta, labels_safe,
                                                                     label_lens, act_lens, Lmax, Q, ws);
      else
        ctc_grad_kernel<false><<<grid_rows, 8 * G, smem_k3, stream>>>(acts, grads, T, B, A, frames_per_cta, labels_safe,
                                                                      label_lens, act_lens, Lmax, Q, ws);
      VOCR_CHECK_LAUNCH();
    }
    return VOCR_OK;
  }
  // label lengths > 191: multi-warp recursion with full lattices
  {
    int Gt = ((Smax + 31) / 32) * 32;
    if (Gt > 512) Gt = 512;
    const size_t smem3 = 32 + sizeof(float) * ((size_t)4 * kCtcChunk * lat_stride(Lmax) + 4 * (size_t)Smax + 64) +
                         sizeof(int) * (size_t)Smax;
    if (smem3 > 48 * 1024) {
      VOCR_REQUIRE(smem3 <= 200 * 1024);
      if (cudaFuncSetAttribute(ctc_alpha_beta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3) !=
          cudaSuccess)
        return VOCR_EXECUTION_FAILED;
    }
    ctc_alpha_beta_kernel<<<B, 2 * Gt, smem3, stream>>>(labels_safe, label_lens, act_lens, T, B, A, Lmax, Gt, ws,
                                                        costs);
    VOCR_CHECK_LAUNCH();
  }
  if (grads && n_rows > 0) {
    int rows_per_tile = 32;
    while (rows_per_tile > 4 && (size_t)rows_per_tile * A * 4 > 32768) rows_per_tile >>= 1;
    const size_t tile_bytes = ((size_t)rows_per_tile * A * 4 + 15) & ~size_t(15);
    const bool base_aligned = (reinterpret_cast<uintptr_t>(acts) % 16 == 0) && (reinterpret_cast<uintptr_t>(grads) % 16 == 0);
    const long long n_tiles = ceil_div64(n_rows, rows_per_tile);
    VOCR_REQUIRE(n_tiles < (1ll << 31));
    const size_t smem4 = 16 + tile_bytes + sizeof(float) * (size_t)kCtcWarps * Smax + 16;
    if (smem4 > 48 * 1024) {
      VOCR_REQUIRE(smem4 <= 200 * 1024);
      if (cudaFuncSetAttribute(ctc_grad_legacy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4) !=
          cudaSuccess)
        return VOCR_EXECUTION_FAILED;
    }
    ctc_grad_legacy_kernel<<<(unsigned)n_tiles, kCtcThreads, smem4, stream>>>(acts, grads, n_rows, T, B, A, rows_per_tile,
                                                                              base_aligned, labels_safe, label_lens,
                                                                              act_lens, Lmax, ws);
    VOCR_CHECK_LAUNCH();
  }
  return VOCR_OK;
}
