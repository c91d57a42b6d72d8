// 3x3 / pad 1 convolutions on the 5th-generation tensor cores: implicit GEMM fed by 4-D TMA tiles of the NHWC
// activations (the conv halo is the TMA unit's out-of-bounds zero fill - no im2col buffer, no index arithmetic in the
// kernel), tcgen05.mma kind::tf32 with TMEM accumulators, error-compensated 3xTF32 (see tc_gemm.cu) so the result
// keeps fp32-level accuracy.  Replaces the FFMA implicit-GEMM kernels of conv.cu for Cin % 32 == 0
// (reference cnnlstm.py:124-134 -> cuDNN).
//
//   forward / data gradient   Z[p, co] = sum_{tap, ci} X[p + tap, ci] * Wn[co, (tap, ci)]
//       A tile = 128 output pixels (BH rows x BW columns of one image) x 32 input channels of one tap:
//                one 4-D box {32 c, BW, BH, 1} at (c0, x0+kx-1, y0+ky-1, b) -> 128 rows x 128 B, K-major SWIZZLE_128B
//       B tile = BN output channels x 32 k: 2-D box of the K-major weight matrix [Cout][9*Cin]
//       K loop = 9 taps x Cin/32 channel chunks (<= 72 k-blocks); hi*hi rotates over 3 TMEM accumulators
//   weight gradient           dW[(tap, ci), co] = sum_p X[p + tap, ci] * dZ[p, co]        (reduction over ALL pixels)
//       both operands are MN-major (channels contiguous, pixels = K rows): boxes {32 c, 32 x, 1, 1}, 128B_ATOM_32B
//       split-K over image rows across CTAs; inside a CTA the reduction is cut into chunks of 8 k-blocks: each chunk
//       accumulates in TMEM, is drained by the epilogue warps into fp32 registers (round-to-nearest adds) while the
//       next chunk runs on another accumulator.  Without this the tensor core's round-toward-zero accumulation biases
//       a 10^5-term sum by ~1e-4 relative; with it the bias stays ~1e-6.
#include "tc_common.cuh"

namespace vocr {

constexpr int kCvThreads = 192;
constexpr int kCvStages = 3;
constexpr int kCvTile = 16384;                 // 128 rows x 128 B
constexpr int kCvStageBytes = 4 * kCvTile;     // A_hi, A_lo, B_hi, B_lo (B uses BN*128 B of its 16 KB)
constexpr int kCvSmemBytes = kCvStages * kCvStageBytes + 1024 + 256;
constexpr uint32_t kCvTmemCols = 512;
constexpr int kCvHiAcc = 3;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct TcConvParams {
  float* z;
  const float* bias;
  int B, H, W, Cin, Cout;
  int BW, BH, tiles_x, tiles_y, BN;
  const int* exp_x;  // FP16 pair operands: device exponents of the activation / weight planes
  const int* exp_w;
  int single;        // 1 = hi planes only, one product per k-step (vocr_set_tc_products(1))
  unsigned* zmax;    // optional: max |z| over the whole output (atomicMax on the bit pattern of the non-negative float)
  // Fused inference epilogue (vocr_tc_conv3x3_bnrelu_f16): a = relu(z * bn_scale[c] + bn_shift[c]) instead of z.
  // z (may be null) then receives a with the strides below; a_hi / a_lo (may be null) receive its FP16 pair planes,
  // dense NHWC, scaled by 2^e with e from the device bound `a_bound` (e is stored to a_exp).
  const float* bn_scale;
  const float* bn_shift;
  long long sB, sH, sW;
  __half* a_hi;
  __half* a_lo;
  const unsigned* a_bound;
  int* a_exp;
};

// epilogue of one float4 group (4 consecutive output channels n..n+3 of one pixel): bias, optional BatchNorm (running
// statistics) + ReLU, fp32 store and / or FP16 pair planes
__device__ __forceinline__ void cv_store4(const TcConvParams& p, float (&v)[4], int n, float* zrow, size_t prow,
                                          float psc, float& vmax) {
  if (p.bias) {
    const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + n));
    v[0] += bv.x; v[1] += bv.y; v[2] += bv.z; v[3] += bv.w;
  }
  if (p.bn_scale) {
    const float4 sc = __ldg(reinterpret_cast<const float4*>(p.bn_scale + n));
    const float4 sh = __ldg(reinterpret_cast<const float4*>(p.bn_shift + n));
    v[0] = fmaxf(fmaf(v[0], sc.x, sh.x), 0.f);
    v[1] = fmaxf(fmaf(v[1], sc.y, sh.y), 0.f);
    v[2] = fmaxf(fmaf(v[2], sc.z, sh.z), 0.f);
    v[3] = fmaxf(fmaf(v[3], sc.w, sh.w), 0.f);
  }
  // max |z|, or max a with the fused epilogue (the measured bound the NEXT layer's analytic bound starts from)
  vmax = fmaxf(vmax, fmaxf(fmaxf(fabsf(v[0]), fabsf(v[1])), fmaxf(fabsf(v[2]), fabsf(v[3]))));
  if (zrow) *reinterpret_cast<float4*>(zrow + n) = make_float4(v[0], v[1], v[2], v[3]);
  if (p.a_hi) {
    uint2 ph, pl;
    pair_pack4(v[0] * psc, v[1] * psc, v[2] * psc, v[3] * psc, ph, pl);
    *reinterpret_cast<uint2*>(p.a_hi + prow + n) = ph;
    *reinterpret_cast<uint2*>(p.a_lo + prow + n) = pl;
  }
}

template <bool F16>
__global__ void __launch_bounds__(kCvThreads, 1)
tc_conv_fwd_kernel(const __grid_constant__ CUtensorMap map_x_hi, const __grid_constant__ CUtensorMap map_x_lo,
                   const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                   TcConvParams p) {
  extern __shared__ unsigned char cv_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(cv_smem_raw) + 1023) &
                                                         ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kCvStages * kCvStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kCvStages;
  uint64_t* tmem_full_bar = bars + 2 * kCvStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kCvStages + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.y * p.BN;
  int tile = blockIdx.x;
  const int tx = tile % p.tiles_x;
  tile /= p.tiles_x;
  const int ty = tile % p.tiles_y;
  const int b = tile / p.tiles_y;
  const int x0 = tx * p.BW, y0 = ty * p.BH;
  using E = TcElem<F16>;
  constexpr int CK = E::kBK;  // input channels per k-block (128 bytes)
  const int chunks = p.Cin / CK;
  const int num_kb = 9 * chunks;
  const uint32_t stage_tx = 2 * kCvTile + 2 * (uint32_t)p.BN * 128u;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kCvStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kCvTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kCvStages;
        const uint32_t ph = (uint32_t)(kb / kCvStages) & 1u;
        mbar_wait_or_trap(&empty_bar[s], ph ^ 1u);
        unsigned char* st = smem + (size_t)s * kCvStageBytes;
        mbar_arrive_expect_tx(&full_bar[s], p.single ? stage_tx / 2 : stage_tx);
        const int tap = kb / chunks, cc = kb - tap * chunks;
        const int ky = tap / 3, kx = tap - ky * 3;
        tma_load_4d(st, &map_x_hi, &full_bar[s], cc * CK, x0 + kx - 1, y0 + ky - 1, b);
        if (!p.single) tma_load_4d(st + kCvTile, &map_x_lo, &full_bar[s], cc * CK, x0 + kx - 1, y0 + ky - 1, b);
        tma_load_2d(st + 2 * kCvTile, &map_w_hi, &full_bar[s], kb * CK, n0);
        if (!p.single) tma_load_2d(st + 3 * kCvTile, &map_w_lo, &full_bar[s], kb * CK, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc =
          (1u << 4) | (E::kFmt << 7) | (E::kFmt << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t tmem_lo = tmem_base + kCvHiAcc * 128;
      uint32_t accum_lo = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kCvStages;
        const uint32_t ph = (uint32_t)(kb / kCvStages) & 1u;
        mbar_wait_or_trap(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t st = smem_u32(smem + (size_t)s * kCvStageBytes);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t a_hi = make_desc(st + ks * 32, 16u, 1024u, 2u);
          const uint64_t a_lo = make_desc(st + kCvTile + ks * 32, 16u, 1024u, 2u);
          const uint64_t b_hi = make_desc(st + 2 * kCvTile + ks * 32, 16u, 1024u, 2u);
          const uint64_t b_lo = make_desc(st + 3 * kCvTile + ks * 32, 16u, 1024u, 2u);
          if (!p.single) {
            E::mma(tmem_lo, a_lo, b_hi, idesc, accum_lo);
            accum_lo = 1;
            E::mma(tmem_lo, a_hi, b_lo, idesc, 1);
          }
          E::mma(tmem_base + (uint32_t)(kb % kCvHiAcc) * 128, a_hi, b_hi, idesc, (kb >= kCvHiAcc || ks > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    mbar_wait_or_trap(tmem_full_bar, 0);
    tc_fence_after();
    const int lane_grp = warp & 3;
    const int r = lane_grp * 32 + lane;            // row of the tile = pixel iy*BW + ix
    const int iy = r / p.BW, ix = r - iy * p.BW;
    const int y = y0 + iy, x = x0 + ix;
    const bool valid = (y < p.H) && (x < p.W);
    float* zrow = p.z ? p.z + (size_t)b * p.sB + (size_t)y * p.sH + (size_t)x * p.sW : nullptr;
    const size_t prow = (((size_t)b * p.H + y) * p.W + x) * p.Cout;
    float psc = 1.f;
    if (p.a_hi) {
      const int e = pair_exponent(__ldg(p.a_bound));
      psc = exp2i(e);
      if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 64) *p.a_exp = e;
    }
    const int n_hi = min(kCvHiAcc, num_kb);
    const int out_shift = F16 ? -(__ldg(p.exp_x) + __ldg(p.exp_w)) : 0;
    float vmax = 0.f;
    for (int cb = 0; cb < p.BN; cb += 32) {
      float acc[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] = 0.f;
      for (int a = 0; a < n_hi + (p.single ? 0 : 1); ++a) {
        const int which = (a == n_hi) ? kCvHiAcc : a;
        uint32_t t[32];
        tmem_ld32(tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(which * 128 + cb), t);
        const float w = (F16 && which == kCvHiAcc) ? 1.f / kPairLoScale : 1.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = fmaf(__uint_as_float(t[j]), w, acc[j]);
      }
      if (F16) {
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = scale_pow2(acc[j], out_shift);
      }
      if (valid) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const int n = n0 + cb + j;
          if (n < p.Cout) {  // Cout % 4 == 0
            float v[4] = {acc[j], acc[j + 1], acc[j + 2], acc[j + 3]};
            cv_store4(p, v, n, zrow, prow, psc, vmax);
          }
        }
      }
    }
    if (p.zmax) {
      vmax = warp_max(vmax);
      if (lane == 0) atomicMax(p.zmax, __float_as_uint(vmax));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kCvTmemCols);
  }
}

// PERSISTENT variant for short reductions (9*Cin/CK <= 24 k-blocks: Cin = 64 / 128 with FP16 pairs): one CTA per SM
// walks over (pixel tile, channel tile) pairs with the accumulators double-buffered in TMEM (set = {hi*hi, lo products}),
// so the epilogue of tile j overlaps the MMAs of tile j+1 and the TMA ring never drains between tiles.  With K this
// short the one-tile-per-CTA kernel above spent about as long in prologue + epilogue as in the k loop.
constexpr int kCvPersistMaxKb = 24;

template <bool F16>
__global__ void __launch_bounds__(kCvThreads, 1)
tc_conv_fwd_persist_kernel(const __grid_constant__ CUtensorMap map_x_hi, const __grid_constant__ CUtensorMap map_x_lo,
                           const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                           TcConvParams p, int n_tiles, int num_tiles) {
  extern __shared__ unsigned char cv_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(cv_smem_raw) + 1023) &
                                                         ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kCvStages * kCvStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kCvStages;
  uint64_t* acc_full = bars + 2 * kCvStages;   // [2]
  uint64_t* acc_empty = acc_full + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  using E = TcElem<F16>;
  constexpr int CK = E::kBK;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunks = p.Cin / CK;
  const int num_kb = 9 * chunks;
  const uint32_t stage_tx = 2 * kCvTile + 2 * (uint32_t)p.BN * 128u;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kCvStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], 128);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kCvTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int pt = tile / n_tiles;
        const int n0 = (tile - pt * n_tiles) * p.BN;
        const int tx = pt % p.tiles_x;
        pt /= p.tiles_x;
        const int ty = pt % p.tiles_y;
        const int b = pt / p.tiles_y;
        const int x0 = tx * p.BW, y0 = ty * p.BH;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % kCvStages;
          const uint32_t ph = (uint32_t)(it / kCvStages) & 1u;
          mbar_wait_or_trap(&empty_bar[s], ph ^ 1u);
          unsigned char* st = smem + (size_t)s * kCvStageBytes;
          mbar_arrive_expect_tx(&full_bar[s], p.single ? stage_tx / 2 : stage_tx);
          const int tap = kb / chunks, cc = kb - tap * chunks;
          const int ky = tap / 3, kx = tap - ky * 3;
          tma_load_4d(st, &map_x_hi, &full_bar[s], cc * CK, x0 + kx - 1, y0 + ky - 1, b);
          if (!p.single) tma_load_4d(st + kCvTile, &map_x_lo, &full_bar[s], cc * CK, x0 + kx - 1, y0 + ky - 1, b);
          tma_load_2d(st + 2 * kCvTile, &map_w_hi, &full_bar[s], kb * CK, n0);
          if (!p.single) tma_load_2d(st + 3 * kCvTile, &map_w_lo, &full_bar[s], kb * CK, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc =
          (1u << 4) | (E::kFmt << 7) | (E::kFmt << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((128u >> 4) << 24);
      int it = 0, j = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++j) {
        const int set = j & 1;
        mbar_wait_or_trap(&acc_empty[set], ((uint32_t)(j >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t tmem_hi = tmem_base + (uint32_t)set * 256, tmem_lo = tmem_hi + 128;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % kCvStages;
          const uint32_t ph = (uint32_t)(it / kCvStages) & 1u;
          mbar_wait_or_trap(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t st = smem_u32(smem + (size_t)s * kCvStageBytes);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t a_hi = make_desc(st + ks * 32, 16u, 1024u, 2u);
            const uint64_t a_lo = make_desc(st + kCvTile + ks * 32, 16u, 1024u, 2u);
            const uint64_t b_hi = make_desc(st + 2 * kCvTile + ks * 32, 16u, 1024u, 2u);
            const uint64_t b_lo = make_desc(st + 3 * kCvTile + ks * 32, 16u, 1024u, 2u);
            const uint32_t acc = (kb > 0 || ks > 0) ? 1u : 0u;
            if (!p.single) {
              E::mma(tmem_lo, a_lo, b_hi, idesc, acc);
              E::mma(tmem_lo, a_hi, b_lo, idesc, 1);
            }
            E::mma(tmem_hi, a_hi, b_hi, idesc, acc);
          }
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&acc_full[set]);
      }
    }
  } else {
    const int lane_grp = warp & 3;
    const int r = lane_grp * 32 + lane;            // row of the tile = pixel iy*BW + ix
    const int iy = r / p.BW, ix = r - iy * p.BW;
    const int out_shift = F16 ? -(__ldg(p.exp_x) + __ldg(p.exp_w)) : 0;
    float vmax = 0.f;
    float psc = 1.f;
    if (p.a_hi) {
      const int e = pair_exponent(__ldg(p.a_bound));
      psc = exp2i(e);
      if (blockIdx.x == 0 && threadIdx.x == 64) *p.a_exp = e;
    }
    int j = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++j) {
      const int set = j & 1;
      int pt = tile / n_tiles;
      const int n0 = (tile - pt * n_tiles) * p.BN;
      const int tx = pt % p.tiles_x;
      pt /= p.tiles_x;
      const int ty = pt % p.tiles_y;
      const int b = pt / p.tiles_y;
      const int y = ty * p.BH + iy, x = tx * p.BW + ix;
      const bool valid = (y < p.H) && (x < p.W);
      float* zrow = p.z ? p.z + (size_t)b * p.sB + (size_t)y * p.sH + (size_t)x * p.sW : nullptr;
      const size_t prow = (((size_t)b * p.H + y) * p.W + x) * p.Cout;
      mbar_wait_or_trap(&acc_full[set], (uint32_t)(j >> 1) & 1u);
      tc_fence_after();
      const uint32_t lane_addr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)set * 256;
#pragma unroll 1
      for (int cb = 0; cb < p.BN; cb += 32) {
        uint32_t th[32], tl[32];
        tmem_ld32(lane_addr + (uint32_t)cb, th);
        tmem_ld32(lane_addr + (uint32_t)(128 + cb), tl);
        if (valid) {
#pragma unroll
          for (int q = 0; q < 32; q += 4) {
            const int n = n0 + cb + q;
            if (n < p.Cout) {  // Cout % 4 == 0
              float v[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float a = p.single ? __uint_as_float(th[q + u])
                                         : fmaf(__uint_as_float(tl[q + u]), F16 ? 1.f / kPairLoScale : 1.f,
                                                __uint_as_float(th[q + u]));
                v[u] = F16 ? scale_pow2(a, out_shift) : a;
              }
              cv_store4(p, v, n, zrow, prow, psc, vmax);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[set]);
    }
    if (p.zmax) {
      vmax = warp_max(vmax);
      if (lane == 0) atomicMax(p.zmax, __float_as_uint(vmax));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kCvTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
struct TcWgradParams {
  float* partial;  // [splits][M][N]
  int B, H, W, Cin, Cout;
  int M, N, BN;
  int rows_per_split;  // image rows (b, y) per grid.z slice
  int xblocks;         // ceil(W / pixels per k-block)
  const int* exp_x;    // FP16 pair operands: device exponents of the x / dz planes
  const int* exp_dz;
  int single;          // 1 = hi planes only, one product per k-step
};
constexpr int kWgChunk = 8;  // k-blocks per TMEM accumulation chunk


template <bool F16>
__global__ void __launch_bounds__(kCvThreads, 1)
tc_conv_wgrad_kernel(const __grid_constant__ CUtensorMap map_x_hi, const __grid_constant__ CUtensorMap map_x_lo,
                     const __grid_constant__ CUtensorMap map_dz_hi, const __grid_constant__ CUtensorMap map_dz_lo,
                     TcWgradParams p) {
  extern __shared__ unsigned char cv_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(cv_smem_raw) + 1023) &
                                                         ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kCvStages * kCvStageBytes);
  uint64_t* full_bar = bars;                          // [stages]
  uint64_t* empty_bar = bars + kCvStages;             // [stages]
  uint64_t* acc_full = bars + 2 * kCvStages;          // [3]
  uint64_t* acc_empty = acc_full + kCvHiAcc;          // [3]
  uint64_t* lo_full = acc_empty + kCvHiAcc;           // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(lo_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * p.BN, m0 = blockIdx.y * 128;
  const int total_rows = p.B * p.H;
  const int r_begin = blockIdx.z * p.rows_per_split;
  const int r_end = min(total_rows, r_begin + p.rows_per_split);
  const int num_kb = max(0, r_end - r_begin) * p.xblocks;
  const int num_chunks = (num_kb + kWgChunk - 1) / kWgChunk;
  using E = TcElem<F16>;
  constexpr int MB = E::kMnBox;   // channels per box (128 bytes) = rows of one (tap, channel group)
  constexpr int PK = E::kBK;      // pixels (k rows) per k-block
  constexpr int NGRP = 128 / MB;  // row groups of the M tile
  const int nb_boxes = p.BN / MB;
  const uint32_t stage_tx = 2 * kCvTile + 2 * (uint32_t)p.BN * 128u;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kCvStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < kCvHiAcc; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], 128);
    }
    mbar_init(lo_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kCvTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // the four 32-row groups of this M tile: (tap, first channel) each
      int tapj[NGRP], cij[NGRP];
      for (int j = 0; j < NGRP; ++j) {
        const int mrow = m0 + MB * j;
        tapj[j] = (mrow < p.M) ? mrow / p.Cin : -1;
        cij[j] = (mrow < p.M) ? mrow % p.Cin : 0;
      }
      int kb = 0;
      for (int r = r_begin; r < r_end; ++r) {
        const int b = r / p.H, y = r - b * p.H;
        for (int xb = 0; xb < p.xblocks; ++xb, ++kb) {
          const int s = kb % kCvStages;
          const uint32_t ph = (uint32_t)(kb / kCvStages) & 1u;
          mbar_wait_or_trap(&empty_bar[s], ph ^ 1u);
          unsigned char* st = smem + (size_t)s * kCvStageBytes;
          mbar_arrive_expect_tx(&full_bar[s], p.single ? stage_tx / 2 : stage_tx);
          for (int j = 0; j < NGRP; ++j) {
            int cx, cy;
            if (tapj[j] >= 0) {
              const int ky = tapj[j] / 3, kx = tapj[j] - ky * 3;
              cx = xb * PK + kx - 1;
              cy = y + ky - 1;
            } else {  // rows beyond 9*Cin: a fully out-of-bounds box delivers zeros (and the expected bytes)
              cx = 0;
              cy = p.H + 4;
            }
            tma_load_4d(st + j * E::kMnBoxBytes, &map_x_hi, &full_bar[s], cij[j], cx, cy, b);
            if (!p.single) tma_load_4d(st + kCvTile + j * E::kMnBoxBytes, &map_x_lo, &full_bar[s], cij[j], cx, cy, b);
          }
          for (int j = 0; j < nb_boxes; ++j) {
            tma_load_4d(st + 2 * kCvTile + j * E::kMnBoxBytes, &map_dz_hi, &full_bar[s], n0 + MB * j, xb * PK, y, b);
            if (!p.single)
              tma_load_4d(st + 3 * kCvTile + j * E::kMnBoxBytes, &map_dz_lo, &full_bar[s], n0 + MB * j, xb * PK, y, b);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (E::kFmt << 7) | (E::kFmt << 10) | (1u << 15) | (1u << 16) |
                             ((uint32_t)(p.BN >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t tmem_lo = tmem_base + kCvHiAcc * 128;
      uint32_t accum_lo = 0;
      for (int c = 0; c < num_chunks; ++c) {
        const int a = c % kCvHiAcc;
        mbar_wait_or_trap(&acc_empty[a], ((uint32_t)(c / kCvHiAcc) & 1u) ^ 1u);
        tc_fence_after();
        const int kb1 = min(num_kb, (c + 1) * kWgChunk);
        for (int kb = c * kWgChunk; kb < kb1; ++kb) {
          const int s = kb % kCvStages;
          const uint32_t ph = (uint32_t)(kb / kCvStages) & 1u;
          mbar_wait_or_trap(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t st = smem_u32(smem + (size_t)s * kCvStageBytes);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint32_t off = ks * E::kMnStep;
            const uint64_t a_hi = make_desc(st + off, E::kMnBoxBytes, E::kMnSbo, E::kMnLayout);
            const uint64_t a_lo = make_desc(st + kCvTile + off, E::kMnBoxBytes, E::kMnSbo, E::kMnLayout);
            const uint64_t b_hi = make_desc(st + 2 * kCvTile + off, E::kMnBoxBytes, E::kMnSbo, E::kMnLayout);
            const uint64_t b_lo = make_desc(st + 3 * kCvTile + off, E::kMnBoxBytes, E::kMnSbo, E::kMnLayout);
            if (!p.single) {
              E::mma(tmem_lo, a_lo, b_hi, idesc, accum_lo);
              accum_lo = 1;
              E::mma(tmem_lo, a_hi, b_lo, idesc, 1);
            }
            E::mma(tmem_base + (uint32_t)a * 128, a_hi, b_hi, idesc, (kb > c * kWgChunk || ks > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&acc_full[a]);
      }
      umma_commit(lo_full);
    }
  } else {
    const int lane_grp = warp & 3;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(lane_grp * 32) << 16);
    float sum[128];
#pragma unroll
    for (int j = 0; j < 128; ++j) sum[j] = 0.f;
    for (int c = 0; c < num_chunks; ++c) {
      const int a = c % kCvHiAcc;
      mbar_wait_or_trap(&acc_full[a], (uint32_t)(c / kCvHiAcc) & 1u);
      tc_fence_after();
#pragma unroll
      for (int cb = 0; cb < 128; cb += 32) {
        if (cb < p.BN) {
          uint32_t t[32];
          tmem_ld32(lane_addr + (uint32_t)(a * 128 + cb), t);
#pragma unroll
          for (int j = 0; j < 32; ++j) sum[cb + j] += __uint_as_float(t[j]);
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[a]);
    }
    if (num_kb > 0 && !p.single) {
      mbar_wait_or_trap(lo_full, 0);
      tc_fence_after();
#pragma unroll
      for (int cb = 0; cb < 128; cb += 32) {
        if (cb < p.BN) {
          uint32_t t[32];
          tmem_ld32(lane_addr + (uint32_t)(kCvHiAcc * 128 + cb), t);
#pragma unroll
          for (int j = 0; j < 32; ++j)
            sum[cb + j] = fmaf(__uint_as_float(t[j]), F16 ? 1.f / kPairLoScale : 1.f, sum[cb + j]);
        }
      }
    }
    if (F16) {
      const int out_shift = -(__ldg(p.exp_x) + __ldg(p.exp_dz));
#pragma unroll
      for (int j = 0; j < 128; ++j) sum[j] = scale_pow2(sum[j], out_shift);
    }
    const int m = m0 + lane_grp * 32 + lane;
    if (m < p.M) {
      float* prow = p.partial + ((size_t)blockIdx.z * p.M + m) * p.N;
#pragma unroll
      for (int cb = 0; cb < 128; cb += 32) {
        if (cb < p.BN) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const int n = n0 + cb + j;
            if (n < p.N)
              *reinterpret_cast<float4*>(prow + n) = make_float4(sum[cb + j], sum[cb + j + 1], sum[cb + j + 2], sum[cb + j + 3]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kCvTmemCols);
  }
}

// sum split-K partials [S][9*Cin][Cout] and write the PyTorch layout dW[Cout][Cin][3][3]
__global__ void __launch_bounds__(256)
tc_wgrad_reduce_kernel(const float* __restrict__ partial, int splits, int Cin, int Cout, float* __restrict__ dw) {
  const int total = 9 * Cin * Cout;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int tap = idx % 9;
    const int ci = (idx / 9) % Cin;
    const int co = idx / (9 * Cin);
    const size_t src = (size_t)(tap * Cin + ci) * Cout + co;
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += partial[(size_t)z * total + src];
    dw[idx] = s;
  }
}

// per-channel sum and sum of squares of z [P][C] accumulated into float64 stats[2C] (BatchNorm statistics)
__global__ void __launch_bounds__(256)
colstats_kernel(const float* __restrict__ z, long long P, int C4, long long rows_per_cta, double* __restrict__ stats) {
  extern __shared__ float s_red[];  // [256][8]
  const int rows = 256 / C4;
  const int c4 = threadIdx.x % C4, r = threadIdx.x / C4;
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  const long long p0 = (long long)blockIdx.x * rows_per_cta, p1 = min(P, p0 + rows_per_cta);
  if (r < rows) {
    for (long long q = p0 + r; q < p1; q += rows) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(z) + q * C4 + c4);
      s1[0] += v.x; s1[1] += v.y; s1[2] += v.z; s1[3] += v.w;
      s2[0] = fmaf(v.x, v.x, s2[0]); s2[1] = fmaf(v.y, v.y, s2[1]);
      s2[2] = fmaf(v.z, v.z, s2[2]); s2[3] = fmaf(v.w, v.w, s2[3]);
    }
  }
  float* mine = s_red + (size_t)threadIdx.x * 8;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    mine[i] = s1[i];
    mine[4 + i] = s2[i];
  }
  __syncthreads();
  const int C = C4 * 4;
  for (int o = threadIdx.x; o < 2 * C; o += 256) {
    const int which = o / C, c = o % C;
    double t = 0.0;
    for (int rr = 0; rr < rows; ++rr) t += (double)s_red[((size_t)rr * C4 + (c >> 2)) * 8 + which * 4 + (c & 3)];
    atomicAdd(&stats[which * C + c], t);
  }
}

}  // namespace vocr

using namespace vocr;

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static void pick_tile(int H, int W, int* BW, int* BH) {
  // 128 output pixels per tile as BH x BW; pick the shape that wastes the fewest out-of-image pixels
  const int cand[3][2] = {{128, 1}, {64, 2}, {32, 4}};
  double best = -1;
  for (int i = 0; i < 3; ++i) {
    const int bw = cand[i][0], bh = cand[i][1];
    const double eff = ((double)W / (ceil_div(W, bw) * bw)) * ((double)H / (ceil_div(H, bh) * bh));
    if (eff > best + 1e-9) {
      best = eff;
      *BW = bw;
      *BH = bh;
    }
  }
}

// z[B,H,W,Cout] = conv3x3_pad1(x) + bias on tensor cores.  x planes (hi, lo) NHWC [B,H,W,Cin]; weight planes
// K-major [Cout][9*Cin] with k = (ky*3+kx)*Cin + ci; Cout % 4 == 0.  TF32 planes: Cin % 32 == 0; FP16 pair planes
// (vocr_split_f16_f32, with their device exponents): Cin % 64 == 0.
template <bool F16>
static int tc_conv_fwd_launch(const void* x_hi, const void* x_lo, const int* exp_x, const void* w_hi, const void* w_lo,
                              const int* exp_w, const float* bias, float* z, int B, int H, int W, int Cin, int Cout,
                              int products, float* zmax, cudaStream_t stream, const TcConvParams* fused = nullptr) {
  constexpr int CK = TcElem<F16>::kBK;
  VOCR_REQUIRE(B >= 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && Cin % CK == 0 && Cout % 4 == 0);
  VOCR_REQUIRE(products == 0 || products == 1 || products == 3);
  if (B == 0) return VOCR_OK;
  VOCR_REQUIRE(x_hi && x_lo && w_hi && w_lo && (z || fused) && (!F16 || (exp_x && exp_w)));
  VOCR_REQUIRE(aligned16(x_hi) && aligned16(x_lo) && aligned16(w_hi) && aligned16(w_lo) && aligned16(z) &&
               (!bias || aligned16(bias)));
  TcConvParams p;
  p.bn_scale = p.bn_shift = nullptr;
  p.a_hi = p.a_lo = nullptr;
  p.a_bound = nullptr;
  p.a_exp = nullptr;
  p.sB = (long long)H * W * Cout; p.sH = (long long)W * Cout; p.sW = Cout;
  if (fused) p = *fused;  // the fused-epilogue fields; everything else is set below
  p.z = z; p.bias = bias; p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.exp_x = exp_x; p.exp_w = exp_w;
  p.zmax = reinterpret_cast<unsigned*>(zmax);
  p.single = (F16 && resolve_tc_products(products) == 1) ? 1 : 0;
  pick_tile(H, W, &p.BW, &p.BH);
  p.tiles_x = ceil_div(W, p.BW);
  p.tiles_y = ceil_div(H, p.BH);
  p.BN = (Cout <= 64) ? 64 : 128;
  CUtensorMap mx_hi, mx_lo, mw_hi, mw_lo;
  bool ok;
  if (F16)
    ok = make_map_nhwc_f16(&mx_hi, x_hi, B, H, W, Cin, CK, p.BW, p.BH) &&
         make_map_nhwc_f16(&mx_lo, x_lo, B, H, W, Cin, CK, p.BW, p.BH) &&
         make_map_2d_f16(&mw_hi, w_hi, Cout, 9LL * Cin, 9LL * Cin, CK, p.BN) &&
         make_map_2d_f16(&mw_lo, w_lo, Cout, 9LL * Cin, 9LL * Cin, CK, p.BN);
  else
    ok = make_map_nhwc(&mx_hi, static_cast<const float*>(x_hi), B, H, W, Cin, CK, p.BW, p.BH, false) &&
         make_map_nhwc(&mx_lo, static_cast<const float*>(x_lo), B, H, W, Cin, CK, p.BW, p.BH, false) &&
         make_map_2d(&mw_hi, static_cast<const float*>(w_hi), Cout, 9LL * Cin, 9LL * Cin, CK, p.BN) &&
         make_map_2d(&mw_lo, static_cast<const float*>(w_lo), Cout, 9LL * Cin, 9LL * Cin, CK, p.BN);
  if (!ok) return VOCR_EXECUTION_FAILED;
  static DeviceLatch attr_latch;
  if (attr_latch.need()) {
    if (cudaFuncSetAttribute(tc_conv_fwd_kernel<F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCvSmemBytes) !=
            cudaSuccess ||
        cudaFuncSetAttribute(tc_conv_fwd_persist_kernel<F16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             kCvSmemBytes) != cudaSuccess)
      return VOCR_EXECUTION_FAILED;
    attr_latch.set();
  }
  const long long tiles = (long long)B * p.tiles_x * p.tiles_y;
  const int n_tiles = ceil_div(Cout, p.BN);
  VOCR_REQUIRE(tiles * n_tiles <= 2147483647LL);
  // (the K limit of the persistent kernel protects the fp32-level accuracy; the single-product mode has no such claim)
  if (9 * (Cin / CK) <= kCvPersistMaxKb || p.single) {
    const int total = (int)(tiles * n_tiles);
    tc_conv_fwd_persist_kernel<F16><<<min(total, kNumSMs), kCvThreads, kCvSmemBytes, stream>>>(mx_hi, mx_lo, mw_hi,
                                                                                                mw_lo, p, n_tiles, total);
    VOCR_CHECK_LAUNCH();
    return VOCR_OK;
  }
  dim3 grid((unsigned)tiles, n_tiles);
  tc_conv_fwd_kernel<F16><<<grid, kCvThreads, kCvSmemBytes, stream>>>(mx_hi, mx_lo, mw_hi, mw_lo, p);
  VOCR_CHECK_LAUNCH();
  return VOCR_OK;
}

extern "C" int vocr_tc_conv3x3_fwd(const float* x_hi, const float* x_lo, const float* w_hi, const float* w_lo,
                                   const float* bias, float* z, int B, int H, int W, int Cin, int Cout,
                                   vocr_stream_t stream_) {
  return tc_conv_fwd_launch<false>(x_hi, x_lo, nullptr, w_hi, w_lo, nullptr, bias, z, B, H, W, Cin, Cout, 3, nullptr,
                                   static_cast<cudaStream_t>(stream_));
}
extern "C" int vocr_tc_conv3x3_fwd_f16(const uint16_t* x_hi, const uint16_t* x_lo, const int32_t* exp_x,
                                       const uint16_t* w_hi, const uint16_t* w_lo, const int32_t* exp_w,
                                       const float* bias, float* z, int B, int H, int W, int Cin, int Cout,
                                       int products, float* zmax, vocr_stream_t stream_) {
  return tc_conv_fwd_launch<true>(x_hi, x_lo, exp_x, w_hi, w_lo, exp_w, bias, z, B, H, W, Cin, Cout, products, zmax,
                                  static_cast<cudaStream_t>(stream_));
}

// Inference form of one Conv + BatchNorm (running statistics) + ReLU unit (reference cnnlstm.py:263-266) in ONE kernel:
// a = relu((conv3x3(x) + bias) * scale[c] + shift[c]), scale / shift from vocr_bn_finalize_f32(training = 0).
// a (fp32, may be NULL) is written with strides (sB, sH, sW) (NHWC, or the time-major sequence layout of the last
// block); a_hi / a_lo (FP16 pair planes of a, dense NHWC, may be NULL) are scaled by 2^e with e from the device scalar
// `bound` >= max |a| (vocr_bn_eval_bound_f32), e is stored to pair_exp[0].  amax (optional device float the caller zeroes)
// receives the measured max a.  Cin % 64 == 0, Cout % 4 == 0.
extern "C" int vocr_tc_conv3x3_bnrelu_f16(const uint16_t* x_hi, const uint16_t* x_lo, const int32_t* exp_x,
                                          const uint16_t* w_hi, const uint16_t* w_lo, const int32_t* exp_w,
                                          const float* bias, const float* scale, const float* shift, float* a,
                                          long long sB, long long sH, long long sW, uint16_t* a_hi, uint16_t* a_lo,
                                          const float* bound, int32_t* pair_exp, float* amax, int B, int H, int W,
                                          int Cin, int Cout, int products, vocr_stream_t stream_) {
  VOCR_REQUIRE(scale && shift && aligned16(scale) && aligned16(shift) && (a || a_hi));
  VOCR_REQUIRE((a_hi == nullptr) == (a_lo == nullptr) && (!a_hi || (bound && pair_exp)));
  VOCR_REQUIRE(sB % 4 == 0 && sH % 4 == 0 && sW % 4 == 0 && (!a_hi || (aligned16(a_hi) && aligned16(a_lo))));
  TcConvParams f;
  f.bn_scale = scale; f.bn_shift = shift;
  f.sB = sB; f.sH = sH; f.sW = sW;
  f.a_hi = reinterpret_cast<__half*>(a_hi); f.a_lo = reinterpret_cast<__half*>(a_lo);
  f.a_bound = reinterpret_cast<const unsigned*>(bound);
  f.a_exp = pair_exp;
  return tc_conv_fwd_launch<true>(x_hi, x_lo, exp_x, w_hi, w_lo, exp_w, bias, a, B, H, W, Cin, Cout, products, amax,
                                  static_cast<cudaStream_t>(stream_), &f);
}

extern "C" size_t vocr_tc_conv3x3_wgrad_workspace_size(int B, int H, int W, int Cin, int Cout) {
  (void)B; (void)H; (void)W;
  return sizeof(float) * (size_t)9 * Cin * Cout * 64 + 256;
}

// dw[Cout,Cin,3,3] = sum over pixels of x(p+tap, ci) * dz(p, co); x and dz given as (hi, lo) NHWC planes.
// TF32 planes: Cin % 32 == 0, Cout % 32 == 0; FP16 pair planes: Cin % 64 == 0, Cout % 64 == 0.
template <bool F16>
static int tc_conv_wgrad_launch(const void* x_hi, const void* x_lo, const int* exp_x, const void* dz_hi,
                                const void* dz_lo, const int* exp_dz, float* dw, int B, int H, int W, int Cin,
                                int Cout, void* workspace, size_t workspace_bytes, int products, cudaStream_t stream) {
  using E = TcElem<F16>;
  constexpr int MB = E::kMnBox, PK = E::kBK;
  VOCR_REQUIRE(B > 0 && H > 0 && W > 0 && Cin % MB == 0 && Cout % MB == 0 && Cin > 0 && Cout > 0);
  VOCR_REQUIRE(x_hi && x_lo && dz_hi && dz_lo && dw && workspace && (!F16 || (exp_x && exp_dz)));
  VOCR_REQUIRE(aligned16(x_hi) && aligned16(x_lo) && aligned16(dz_hi) && aligned16(dz_lo) && aligned16(workspace));
  TcWgradParams p;
  p.partial = static_cast<float*>(workspace);
  p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.M = 9 * Cin; p.N = Cout;
  p.exp_x = exp_x; p.exp_dz = exp_dz;
  VOCR_REQUIRE(products == 0 || products == 1 || products == 3);
  p.single = (F16 && resolve_tc_products(products) == 1) ? 1 : 0;
  p.BN = (Cout <= 64) ? 64 : 128;
  p.xblocks = ceil_div(W, PK);
  const int tiles = ceil_div(p.M, 128) * ceil_div(p.N, p.BN);
  const int total_rows = B * H;
  int splits = max(1, min(64, min(total_rows, (2 * kNumSMs) / tiles)));
  p.rows_per_split = ceil_div(total_rows, splits);
  splits = ceil_div(total_rows, p.rows_per_split);
  VOCR_REQUIRE(sizeof(float) * (size_t)p.M * p.N * splits <= workspace_bytes);
  CUtensorMap mx_hi, mx_lo, md_hi, md_lo;
  bool ok;
  if (F16)
    ok = make_map_nhwc_f16(&mx_hi, x_hi, B, H, W, Cin, MB, PK, 1) && make_map_nhwc_f16(&mx_lo, x_lo, B, H, W, Cin, MB, PK, 1) &&
         make_map_nhwc_f16(&md_hi, dz_hi, B, H, W, Cout, MB, PK, 1) && make_map_nhwc_f16(&md_lo, dz_lo, B, H, W, Cout, MB, PK, 1);
  else
    ok = make_map_nhwc(&mx_hi, static_cast<const float*>(x_hi), B, H, W, Cin, MB, PK, 1, true) &&
         make_map_nhwc(&mx_lo, static_cast<const float*>(x_lo), B, H, W, Cin, MB, PK, 1, true) &&
         make_map_nhwc(&md_hi, static_cast<const float*>(dz_hi), B, H, W, Cout, MB, PK, 1, true) &&
         make_map_nhwc(&md_lo, static_cast<const float*>(dz_lo), B, H, W, Cout, MB, PK, 1, true);
  if (!ok) return VOCR_EXECUTION_FAILED;
  static DeviceLatch attr_latch;
  if (attr_latch.need()) {
    if (cudaFuncSetAttribute(tc_conv_wgrad_kernel<F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCvSmemBytes) !=
        cudaSuccess)
      return VOCR_EXECUTION_FAILED;
    attr_latch.set();
  }
  dim3 grid(ceil_div(p.N, p.BN), ceil_div(p.M, 128), splits);
  tc_conv_wgrad_kernel<F16><<<grid, kCvThreads, kCvSmemBytes, stream>>>(mx_hi, mx_lo, md_hi, md_lo, p);
  VOCR_CHECK_LAUNCH();
  tc_wgrad_reduce_kernel<<<min(ceil_div(p.M * p.N, 256), 4 * kNumSMs), 256, 0, stream>>>(p.partial, splits, Cin, Cout, dw);
  VOCR_CHECK_LAUNCH();
  return VOCR_OK;
}

extern "C" int vocr_tc_conv3x3_wgrad(const float* x_hi, const float* x_lo, const float* dz_hi, const float* dz_lo,
                                     float* dw, int B, int H, int W, int Cin, int Cout, void* workspace,
                                     size_t workspace_bytes, vocr_stream_t stream_) {
  return tc_conv_wgrad_launch<false>(x_hi, x_lo, nullptr, dz_hi, dz_lo, nullptr, dw, B, H, W, Cin, Cout, workspace,
                                     workspace_bytes, 3, static_cast<cudaStream_t>(stream_));
}
extern "C" int vocr_tc_conv3x3_wgrad_f16(const uint16_t* x_hi, const uint16_t* x_lo, const int32_t* exp_x,
                                         const uint16_t* dz_hi, const uint16_t* dz_lo, const int32_t* exp_dz, float* dw,
                                         int B, int H, int W, int Cin, int Cout, void* workspace,
                                         size_t workspace_bytes, int products, vocr_stream_t stream_) {
  return tc_conv_wgrad_launch<true>(x_hi, x_lo, exp_x, dz_hi, dz_lo, exp_dz, dw, B, H, W, Cin, Cout, workspace,
                                    workspace_bytes, products, static_cast<cudaStream_t>(stream_));
}

// stats[0:C] += sum_p z[p,c], stats[C:2C] += sum_p z[p,c]^2 (float64).  C % 4 == 0, C <= 1024.
extern "C" int vocr_colstats_f32(const float* z, long long P, int C, double* stats, vocr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VOCR_REQUIRE(P >= 0 && C > 0 && C % 4 == 0 && C <= 1024 && stats);
  if (P == 0) return VOCR_OK;
  VOCR_REQUIRE(z && aligned16(z));
  const int C4 = C / 4, rows = 256 / C4;
  long long rows_per_cta = ceil_div64(P, (long long)kNumSMs * 4);
  rows_per_cta = max((long long)rows * 8, ceil_div64(rows_per_cta, rows) * rows);
  const int grid = (int)ceil_div64(P, rows_per_cta);
  colstats_kernel<<<grid, 256, sizeof(float) * 256 * 8, stream>>>(z, P, C4, rows_per_cta, stats);
  VOCR_CHECK_LAUNCH();
  return VOCR_OK;
}
