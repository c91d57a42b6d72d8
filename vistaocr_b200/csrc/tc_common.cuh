// Shared device/host helpers of the tcgen05 / TMA kernels (tc_gemm.cu, tc_conv.cu).
#pragma once
#include <cuda.h>
#include <cstdlib>
#include "common.cuh"
#include "pair_f16.cuh"

namespace vocr {

// api.cu: per-call arithmetic mode, 3 = compensated products, 1 = hi planes only, 0 = the vocr_set_tc_products default
int resolve_tc_products(int products);

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// ---- thread-block clusters: CTAs that work on neighbouring output tiles share operand tiles ---------------------------
// A tile that several CTAs of a cluster need is fetched ONCE from L2: each of them loads 1/n-th of it and the TMA unit
// multicasts that part into the same shared-memory offset of every CTA in `mask` (and signals the mbarrier at the same
// offset there), so the L2 -> SM traffic of the tile is divided by n.
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster "
      "[%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ uint32_t tc_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void tc_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrives (once the MMAs issued so far have retired) on the mbarrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_c, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_c),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_c, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_c),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Operand element type of a kernel instantiation.  F16 = false: TF32 planes (hi, lo), K = 8 per instruction.
// F16 = true: FP16 pair planes (hi, lo * 2^11) of x * 2^e, K = 16 per instruction - the same 32 bytes of K per
// instruction and per swizzle row, so every tile keeps its byte geometry and the k loop covers twice the elements.
template <bool F16>
struct TcElem {
  static constexpr int kBK = F16 ? 64 : 32;                       // K elements per 128-byte swizzle row (one k-block)
  static constexpr uint32_t kFmt = F16 ? 0u : 2u;                 // instruction-descriptor operand format
  // MN-major operand tiles (the reduction index is the row): TMA box / UMMA descriptor constants
  static constexpr int kMnBox = F16 ? 64 : 32;                    // M/N elements per box (128 bytes)
  static constexpr uint32_t kMnBoxBytes = F16 ? 8192u : 4096u;    // box = kMnBox x kBK k rows
  static constexpr uint32_t kMnSbo = F16 ? 1024u : 512u;          // stride between k groups (8 / 4 rows of 128 B)
  static constexpr uint32_t kMnStep = F16 ? 2048u : 1024u;        // one instruction's k rows (16 / 8) x 128 B
  static constexpr uint32_t kMnLayout = F16 ? 2u : 1u;            // SWIZZLE_128B / SWIZZLE_128B_BASE32B
  static __device__ __forceinline__ void mma(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    if constexpr (F16) umma_f16(tmem_c, da, db, idesc, acc);
    else umma_tf32(tmem_c, da, db, idesc, acc);
  }
};
// UMMA shared-memory descriptor (sm_100 format, cute/arch/mma_sm100_desc.hpp): start address >> 4 in bits [0,14),
// leading byte offset >> 4 in [16,30), stride byte offset >> 4 in [32,46), version 1 in [46,48), layout type
// SWIZZLE_128B = 2 in [61,64).  K-major SW128 tile: rows of 128 B, 8-row atoms of 1024 B -> SBO = 1024, LBO unused (1).
// MN-major 32-bit operands have exactly one legal layout, SWIZZLE_128B_BASE32B = 1 ("128-B swizzle, 32-B atomicity",
// TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): atoms of (32 elements of M/N = 128 B) x (4 k rows) = 512 B; SBO = stride
// between 4-row k groups (512 B inside a TMA box of 32 rows), LBO = stride between M/N groups (one 4096-B box).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}


__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
}
// 32 TMEM lanes (this warp's quarter) x 32 consecutive columns -> 32 registers per thread
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&t)[32]) {
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
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// VOCR_TC_CLUSTER=0 launches every tensor-core kernel without clusters (A/B measurements; same results either way)
inline bool tc_clusters_enabled() {
  static const bool on = [] {
    const char* e = getenv("VOCR_TC_CLUSTER");
    return !(e && e[0] == '0');
  }();
  return on;
}

// same load without the wait: issue several, then one tcgen05.wait::ld
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&t)[32]) {
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

// ---- host: tensor maps through the driver entry point (no link-time dependency on libcuda) ---------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D fp32 tensor [rows][cols] with leading dimension ld (elements); box = {box_cols, box_rows}.
// mn_major selects the 128B_ATOM_32B swizzle (the only legal layout of MN-major 32-bit UMMA operands).
inline bool make_map_2d(CUtensorMap* map, const float* base, long long rows, long long cols, long long ld, int box_cols,
                        int box_rows, bool mn_major = false) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// FP16 planes: 2-D [rows][cols] with leading dimension ld (elements, multiple of 8); SWIZZLE_128B for both majors.
inline bool make_map_2d_f16(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld,
                            int box_cols, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
inline bool make_map_nhwc_f16(CUtensorMap* map, const void* base, int B, int H, int W, int C, int box_c, int box_w,
                              int box_h) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// 4-D NHWC fp32 activation tensor [B][H][W][C] (dense); box = {box_c, box_w, box_h, 1}.  Out-of-range coordinates
// (the conv halo) are zero-filled by the TMA unit.
inline bool make_map_nhwc(CUtensorMap* map, const float* base, int B, int H, int W, int C, int box_c, int box_w,
                          int box_h, bool mn_major) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace vocr
