// Shared device/host helpers for the vistaocr_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/vistaocr_b200.h"

#define VOCR_CHECK_LAUNCH()                                   \
  do {                                                        \
    cudaError_t e__ = cudaGetLastError();                     \
    if (e__ != cudaSuccess) return VOCR_EXECUTION_FAILED;     \
  } while (0)

#define VOCR_REQUIRE(cond)                   \
  do {                                       \
    if (!(cond)) return VOCR_INVALID_VALUE;  \
  } while (0)

#include <atomic>

namespace vocr {

// cudaFuncSetAttribute is PER DEVICE: a process that drives several GPUs through this ABI must set the attributes once
// on each of them.  One latch per call site; bit d = done on device d (setting twice from two threads is harmless).
struct DeviceLatch {
  std::atomic<unsigned long long> done[4];
  bool need() const {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 256) return true;
    return ((done[d >> 6].load(std::memory_order_acquire) >> (d & 63)) & 1ull) == 0;
  }
  void set() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 256) return;
    done[d >> 6].fetch_or(1ull << (d & 63), std::memory_order_release);
  }
};

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs
constexpr float kNegInf = -INFINITY;

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline long long ceil_div64(long long a, long long b) { return (a + b - 1) / b; }

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// arg-max ordering of numpy / the reference decoder: NaN is maximal, otherwise plain '>' (the first maximum wins)
__device__ __forceinline__ bool argmax_gt(float a, float b) { return (a > b) || (a != a && b == b); }

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- mbarrier + 1-D bulk async copy (TMA engine, no tensor map) -----------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a descriptor/size bug must surface as an error, never as a hung GPU box.
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity, uint32_t max_spins = (1u << 24)) {
#pragma unroll 1
  for (uint32_t i = 0; i < max_spins; ++i)
    if (mbar_try_wait(bar, parity)) return true;
  return false;
}
// ... and the loud form: a wait that never completes traps (sticky launch error on the host side).
__device__ __forceinline__ void mbar_wait_or_trap(uint64_t* bar, uint32_t parity) {
  if (!mbar_wait_bounded(bar, parity)) asm volatile("trap;");
}
// global -> shared bulk copy; dst/src 16-B aligned, bytes % 16 == 0.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---- programmatic dependent launch (PDL): the next kernel of a chain is launched while this one still runs -------------
// pdl_trigger(): lets the dependent grid be scheduled (its CTAs start as SM slots free up);  pdl_wait(): the dependent
// grid blocks here until the primary grid has COMPLETED and its memory is visible - call before the first access to
// anything the primary writes.  Both are no-ops when the launch carries no PDL attribute.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ float4 ldg_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream_f4(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

// host: launch `kernel` as the programmatic dependent of the previous launch on `stream`
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace vocr

namespace vocr {
// Stage `rows_here` consecutive rows of `A` floats (a flat, contiguous span of global memory) into shared memory.
// Fast path: one 1-D bulk async copy completed on an mbarrier; partial/misaligned tiles fall back to plain loads.
// `smem` layout: [0,16) mbarrier, [16, ...) tile.  All threads of the CTA must call; returns the tile pointer.
__device__ __forceinline__ float* stage_row_tile(unsigned char* smem, const float* __restrict__ gsrc, int n_floats,
                                                 bool base_aligned) {
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  float* tile = reinterpret_cast<float*>(smem + 16);
  const uint32_t bytes = (uint32_t)n_floats * 4u;
  if (base_aligned && (bytes % 16u) == 0u && bytes > 0u) {
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
    for (int i = threadIdx.x; i < n_floats; i += blockDim.x) tile[i] = __ldg(gsrc + i);
    __syncthreads();
  }
  return tile;
}
}  // namespace vocr
