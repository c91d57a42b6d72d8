// Probes (GPU box only) for the cluster-resident BiLSTM kernel:
//  K1  tcgen05.mma with the A operand in TMEM (M = 128, K = 16, kind::f16) and a K-major SWIZZLE_NONE B tile of N = 16
//      rows with LBO = 512 / SBO = 128: is A[m][k] at lane m, column k/2, half k%2, and B where the canonical layout says?
//  K2  a 16-CTA cluster in which every CTA publishes 2 KB to global memory and multicasts it into all 16 CTAs' shared
//      memory with ONE cp.async.bulk ... .multicast::cluster (mbarrier complete_tx on every destination).
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include "../vistaocr_b200/csrc/tc_common.cuh"
using namespace vocr;
namespace cg = cooperative_groups;

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_c, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_c), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc)
      : "memory");
}

__global__ void __launch_bounds__(128, 1) k1(float* out, int K0) {
  extern __shared__ unsigned char raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  unsigned char* B = smem;  // no-swizzle K-major: [kchunk][plane][2 row groups][8 rows][16 B]; plane 0 used
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 4096);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 1024; i += 128) reinterpret_cast<uint32_t*>(B)[i] = 0;
  __syncthreads();
  if (tid < 16) {
    const int n = tid, k = K0;
    *reinterpret_cast<__half*>(B + (k / 8) * 512 + (n / 8) * 128 + (n % 8) * 16 + (k % 8) * 2) = __float2half((float)(n + 1));
  }
  if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(slot, 64);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  // A: lane m, columns 32..39 (K = 16 halves = 8 columns); A[m][K0] = m + 1
  {
    uint32_t v[8];
    for (int c = 0; c < 8; ++c) v[c] = 0;
    const uint32_t h = (uint32_t)__half_as_ushort(__float2half((float)(tid + 1)));
    v[K0 / 2] = (K0 & 1) ? (h << 16) : h;
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + 32;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t db = make_desc(smem_u32(B), 512, 128, 0);  // LBO 512 (k chunk stride), SBO 128 (row group stride), no swizzle
    umma_f16_ts(tmem, tmem + 32, db, idesc, 0);
    umma_commit(bar);
  }
  mbar_wait_or_trap(bar, 0);
  tc_fence_after();
  uint32_t t[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]), "=r"(t[8]),
                 "=r"(t[9]), "=r"(t[10]), "=r"(t[11]), "=r"(t[12]), "=r"(t[13]), "=r"(t[14]), "=r"(t[15])
               : "r"(tmem + ((uint32_t)(warp * 32) << 16)));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int c = 0; c < 16; ++c) out[tid * 16 + c] = __uint_as_float(t[c]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

constexpr int kPiece = 2048, kCl = 16;
__global__ void __launch_bounds__(128, 1) k2(unsigned char* gbuf, int* result, long long* cycles, int rounds) {
  extern __shared__ unsigned char raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  unsigned char* tile = smem;                       // [2 buffers][16 pieces][2 KB]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + 2 * kCl * kPiece);  // [2]
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();
  const int tid = threadIdx.x;
  if (tid == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); mbar_fence_init(); }
  cluster.sync();
  unsigned char* mine = gbuf + ((size_t)blockIdx.x) * kPiece * 2;  // two alternating global pieces per CTA
  int ok = 1;
  long long t0 = clock64();
  for (int it = 0; it < rounds; ++it) {
    const int buf = it & 1;
    if (tid == 0) mbar_arrive_expect_tx(&full[buf], kCl * kPiece);
    // "epilogue": every thread writes 16 B of this CTA's piece (value = rank + it)
    uint4 v = make_uint4(rank + it, rank + it, rank + it, rank + it);
    reinterpret_cast<uint4*>(mine + buf * kPiece)[tid] = v;
    __syncthreads();
    if (tid == 0) {
      asm volatile("fence.proxy.async;" ::: "memory");
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
          ::"r"(smem_u32(tile + (size_t)buf * kCl * kPiece + rank * kPiece)), "l"(mine + buf * kPiece), "r"(kPiece),
          "r"(smem_u32(&full[buf])), "h"((unsigned short)0xffff)
          : "memory");
    }
    mbar_wait_or_trap(&full[buf], (it >> 1) & 1);
    for (int j = 0; j < kCl; ++j) {
      const uint4 g = reinterpret_cast<const uint4*>(tile + (size_t)buf * kCl * kPiece + j * kPiece)[tid];
      if (g.x != (unsigned)(j + it) || g.w != (unsigned)(j + it)) ok = 0;
    }
    __syncthreads();
  }
  long long t1 = clock64();
  if (!ok) atomicExch(result, 0);
  if (tid == 0 && blockIdx.x == 0) *cycles = t1 - t0;
  cluster.sync();
}

int main() {
  float* d;
  cudaMalloc(&d, 128 * 16 * 4);
  static float h[128 * 16];
  cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 1024);
  for (int K0 : {0, 1, 5, 8, 15}) {
    cudaMemset(d, 0, sizeof(h));
    k1<<<1, 128, 8 * 1024>>>(d, K0);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 16; ++n) bad += h[m * 16 + n] != (float)((m + 1) * (n + 1));
    printf("K1 K0=%2d: %s, mismatches %d  (D[0][0..3] = %g %g %g %g, D[5][2] = %g)\n", K0, cudaGetErrorString(e), bad, h[0], h[1],
           h[2], h[3], h[5 * 16 + 2]);
  }
  // K2
  unsigned char* g;
  int* res;
  long long* cyc;
  const int clusters = 8, rounds = 200;
  cudaMalloc(&g, (size_t)clusters * kCl * kPiece * 2);
  cudaMalloc(&res, 4);
  cudaMalloc(&cyc, 8);
  int one = 1;
  cudaMemcpy(res, &one, 4, cudaMemcpyHostToDevice);
  const size_t smem = 2 * kCl * kPiece + 1024 + 64 + 128 * 1024;  // + 128 KB ballast: the real kernel's footprint
  printf("K2 setattr smem: %s\n", cudaGetErrorString(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
  printf("K2 setattr nonportable: %s\n", cudaGetErrorString(cudaFuncSetAttribute(k2, cudaFuncAttributeNonPortableClusterSizeAllowed, 1)));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(clusters * kCl);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = kCl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int nclusters = -1;
  cudaError_t eo = cudaOccupancyMaxActiveClusters(&nclusters, k2, &cfg);
  printf("K2 max active clusters of 16 at %zu B smem: %d (%s)\n", smem, nclusters, cudaGetErrorString(eo));
  cudaError_t el = cudaLaunchKernelEx(&cfg, k2, g, res, cyc, rounds);
  cudaError_t es = cudaDeviceSynchronize();
  int r = -1; long long c = 0;
  cudaMemcpy(&r, res, 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("K2 launch %s / sync %s: data %s, %lld cycles per round (write 2 KB -> multicast to 16 CTAs -> 32 KB landed -> verified)\n",
         cudaGetErrorString(el), cudaGetErrorString(es), r == 1 ? "OK" : "WRONG", c / rounds);
  return 0;
}
