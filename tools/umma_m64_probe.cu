// Probe (GPU box only): where does tcgen05.mma cta_group::1 kind::f16 with M = 64 put row r, column n of D in TMEM?
// A[r][k] = (k == 0) ? r + 1 : 0, B[n][k] = (k == 0) ? n + 1 : 0  ->  D[r][n] = (r + 1) * (n + 1).
// All four warps dump their 32 lanes x 32 columns; the host prints, for every TMEM lane, the row it holds.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/probe tools/umma_m64_probe.cu && /tmp/probe
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "../vistaocr_b200/csrc/tc_common.cuh"
using namespace vocr;

__global__ void __launch_bounds__(128, 1) probe(float* out, int M, int N) {
  extern __shared__ unsigned char raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __half* A = reinterpret_cast<__half*>(smem);            // [128 rows][64 k] K-major SW128: 16 KB
  __half* B = reinterpret_cast<__half*>(smem + 16384);    // [64 rows][64 k]: 8 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 8192);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (16384 + 8192) / 2; i += 128) A[i] = __float2half(0.f);
  __syncthreads();
  auto put = [](__half* T, int r, int k, float v) {
    const int off = (r / 8) * 1024 + (r % 8) * 128 + (((k / 8) ^ (r % 8)) * 16) + (k % 8) * 2;
    *reinterpret_cast<__half*>(reinterpret_cast<unsigned char*>(T) + off) = __float2half(v);
  };
  if (tid < M) put(A, tid, 0, (float)(tid + 1));
  if (tid < N) put(B, tid, 0, (float)(tid + 1));
  if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(slot, 64);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint64_t da = make_desc(smem_u32(A), 16, 1024, 2), db = make_desc(smem_u32(B), 16, 1024, 2);
    umma_f16(tmem, da, db, idesc, 0);
    umma_commit(bar);
  }
  mbar_wait_or_trap(bar, 0);
  tc_fence_after();
  uint32_t t[32];
  tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), t);
  for (int c = 0; c < 32; ++c) out[tid * 32 + c] = __uint_as_float(t[c]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

int main() {
  float* d;
  cudaMalloc(&d, 128 * 32 * 4);
  static float h[128 * 32];
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int M : {64, 128}) {
    const int N = 32;
    cudaMemset(d, 0, sizeof(h));
    probe<<<1, 128, 40 * 1024>>>(d, M, N);
    cudaError_t e = cudaDeviceSynchronize();
    printf("M=%d N=%d: %s\n", M, N, cudaGetErrorString(e));
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    for (int lane = 0; lane < 128; ++lane) {
      // D[r][n] = (r+1)(n+1): column 0 gives r+1, column 1 must be twice that
      const float v0 = h[lane * 32 + 0], v1 = h[lane * 32 + 1], v31 = h[lane * 32 + 31];
      printf("lane %3d: row %3d  (col1/col0 = %.2f, col31/col0 = %.2f)%s", lane, (int)v0 - 1, v0 ? v1 / v0 : 0.f,
             v0 ? v31 / v0 : 0.f, (lane % 4 == 3) ? "\n" : "   ");
    }
  }
  return 0;
}
