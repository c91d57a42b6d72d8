// Microbenchmark: issue rate of legacy mma.sync on sm_100a (tf32 m16n8k8 vs f16 m16n8k16), 8 warps per SM, 1 CTA per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int KIND, int ILP>
__global__ void __launch_bounds__(256, 1) rate_kernel(float* out, int iters) {
  float d[ILP][4];
  uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 11u}, b[2] = {threadIdx.x * 5u, 13u};
#pragma unroll
  for (int j = 0; j < ILP; ++j) d[j][0] = d[j][1] = d[j][2] = d[j][3] = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < ILP; ++j) {
      if (KIND == 0)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[j][0]), "+f"(d[j][1]), "+f"(d[j][2]), "+f"(d[j][3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
      else if (KIND == 1)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[j][0]), "+f"(d[j][1]), "+f"(d[j][2]), "+f"(d[j][3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
      else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[j][0]), "+f"(d[j][1]), "+f"(d[j][2]), "+f"(d[j][3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < ILP; ++j) s += d[j][0] + d[j][1] + d[j][2] + d[j][3];
  out[blockIdx.x * 256 + threadIdx.x] = s;
}
template <int KIND, int ILP>
void run(const char* name, int k) {
  float* out;
  cudaMalloc(&out, 148 * 256 * 4);
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  rate_kernel<KIND, ILP><<<148, 256>>>(out, 100);
  cudaEventRecord(e0);
  rate_kernel<KIND, ILP><<<148, 256>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double mmas_per_sm = 8.0 * ILP * iters;
  double fma = mmas_per_sm * 16 * 8 * k;
  printf("%s ILP=%d: %.3f ms, %.1f ns per MMA per SM (8 warps), %.0f FMA/ns/SM, %.1f TFLOP/s chip\n", name, ILP, ms,
         ms * 1e6 / mmas_per_sm, fma / (ms * 1e6), 2 * fma * 148 / (ms * 1e-3) / 1e12);
  cudaFree(out);
}
int main() {
  run<0, 4>("tf32 m16n8k8 ", 8);
  run<0, 8>("tf32 m16n8k8 ", 8);
  run<1, 4>("f16  m16n8k16", 16);
  run<1, 8>("f16  m16n8k16", 16);
  run<2, 8>("bf16 m16n8k16", 16);
  return 0;
}
