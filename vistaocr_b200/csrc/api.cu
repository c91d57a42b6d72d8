// Version and status text of the C ABI (include/vistaocr_b200.h).
#include "common.cuh"

extern "C" int vocr_version(void) { return 1000; }

extern "C" const char* vocr_status_string(int status) {
  switch (status) {
    case VOCR_OK: return "success";
    case VOCR_MEMOPS_FAILED: return "cuda memcpy or memset failed";
    case VOCR_INVALID_VALUE: return "invalid value";
    case VOCR_EXECUTION_FAILED: return "execution failed";
    default: return "unknown error";
  }
}

// Arithmetic mode of the FP16-pair tensor-core kernels (vocr_tc_gemm_f16x3, vocr_tc_conv3x3_fwd_f16,
// vocr_tc_conv3x3_wgrad_f16), process-wide like a math-mode flag:
//   3 (default)  three error-compensated products per k-step: fp32-level accuracy (1e-5)
//   1            one product on the hi planes: fp16 operands (11-bit mantissa, per-tensor power-of-two scaling),
//                fp32 accumulation - the reduced-precision mode of BASELINE.json's cfg3 ("bf16 training")
// Every entry point takes the mode PER CALL (`products`: 3, 1, or 0 = this process-wide default), so two models / two
// threads of one process can run different modes without sharing mutable state on the launch path.
namespace vocr {
std::atomic<int> g_tc_products{3};
int resolve_tc_products(int products) {
  return products == 0 ? g_tc_products.load(std::memory_order_relaxed) : products;
}
}
extern "C" int vocr_set_tc_products(int n) {
  if (n != 1 && n != 3) return VOCR_INVALID_VALUE;
  vocr::g_tc_products.store(n, std::memory_order_relaxed);
  return VOCR_OK;
}
extern "C" int vocr_get_tc_products(void) { return vocr::g_tc_products.load(std::memory_order_relaxed); }
