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
