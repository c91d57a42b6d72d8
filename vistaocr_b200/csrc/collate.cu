// Device-side batch assembly: the pixel work of the reference's SortByWidthCollater
// (src/datautils.py:61-176) - copy B ragged line images [C,H,w_i] into one zero-padded batch tensor [B,C,H,Wout] in
// width-sorted order - plus the concatenation of the int32 label sequences.  The host uploads the RAGGED pixels once
// (sum of widths, not B * max width) and computes the stable descending order of the B width keys (it needs the sorted
// widths on the host anyway: CnnOcrModel.forward turns them into sequence lengths there).  HBM-bound copy kernel,
// coalesced along x; one CTA row-group per (sorted sample, channel*row).
#include "common.cuh"

namespace vocr {

__global__ void __launch_bounds__(256)
collate_lines_kernel(const float* __restrict__ packed, const long long* __restrict__ img_offsets,
                     const int32_t* __restrict__ img_widths, const int32_t* __restrict__ order, int CH, int Wout,
                     float* __restrict__ out) {
  const int b = blockIdx.y;                    // position in the sorted batch
  const int src = order[b];
  const int w = min(img_widths[src], Wout);
  const float* sp = packed + img_offsets[src];
  float* dp = out + (size_t)b * CH * Wout;
  const long long total = (long long)CH * Wout;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / Wout), x = (int)(i - (long long)r * Wout);
    dp[i] = (x < w) ? __ldg(sp + (size_t)r * img_widths[src] + x) : 0.f;
  }
}

// labels_out = concatenation of the label sequences in sorted order; label_lens_out[b] = length of the b-th sorted one
__global__ void __launch_bounds__(256)
collate_labels_kernel(const int32_t* __restrict__ packed_labels, const int32_t* __restrict__ label_offsets,
                      const int32_t* __restrict__ order, int B, int32_t* __restrict__ labels_out,
                      int32_t* __restrict__ label_lens_out) {
  __shared__ int s_off[1025];
  if (threadIdx.x == 0) {
    int run = 0;
    for (int b = 0; b < B; ++b) {
      s_off[b] = run;
      run += label_offsets[order[b] + 1] - label_offsets[order[b]];
    }
    s_off[B] = run;
  }
  __syncthreads();
  for (int b = 0; b < B; ++b) {
    const int src0 = label_offsets[order[b]];
    const int n = s_off[b + 1] - s_off[b];
    if (threadIdx.x == 0) label_lens_out[b] = n;
    for (int j = threadIdx.x; j < n; j += blockDim.x) labels_out[s_off[b] + j] = packed_labels[src0 + j];
  }
}

}  // namespace vocr

using namespace vocr;

extern "C" int vocr_collate_lines_f32(const float* packed, const long long* img_offsets, const int32_t* img_widths,
                                      const int32_t* order, int B, int C, int H, int Wout, float* out,
                                      const int32_t* packed_labels, const int32_t* label_offsets, int32_t* labels_out,
                                      int32_t* label_lens_out, vocr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VOCR_REQUIRE(B >= 0 && B <= 1024 && C > 0 && H > 0 && Wout >= 0);
  if (B == 0) return VOCR_OK;
  VOCR_REQUIRE(img_offsets && img_widths && order && (Wout == 0 || (packed && out)));
  if (Wout > 0) {
    const long long per = (long long)C * H * Wout;
    dim3 grid((unsigned)min((long long)64, ceil_div64(per, 256)), B);
    collate_lines_kernel<<<grid, 256, 0, stream>>>(packed, img_offsets, img_widths, order, C * H, Wout, out);
    VOCR_CHECK_LAUNCH();
  }
  if (labels_out) {
    VOCR_REQUIRE(packed_labels && label_offsets && label_lens_out);
    collate_labels_kernel<<<1, 256, 0, stream>>>(packed_labels, label_offsets, order, B, labels_out, label_lens_out);
    VOCR_CHECK_LAUNCH();
  }
  return VOCR_OK;
}
