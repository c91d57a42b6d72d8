// Line-image pre-processing in front of the hot path, on the device:
//   [ConvertGray]  cv2.cvtColor(BGR2GRAY)          reference src/imagetransforms.py:411-416
//   Scale(new_h)   cv2.resize, see below            reference src/imagetransforms.py:453-507
//   [InvertBlackWhite]  255 - v                     reference src/imagetransforms.py:383-385
//   ToTensor       float(v) / 255                   reference src/imagetransforms.py:423-434
//   width floor of 15 px padded with ones           reference src/ocr_dataset.py:174-180
//   zero padding to the batch width, width-sorted   reference src/datautils.py:61-176 (SortByWidthCollater)
// fused into ONE kernel that reads the raw decoded uint8 pixels of B ragged images and writes the padded float batch
// [B,1,H,Wout]: the resized uint8 images and the per-image float tensors of the reference never exist.
//
// Scale: the reference passes its INTER_CUBIC default as cv2.resize's third POSITIONAL argument, which is `dst`, so
// OpenCV runs its default INTER_LINEAR.  The 8-bit bilinear path of OpenCV (modules/imgproc/src/resize.cpp) is integer
// arithmetic and is restated bit for bit (oracle/preproc_ref.py, pinned against cv2):
//   scale = 1 / (dst / src) in float64;  f = float32((d + 0.5) * scale - 0.5);  s = floor(f);  f -= s;  x borders reset
//   (s, f) to (0, 0) / (w - 1, 0), y clamps the two rows;  weights round_half_even(f * 2048) as int16;
//   horizontal pass D = S[s] * a0 + S[s + 1] * a1 (int32);  vertical pass ((b0 * (D0 >> 4)) >> 16) + ((b1 * (D1 >> 4)) >> 16)
//   + 2) >> 2;  exact 2x down-scaling in both directions is OpenCV's INTER_AREA fast path (s00 + s01 + s10 + s11 + 2) >> 2.
// HBM-bound byte work: a thread owns one output column (its horizontal taps are computed once, in registers) and walks
// the H output rows; stores are coalesced along x, the source bytes of a row pair are shared through L1.
#include "common.cuh"

namespace vocr {

struct Tap {
  int s;       // first source index
  int a0, a1;  // 11-bit fixed-point weights
};

// OpenCV's coefficient computation for output index d of a dst-long axis resized from src (explicit _rn intrinsics: no
// FMA contraction, the float64 / float32 roundings are OpenCV's)
__device__ __forceinline__ void linear_coeff(int d, int dst, int src, int& s, float& f) {
  const double inv_scale = __ddiv_rn((double)dst, (double)src);
  const double scale = __ddiv_rn(1.0, inv_scale);
  const float ff = __double2float_rn(__dsub_rn(__dmul_rn(__dadd_rn((double)d, 0.5), scale), 0.5));
  s = __float2int_rd(ff);
  f = __fsub_rn(ff, (float)s);
}
__device__ __forceinline__ int fix11(float c) {
  const int v = __float2int_rn(__fmul_rn(c, 2048.f));  // cvRound: round half to even
  return max(-32768, min(32767, v));
}

template <int CN>
__device__ __forceinline__ int load_gray(const uint8_t* __restrict__ p, long long idx) {
  if (CN == 1) return (int)__ldg(p + idx);
  const uint8_t* q = p + idx * 3;  // BGR -> gray, OpenCV's 15-bit coefficients
  return ((int)__ldg(q) * 3735 + (int)__ldg(q + 1) * 19235 + (int)__ldg(q + 2) * 9798 + (1 << 14)) >> 15;
}

template <int CN>
__global__ void __launch_bounds__(256)
scale_lines_kernel(const uint8_t* __restrict__ packed, const long long* __restrict__ img_offsets,
                   const int32_t* __restrict__ src_h, const int32_t* __restrict__ src_w,
                   const int32_t* __restrict__ dst_w, const int32_t* __restrict__ order, int H, int Wout, int invert,
                   int min_width, float* __restrict__ out) {
  extern __shared__ int s_rows[];  // [H][4]: r0, r1, b0, b1
  const int b = blockIdx.y;        // position in the (sorted) batch
  const int i = order ? order[b] : b;
  const int h = src_h[i], w = src_w[i], dw = min(dst_w[i], Wout);
  const uint8_t* src = packed + img_offsets[i];
  float* dst = out + (size_t)b * H * Wout;
  const bool same = (dw == w) && (H == h);
  const bool area2 = !same && (w == 2 * dw) && (h == 2 * H);  // scale == 2 exactly in both directions
  for (int y = threadIdx.x; y < H; y += blockDim.x) {
    int sy;
    float fy;
    linear_coeff(y, H, h, sy, fy);
    s_rows[4 * y + 0] = max(0, min(sy, h - 1));
    s_rows[4 * y + 1] = max(0, min(sy + 1, h - 1));
    s_rows[4 * y + 2] = fix11(1.f - fy);
    s_rows[4 * y + 3] = fix11(fy);
  }
  __syncthreads();
  const int pad_to = min(Wout, max(dw, min_width));
  for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < Wout; x += gridDim.x * blockDim.x) {
    if (x >= dw) {  // ones up to the width floor, zeros beyond (batch padding)
      const float v = (x < pad_to) ? 1.f : 0.f;
      for (int y = 0; y < H; ++y) dst[(size_t)y * Wout + x] = v;
      continue;
    }
    Tap tp;
    {
      float fx;
      linear_coeff(x, dw, w, tp.s, fx);
      if (tp.s < 0) tp.s = 0, fx = 0.f;
      if (tp.s >= w - 1) tp.s = w - 1, fx = 0.f;
      tp.a0 = fix11(1.f - fx);
      tp.a1 = fix11(fx);
    }
    const int s1 = min(tp.s + 1, w - 1);  // weight 0 wherever this clamps
    for (int y = 0; y < H; ++y) {
      int v;
      if (same) {
        v = load_gray<CN>(src, (long long)y * w + x);
      } else if (area2) {
        const long long r0 = (long long)(2 * y) * w + 2 * x, r1 = r0 + w;
        v = (load_gray<CN>(src, r0) + load_gray<CN>(src, r0 + 1) + load_gray<CN>(src, r1) + load_gray<CN>(src, r1 + 1) + 2) >> 2;
      } else {
        const long long r0 = (long long)s_rows[4 * y + 0] * w, r1 = (long long)s_rows[4 * y + 1] * w;
        const int b0 = s_rows[4 * y + 2], b1 = s_rows[4 * y + 3];
        const int d0 = load_gray<CN>(src, r0 + tp.s) * tp.a0 + load_gray<CN>(src, r0 + s1) * tp.a1;
        const int d1 = load_gray<CN>(src, r1 + tp.s) * tp.a0 + load_gray<CN>(src, r1 + s1) * tp.a1;
        v = (((b0 * (d0 >> 4)) >> 16) + ((b1 * (d1 >> 4)) >> 16) + 2) >> 2;
      }
      v &= 255;
      if (invert) v = 255 - v;
      dst[(size_t)y * Wout + x] = __fdiv_rn((float)v, 255.f);
    }
  }
}

}  // namespace vocr

using namespace vocr;

// out[B,1,H,Wout] (fp32) = pre-processed line images in batch order b -> source image order[b] (NULL = identity).
//   packed: the decoded uint8 images back to back (gray: h*w bytes; channels = 3: h*w*3 bytes, BGR interleaved, converted
//   to gray like cv2.cvtColor);  img_offsets[B] byte offsets;  src_h, src_w, dst_w [B] (dst_w = Scale's
//   int(w * float(H / h)), computed by the caller in float64);  columns [dst_w, max(dst_w, min_width)) = 1, the rest 0.
extern "C" int vocr_scale_lines_u8(const uint8_t* packed, const long long* img_offsets, const int32_t* src_h,
                                   const int32_t* src_w, const int32_t* dst_w, const int32_t* order, int B, int channels,
                                   int H, int Wout, int invert, int min_width, float* out, vocr_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VOCR_REQUIRE(B >= 0 && H > 0 && H <= 4096 && Wout >= 0 && (channels == 1 || channels == 3) && min_width >= 0);
  if (B == 0 || Wout == 0) return VOCR_OK;
  VOCR_REQUIRE(packed && img_offsets && src_h && src_w && dst_w && out && B <= 65535);
  dim3 grid((unsigned)ceil_div(Wout, 256), (unsigned)B);
  const size_t smem = sizeof(int) * 4 * (size_t)H;
  if (channels == 1)
    scale_lines_kernel<1><<<grid, 256, smem, stream>>>(packed, img_offsets, src_h, src_w, dst_w, order, H, Wout, invert,
                                                       min_width, out);
  else
    scale_lines_kernel<3><<<grid, 256, smem, stream>>>(packed, img_offsets, src_h, src_w, dst_w, order, H, Wout, invert,
                                                       min_width, out);
  VOCR_CHECK_LAUNCH();
  return VOCR_OK;
}
