/*
 * vistaocr_b200 — C ABI of the B200-native (sm_100a) line-recognition hot path of isi-vista/VistaOCR.
 *
 * The reference has no FFI of its own for this path: it reaches native code through three Python seams
 * (SURVEY.md §8b).  Every entry point below names the reference interface it replaces (file:line relative to
 * the reference tree).  Conventions, identical for all entry points:
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless its name ends in `_host`;
 *   - the caller owns every buffer, including workspaces; no entry point allocates or synchronises;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); distinct streams are thread-safe;
 *   - the return value is a vocr_status_t (0 = success), mirroring warp-ctc's ctcStatus_t
 *     {SUCCESS, MEMOPS_FAILED, INVALID_VALUE, EXECUTION_FAILED, UNKNOWN_ERROR};
 *   - activations inside the CNN are NHWC fp32 ("pixels x channels"), sequences are time-major [T,B,*].
 */
#ifndef VISTAOCR_B200_H_
#define VISTAOCR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  VOCR_OK = 0,
  VOCR_MEMOPS_FAILED = 1,
  VOCR_INVALID_VALUE = 2,
  VOCR_EXECUTION_FAILED = 3,
  VOCR_UNKNOWN_ERROR = 4
} vocr_status_t;

typedef void* vocr_stream_t; /* cudaStream_t */

/* ABI version (major*1000 + minor) and status text. */
int vocr_version(void);
const char* vocr_status_string(int status);

/* Arithmetic mode of the FP16-pair tensor-core entry points (vocr_tc_gemm_f16x3, vocr_tc_conv3x3_fwd_f16,
 * vocr_tc_conv3x3_wgrad_f16): 3 = three error-compensated products per k-step, fp32-level accuracy; 1 = one product on
 * the hi planes only (fp16 operands, fp32 accumulation) - the reduced-precision mode of BASELINE.json's cfg3.  Those
 * entry points take the mode PER CALL (`products`); the setter below only changes what `products = 0` means (an atomic
 * process default, 3 initially) - no launch path reads shared mutable state otherwise.  The reference has no counterpart
 * (it trains in fp32, src/train_cnn_lstm.py). */
int vocr_set_tc_products(int n);
int vocr_get_tc_products(void);

/* ------------------------------------------------------------------------------------------------------------
 * Greedy CTC decode.  Replaces ArgmaxDecoder.decode (src/decoder.py:116-185) and its twin
 * CnnOcrModel.decode_without_lm (src/models/cnnlstm.py:479-541): per-frame argmax over the alphabet on RAW
 * logits (first index wins ties, NaN counts as maximal like numpy), blank (0) / low-confidence
 * (max < thresh, float32 compare) frames reset the previous character, repeats collapse, frames t >= lens[b]
 * are ignored.
 *   logits  [T,B,A] fp32 contiguous      lens [B] int32
 *   canon   [A] int32 or NULL: canonical index per symbol (two alphabet entries with the same string
 *           collapse like the reference's string compare); NULL = identity
 *   path    [B,T] int32 out: per-frame label after blank/threshold mapping, -1 for t >= lens[b]
 *           (the "CTC alignment" of src/utils/visualization.py:111-157)
 *   labels  [B,ld] int32 out: collapsed label sequence;  counts [B] int32 out: its length
 * ---------------------------------------------------------------------------------------------------------- */
int vocr_greedy_decode_f32(const float* logits, int T, int B, int A, const int32_t* lens, float thresh,
                           const int32_t* canon, int32_t* path, int32_t* labels, int32_t* counts, int ld,
                           vocr_stream_t stream);
/* The two halves separately, for the fused inference path (the logits never reach HBM):
 * vocr_tc_gemm_f16x3_argmax: the prob layer (src/models/cnnlstm.py:152-154; A [T*B,K] activations and W [N,K] weights as
 *   K-major FP16 pair planes, N <= 128, K <= 1536, row m = t*B + b) with the per-frame arg-max / blank / threshold mapping
 *   above in the GEMM epilogue: path[b*T + t] as vocr_greedy_decode_f32 writes it.  C (the logits, ldc) may be NULL.
 * vocr_ctc_collapse_i32: path -> labels / counts (collapse repeats, drop blanks, canon as above). */
int vocr_tc_gemm_f16x3_argmax(int M, int N, int K, const uint16_t* a_hi, const uint16_t* a_lo, int lda,
                              const int32_t* exp_a, const uint16_t* b_hi, const uint16_t* b_lo, int ldb,
                              const int32_t* exp_b, float* C, int ldc, const float* bias, const int32_t* lens, int T, int B,
                              float thresh, int32_t* path, int products, vocr_stream_t stream);
int vocr_ctc_collapse_i32(const int32_t* path, int T, int B, const int32_t* lens, const int32_t* canon, int32_t* labels,
                          int32_t* counts, int ld, vocr_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * CTC loss forward + gradient.  Replaces warpctc_pytorch.CTCLoss / warp-ctc compute_ctc_loss +
 * get_workspace_size (call sites src/train_cnn_lstm.py:12,52,138,358).  Blank = 0, softmax inside (acts are
 * raw), log-space fp32.  costs[b] = -log p(labels_b | acts[:act_lens[b], b, :]); grads = softmax - occupancy
 * for t < act_lens[b], exactly 0 beyond; infeasible utterances (act_len < L + repeats) give cost 0 / grad 0.
 *   acts [T,B,A] fp32;  grads [T,B,A] fp32 out or NULL (loss only)
 *   labels: concatenated int32 labels (sum of label_lens), values in [1,A);  label_lens, act_lens: [B] int32
 *   costs [B] fp32 out;  workspace: vocr_ctc_workspace_size(...) bytes, 256-B aligned
 * ---------------------------------------------------------------------------------------------------------- */
size_t vocr_ctc_workspace_size(int T, int B, int A, int max_label_len);
int vocr_ctc_loss_f32(const float* acts, float* grads, const int32_t* labels, const int32_t* label_lens,
                      const int32_t* act_lens, int T, int B, int A, int max_label_len, float* costs,
                      void* workspace, size_t workspace_bytes, vocr_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Dense fp32 GEMM with fused epilogue.  Replaces the cuBLAS calls behind nn.Linear in bridge_layer / prob_layer
 * (src/models/cnnlstm.py:143-146,152-154,278,294), the LSTM input projections (nn.LSTM, :148-149) and their
 * backward GEMMs.  Row-major.  C[M,N] = op(A)[M,K] * op(B)[K,N] (+ bias[n]) (+ C if accumulate) (ReLU if relu).
 *   transa = 0: A is [M,K] (lda >= K)   transa = 1: A is stored [K,M] (lda >= M)
 *   transb = 0: B is [K,N] (ldb >= N)   transb = 1: B is stored [N,K] (ldb >= K)   (nn.Linear weight layout)
 * vocr_colsum_f32: out[c] (+)= sum_r x[r*ld + c]  (bias gradients).
 * ---------------------------------------------------------------------------------------------------------- */
int vocr_gemm_f32(int transa, int transb, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                  float* C, int ldc, const float* bias, int relu, int accumulate, vocr_stream_t stream);
int vocr_colsum_f32(const float* x, long long rows, int cols, int ld, float* out, int accumulate,
                    vocr_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * CNN feature extractor, NHWC fp32.  Replaces cuDNN conv / ATen pooling / cuDNN BatchNorm behind
 * CnnOcrModel.rapid_ds and .cnn (src/models/cnnlstm.py:114-134,263-266,270-271).
 *
 * vocr_conv_weight_layout_f32: W[Cout,Cin,3,3] (state_dict layout) -> wk[(ky,kx,ci),co] (forward) and/or
 *   wd[(ky,kx,co),ci] = W[co,ci,2-ky,2-kx] (data gradient); either output may be NULL.
 * vocr_conv3x3_fwd_f32: z[B,H,W,Cout] = conv3x3_pad1(x[B,H,W,Cin]; wk) + bias.  If stats != NULL the per-channel
 *   sum and sum of squares of z are ACCUMULATED into stats[0:Cout], stats[Cout:2Cout] (float64; caller zeroes).
 *   zmax (optional device float the caller zeroes) receives max |z| (see vocr_tc_conv3x3_fwd_f16).
 *   The data gradient is the same call with (x := dz, wk := wd, Cin/Cout swapped, bias = stats = zmax = NULL).
 * vocr_conv3x3_wgrad_f32: dw[Cout,Cin,3,3] = sum over pixels (state_dict layout, deterministic split-K).
 * vocr_rds_fwd_f32: rapid-downsample stage, y[B,H/2,W/2,16] = maxpool2x2(relu(conv3x3(x[B,H,W,Cin]; wk[9*Cin,16])
 *   + bias)); arg (uint8 per output, may be NULL) records the winning window position for the backward pass.
 * vocr_rds_unpool_f32: dpre[B,H,W,16] = gradient w.r.t. the conv output of that stage (zeros where the position
 *   lost the max or ReLU was inactive); conv gradients then use vocr_conv3x3_wgrad_f32 / _fwd_f32.
 * ---------------------------------------------------------------------------------------------------------- */
int vocr_conv_weight_layout_f32(const float* w, int Cin, int Cout, float* wk, float* wd, vocr_stream_t stream);
int vocr_conv3x3_fwd_f32(const float* x, const float* wk, const float* bias, float* z, int B, int H, int W, int Cin,
                         int Cout, double* stats, float* zmax, vocr_stream_t stream);
size_t vocr_conv3x3_wgrad_workspace_size(int B, int H, int W, int Cin, int Cout);
int vocr_conv3x3_wgrad_f32(const float* x, const float* dz, float* dw, int B, int H, int W, int Cin, int Cout,
                           void* workspace, size_t workspace_bytes, vocr_stream_t stream);
int vocr_rds_fwd_f32(const float* x, const float* wk, const float* bias, float* y, uint8_t* arg, int B, int H, int W,
                     int Cin, vocr_stream_t stream);
int vocr_rds_unpool_f32(const float* dy, const float* y, const uint8_t* arg, float* dpre, int B, int H, int W,
                        vocr_stream_t stream);
/* first stage (Cin = 1, no data gradient): dw[16,1,3,3] and db[16] straight from the pooled gradient; ws: double[160] */
int vocr_rds_wgrad_c1_f32(const float* x, const float* dy, const float* y, const uint8_t* arg, float* dw, float* db,
                          int B, int H, int W, double* ws, vocr_stream_t stream);
/* 16 -> 16 channel 3x3 / pad 1 convolution of the second rapid-downsample stage (src/models/cnnlstm.py:96-121 at line height
 * 120) and its weight / bias gradient as direct FFMA kernels: with 16 channels on both sides the implicit-GEMM kernels
 * waste most of their tiles.  x, z, dz [B,H,W,16] NHWC; wk [144][16] = [(ky,kx,ci)][co] (the data gradient passes the
 * flipped matrix wd of vocr_conv_weight_layout_f32); dw [16,16,3,3], db [16]; ws: float64[2320] scratch. */
int vocr_conv3x3_c16_fwd_f32(const float* x, const float* wk, const float* bias, float* z, int B, int H, int W,
                             vocr_stream_t stream);
int vocr_conv3x3_c16_wgrad_f32(const float* x, const float* dz, float* dw, float* db, int B, int H, int W, double* ws,
                               vocr_stream_t stream);

/* BatchNorm2d(eps, momentum) + ReLU (src/models/cnnlstm.py:263-266).
 * vocr_bn_finalize_f32: training != 0: batch statistics from stats (see conv fwd) over `count` pixels, running
 *   stats updated in place (unbiased variance); else running stats.  Emits scale = gamma*invstd,
 *   shift = beta - mean*scale and (optionally) save_mean / save_invstd for the backward pass.
 * vocr_bn_relu_apply_f32: a[b*sB + y*sH + x*sW + c] = relu(z[b,y,x,c]*scale[c] + shift[c])  (strides in elements,
 *   c contiguous; a may be NULL when only the FP16 pair planes are wanted - inference, consumer = a tensor-core conv) - the last CNN block writes the [T,B,h*C] sequence layout directly (replaces the
 *   permute+contiguous of cnnlstm.py:276).  a_hi / a_lo (both or neither; same layout as a) optionally receive the
 *   TF32 split planes the tensor-core conv of the next block reads (saves a vocr_split_tf32_f32 pass); dz_hi / dz_lo
 *   of vocr_bn_relu_bwd_f32 likewise.
 * vocr_bn_relu_bwd_f32: da (same strided layout) -> dz[B,H,W,C] (may be NULL when only the FP16 pair planes dz_hi16 /
 *   dz_lo16 are wanted: both gradient convolutions then read the planes), dgamma, dbeta, dbias (= sum dz, the gradient of
 *   the conv bias in front of the BatchNorm; may be NULL).  red_ws: float64[3*C] scratch.
 * FP16 pair planes (see vocr_split_f16_f32) come out of the same kernels without an extra pass over the tensor:
 *   vocr_bn_finalize_f32 aux (device float[2], optional): [0] = an upper bound of the activations - analytic with batch
 *     statistics (|xhat| <= sqrt(count)); with running statistics max_c(|scale_c| zmax + |shift_c|) when the conv's
 *     measured zmax = max |z| is passed, 0 otherwise -, [1] = max_c |scale_c|.
 *   vocr_bn_relu_apply_f32 a_hi16 / a_lo16 / bound (= aux) / pair_exp (device int32, receives the exponent).
 *   vocr_bn_relu_bwd_f32 dz_hi16 / dz_lo16 / scale_max (= aux + 1) / pair_state (device int32[2]: [0] receives the
 *     exponent, [1] scratch); the bound max|scale| * max|g| * (2 + sqrt(N)) is formed on the device.
 */
int vocr_bn_finalize_f32(const double* stats, long long count, const float* gamma, const float* beta,
                         float* running_mean, float* running_var, float momentum, float eps, int training,
                         float* scale, float* shift, float* save_mean, float* save_invstd, int C, float* aux,
                         const float* zmax, vocr_stream_t stream);
int vocr_bn_relu_apply_f32(const float* z, const float* scale, const float* shift, float* a, float* a_hi, float* a_lo,
                           int B, int H, int W, int C, long long sB, long long sH, long long sW, uint16_t* a_hi16,
                           uint16_t* a_lo16, const float* bound, int32_t* pair_exp, vocr_stream_t stream);
int vocr_bn_relu_bwd_f32(const float* da, const float* z, const float* scale, const float* shift,
                         const float* save_mean, const float* save_invstd, int training, int B, int H, int W, int C,
                         long long sB, long long sH, long long sW, float* dz, float* dz_hi, float* dz_lo,
                         float* dgamma, float* dbeta, float* dbias, double* red_ws, uint16_t* dz_hi16,
                         uint16_t* dz_lo16, const float* scale_max, int32_t* pair_state, vocr_stream_t stream);

/* FractionalMaxPool2d(2, output_ratio=(0.5,0.7)) with explicit per-(sample,channel) samples[B,C,2]
 * (src/models/cnnlstm.py:127,130; ATen interval rule, random in train AND eval).  idx (int32, may be NULL) keeps the
 * flat input position h*W+w of each winner; backward scatters dy through idx into dx (zeroed inside) in four
 * race-free passes over the (row parity, column parity) classes of the windows: deterministic, no atomics. */
int vocr_fracpool_fwd_f32(const float* x, const float* samples, float* y, int32_t* idx, int B, int H, int W, int C,
                          int Ho, int Wo, vocr_stream_t stream);
int vocr_fracpool_bwd_f32(const float* dy, const int32_t* idx, float* dx, int B, int H, int W, int C, int Ho,
                          int Wo, vocr_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Bidirectional LSTM layer recurrence (persistent cooperative kernels).  Replaces the cuDNN RNN behind
 * nn.LSTM(bidirectional) on a packed sequence (src/models/cnnlstm.py:148-149,285-290).  PyTorch gate order i,f,g,o.
 *   xproj [T,B,2,4H] = x W_ih^T + b_ih + b_hh for both directions (one vocr_gemm_f32);  whh [2,4H,H];
 *   lens [B] int32 (device): sample b is processed for t < lens[b]; the reverse direction starts at lens[b]-1;
 *   out [T,B,2H] (forward | reverse), zero for t >= lens[b];  Tmax = max(lens) <= T.
 *   gates [T,B,2,4H] / cst [T,B,2,H]: activated gates and cell states saved for backward (NULL for inference).
 * backward: dout [T,B,2H] -> dgates [T,B,2,4H] = gradient w.r.t. xproj (zero beyond lens); the weight / input
 * gradients are GEMMs over dgates (see vistaocr_b200/ops.py).  H <= 512.  dgates_absmax (device float, optional)
 * receives max |dgates|: the operand bound of those GEMMs, so that their FP16-pair split needs no abs-max pass.
 * workspace: vocr_bilstm_workspace_size(T = Tmax, B, H, backward) bytes - the forward pass exchanges h_t between the CTAs
 * of a direction through one sentinel-initialised slot per step (no flags, no fences), hence the dependence on T.
 * ---------------------------------------------------------------------------------------------------------- */
size_t vocr_bilstm_workspace_size(int T, int B, int H, int backward);
int vocr_bilstm_fwd_f32(const float* xproj, const float* whh, const int32_t* lens, float* out, float* gates,
                        float* cst, int T, int B, int H, int Tmax, void* workspace, size_t workspace_bytes,
                        vocr_stream_t stream);
int vocr_bilstm_bwd_f32(const float* dout, const float* whh, const int32_t* lens, const float* gates,
                        const float* cst, float* dgates, float* dgates_absmax, int T, int B, int H, int Tmax,
                        void* workspace, size_t workspace_bytes, vocr_stream_t stream);

/* Inter-layer LSTM dropout (nn.LSTM(dropout=p), src/models/cnnlstm.py:148-149; p = 0.5 at src/train_cnn_lstm.py:331;
 * training only, on the output of every layer but the last).  y[i] = keep[i] ? x[i] * 1/(1-p) : 0 (in place allowed).
 * keep comes from mask_in (uint8[n], 1 = keep) when given - the parity tests inject the oracle's mask - else from a
 * Philox4x32-10 counter stream: element i uses word i&3 of Philox(counter {i>>2, offset}, key seed), kept iff
 * word >= floor(p 2^32).  Nothing is stored: the backward pass is the same call on dy with the same (seed, offset).
 * rng (device uint64[2] = {seed, offset}, optional) replaces `seed` and is ADDED to `offset`, so a captured CUDA graph
 * draws a fresh mask per replay (vocr_rng_advance bumps rng[1] on the stream); rng_used (device uint64[2], optional)
 * receives the effective pair for the backward call.  mask_out (optional) receives the keep mask; x = y = NULL with
 * mask_out set only writes the mask. */
int vocr_dropout_f32(const float* x, float* y, long long n, float p, const unsigned long long* rng,
                     unsigned long long seed, unsigned long long offset, unsigned long long* rng_used,
                     const uint8_t* mask_in, uint8_t* mask_out, vocr_stream_t stream);
int vocr_rng_advance(unsigned long long* rng, unsigned long long inc, vocr_stream_t stream);

/* Fused element-wise gradient clamp to [-clamp,clamp] + torch.optim.Adam update over a flat parameter buffer
 * (src/train_cnn_lstm.py:143-149,363).  step >= 1; clamp <= 0 disables; grad_scale pre-multiplies the gradient. */
int vocr_clamp_adam_f32(float* p, const float* g, float* m, float* v, long long n, int step, float lr, float beta1,
                        float beta2, float eps, float weight_decay, float clamp, float grad_scale,
                        vocr_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Tensor-core GEMM (TMA + tcgen05.mma kind::tf32 + TMEM), error-compensated 3xTF32: fp32-level accuracy at tensor-core
 * speed.  Same role as vocr_gemm_f32 (src/models/cnnlstm.py:143-154 -> cuBLAS).  Operands arrive pre-split by
 * vocr_split_tf32_f32 (hi = rna_tf32(x), lo = x - hi; both planes keep x's layout):
 *   a_mn = 0: A planes [M,K] row-major (lda)      a_mn = 1: A planes stored [K,M] (lda)
 *   b_mn = 0: B planes [N,K] row-major (ldb)      b_mn = 1: B planes stored [K,N] (ldb)
 * lda, ldb multiples of 4; plane bases 16-B aligned.  C[M,N] (ldc) = op(A) op(B) (+bias[n]) (+C) (ReLU).
 * workspace (optional): lets long reductions with few output tiles (weight gradients) run split-K over all SMs.
 * ---------------------------------------------------------------------------------------------------------- */
int vocr_split_tf32_f32(const float* x, float* hi, float* lo, long long n, vocr_stream_t stream);
int vocr_tc_gemm_tf32x3(int a_mn, int b_mn, int M, int N, int K, const float* a_hi, const float* a_lo, int lda,
                        const float* b_hi, const float* b_lo, int ldb, float* C, int ldc, const float* bias, int relu,
                        int accumulate, void* workspace, size_t workspace_bytes, vocr_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * FP16 pair operands: the same error-compensated three-product scheme on tcgen05.mma kind::f16 - twice the K per
 * instruction and half the operand bytes of the TF32 planes, same fp32-level accuracy (22 significant bits).
 * vocr_split_f16_f32: x[n] -> hi = fp16(x 2^e), lo = fp16((x 2^e - hi) 2^11), e chosen so that bound 2^e lies in
 *   [2^14, 2^15).  state: device int32[2], [0] receives e, [1] is scratch.  bound: optional device float >= max|x|
 *   (when NULL an absmax pass computes it).  Everything stays on the stream - no host synchronisation.
 * vocr_tc_gemm_f16x3 / vocr_tc_conv3x3_fwd_f16 / vocr_tc_conv3x3_wgrad_f16: as their TF32 namesakes, taking the planes
 *   and the device exponent of each operand; lda / ldb multiples of 8, Cin % 64 == 0 (and Cout % 64 == 0 for wgrad).
 *   zmax (vocr_tc_conv3x3_fwd_f16, optional device float the caller zeroes): the epilogue max-es |z| into it, which lets
 *   vocr_bn_finalize_f32 bound the activations of an eval-mode block so that vocr_bn_relu_apply_f32 can write the next
 *   layer's operand planes without an abs-max pass.
 *   products: arithmetic mode of THIS call - 3 = three compensated products (fp32-level), 1 = hi planes only (fp16
 *   operands, fp32 accumulation), 0 = the process default of vocr_set_tc_products.
 * ---------------------------------------------------------------------------------------------------------- */
int vocr_split_f16_f32(const float* x, long long n, const float* bound, int32_t* state, uint16_t* hi, uint16_t* lo,
                       vocr_stream_t stream);
/* 3x3 / pad 1 patches of a narrow NHWC activation (C % 4 == 0, meant for C < 64) as FP16 pair planes
 * cols[B*H*W][9*C], cols[p][tap*C + c] = x[p + tap][c]: the conv of such a layer and its weight gradient then run as
 * K = 9C GEMMs on vocr_tc_gemm_f16x3 (cols W^T and cols^T dz).  bound / state as in vocr_split_f16_f32. */
int vocr_im2col3x3_f16(const float* x, int B, int H, int W, int C, const float* bound, int32_t* state, uint16_t* hi,
                       uint16_t* lo, vocr_stream_t stream);
int vocr_tc_gemm_f16x3(int a_mn, int b_mn, int M, int N, int K, const uint16_t* a_hi, const uint16_t* a_lo, int lda,
                       const int32_t* exp_a, const uint16_t* b_hi, const uint16_t* b_lo, int ldb, const int32_t* exp_b,
                       float* C, int ldc, const float* bias, int relu, int accumulate, void* workspace,
                       size_t workspace_bytes, int products, vocr_stream_t stream);
int vocr_tc_conv3x3_fwd_f16(const uint16_t* x_hi, const uint16_t* x_lo, const int32_t* exp_x, const uint16_t* w_hi,
                            const uint16_t* w_lo, const int32_t* exp_w, const float* bias, float* z, int B, int H, int W,
                            int Cin, int Cout, int products, float* zmax, vocr_stream_t stream);
int vocr_tc_conv3x3_wgrad_f16(const uint16_t* x_hi, const uint16_t* x_lo, const int32_t* exp_x, const uint16_t* dz_hi,
                              const uint16_t* dz_lo, const int32_t* exp_dz, float* dw, int B, int H, int W, int Cin,
                              int Cout, void* workspace, size_t workspace_bytes, int products, vocr_stream_t stream);
/* Inference form of one Conv + BatchNorm + ReLU unit (src/models/cnnlstm.py:263-266 with running statistics) as ONE
 * kernel: the convolution's epilogue applies  a = relu((conv3x3(x) + bias) * scale[c] + shift[c])  (scale / shift from
 * vocr_bn_finalize_f32 with training = 0) and writes a with strides (sB, sH, sW) - NHWC or the time-major sequence
 * layout of the last block - and / or the FP16 pair planes of a (dense NHWC) that the next tensor-core convolution reads;
 * either output may be NULL.  The planes are scaled by 2^e, e from the device scalar `bound` >= max |a| and stored to
 * pair_exp[0]; vocr_bn_eval_bound_f32 derives such a bound from the weights alone (aux[0] = max_c |scale_c| (sum_k
 * |w_ck| xbound + |bias_c|) + |shift_c|, atomicMax into aux[0], which vocr_bn_finalize_f32 zeroes; w [C][K] fp32).
 * amax (optional device float the caller zeroes) receives the MEASURED max a: the next block's analytic bound starts from
 * it, so the looseness of the sum-of-magnitudes bound does not compound from block to block.
 * Replaces conv -> z, vocr_bn_relu_apply_f32(z) of the eval path: no fp32 z round trip through HBM. */
int vocr_bn_eval_bound_f32(const float* w, const float* bias, const float* scale, const float* shift,
                           const float* xbound, int C, int K, float* aux, vocr_stream_t stream);
int vocr_tc_conv3x3_bnrelu_f16(const uint16_t* x_hi, const uint16_t* x_lo, const int32_t* exp_x, const uint16_t* w_hi,
                               const uint16_t* w_lo, const int32_t* exp_w, const float* bias, const float* scale,
                               const float* shift, float* a, long long sB, long long sH, long long sW, uint16_t* a_hi,
                               uint16_t* a_lo, const float* bound, int32_t* pair_exp, float* amax, int B, int H, int W,
                               int Cin, int Cout, int products, vocr_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * 3x3 convolutions on the tensor cores (4-D TMA implicit GEMM + tcgen05 3xTF32), same math as vocr_conv3x3_fwd_f32 /
 * vocr_conv3x3_wgrad_f32 (src/models/cnnlstm.py:124-134 -> cuDNN).  Activations arrive as (hi, lo) TF32 planes
 * (vocr_split_tf32_f32), NHWC.
 *   fwd:   Cin % 32 == 0, Cout % 4 == 0; weight planes K-major [Cout][9*Cin], k = (ky*3+kx)*Cin + ci.  The data
 *          gradient is the same call on dz with the flipped / transposed weight matrix [Cin][9*Cout].
 *   wgrad: Cin % 32 == 0, Cout % 32 == 0; dw in state_dict layout [Cout,Cin,3,3]; chunked TMEM accumulation keeps
 *          the 10^5..10^6-term pixel reduction at fp32 accuracy.
 * vocr_colstats_f32: BatchNorm statistics, stats[0:C] += sum_p z, stats[C:2C] += sum_p z^2 (float64).
 * ---------------------------------------------------------------------------------------------------------- */
int vocr_tc_conv3x3_fwd(const float* x_hi, const float* x_lo, const float* w_hi, const float* w_lo, const float* bias,
                        float* z, int B, int H, int W, int Cin, int Cout, vocr_stream_t stream);
size_t vocr_tc_conv3x3_wgrad_workspace_size(int B, int H, int W, int Cin, int Cout);
int vocr_tc_conv3x3_wgrad(const float* x_hi, const float* x_lo, const float* dz_hi, const float* dz_lo, float* dw,
                          int B, int H, int W, int Cin, int Cout, void* workspace, size_t workspace_bytes,
                          vocr_stream_t stream);
int vocr_colstats_f32(const float* z, long long P, int C, double* stats, vocr_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Device-side batch assembly.  Replaces the pixel / label copying of SortByWidthCollater
 * (src/datautils.py:61-176): B ragged images [C,H,w_i], uploaded back to back (`packed`, element offsets
 * `img_offsets`, tensor widths `img_widths`), land in the zero-padded batch out[B,C,H,Wout] in the order given by
 * `order` (sorted position -> original index; the stable descending sort of the B width keys is host work); the label
 * sequences (packed_labels, label_offsets[B+1]) are concatenated in the same order (labels_out may be NULL).  B <= 1024.
 * ---------------------------------------------------------------------------------------------------------- */
int vocr_collate_lines_f32(const float* packed, const long long* img_offsets, const int32_t* img_widths,
                           const int32_t* order, int B, int C, int H, int Wout, float* out,
                           const int32_t* packed_labels, const int32_t* label_offsets, int32_t* labels_out,
                           int32_t* label_lens_out, vocr_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Line-image pre-processing.  Replaces, for a batch of raw decoded uint8 line images, the chain
 * [ConvertGray] -> Scale(new_h=H) -> [InvertBlackWhite] -> ToTensor (src/imagetransforms.py:411-416,453-507,383-385,
 * 423-434; assembled in src/decode_testset.py:48-65 and src/train_cnn_lstm.py:263-279), the 15-px width floor padded
 * with ones (src/ocr_dataset.py:174-180) and the zero-padded, ordered batch of SortByWidthCollater
 * (src/datautils.py:61-176).  Scale is bit-exact with what the reference's cv2.resize call computes: OpenCV's 8-bit
 * INTER_LINEAR (the reference's INTER_CUBIC argument lands on cv2.resize's `dst` parameter), incl. the INTER_AREA
 * re-route of exact 2x down-scaling.
 *   packed: images back to back, uint8, gray (channels = 1) or BGR interleaved (channels = 3, converted to gray);
 *   img_offsets [B] byte offsets;  src_h, src_w [B];  dst_w [B] = int(w * float(H / h)) from the caller (float64);
 *   order [B] batch position -> image index, or NULL;  out [B,1,H,Wout] fp32: resized / inverted / 255 in columns
 *   [0, dst_w), ones in [dst_w, max(dst_w, min_width)), zeros beyond.
 * ---------------------------------------------------------------------------------------------------------- */
int vocr_scale_lines_u8(const uint8_t* packed, const long long* img_offsets, const int32_t* src_h,
                        const int32_t* src_w, const int32_t* dst_w, const int32_t* order, int B, int channels, int H,
                        int Wout, int invert, int min_width, float* out, vocr_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * LM-decode front end.  Replaces the host part of LmDecoder.decode (src/decoder.py:61-101):
 * log_softmax over the alphabet, remap of model symbols to LM units (inv[u] = model index of unit u or -1, missing
 * units get `fill` = log(1e-10)), sliced per line to its valid frames, float64.  out [sum_b lens[b], U],
 * row_offsets[b] = first row of line b.  The WFST decoder itself (external EESEN) is out of scope.
 * ---------------------------------------------------------------------------------------------------------- */
int vocr_lm_frontend_f32(const float* logits, int T, int B, int A, const int32_t* lens, const long long* row_offsets,
                         const int32_t* inv, int U, double fill, double* out, vocr_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Batched edit distance.  Replaces the NumPy double loop of edit_distance
 * (src/textutils.py:264-287) behind compute_cer_wer (:326-351): P pairs of int32 sequences, unit costs, exact.
 *   a_flat/a_off[P+1], b_flat/b_off[P+1]: concatenated sequences + offsets;  max_n/max_m: length bounds;  dist [P].
 * ---------------------------------------------------------------------------------------------------------- */
int vocr_edit_distance_i32(const int32_t* a_flat, const int32_t* a_off, const int32_t* b_flat, const int32_t* b_off,
                           int P, int max_n, int max_m, int32_t* dist, vocr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VISTAOCR_B200_H_ */
