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

#ifdef __cplusplus
}
#endif
#endif /* VISTAOCR_B200_H_ */
