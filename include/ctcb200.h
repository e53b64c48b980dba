/* ctcb200 -- B200-native (sm_100a) CTC loss + gradient w.r.t. logits.  C ABI.
 *
 * Drop-in boundary for the CTC branch of the joint CTC/attention objective of
 * zqs01/ASR_chinese_e2e.  The reference is pure Python and has no FFI and no CTC
 * call site; each entry point below cites the reference interface
 * whose tensors it consumes / the torch op it replaces:
 *
 *   logits   f32 [B,T,V] batch-major, contiguous   <- Predictor/Utils/loss.py:10 ("pred: N x T x C"),
 *                                                     encoder tap Predictor/Models/transformer_official.py:76
 *   targets  i64 [B,Umax] padded with 0 (or 1-D)   <- data/data_loader/ai_shell_1.py:75-88 (tgt_for_input),
 *                                                     Predictor/data_handler/padder.py:6-27
 *   in_len / tgt_len  i64 [B]                      <- ai_shell_1.py:80-84 (wave_len, tgt_len)
 *   blank = 0 = PAD                                <- Predictor/data_handler/vocab.py:10,17
 *   loss consumer                                  <- Predictor/Models/transformer_official.py:83-104
 *                                                     (cal_metrics / iterate: loss.backward()),
 *                                                     Trainer/trainer11.py:57,73-74
 *
 * Semantics (the op BASELINE.json names):
 *   F.ctc_loss(F.log_softmax(logits,-1).transpose(0,1), targets, in_len, tgt_len,
 *              blank, reduction, zero_infinity)  and its gradient w.r.t. logits.
 *
 * Rules of the boundary
 *   - plain C types only; every pointer except host out-params is a DEVICE pointer;
 *   - the caller owns every buffer including the workspace; the library allocates
 *     nothing on the device and retains no pointer after return;
 *   - all work is enqueued asynchronously on `stream`; no internal synchronisation;
 *   - return 0 on success, a negative ctcb200_error on bad arguments, a positive
 *     cudaError_t if a launch failed.  No exceptions, no exit(), no CPU fallback;
 *   - data-dependent conditions are not errors: an infeasible utterance yields
 *     nll=+inf (0 with zero_infinity) exactly like torch.  Invalid lengths / labels
 *     are clamped for memory safety and reported in a device status word
 *     (ctcb200_read_status, debug use: it synchronises the stream);
 *   - deterministic: the only floating-point atomics are the sparse occupancy corrections of the fused
 *     gradient, one per (frame, class) at most, so repeated calls are bit-identical.  Occupancies below
 *     2^-40 (env CTCB200_OCC_SKIP_BITS) are not applied: < 1e-12 of the utterance's gradient scale.
 *
 * Limits: V >= 2, 1 <= T, 0 <= Umax <= 255, V*4 bytes must fit a shared-memory
 * stage (V <= 16384); logits / grad_logits base addresses 16-byte aligned.
 */
#ifndef CTCB200_H_
#define CTCB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTCB200_VERSION 100 /* 0.1.0 */

typedef void *ctcb200_stream_t; /* cudaStream_t */
typedef void *ctcb200_event_t;  /* cudaEvent_t  */

enum ctcb200_error {
    CTCB200_OK = 0,
    CTCB200_ERR_NULL = -1,      /* a required pointer is NULL */
    CTCB200_ERR_SHAPE = -2,     /* B<0, T<1, V<2 or V too large for a shared-memory stage */
    CTCB200_ERR_BLANK = -3,     /* blank not in [0,V) */
    CTCB200_ERR_UMAX = -4,      /* Umax<0 or Umax>255 */
    CTCB200_ERR_ALIGN = -5,     /* logits/grad base not 16-byte aligned, workspace not 256-byte aligned */
    CTCB200_ERR_REDUCTION = -6, /* unknown reduction code */
    CTCB200_ERR_WORKSPACE = -7, /* workspace_bytes smaller than ctcb200_workspace_bytes() */
    CTCB200_ERR_NO_DEVICE = -8, /* no CUDA device / not an sm_100 device */
    CTCB200_ERR_OPTION = -9     /* ctcb200_set_option / _get_option: unknown name */
};

enum ctcb200_reduction { CTCB200_REDUCE_NONE = 0, CTCB200_REDUCE_MEAN = 1, CTCB200_REDUCE_SUM = 2 };

/* The `zero_infinity` argument of the forward-type calls is a small bit field: bit 0 = zero_infinity,
 * bit 1 = also record the per-frame argmax class for ctcb200_greedy_decode (costs ~5 % of the sweep),
 * bit 2 = run the log-space alpha/beta recursion for every utterance (the algorithm BASELINE.json's north_star
 *         names; by default it is the fallback of the faster linear-domain recursion). */
enum ctcb200_forward_flags { CTCB200_FLAG_ZERO_INFINITY = 1, CTCB200_FLAG_DECODE = 2, CTCB200_FLAG_LATTICE_LOG = 4 };

/* bits of the device status word */
enum ctcb200_status_bits {
    CTCB200_ST_BAD_INPUT_LENGTH = 1,  /* in_len[b] outside [0,T]: clamped */
    CTCB200_ST_BAD_TARGET_LENGTH = 2, /* tgt_len[b] outside [0,Umax]: clamped */
    CTCB200_ST_BAD_LABEL = 4          /* label outside [0,V) or equal to blank */
};

int ctcb200_version(void);
const char *ctcb200_strerror(int code);

/* Developer tunables (launch shapes, experiment switches; DESIGN.md section 7 lists the names).  Each is read from
 * its CTCB200_* environment variable once, when the library is loaded; afterwards these two calls are the only way
 * to change or read one.  Process-wide, not synchronised: set them before issuing work from several threads. */
int ctcb200_set_option(const char *name, int value);
int ctcb200_get_option(const char *name, int *value);

/* Host-only arithmetic (no CUDA call): bytes of workspace the calls below need. */
int ctcb200_workspace_bytes(int B, int T, int V, int Umax, size_t *out_bytes);

/* Forward for training (replaces F.log_softmax + aten::_ctc_loss):
 *   prep + fused log-softmax/label-gather sweep + alpha/beta lattice recursion.
 * Writes nll[B] (per-utterance negative log-likelihood, zero_infinity applied) and, if
 * loss_sums != NULL, loss_sums[4] = { sum_b nll_b / max(U_b,1), sum_b nll_b, B, 'mean' loss }
 * ([0],[2]: the normaliser pair a data-parallel all-reduce combines; [3] = [0]/B, or
 * [0]*inv_batch in ctcb200_loss_grad).
 * Leaves the state occupancies in `workspace` for ctcb200_backward.
 * targets_stride: row stride of the [B,Umax] target matrix, or 0 for 1-D concatenated
 * targets (then targets_numel bounds the reads).
 * sweep_done: optional caller-owned event, recorded on `stream` right after the HBM-bound sweep
 * (before the latency-bound lattice kernel) so that a caller pipelining several utterance chunks
 * over two streams can start the next chunk's sweep under this chunk's lattice; NULL = none. */
int ctcb200_forward(const float *logits, const int64_t *targets, int64_t targets_stride,
                    int64_t targets_numel, const int64_t *in_len, const int64_t *tgt_len,
                    int B, int T, int V, int Umax, int blank, int zero_infinity,
                    float *nll, float *loss_sums, void *workspace, size_t workspace_bytes,
                    ctcb200_stream_t stream, ctcb200_event_t sweep_done);

/* Evaluation fast path (forward only, Trainer11.evaluate, trainer11.py:114-129): one sweep of
 * the logits; no occupancies are kept. Same outputs as ctcb200_forward. */
int ctcb200_loss_only(const float *logits, const int64_t *targets, int64_t targets_stride,
                      int64_t targets_numel, const int64_t *in_len, const int64_t *tgt_len,
                      int B, int T, int V, int Umax, int blank, int zero_infinity,
                      float *nll, float *loss_sums, void *workspace, size_t workspace_bytes,
                      ctcb200_stream_t stream, ctcb200_event_t sweep_done);

/* Backward (replaces aten::_ctc_loss_backward + _log_softmax_backward_data):
 *   grad_logits[b,t,v] = g_b * (softmax(logits)[b,t,v] - occupancy[b,t,v]),  0 for t >= in_len[b].
 * g_b = grad_out[b*grad_out_stride] * (reduction==MEAN ? inv_batch / max(U_b,1) : 1);
 * grad_out_stride is 0 for a scalar upstream gradient ('mean'/'sum') and 1 for 'none'.
 * inv_batch = 1 / (global batch size): pass 1/(world_size*B) when the batch is sharded.
 * `workspace` must be the one ctcb200_forward filled for the same inputs. */
int ctcb200_backward(const float *logits, const int64_t *targets, int64_t targets_stride,
                     int64_t targets_numel, const float *grad_out, int64_t grad_out_stride,
                     int reduction, float inv_batch, int B, int T, int V, int Umax, int blank,
                     int zero_infinity, float *grad_logits, const void *workspace,
                     size_t workspace_bytes, ctcb200_stream_t stream);

/* Loss AND gradient in TWO sweeps of [B,T,V] instead of three (the fast path of the autograd op).
 * The sweep that computes the log-softmax statistics also writes the dense part of the gradient,
 * g_b * softmax(logits), from the registers that hold the row; after the lattice a sparse kernel adds
 * -g_b * occupancy to the <= U_b+1 label columns of each frame and zero/NaN-fills utterances without
 * a valid alignment.  g_b is computed for an upstream gradient of 1:
 * g_b = (reduction==MEAN ? inv_batch / max(U_b,1) : 1); use ctcb200_rescale_grad afterwards if the
 * real upstream gradient differs.  Same outputs as ctcb200_forward + ctcb200_backward(grad_out=1). */
int ctcb200_loss_grad(const float *logits, const int64_t *targets, int64_t targets_stride,
                      int64_t targets_numel, const int64_t *in_len, const int64_t *tgt_len,
                      int B, int T, int V, int Umax, int blank, int zero_infinity, int reduction,
                      float inv_batch, float *nll, float *loss_sums, float *grad_logits,
                      void *workspace, size_t workspace_bytes, ctcb200_stream_t stream,
                      ctcb200_event_t sweep_done);

/* The same call, one stage at a time, for callers that pipeline utterance chunks over two streams
 * (the lattice is latency-bound and should run under another chunk's sweep):
 * stages is a bit mask of  1 = prep + fused sweep,  2 = lattice,  4 = sparse patch; each stage must
 * be ordered after the previous stage of the same workspace by the caller (events).  For profiling, 8 = prep alone
 * and 16 = fused sweep alone split stage 1 in two (bench.py times the dominant kernel by itself this way). */
int ctcb200_loss_grad_stages(int stages, const float *logits, const int64_t *targets,
                             int64_t targets_stride, int64_t targets_numel, const int64_t *in_len,
                             const int64_t *tgt_len, int B, int T, int V, int Umax, int blank,
                             int zero_infinity, int reduction, float inv_batch, float *nll,
                             float *loss_sums, float *grad_logits, void *workspace,
                             size_t workspace_bytes, ctcb200_stream_t stream);

/* Speculative-gradient support.  A caller that wants loss AND gradient from one pass may run
 * ctcb200_backward right after ctcb200_forward with grad_out == 1 (before autograd has produced the
 * real upstream gradient) and fix the result up later: this multiplies utterance b's gradient slab
 * in place by grad_out[b*stride] / applied_in[b] and records grad_out in applied_out[b].  When the
 * two are equal (loss.backward() with an upstream gradient of 1) the kernel exits without touching
 * memory, so the usual cost is one empty launch instead of a third sweep.  applied_in[b] == 0 with a different
 * grad_out cannot be rescaled (the slab holds zeros): the slab becomes NaN -- recompute it with ctcb200_backward. */
int ctcb200_rescale_grad(float *grad_logits, const float *grad_out, int64_t grad_out_stride,
                         const float *applied_in, float *applied_out, int B, int T, int V,
                         ctcb200_stream_t stream);

/* On-device greedy (best-path) CTC decode + token-level edit distance: replaces the
 * per-step D->H sync and Python Levenshtein loop of cal_metrics (transformer_official.py:87-91,
 * Predictor/Utils/score.py:4-13) for a CTC-branch character error rate.  Uses the per-frame argmax the
 * sweep of ctcb200_forward / _loss_only / _loss_grad left in `workspace` (call them with
 * zero_infinity | CTCB200_FLAG_DECODE).  edit_out[B]: Levenshtein
 * distance between the collapsed best path and the labels; hyp_len_out[B]; hyp_out (optional)
 * int64 [B,T], padded with blank. */
int ctcb200_greedy_decode(const int64_t *targets, int64_t targets_stride, int64_t targets_numel,
                          int B, int T, int V, int Umax, int blank, const void *workspace,
                          size_t workspace_bytes, int *edit_out, int *hyp_len_out, int64_t *hyp_out,
                          ctcb200_stream_t stream);

/* The reference's `cer` metric on the device (cal_metrics, transformer_official.py:87-91 ->
 * Vocab.convert_id2str, vocab.py:74-78 -> calculate_cer, Predictor/Utils/score.py:4-13): per row b the edit distance
 * between hyp[b, 0..L) and gold[b, 0..L) (int64 id matrices with the given row strides), ids == pad dropped.
 *   mode 0: token-level Levenshtein distance;
 *   mode 1: distance between the space-joined strings (what python-Levenshtein returns for the reference's
 *           single-character tokens): insertions / deletions of a token also cost its separator.
 * edit_out[B]; words_out[B] = max(#gold tokens, 1) = len(gold_string.split(' ')), the reference's denominator.
 * L <= 512 (mode 1) / 1024 (mode 0). */
int ctcb200_edit_distance(const int64_t *hyp, int64_t hyp_stride, const int64_t *gold, int64_t gold_stride, int B,
                          int L, int pad, int mode, int *edit_out, int *words_out, ctcb200_stream_t stream);

/* Attention-branch loss on the same sweep machinery: cross-entropy with optional label
 * smoothing over pred[rows, V] logits, rows whose gold == ignore_index skipped -- the reference's
 * cal_loss / calculate_loss (Predictor/Utils/loss.py:26-76; called with pred [N,T,C] flattened, PAD=0):
 *   smoothing == 0: F.cross_entropy(pred, gold, ignore_index, 'mean')
 *   smoothing  > 0: target = (1-eps) on the label and eps/V elsewhere; sum_rows(-sum_c target*logp) / n_word
 * loss_out[2] = { weight * loss, n_word }.  If grad != NULL the same sweep writes
 * d(weight*loss)/d pred (zeros for ignored rows); n_word is counted on the device (no host sync). */
int ctcb200_ce_workspace_bytes(int64_t rows, size_t *out_bytes);
int ctcb200_ce_loss_grad(const float *pred, const int64_t *gold, int64_t rows, int V, int ignore_index,
                         float smoothing, float weight, float *loss_out, float *grad, void *workspace,
                         size_t workspace_bytes, ctcb200_stream_t stream);

/* ---- f1: the CTC head fused with the loss -------------------------------------------------
 * logits = enc[B*T, K] x weight[V, K]^T + bias[V]  (the nn.Linear(d_model, vocab) the joint model attaches to the
 * encoder output, Predictor/Models/transformer_official.py:76) is computed tile by tile on the tensor cores
 * (tcgen05.mma kind::tf32, fp32 accumulators in TMEM, operands staged by TMA) and consumed in the epilogue: the
 * [B,T,V] logits tensor is never written to or read from HBM.
 *   ctcb200_head_loss       forward / evaluation: per-row log-sum-exp + label gather in the GEMM epilogue, then the
 *                           alpha/beta lattice.  Same outputs as ctcb200_loss_only on F.linear(enc, weight, bias).
 *   ctcb200_head_loss_grad  training: the same forward, the lattice with occupancies, then a second GEMM pass that
 *                           recomputes each logits tile and writes d loss / d logits = g_b * (softmax - occupancy)
 *                           (zeros for padded frames) to dlogits[B*T, dlogits_pitch] (pitch: multiple of 4, V <=
 *                           pitch < V + 32, so rows stay 16-byte aligned; columns >= V are written as zeros).  The
 *                           caller forms d weight = dlogits^T x enc, d enc = dlogits x weight, d bias = column sums
 *                           with plain library GEMMs.  g_b as in ctcb200_loss_grad (upstream gradient 1).
 * precision: CTCB200_HEAD_3XTF32 -- operands split as hi + lo (tf32 each), hi*hi + hi*lo + lo*hi accumulated in fp32:
 *            fp32-GEMM grade (meets the 1e-5 relative bar on the per-utterance loss); CTCB200_HEAD_TF32 -- one pass,
 *            10-bit mantissa operands (like torch.backends.cuda.matmul.allow_tf32), 3x fewer tensor-core passes.
 * K (d_model) must be a multiple of 32; enc and weight 16-byte aligned, row-major, contiguous.  bias may be NULL.
 * The workspace (ctcb200_head_workspace_bytes) holds the CTC workspace plus, for 3XTF32, the split operands. */
enum ctcb200_head_precision { CTCB200_HEAD_3XTF32 = 0, CTCB200_HEAD_TF32 = 1 };

int ctcb200_head_workspace_bytes(int B, int T, int V, int K, int Umax, int precision, size_t *out_bytes);

int ctcb200_head_loss(const float *enc, const float *weight, const float *bias, const int64_t *targets,
                      int64_t targets_stride, int64_t targets_numel, const int64_t *in_len, const int64_t *tgt_len, int B,
                      int T, int V, int K, int Umax, int blank, int flags, int precision, float *nll, float *loss_sums,
                      void *workspace, size_t workspace_bytes, ctcb200_stream_t stream);

int ctcb200_head_loss_grad(const float *enc, const float *weight, const float *bias, const int64_t *targets,
                           int64_t targets_stride, int64_t targets_numel, const int64_t *in_len, const int64_t *tgt_len,
                           int B, int T, int V, int K, int Umax, int blank, int flags, int precision, int reduction,
                           float inv_batch, float *nll, float *loss_sums, float *dlogits, int64_t dlogits_pitch,
                           void *workspace, size_t workspace_bytes, ctcb200_stream_t stream);

/* The two parameter-gradient GEMMs of the fused head, on the tensor cores with fp32-grade accuracy (3xTF32: every fp32
 * operand tile is split into hi + lo tf32 halves inside the shared-memory ring, three tcgen05.mma passes accumulate in
 * fp32) -- what torch.autograd would run as two fp32 SIMT GEMMs for the nn.Linear it replaces:
 *   d_enc[B*T, K]   = dlogits[B*T, V] x weight[V, K]          (NULL to skip)
 *   d_weight[V, K]  = dlogits[B*T, V]^T x enc[B*T, K]         (NULL to skip; needs the workspace: two partial sums)
 * dlogits is the buffer ctcb200_head_loss_grad wrote (row pitch dlogits_pitch floats, columns >= V zero).
 * d_bias is the column sum of dlogits (left to the caller).  K % 32 == 0; all pointers 16-byte aligned. */
int ctcb200_head_param_grads_workspace_bytes(int V, int K, size_t *out_bytes);
int ctcb200_head_param_grads(const float *dlogits, int64_t dlogits_pitch, const float *enc, const float *weight, int B,
                             int T, int V, int K, float *d_enc, float *d_weight, void *workspace, size_t workspace_bytes,
                             ctcb200_stream_t stream);

/* Debug: copies the device status word to *host_status (synchronises `stream`). */
int ctcb200_read_status(const void *workspace, int *host_status, ctcb200_stream_t stream);

/* Debug: which recursion the lattice kernel of the last call used (synchronises `stream`).
 * host_stats[0] = utterances that ran the log-space recursion (the rest ran the faster linear-domain one),
 * host_stats[1] = of those, utterances that were first tried in the linear domain and whose likelihood
 * underflowed to zero there (this includes every infeasible utterance). */
int ctcb200_read_lattice_stats(const void *workspace, int *host_stats, ctcb200_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CTCB200_H_ */
