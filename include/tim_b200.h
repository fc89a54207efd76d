/* libtim_b200 — C ABI of the B200-native TIM encoder: forward, and the training leg (backward + gradient all-reduce).
 *
 * The reference (JacobChalk/TIM) has no FFI: its seam for this path is the Python nn.Module
 *     TIM.forward(inputs, forward_type, ...)      recognition/time_interval_machine/models/tim.py:174-191
 *                                                 detection/time_interval_machine/models/tim.py:415-430
 * plus the state_dict key layout (utils/checkpoint.py:17-36). Each entry point below names the reference
 * code it replaces; tim_b200/plugin.py binds them with ctypes and INTEGRATION.md shows the reference-side stub.
 *
 * Conventions
 *  - plain pointers and sizes only; all device buffers passed in are owned by the caller (PyTorch) and only
 *    borrowed for the call. The context owns its packed weights and its workspace.
 *  - every function returns TIM_OK (0) or a negative tim_status; tim_last_error() gives the message. The library
 *    never calls exit()/abort() and has NO CPU fallback: without a B200 (sm_100) device tim_create() fails.
 *  - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream). Device entry points
 *    only enqueue work on it and never synchronise; one host thread per context.
 */
#ifndef TIM_B200_H
#define TIM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TIM_ABI_VERSION 2

typedef enum {
    TIM_OK = 0,
    TIM_ERR_INVALID = -1,      /* bad argument / unsupported configuration */
    TIM_ERR_CUDA = -2,         /* a CUDA runtime / driver call failed */
    TIM_ERR_NO_DEVICE = -3,    /* no sm_100 device: the library refuses to run (no fallback) */
    TIM_ERR_WEIGHTS = -4,      /* unknown key, wrong shape, or weights missing at forward time */
    TIM_ERR_NOMEM = -5
} tim_status;

enum { TIM_FP32 = 0, TIM_BF16 = 1, TIM_FP16 = 2 };                      /* compute_dtype */
enum { TIM_AUDIO_VISUAL = 0, TIM_VISUAL = 1, TIM_AUDIO = 2 };           /* modalities */
enum { TIM_RECOGNITION = 0, TIM_DETECTION = 1 };                        /* variant */

/* Constructor arguments of the reference TIM (tim.py:18-34) flattened to ints.
 * n_* = 0 means "head absent" (follows head.py's isinstance(num_class...) logic, see tim_b200/config.py). */
typedef struct {
    int32_t variant;            /* TIM_RECOGNITION | TIM_DETECTION (detection adds reg heads, shares query rows between heads) */
    int32_t d_model;            /* transformer width is 2*d_model (tim.py:115-121) */
    int32_t nhead;
    int32_t num_layers;
    int32_t ff_dim;             /* d_model * feedforward_scale */
    int32_t vis_dim, aud_dim;   /* input feature widths */
    int32_t num_feats;          /* feature tokens per modality */
    int32_t input_modality;     /* which embedders exist */
    int32_t data_modality;      /* which CLS queries / heads exist */
    int32_t include_verb_noun;  /* recognition: separate verb/noun CLS token groups (encodings.py:166-171) */
    int32_t n_verb, n_noun, n_action, n_audio;
    int32_t compute_dtype;      /* TIM_FP32: CUDA-core fp32 everywhere (parity mode, <=1e-5 rel);
                                   TIM_BF16 / TIM_FP16: 16-bit tcgen05 operands, fp32 accumulate / softmax / LayerNorm / residual */
} tim_config;

/* Caller-allocated outputs of the encoder; NULL where the reference returns None. fp32, contiguous.
 * Shapes follow head.py's flatten(0,1): [B*Qv, n_*] / [B*Qa, n_audio]; reg [B*Q, 2]; feats [B, F_tot, 2*d_model]. */
typedef struct {
    float* verb;
    float* noun;
    float* action;
    float* audio;
    float* reg_visual;
    float* reg_audio;
    float* feats;               /* x[:, :num_feats] of tim.py:172 (may be NULL to skip) */
} tim_outputs;

typedef struct tim_ctx tim_ctx;

int tim_abi_version(void);

/* One context per (process, device); replaces TIM.__init__/_create_model (tim.py:17-145) for the hot path. */
int tim_create(tim_ctx** out, const tim_config* cfg, int device);
void tim_destroy(tim_ctx* ctx);
const char* tim_last_error(const tim_ctx* ctx);        /* ctx may be NULL: error of the last failed tim_create() */

/* Load one parameter by its reference state_dict key (checkpoint contract, utils/checkpoint.py:17-36).
 * `data` is a device pointer to contiguous fp32 with the reference's shape; the library packs its own copy
 * (16-bit operand copy for GEMM weights, 1/sqrt(head_dim)*log2(e) folded into the q rows of in_proj).
 * Keys off the hot path (drloc_mlp.*, pool.*) are accepted and ignored. Call again after optimizer.step(). */
int tim_set_weight(tim_ctx* ctx, const char* key, const float* data, const int64_t* shape, int ndim, void* stream);
int tim_weights_missing(const tim_ctx* ctx, char* buf, size_t buflen);   /* number of required keys not yet set; names into buf */

/* forward_type == "time_mlp" (tim.py:181-182): times [B, T, 2] fp32 -> out [B, T, d_model] fp32 (device pointers). */
int tim_time_mlp_fwd(tim_ctx* ctx, const float* times, float* out, int B, int T, void* stream);

/* forward_type == "encoder" (tim.py:147-172; detection forward_inference :378-400 after its own time_mlp):
 * vis [B, F, vis_dim], aud [B, F, aud_dim] (NULL if the modality is absent), time_enc [B, T, d_model], all fp32 device.
 * T = F_tot + Qv + Qa (time rows: vis feats, aud feats, visual queries, audio queries). */
int tim_encoder_fwd(tim_ctx* ctx, const float* vis, const float* aud, const float* time_enc, int B, int T, int Qv, int Qa,
                    const tim_outputs* outs, void* stream);

/* The same forward with the input windows gathered ON THE DEVICE from feature banks resident in HBM (SURVEY.md section 8f row 4;
 * the reference's loader gathers feats[video][feat_indices, aug_indices] on the host, datasets/sliding_window.py:356-375, and
 * ships [B, F, D] fp32 over PCIe every step). vis_rows / aud_rows [B * num_feats] (device, int64) give the bank row of every
 * feature token, clip-major. Banks may be fp32 or 16-bit. A row index outside its bank (the reference's host-side indexing raises
 * IndexError) reads as zeros and is COUNTED: tim_index_check() blocks for the verdict on the last indexed forward (TIM_ERR_INVALID and a
 * message with the count), and any later indexed forward on the context fails if an earlier one saw bad indices. */
typedef struct {
    const void* vis_bank;           /* [vis_bank_rows, vis_dim], device; NULL if the modality is absent */
    const void* aud_bank;           /* [aud_bank_rows, aud_dim] */
    const int64_t* vis_rows;
    const int64_t* aud_rows;
    int64_t vis_bank_rows, aud_bank_rows;
    int32_t bank_dtype;             /* TIM_FP32 | TIM_BF16 | TIM_FP16 */
} tim_feature_bank;
int tim_encoder_fwd_indexed(tim_ctx* ctx, const tim_feature_bank* bank, const float* time_enc, int B, int T, int Qv, int Qa,
                            const tim_outputs* outs, void* stream);
int tim_index_check(tim_ctx* ctx);

/* End-to-end call on HOST buffers (pinned or pageable): time_mlp + encoder with the H2D input copies and D2H
 * result copies inside the call, chunked over clips and overlapped on three internal streams. Blocks until the
 * outputs are in host memory. `times` is [B, T, 2]; outputs as in tim_outputs but host pointers.
 * clips_per_chunk is the TARGET chunk (<= 0 or > B: the whole batch): no chunk is larger; the first and last chunks are tapered
 * (a quarter, a half of it) so that the un-overlapped first H2D / last D2H are short, and on the 16-bit path chunk sizes are
 * rounded down to whole waves of GEMM row tiles (num_sms / gcd(num_sms, E / 256) tiles of 256 token rows).
 * h2d_bytes / d2h_bytes (optional) receive the bytes moved. */
int tim_forward_host(tim_ctx* ctx, const float* vis, const float* aud, const float* times, int B, int T, int Qv, int Qa,
                     const tim_outputs* host_outs, int clips_per_chunk, uint64_t* h2d_bytes, uint64_t* d2h_bytes);
/* The same call with the host-side formats and the ordering made explicit:
 *  in_dtype   TIM_FP32, or the context's 16-bit compute dtype: vis / aud then point to a 16-bit HOST feature bank slice. The
 *             16-bit paths round the features to the operand type before the embedder GEMM anyway, so results are bit-identical
 *             to the fp32 call while half the bytes cross PCIe and the cast pass disappears.
 *  out_dtype  TIM_FP32, or TIM_FP16: every output buffer of host_outs is then __half (logits rounded once, on the device).
 *  stream     the caller's stream (torch.cuda.current_stream): work enqueued there before the call (weight packing, an earlier
 *             forward) is waited for through an event; nothing else on the device is synchronised.
 * The precision guard of the folded-LayerNorm flow acts inside the call: if a chunk flags rows outside the folded path's error
 * bound, the context switches to the un-folded flow and the whole call is recomputed before it returns. */
int tim_forward_host_ex(tim_ctx* ctx, const void* vis, const void* aud, const float* times, int B, int T, int Qv, int Qa,
                        const tim_outputs* host_outs, int clips_per_chunk, int in_dtype, int out_dtype, void* stream,
                        uint64_t* h2d_bytes, uint64_t* d2h_bytes);

/* The chunk sizes tim_forward_host uses for a batch of B clips (pure host arithmetic, no device needed): writes up to max_out
 * sizes to out and returns the number of chunks (negative tim_status on bad arguments). rows_per_clip = token rows per clip
 * (tim_seq_len), E = 2 * d_model, sixteen_bit = compute dtype is not TIM_FP32. */
int tim_host_chunk_schedule(int B, int clips_per_chunk, int rows_per_clip, int E, int num_sms, int sixteen_bit, int* out, int max_out);

/* Introspection used by bench.py / tests. */
size_t tim_workspace_bytes(const tim_ctx* ctx);         /* bytes currently held by the context's workspace */
uint64_t tim_launch_count(const tim_ctx* ctx);          /* kernels launched by this context so far */
int tim_seq_len(const tim_config* cfg, int Qv, int Qa); /* S = F_tot + query tokens (pure host arithmetic) */
/* 1 while the encoder LayerNorms are folded into the GEMMs around them (16-bit path), 0 when the
 * un-folded flow is in use: fp32 mode, shapes the CTA-pair GEMM does not cover, TIM_B200_FOLD=0, or after the precision
 * guard saw residual-stream rows with |mean| > 8 std in an earlier forward of this context. */
int tim_fold_active(const tim_ctx* ctx);
/* Device entry points never synchronise, so the guard's verdict on a forward is known only after it ran: tim_fold_check blocks until
 * the most recent tim_encoder_fwd / _indexed of this context has finished its check and returns 1 if that forward took the folded
 * flow AND flagged such rows - its outputs should be recomputed by calling the forward again (the context has switched to the
 * un-folded flow) - else 0. tim_forward_host(_ex) does this internally. */
int tim_fold_check(tim_ctx* ctx);

/* Detection query labelling (detection/time_interval_machine/models/tim.py:186-270 get_query_ious + label_queries): for every
 * query [B, Nq, 2] the ground-truth segment [B, Na, 2] of maximal IoU (first maximum; computed after the reference's shift by
 * |min(min_a start, 0)|, which is also applied to the returned segment), its labels gt_labels [B, Na, Nl] and the IoU. Queries
 * with IoU < iou_threshold get targets = +inf and label ids = -1. All pointers are device pointers; results are bit-identical
 * to the reference's fp32 arithmetic. Stateless: errors are reported through tim_last_error(NULL). */
int tim_label_queries(const float* queries, const float* gt_segs, const int64_t* gt_labels, int B, int Nq, int Na, int Nl,
                      float iou_threshold, float* targets /*[B*Nq,2]*/, int64_t* label_ids /*[B*Nq,Nl]*/, float* ious /*[B*Nq]*/,
                      void* stream);
/* assign_positive_labels (tim.py:158-185): out[row, c] = one_hot(id, C + 1)[c] * smoothing + (1 - smoothing) / (C + 1) for
 * c < C = num_classes, id = label_ids[row * stride + col] with -1 mapped to the dropped class C. out is [rows, C] fp32. */
int tim_smooth_labels(const int64_t* label_ids, int stride, int col, int64_t rows, int num_classes, double smoothing, float* out,
                      void* stream);

/* Detection post-processing: 1-D soft-NMS / NMS over G independent groups of proposals (a group = one class of one video), one
 * launch for all groups. Replaces the reference's scalar CPU extension and the per-class / per-video loops that drive it:
 *   tim_softnms_1d  <- nms_1d_cpu.softnms (detection/eval_detection/csrc/nms_cpu.cpp:67-160) under SoftNMSop (nms.py:35-61)
 *   tim_nms_1d      <- nms_1d_cpu.nms     (csrc/nms_cpu.cpp:19-58) under NMSop (nms.py:7-33: scores <= min_score are dropped
 *                      first when min_score > 0, at most max_num picks when max_num > 0)
 * segs [N,2], scores [N] fp32; group g owns rows group_offsets[g] .. group_offsets[g+1] (int64 [G+1], ascending). For group g the
 * picks are written in pick order (= descending final score) to rows group_offsets[g] .. + kept[g] of dets [N,3] = (start, end,
 * score after decay) and inds [N] (index of the pick INSIDE its group); rows beyond kept[g] are left untouched. method: 0 vanilla,
 * 1 linear, 2 gaussian exp(-iou^2 / sigma). Pick order and indices equal the reference's, including its order-dependent handling
 * of equal scores (tim_nms_1d breaks ties by input index, i.e. a stable descending sort; the reference's torch.sort leaves that
 * order unspecified). Scores must not be NaN. workspace: tim_nms_workspace_bytes(N) bytes of device memory. All pointers are
 * device pointers; stateless, errors through tim_last_error(NULL). */
size_t tim_nms_workspace_bytes(int64_t N);
int tim_softnms_1d(const float* segs, const float* scores, const int64_t* group_offsets, int G, int64_t N, float iou_threshold,
                   float sigma, float min_score, int method, float* dets /*[N,3]*/, int64_t* inds /*[N]*/, int* kept /*[G]*/,
                   void* workspace, size_t workspace_bytes, void* stream);
int tim_nms_1d(const float* segs, const float* scores, const int64_t* group_offsets, int G, int64_t N, float iou_threshold,
               float min_score, int64_t max_num, float* dets /*[N,3]*/, int64_t* inds /*[N]*/, int* kept /*[G]*/, void* workspace,
               size_t workspace_bytes, void* stream);

/* Detection post-processing in front of the NMS, on the device (all pointers are device pointers; stateless, errors through
 * tim_last_error(NULL)).
 * tim_det_decode <- FeatureMeter.update (detection/time_interval_machine/utils/meters.py:652-724): preds = sigmoid(logits) [R,C]
 *   and proposals[r, :] = double(fp32(clamp(reg[r, :], 0, max_time) * win_size)) + win_start[r / rows_per_window] (seconds of the
 *   video; float64 like the reference's tensor, whose window metadata collates to float64). Either output may be NULL.
 * tim_det_count / tim_det_emit <- the thresholding loop of detection/eval_detection/format_predictions.py:103-125: proposals are
 *   rounded to 3 decimals (numpy: multiply, rint, divide in double), rows with end - start <= 0 are dropped, every class whose
 *   score exceeds score_threshold gives one detection. counts[r] = detections of row r; with offsets = the exclusive prefix sum of
 *   counts (formed by the caller), tim_det_emit writes them in (row, class) order: source row, class, score, fp32 segment. */
int tim_det_decode(const float* logits, const float* reg, const double* win_start, int rows_per_window, int64_t R, int C, float win_size,
                   float max_time, float* preds /*[R,C] or NULL*/, double* proposals /*[R,2] or NULL*/, void* stream);
int tim_det_count(const float* preds, const double* proposals, int64_t R, int C, float score_threshold, int* counts /*[R]*/, void* stream);
int tim_det_emit(const float* preds, const double* proposals, int64_t R, int C, float score_threshold, const int64_t* offsets /*[R]*/,
                 int64_t* out_row, int64_t* out_cls, float* out_score, float* out_seg /*[n,2]*/, void* stream);

/* ---------------------------------------------------------------------------------------------------------------------
 * Training leg. Replaces what torch.autograd + DistributedDataParallel do for this path when the
 * reference trains: recognition/scripts/train.py:190-260 (forward under autocast), :354-366 (GradScaler.scale(loss).backward(),
 * optimizer step), recognition/time_interval_machine/models/build.py:58-63 (DDP bucketed gradient all-reduce),
 * detection/time_interval_machine/models/tim.py:272-337 (forward_train). Dropout: tim_set_dropout (all six nn.Dropout sites of the
 * reference, statistically equivalent masks from a counter-based hash instead of torch's Philox stream).
 *
 *  tim_train_enable      allocates the transposed weight copies the input-gradient GEMMs read; every weight must be set (again)
 *                        with tim_set_weight afterwards.
 *  tim_bind_grad         gradient destination (device fp32, 16-byte aligned, the parameter's shape) of one state_dict key. The
 *                        backward ACCUMULATES into it (+=), like autograd into param.grad; the caller zeroes it (zero_grad).
 *                        Binding all keys to slices of ONE flat buffer makes the data-parallel exchange a single collective.
 *  tim_time_mlp_fwd_train / tim_time_mlp_bwd        "time_mlp" forward that keeps its activations, and its backward
 *                        (d_out [B, T, d_model] = gradient of the time encodings, i.e. d_time_enc of tim_encoder_bwd plus
 *                        whatever else the loss sent there).
 *  tim_encoder_fwd_train / tim_encoder_bwd          "encoder" forward that keeps its activations (one outstanding forward per
 *                        context), and its backward: grad_outs holds the gradients of the outputs (NULL where the loss does not
 *                        touch an output), d_time_enc [B, T, d_model] receives the gradient w.r.t. the time encodings. Gradients
 *                        w.r.t. the input features are not produced (they are data).
 *  tim_comm_unique_id / tim_comm_init / tim_allreduce_grads     the one data-path collective: ncclAllReduce (average) over the
 *                        flat gradient buffer on `stream`, on a communicator of the library's own (NCCL is dlopen'ed: the copy
 *                        torch already mapped). Rank 0 creates the 128-byte id, the caller ships it to the other ranks.
 * 16-bit modes use 16-bit gradient operands with fp32 accumulation, exactly like the reference's autocast backward: with fp16
 * keep the reference's GradScaler (train.py:354-366) so that small gradients do not underflow. */
int tim_train_enable(tim_ctx* ctx);
int tim_bind_grad(tim_ctx* ctx, const char* key, float* dst);
/* Dropout of the NEXT tim_encoder_fwd_train (and of its backward): p_feat = the embedders' input dropout (helpers/encodings.py:141,149),
 * p_seq = dropout of the assembled token sequence (:177), p_enc = the transformer's four sites - attention probabilities, dropout1,
 * the FFN's inner dropout, dropout2 (helpers/transformers.py:73-82, 102-108). No mask is stored: keep decisions are a pure function
 * of (seed, site, layer, element) - one 32-bit hash per pair of elements, described next to DropSite in tim_b200/csrc/kernels.h -
 * and are re-evaluated by the backward. Pass a fresh seed every step. All zero (the default) = no dropout. 16-bit modes need
 * head_dim 64 or 128 for p_enc > 0 (the tcgen05 attention kernels carry the probability dropout). */
int tim_set_dropout(tim_ctx* ctx, float p_feat, float p_seq, float p_enc, uint64_t seed);
int tim_time_mlp_fwd_train(tim_ctx* ctx, const float* times, float* out, int B, int T, void* stream);
int tim_time_mlp_bwd(tim_ctx* ctx, const float* d_out, void* stream);
int tim_encoder_fwd_train(tim_ctx* ctx, const float* vis, const float* aud, const float* time_enc, int B, int T, int Qv, int Qa,
                          const tim_outputs* outs, void* stream);
int tim_encoder_bwd(tim_ctx* ctx, const tim_outputs* grad_outs, float* d_time_enc, void* stream);
size_t tim_train_tape_bytes(const tim_ctx* ctx);        /* bytes held by the saved activations */
int tim_comm_unique_id(void* out128);
int tim_comm_init(tim_ctx* ctx, const void* id128, int rank, int world);
int tim_allreduce_grads(tim_ctx* ctx, float* buf, size_t n, void* stream);

/* Live per-kernel-class timing (bench.py's roofline object): between begin and end every launch is bracketed by a
 * CUDA-event pair on its own stream. Classes: 0 GEMM (tcgen05 / fp32 SIMT) other than 5 and 6, 1 attention, 2 LayerNorm /
 * row statistics (forward and backward), 3 token assembly, 4 other row kernels, 5 encoder GEMMs with a folded LayerNorm in front
 * (in_proj, linear1: tensor-bound), 6 encoder GEMMs that also write the residual stream (out_proj, linear2: 12 - 14 KB of HBM
 * traffic per row), 7 input-gradient GEMMs, 8 weight-gradient GEMMs, 9 attention backward.
 * end() synchronises the device and fills ms / algorithmic FLOPs / launch counts. */
#define TIM_PROFILE_CLASSES 10
int tim_profile_begin(tim_ctx* ctx);
int tim_profile_end(tim_ctx* ctx, double* ms, double* flops, uint64_t* count, int n_classes);

/* Single-kernel test hooks (tests/ only): C[M,N] = act(A[M,K] W[N,K]^T + bias) (+resid), fp32 in / fp32 out,
 * computed through the selected compute path (the 16-bit paths cast A and W on device first). */
int tim_test_linear(int compute_dtype, const float* A, const float* W, const float* bias, const float* resid, float* out,
                    int M, int N, int K, int act, void* stream);
/* timing hook (tools/gemm_bench.py): `iters` back-to-back launches of one linear layer on 16-bit device operands
 * A16 [M,K], W16 [N,K]; out is T [M,N] (out_fp32 = 0) or float [M,N]; version 1 = single-CTA kernel, 2 = CTA-pair kernel. */
int tim_bench_linear(int compute_dtype, const void* A16, const void* W16, const float* bias, const float* resid, void* out,
                     int M, int N, int K, int act, int out_fp32, int version, int iters, float* ms_per_iter);
/* attention over a two-stream qkv buffer [(B*Ft + B*Qt), 3*H*hd] fp32 -> out [(B*Ft + B*Qt), H*hd] fp32.
 * The q columns must already carry the hd^-0.5 * log2(e) factor that tim_set_weight folds into in_proj. */
int tim_test_attention(int compute_dtype, const float* qkv, float* out, int B, int Ft, int Qt, int H, int hd, void* stream);

/* dW[N, K] (fp32, accumulated into) += dY[M, N]^T X[M, K] through the selected compute path; splits <= 0: chosen by the library */
int tim_test_wgrad(int compute_dtype, const float* dY, const float* X, float* dW, int M, int N, int K, int splits, void* stream);
int tim_bench_wgrad(int compute_dtype, const void* dY16, const void* X16, float* dW, int M, int N, int K, int splits, int iters,
                    float* ms_per_iter);
/* attention backward at the kernel boundary: qkv / dqkv [(B*Ft + B*Qt), 3*H*hd], dO [.., H*hd], fp32 device buffers; dq is scaled by
 * qscale (ln 2: gradient w.r.t. the stored, pre-scaled q; hd^-0.5: w.r.t. the un-scaled in_proj output) */
int tim_test_attention_bwd(int compute_dtype, const float* qkv, const float* dO, float* dqkv, int B, int Ft, int Qt, int H, int hd,
                           float qscale, void* stream);

/* debug hook (tools/attn_roles.py): a device buffer [>= 148][16] of 64-bit counters to which every launch of the decoupled attention
 * forward kernel adds, per CTA, the cycles one warp of each role spent waiting (see attention_umma4.cu: PR_*); NULL switches it off.
 * Process-wide, not thread-safe: a measurement aid, not part of the data path. */
int tim_debug_role_prof(unsigned long long* dev_buf);

/* timing hook (tools/attn_bench.py): `iters` back-to-back attention launches on a 16-bit two-stream qkv buffer
 * [(B*Ft + B*Qt), 3*H*hd] -> out16 [(B*Ft + B*Qt), H*hd]; version 1 = warp-MMA kernel, 2 = tcgen05 kernel (where supported). */
int tim_bench_attention(int compute_dtype, const void* qkv16, void* out16, int B, int Ft, int Qt, int H, int hd, int version,
                        int iters, float* ms_per_iter);

#ifdef __cplusplus
}
#endif
#endif /* TIM_B200_H */
