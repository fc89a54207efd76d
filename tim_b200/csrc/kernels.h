// Host-side launch interface of the TIM kernels (implemented in the .cu files of this directory).
// Plain pointers + sizes only; no torch types.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tim {

// message returned by tim_last_error(NULL) (entry points without a context; defined in api.cu)
void set_global_error(const char* msg);

enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2 };

// Dropout of the training leg (the reference's nn.Dropout sites: helpers/encodings.py:141,149,177, helpers/transformers.py:73-82,
// 102-108). No mask tensor exists: the keep decision of element e of a site is a pure function of (seed, site, layer, e), evaluated
// wherever the forward or the backward needs it - one 32-bit integer hash (lowbias32) per PAIR of adjacent elements, 16 bits each:
//     h = lowbias32((e >> 1) ^ key);  keep(e) = ((e & 1) ? h >> 16 : h & 0xffff) >= thr;  kept values are scaled by 65536 / (65536 - thr)
// with thr = round(p * 65536). tests/ and oracle/tim_oracle_bwd.py restate the same function in numpy, so gradient parity is
// checked WITH dropout, mask for mask. e is the element's flat index in the library's own layout (two-stream token rows).
struct DropSite {
    uint32_t key = 0;
    uint32_t thr = 0;       // 0: no dropout at this site
    float scale = 1.0f;
};
enum DropSiteId { DROP_FEAT_VIS = 1, DROP_FEAT_AUD = 2, DROP_SEQ = 3, DROP_ATTN = 4, DROP_SUB1 = 5, DROP_FFN = 6, DROP_SUB2 = 7 };
inline DropSite make_drop_site(float p, uint32_t seed, uint32_t site, uint32_t layer) {
    DropSite d;
    if (p <= 0.0f) return d;
    uint32_t thr = static_cast<uint32_t>(p * 65536.0f + 0.5f);
    if (thr > 65535u) thr = 65535u;
    d.thr = thr;
    d.scale = 65536.0f / static_cast<float>(65536u - thr);
    d.key = seed ^ (site * 0x9E3779B9u) ^ (layer * 0x85EBCA6Bu);
    return d;
}
constexpr int DROP_ATTN_KW = 130;    // keys per row in the attention-probability index: feature keys 0 .. Ft - 1, own key at Ft (Ft <= 128)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: remember the largest size requested per device, so a
// process that drives several GPUs (several contexts) sets it on each of them.
struct SmemAttrCache { size_t set[64] = {}; };
template <typename K>
inline cudaError_t ensure_dynamic_smem(K kern, size_t bytes, SmemAttrCache& cache) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (bytes > cache.set[dev]) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
        if (e != cudaSuccess) return e;
        cache.set[dev] = bytes;
    }
    return cudaSuccess;
}

// How the rows of a linear layer's A operand / output are laid out.
// Logical rows are (group g, row r), g < G, r < R. One tile covers box_g groups x box_r rows (<= 128 rows).
//   A row in memory   : g * a_group_rows + a_row_off + r      (SIMT path; the UMMA path bakes this into a 3-D tensor map
//                                                               (K, a_group_rows, G) and uses coordinates (k, a_row_off + r, g))
//   output/resid row  : g * out_group_rows + out_row_off + r
// A plain [M, K] matrix is G = 1, R = M, box_r = 128, box_g = 1.
struct RowMap {
    int G, R;
    int box_r, box_g;
    int a_group_rows, a_row_off;
    int out_group_rows, out_row_off;
};
inline RowMap plain_rows(int M) { return RowMap{1, M, 128, 1, M, 0, M, 0}; }

// out = act(acc + bias) (+ resid), stored as fp32 or as the 16-bit operand type.
struct Epilogue {
    const float* bias;    // [N] or nullptr
    const float* resid;   // fp32 [rows, ldr] or nullptr (added after the activation)
    int ldr;
    // optional "LayerNorm on read" of the residual (16-bit tcgen05 kernels only): when rstats != nullptr the residual
    // buffer holds the PRE-LayerNorm rows z and the value added is (z - mean) * rstd * rgamma[n] + rbeta[n], with
    // (mean, rstd) of the row from rstats[row]. Saves materialising the fp32 LayerNorm output (transformers.py:105,109).
    const float2* rstats;
    const float* rgamma;
    const float* rbeta;
    void* out;            // [rows, ldo]
    int ldo;
    int out_fp32;         // 1: float*, 0: T*
    int act;              // Act
    // training-leg fusions of the CTA-pair kernel (16-bit output only; gemm_umma2.cu modes 8 / 9):
    void* out_act;        // when set: `out` receives T(acc + bias) (the pre-activation kept for the backward) and out_act [rows, ldo]
                          // receives T(act(float(that rounded value))) - linear1 + GELU of the training forward in one launch
    const void* dact_of;  // when set (T [rows, ldo]): out = T((acc + bias) * gelu'(float(dact_of))) - linear2's dgrad with the GELU
                          // backward in its epilogue
};

// ---- tcgen05 GEMM: C[rows, N] = A[rows, K] * W[N, K]^T, 16-bit operands, fp32 accumulate in TMEM ----
struct UmmaParams {
    CUtensorMap tmA;      // 3-D (K, a_group_rows, G), box (64, box_r, box_g), SWIZZLE_128B
    CUtensorMap tmB;      // 2-D (K, N), box (64, block_n), SWIZZLE_128B
    int N, K;
    RowMap rm;
    Epilogue ep;
    int tiles_r, tiles_m, tiles_n;
};
int umma_block_n(int N);   // tile width chosen for an N-column weight (64 / 128 / 256)
template <typename T>
cudaError_t launch_linear_umma(UmmaParams p, int block_n, int num_sms, cudaStream_t s);

// ---- CTA-pair tcgen05 GEMM (cta_group::2, 256 x 256 tiles, TMA-store epilogue): plain [M, K] x [N, K]^T only ----
struct Umma2Params {
    CUtensorMap tmA;      // 2-D (K, M), box (64, 128), SWIZZLE_128B
    CUtensorMap tmB;      // 2-D (K, N), box (64, 128), SWIZZLE_128B
    CUtensorMap tmOut;    // 2-D (N, M): 16-bit box (64, 32) / fp32 box (32, 32), SWIZZLE_128B
    CUtensorMap tmRes;    // fp32 (N, M), box (32, 32), SWIZZLE_128B (modes 2 / 3 only)
    const float* bias;    // [N]
    const float2* rstats; // mode 3: per-row (mean, rstd) of the residual rows; rgamma / rbeta [N]
    const float* rgamma;
    const float* rbeta;
    // folded-LayerNorm pair (modes 5 / 6); rstats doubles as the statistics of the residual rows (5) / of the A rows (6)
    CUtensorMap tmOut16;  // mode 5: 16-bit copy of the output, (N, M), box (64, 32), SWIZZLE_128B
    // mode 7 (two-plane residual stream): tmOut / tmRes are the HI planes, tmOutLo / tmResLo the LO planes; all four 16-bit (N, M),
    // box (32, 32), SWIZZLE_64B
    CUtensorMap tmOutLo, tmResLo;
    float2* opart;        // mode 5: [2 * tiles_n][M] partial (sum, sum of squares) of the rows written
    const float* cs;      // mode 6: [N] row sums of the gamma-folded 16-bit weight
    int M, N, K;
    int tiles_m, tiles_n;
};
bool umma2_supported(int M, int N, int K);
// programmatic dependent launch of the pair GEMM and the tcgen05 attention forward (ptx.cuh: pdl_wait); TIM_B200_PDL=0 switches it off.
// Measured (profiles/r02z, r03a): e2e call 25.0-25.6 -> 23.9-24.4 ms, device-resident step -0.1 ... -0.4 ms; extending it to the row kernels
// and the single-CTA GEMM gave nothing and cost 0.3 ms of the device-resident step, the training kernels showed no change - both left out.
bool pdl_enabled();
// <<<grid, block, smem, s>>> with the programmatic-stream-serialization attribute when `pdl`
// (every kernel launched this way executes pdl_wait() before its first global access)
template <typename... KA, typename... A>
inline cudaError_t launch_maybe_pdl(void (*kern)(KA...), unsigned grid, unsigned block, size_t smem, cudaStream_t s, bool pdl, A&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KA>(args)...);
}
// mode 0: out = T(act(acc + bias)); 1: out = float(act(acc + bias)); 2: out = float(act(acc + bias) + resid);
// mode 3: out = float(act(acc + bias) + LayerNorm(resid)) with the row statistics given (Epilogue::rstats)
// modes 5 / 6: the folded-LayerNorm producer / consumer pair, see gemm_umma2.cu; mode 7: producer on the two-plane residual stream;
// mode 8: two 16-bit outputs (pre-activation in tmOut, activation in tmOut16); mode 9: 16-bit out = (acc + bias) * gelu'(tmRes tile, 16-bit)
template <typename T>
cudaError_t launch_linear_umma2(Umma2Params p, int mode, int act, int num_sms, cudaStream_t s);

// ---- fp32 SIMT GEMM with the same row mapping / epilogue (fp32 parity mode) ----
cudaError_t launch_linear_simt(const float* A, int lda, const float* W, int N, int K, RowMap rm, Epilogue ep,
                               cudaStream_t s);

// ---- attention over the two-stream token layout ----
// qkv: [B*Ft + B*Qt, 3E]; feature rows of clip b at b*Ft.., query rows at B*Ft + b*Qt..
// q columns are pre-scaled by head_dim^-0.5 * log2(e) (folded into the packed in_proj weights).
// Every row attends to the Ft feature keys of its clip; query rows additionally to their own key.
template <typename T>
cudaError_t launch_attention_mma(const T* qkv, T* out, int B, int Ft, int Qt, int H, int hd, cudaStream_t s);
// tcgen05 version (head_dim 64 / 128 / 192, Ft <= 128): S = Q K_f^T and O = P V_f on the tensor cores, TMEM accumulators,
// TMA loads / stores through 3-D maps over (columns, rows of a clip, clips). All maps use 128-byte swizzle.
struct AttnUmmaParams {
    CUtensorMap tmKV;     // qkv feature rows  (3E, Ft, B), box (64, Fp, 1)   - K_f / V_f of one (clip, head)
    CUtensorMap tmQf;     // qkv feature rows  (3E, Ft, B), box (64, 128, 1)  - Q tile of the feature rows
    CUtensorMap tmQq;     // qkv query rows    (3E, Qt, B), box (64, 128, 1)
    CUtensorMap tmOf;     // out feature rows  (E, Ft, B),  box (64, 32, 1)
    CUtensorMap tmOq;     // out query rows    (E, Qt, B),  box (64, 32, 1)
    const void* qkv;      // raw pointer for the own-key / own-value rows of query tokens
    int B, Ft, Qt, H;
    int Fp;               // Ft rounded up to 16 (MMA N of S, MMA K of P.V)
    int tiles_q;          // 128-row query tiles per clip
    int tpu, chunks;      // row tiles per work unit, work units per (clip, head)
    int num_units;
    int pf_mode, pf_tiles; // L2 prefetch policy of the TMA producer (see launch_attention_umma)
    DropSite drop;         // attention-probability dropout (training forward; attention_umma.cu / attention_umma4.cu)
    unsigned long long* prof;  // role profile (tim_debug_role_prof): [CTA][16] cycle counters, NULL = off (attention_umma4.cu)
};
// device buffer [>= 148][16] of 64-bit counters the decoupled attention kernel adds its per-role wait / busy cycles to (NULL = off)
void set_attention_role_prof(unsigned long long* dev_buf);
bool attention_umma_supported(int Ft, int hd);
size_t attention_umma_smem(int Ft, int hd);
// fills Fp / tiles_q / tpu / chunks / num_units from B, Ft, Qt, H (the maps and qkv must already be set)
template <typename T>
cudaError_t launch_attention_umma(AttnUmmaParams p, int hd, int num_sms, cudaStream_t s);

// decoupled-pipeline form (attention_umma4.cu): event-driven MMA issue, own output staging tile, stages released at P.V completion;
// same parameter block and tensor maps. head_dim 64 / 128 and K_f + V_f + two stages + staging within 227 KB
bool attention_umma4_supported(int Ft, int hd);
template <typename T>
cudaError_t launch_attention_umma4(AttnUmmaParams p, int hd, int num_sms, cudaStream_t s);

cudaError_t launch_attention_simt(const float* qkv, float* out, int B, int Ft, int Qt, int H, int hd, cudaStream_t s, DropSite drop = DropSite());
size_t attention_simt_smem(int Ft, int hd);

// ---- elementwise / row kernels ----
// relu(times[M,2] * W[d,2]^T + b) -> out[M, d]
template <typename TO>
cudaError_t launch_time_l1(const float* times, const float* W, const float* b, TO* out, int M, int d, cudaStream_t s);

// LayerNorm over rows of `in` [M, n] (row stride ldi); writes fp32 (out32, stride ld32) and/or T (out16, stride ld16)
// and/or the row statistics (mean, rstd) to stats[M]
template <typename T>
cudaError_t launch_layernorm(const float* in, int ldi, const float* gamma, const float* beta, float* out32, int ld32,
                             T* out16, int ld16, int M, int n, cudaStream_t s, float2* stats = nullptr);

// LayerNorm over rows held as two 16-bit planes (z = hi + lo, the two-plane residual stream of gemm_umma2.cu mode 7); out16 may alias hi
template <typename T>
cudaError_t launch_layernorm_planes(const T* hi, const T* lo, int ldi, const float* gamma, const float* beta, float* out32, int ld32,
                                    T* out16, int ld16, int M, int n, cudaStream_t s);

// Token assembly (encodings.py forward): see elementwise.cu for the row plan
struct TokenGroup {          // a run of `count` query tokens per clip
    const float* cls;        // [d] CLS parameter (left half)
    const float* mod;        // [E] modality encoding or nullptr
    int te_off;              // first time-encoding row (inside a clip's T rows) feeding the right half
    int count;
};
struct AssembleParams {
    int B, d, T;             // clips, d_model, time rows per clip
    int Fv, Fa;              // feature tokens per modality (0 if absent)
    const float* emb_v;      // [B*Fv, d] pre-LayerNorm embedder output (after GELU)
    const float* emb_a;      // [B*Fa, d]
    const float* ln_v_g; const float* ln_v_b;
    const float* ln_a_g; const float* ln_a_b;
    const float* mod_v; const float* mod_a;   // [E] or nullptr
    const float* te;         // [B, T, d] time encodings
    int n_groups;
    TokenGroup groups[4];
    int Qt;                  // query tokens per clip = sum(groups.count)
    float* x32;              // [B*Ft + B*Qt, E] fp32 residual stream (two-stream layout)
    void* x16;               // same rows, operand dtype (nullptr in fp32 mode)
    void* xlo;               // optional LO plane of the two-plane residual stream: T(x - float(T(x)))   (x16 is the HI plane)
};
template <typename T>
cudaError_t launch_assemble(const AssembleParams& p, cudaStream_t s);

// fp32 -> T cast, optional scaling of the first `scale_rows` rows (n = rows * cols, contiguous)
template <typename T>
cudaError_t launch_cast(const float* in, T* out, size_t rows, size_t cols, size_t scale_rows, float scale, cudaStream_t s);
cudaError_t launch_scale_copy(const float* in, float* out, size_t rows, size_t cols, size_t scale_rows, float scale,
                              cudaStream_t s);

// out[m, :] = TO(bank[rows[m], :]); bank_dtype 0 fp32 / 1 bf16 / 2 fp16; rows outside [0, bank_rows) give zeros
template <typename TO>
cudaError_t launch_gather_rows(const void* bank, int bank_dtype, long long bank_rows, const long long* rows, TO* out, long long M, int D,
                               cudaStream_t s, int* bad = nullptr);

// (mean, rstd) per row from the partial sums a mode-5 GEMM wrote: part [P][M] (sum, sum of squares), width columns in total;
// *alarm (device int, optional) is set when some row has mean^2 > alarm_ratio * var
cudaError_t launch_row_stats_finalize(const float2* part, int P, int width, float2* stats, int M, float alarm_ratio, int* alarm,
                                      cudaStream_t s);

// LayerNorm affine folded into the following linear layer's weight (see fold_ln_kernel); W, bias: fp32 masters
template <typename T>
cudaError_t launch_fold_ln(const float* W, const float* bias, const float* gamma, const float* beta, T* Wf, float* cs, float* bw,
                           int N, int K, int scale_rows, float scale, cudaStream_t s);

// last regression layer: sigmoid(h[rows, K] * W[2, K]^T + b) -> out[rows, 2] fp32
template <typename T>
cudaError_t launch_reg_final(const T* h, int ldh, const float* W, const float* b, float* out, int rows, int K,
                             cudaStream_t s);

// =====================================================================================================================
// training leg (backward kernels)
// =====================================================================================================================
// ---- weight gradient: dW[N, K] += dY[M, N]^T * X[M, K] (gemm_wgrad.cu) ----
struct WgradParams {
    CUtensorMap tmA;      // dY: 2-D (N, M) 16-bit, box (64, 64), SWIZZLE_128B
    CUtensorMap tmB;      // X : 2-D (K, M) 16-bit, box (64, 64), SWIZZLE_128B
    CUtensorMap tmOut;    // dW: 2-D (K, N) fp32,   box (32, 32), SWIZZLE_128B (TMA reduce-add target)
    int M, N, K;
    int tiles_n, tiles_k, splits, kb_per_split;      // filled by the launcher (splits <= 0: chosen by wgrad_pick_splits)
};
bool wgrad_umma_supported(int M, int N, int K, int ldy, int ldx, int ldw);
int wgrad_pick_splits(int M, int N, int K, int num_sms);
template <typename T>
cudaError_t launch_wgrad_umma(WgradParams p, int num_sms, cudaStream_t s);
// CUDA-core version (fp32 parity mode: TA = float; also any shape the tensor-core kernel does not take)
template <typename TA>
cudaError_t launch_wgrad_simt(const TA* dY, int ldy, const TA* X, int ldx, float* dW, int ldw, int M, int N, int K, cudaStream_t s);

// ---- attention backward over the two-stream layout (attention_bwd.cu); qscale: see the file header ----
size_t attention_bwd_stats_bytes(int B, int Ft, int Qt, int H);
template <typename T>
cudaError_t launch_attention_bwd(const T* qkv, const T* dO, T* dqkv, void* stats, int B, int Ft, int Qt, int H, int hd, float qscale,
                                 cudaStream_t s);
// tcgen05 version (head_dim 64 / 128, Ft <= 128), attention_bwd_umma.cu. All maps: 16-bit, 128-byte swizzle, 3-D (columns, rows of a clip, clips)
struct AttnBwdUmmaParams {
    CUtensorMap tmKV;     // qkv feature rows (3E, Ft, B), box (64, Fp, 1)   - K_f / V_f of one (clip, head)
    CUtensorMap tmQf;     // qkv feature rows (3E, Ft, B), box (64, 128, 1)  - Q tile of the feature rows
    CUtensorMap tmQq;     // qkv query rows   (3E, Qt, B), box (64, 128, 1)
    CUtensorMap tmDf;     // dO feature rows  (E, Ft, B),  box (64, 128, 1)
    CUtensorMap tmDq;     // dO query rows    (E, Qt, B),  box (64, 128, 1)
    CUtensorMap tmGf;     // dqkv feature rows (3E, Ft, B), box (64, 32, 1)  - dQ slabs of the epilogue warps (bulk stores)
    CUtensorMap tmGq;     // dqkv query rows   (3E, Qt, B), box (64, 32, 1)
    const void* qkv; const void* dO; void* dqkv;
    int B, Ft, Qt, H;
    int Fp, tiles_q, num_units;       // filled by the launcher
    float qscale;
    DropSite drop;                    // attention-probability dropout of the forward this is the backward of
};
bool attention_bwd_umma_supported(int Ft, int hd);
template <typename T>
cudaError_t launch_attention_bwd_umma(AttnBwdUmmaParams p, int hd, int num_sms, cudaStream_t s);

size_t attention_bwd_simt_smem(int Ft, int hd);
cudaError_t launch_attention_bwd_simt(const float* qkv, const float* dO, float* dqkv, int B, int Ft, int Qt, int H, int hd, float qscale,
                                      cudaStream_t s, DropSite drop = DropSite());

// ---- row / elementwise kernels (train_rows.cu) ----
// LayerNorm backward; dy is overwritten by dz, dz16 (optional) receives its operand copy; dgamma / dbeta / dbias accumulate (+=)
// drop: dropout of the sub-layer output that was added into z (dropout1 / dropout2): dz16 and dbias then carry dz o mask (the
// gradient w.r.t. the sub-layer's output), dy keeps the un-masked dz (the residual branch)
template <typename T>
cudaError_t launch_ln_bwd(float* dy, int ldd, const float* z, int ldz, const float* gamma, T* dz16, int ld16, float* dgamma, float* dbeta,
                          float* dbias, int M, int n, cudaStream_t s, DropSite drop = DropSite());
// out = d * f'(a); mode 0: erf-GELU, a = pre-activation; mode 1: ReLU, a = post-activation. dbias (optional) += column sums of out
template <typename TD, typename TA, typename TO>
cudaError_t launch_act_bwd(int mode, const TD* d, const TA* a, TO* out, int rows, int cols, float* dbias, cudaStream_t s, DropSite drop = DropSite());
template <typename T>
cudaError_t launch_gelu_fwd(const T* u, T* h, size_t n, cudaStream_t s, DropSite drop = DropSite());
// out = T(in o mask) over a flat fp32 array (input-feature dropout fused into the operand cast); n even
template <typename T>
cudaError_t launch_drop_cast(const float* in, T* out, size_t n, DropSite drop, cudaStream_t s);
// x32 (and its 16-bit copy x16, optional) *= mask, in place, flat index = element index (token dropout; gradient masking)
template <typename T>
cudaError_t launch_drop_apply(float* x32, T* x16, size_t n, DropSite drop, cudaStream_t s);
// z = R(resid) + a o mask, fp32 [M, n]: the sub-layer output a (fp32, no residual yet) is dropped and added to the residual;
// R = LayerNorm-on-read with (rstats, rgamma, rbeta) or the identity when rstats == nullptr. z may alias a.
cudaError_t launch_residual_drop(const float* a, const float* resid, const float2* rstats, const float* rgamma, const float* rbeta, float* z,
                                 int M, int n, DropSite drop, cudaStream_t s);
// out[c] += sum_{g < G, r < R} x[(g * group_rows + row_off + r) * ld + col_off + c],  c < ncols
template <typename T>
cudaError_t launch_colsum(const T* x, int ld, int G, int group_rows, int row_off, int R, int col_off, int ncols, float* out, cudaStream_t s);
// Wt[k, n] = T(W[n, k]) (n < N), 0 for N <= n < Np
template <typename T>
cudaError_t launch_transpose_pack(const float* W, T* Wt, int N, int K, int Np, cudaStream_t s);
template <typename T>
cudaError_t launch_cast_pad(const float* in, T* out, size_t rows, int C, int Cp, cudaStream_t s);
template <typename T>
cudaError_t launch_gather_group(const T* x, T* out, int B, int Qt, int off, int Q, int E, cudaStream_t s);
cudaError_t launch_scatter_add_group(const float* in, float* dx, int B, int Qt, int off, int Q, int E, cudaStream_t s);
struct AssembleBwdParams {
    int B, d, T, Fv, Fa, Qt;
    const float* dtok;       // [B*Ft + B*Qt, E] gradient w.r.t. the assembled tokens (two-stream layout)
    float* dte;              // [B, T, d] gradient w.r.t. the time encodings (written, not accumulated)
    float* demb_v;           // [B*Fv, d] gradient w.r.t. LayerNorm(GELU(embedder)) rows
    float* demb_a;           // [B*Fa, d]
    int n_groups;
    TokenGroup groups[4];
};
cudaError_t launch_assemble_bwd(const AssembleBwdParams& p, cudaStream_t s);
template <typename T>
cudaError_t launch_time_l0_bwd(const T* d1, const float* times, float* dW0, int M, int d, cudaStream_t s);
template <typename T>
cudaError_t launch_reg_final_bwd(const float* dout, const float* y, const T* h, const float* W4, float* dW4, float* db4, T* dh, float* db2,
                                 int rows, int K, cudaStream_t s);
cudaError_t launch_axpy(float* y, const float* x, size_t n, cudaStream_t s);

}  // namespace tim
