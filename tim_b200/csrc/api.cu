// C ABI of libtim_b200 (declared in include/tim_b200.h): context, weight packing keyed by the reference's
// state_dict names, and the host-side schedule of the forward (which kernels run in which order on which buffers).
// The arithmetic lives in gemm_umma.cu / gemm_simt.cu / attention.cu / elementwise.cu.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <map>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/tim_b200.h"
#include "kernels.h"

using namespace tim;

namespace {

thread_local std::string g_create_error;

}  // namespace
namespace tim {
void set_global_error(const char* msg) { g_create_error = msg ? msg : ""; }
}  // namespace tim
namespace {

struct LinearW {
    int N = 0, K = 0;
    void* w = nullptr;          // [N, K] in the compute dtype
    float* bias = nullptr;      // [N] fp32
    CUtensorMap tmB;            // 16-bit modes only: box (64, block_n) for the single-CTA kernel
    CUtensorMap tmB2;           // box (64, 128) for the CTA-pair kernel
    bool has_tmB2 = false;
    int block_n = 0;
    int scale_rows = 0;         // leading rows (and bias entries) multiplied by `scale` when packed (q part of in_proj)
    float scale = 1.0f;
    // folded-LayerNorm variant (in_proj of layers >= 1, linear1): fp32 masters + gamma-folded 16-bit weight, see fold_ln_kernel
    bool foldable = false;
    float* w32 = nullptr;       // [N, K] fp32 master (unscaled)
    float* b32 = nullptr;       // [N] fp32 master (unscaled)
    void* wf = nullptr;         // [N, K] folded, compute dtype
    float* cs = nullptr;        // [N]
    float* bwf = nullptr;       // [N]
    CUtensorMap tmB2f;          // CTA-pair map of wf
    // training leg: transposed copy [K, Np] (Np = N rounded up to 8, zero-padded; unscaled) = the K-major "weight" of the dgrad GEMM
    void* wt = nullptr;
    int Np = 0;
    CUtensorMap tmBt, tmBt2;
    bool has_tmBt2 = false;
    int block_n_t = 0;
    bool wt_set = false;
};

struct Slot {                   // one state_dict key
    enum Kind { LIN_W, LIN_B, VEC } kind;
    LinearW* lin = nullptr;
    float* vec = nullptr;       // VEC destination
    std::vector<int64_t> shape;
    bool set = false;
};

struct Layer {
    LinearW in_proj, out_proj, lin1, lin2;
    float *n1g, *n1b, *n2g, *n2b;
};

struct RegHead {
    LinearW l0, l2;
    float* w4 = nullptr;        // [2, E/2]
    float* b4 = nullptr;        // [2]
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

struct tim_train_state;
struct ncclUniqueIdBlob { char internal[128]; };

struct tim_ctx {
    tim_config cfg;
    tim_train_state* train = nullptr;   // training leg (train.inl); allocated by tim_train_enable
    int class_override = -1;            // profiling class forced onto run_linear launches (dgrad GEMMs)
    int wgrad_splits = 0;               // 0: chosen per shape (TIM_B200_WGRAD_SPLITS overrides)
    bool train_fuse = true;             // training leg: linear1 + GELU and linear2-dgrad + GELU' as single launches (TIM_B200_TRAIN_FUSE=0: separate row kernels)
    int device = 0, num_sms = 0;
    int d = 0, E = 0, FF = 0, H = 0, hd = 0, L = 0, F = 0, Fv = 0, Fa = 0, Ft = 0;
    bool vis_data = false, aud_data = false, vn_tokens = false;
    size_t esize = 4;           // bytes per element of the compute dtype
    std::string err;
    uint64_t launches = 0;
    EncodeTiledFn encode = nullptr;
    int gemm_version = 2;       // 2: CTA-pair kernel where the shape allows, 1: single-CTA kernel only (TIM_B200_GEMM=1)
    int attn_version = 4;       // 4: tcgen05 attention, decoupled pipeline (attention_umma4.cu, default where it fits, else 2); 2: the r01 form (attention_umma.cu); 1: warp-MMA attention only (TIM_B200_ATTN)
    bool fold_ln = false;       // encoder LayerNorms folded into the GEMMs around them (16-bit path, CTA-pair kernel shapes; TIM_B200_FOLD=0 disables)
    bool fold_dirty = true;     // a weight changed since the folded copies were made
    bool planes = true;         // folded flow keeps the residual stream as two 16-bit planes (mode 7); TIM_B200_PLANES=0: fp32 + 16-bit copy (mode 5)
    // precision guard of the folded path: the finalize kernel flags rows with |mean| > 8 std (relative error of the folded
    // operand x8); the flag is copied to pinned host memory at the end of every forward and read, without a synchronisation,
    // at the start of the next one - from then on this context takes the un-folded flow.
    int* fold_alarm_dev = nullptr;
    volatile int* fold_alarm_host = nullptr;
    bool fold_tripped = false;
    bool last_folded = false;           // the most recent encoder forward took the folded flow
    cudaEvent_t fold_event = nullptr;   // recorded behind the alarm's D2H copy of that forward (tim_fold_check waits on it)
    // out-of-range rows of tim_encoder_fwd_indexed: counted on the device, copied to pinned host memory behind the forward
    int* idx_alarm_dev = nullptr;
    volatile int* idx_alarm_host = nullptr;
    cudaEvent_t idx_event = nullptr;
    std::vector<cudaEvent_t> host_events;   // reused by tim_forward_host_ex (three per chunk) + one for the caller's stream

    // optional live profiling: CUDA-event pairs around every launch, accumulated per kernel class
    bool profiling = false;
    int cur_class = 0;
    double cur_flops = 0.0;
    struct ProfRec { cudaEvent_t a, b; int cls; double flops; };
    std::vector<ProfRec> prof;
    std::vector<cudaEvent_t> ev_pool;

    // weights
    std::map<std::string, Slot> slots;
    std::vector<void*> allocs;
    float *t0w = nullptr, *t0b = nullptr, *tlg = nullptr, *tlb = nullptr;
    LinearW t2, t4, emb_v, emb_a;
    float *lnv_g = nullptr, *lnv_b = nullptr, *lna_g = nullptr, *lna_b = nullptr;
    float *mod_v = nullptr, *mod_a = nullptr;
    float *cls_verb = nullptr, *cls_noun = nullptr, *cls_action = nullptr, *cls_audio = nullptr;
    std::vector<Layer> layers;
    LinearW h_verb, h_noun, h_action, h_audio;
    RegHead reg_v, reg_a;

    // workspace arena (grow-only)
    uint8_t* ws = nullptr;
    size_t ws_bytes = 0;
    // staging for tim_forward_host
    uint8_t* io = nullptr;
    size_t io_bytes = 0;
    cudaStream_t s_h2d = nullptr, s_comp = nullptr, s_d2h = nullptr;

    int fail(int code, const char* fmt, ...) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof(buf), fmt, ap);
        va_end(ap);
        err = buf;
        return code;
    }
};

namespace {

cudaEvent_t prof_event(tim_ctx* c) {
    cudaEvent_t e = nullptr;
    if (!c->ev_pool.empty()) { e = c->ev_pool.back(); c->ev_pool.pop_back(); }
    else cudaEventCreate(&e);
    return e;
}

#define CU_OK(ctx, expr)                                                                                   \
    do {                                                                                                   \
        cudaError_t _e = (expr);                                                                           \
        if (_e != cudaSuccess)                                                                             \
            return (ctx)->fail(TIM_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

// kernel classes for tim_profile_*: 0 tcgen05/SIMT GEMM, 1 attention, 2 LayerNorm, 3 token assembly, 4 other row kernels,
// 5 folded-LayerNorm consumer GEMMs (in_proj, linear1), 6 producer GEMMs (out_proj, linear2)
#define LAUNCH_C(ctx, cls_, flops_, stream_, expr)                                                         \
    do {                                                                                                   \
        tim_ctx::ProfRec _pr;                                                                              \
        const bool _prof = (ctx)->profiling;                                                               \
        if (_prof) { _pr.a = prof_event(ctx); _pr.b = prof_event(ctx);                                     \
                     _pr.cls = ((ctx)->class_override >= 0 && (cls_) != 2) ? (ctx)->class_override : (cls_); _pr.flops = (flops_); \
                     cudaEventRecord(_pr.a, (stream_)); }                                                  \
        cudaError_t _e = (expr);                                                                           \
        (ctx)->launches++;                                                                                 \
        if (_prof) { cudaEventRecord(_pr.b, (stream_)); (ctx)->prof.push_back(_pr); }                      \
        if (_e != cudaSuccess)                                                                             \
            return (ctx)->fail(TIM_ERR_CUDA, "launch %s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)
#define LAUNCH(ctx, expr) LAUNCH_C(ctx, 4, 0.0, s, expr)

#define TIM_TRY(expr)              \
    do {                           \
        int _r = (expr);           \
        if (_r != TIM_OK) return _r; \
    } while (0)

int dev_alloc(tim_ctx* c, void** p, size_t bytes) {
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) return c->fail(TIM_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    c->allocs.push_back(*p);
    return TIM_OK;
}

int get_encode_fn(tim_ctx* c) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn)
        return c->fail(TIM_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable: %s", cudaGetErrorString(e));
    c->encode = reinterpret_cast<EncodeTiledFn>(fn);
    return TIM_OK;
}

// 16-bit K-major operand map: dims (K, rows, groups), box (64, box_r, box_g), 128-byte swizzle, zero OOB fill.
int make_tmap(tim_ctx* c, CUtensorMap* tm, const void* base, int K, long long rows, long long groups, int box_r, int box_g) {
    if (K % 8) return c->fail(TIM_ERR_INVALID, "tcgen05 path needs K %% 8 == 0 (got K=%d)", K);
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return c->fail(TIM_ERR_INVALID, "operand base not 16-byte aligned");
    const CUtensorMapDataType dt = c->cfg.compute_dtype == TIM_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(groups)};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(K) * 2, static_cast<cuuint64_t>(K) * 2 * static_cast<cuuint64_t>(rows)};
    cuuint32_t box[3] = {64, static_cast<cuuint32_t>(box_r), static_cast<cuuint32_t>(box_g)};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = c->encode(tm, dt, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return c->fail(TIM_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) K=%d rows=%lld groups=%lld box=(%d,%d)", static_cast<int>(r), K, rows, groups, box_r, box_g);
    return TIM_OK;
}

// generic row-major 2-D map: `inner` elements per row, `rows` rows, row pitch `pitch_bytes`; 128-byte swizzle
int make_tmap_2d(tim_ctx* c, CUtensorMap* tm, const void* base, CUtensorMapDataType dt, int esize, long long inner, long long rows,
                 long long pitch_bytes, int box_inner, int box_rows) {
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (pitch_bytes & 15) != 0 || box_inner * esize != 128)
        return c->fail(TIM_ERR_INVALID, "TMA map: base / pitch must be 16-byte aligned and the box 128 bytes wide");
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(rows)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(pitch_bytes)};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(box_inner), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = c->encode(tm, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return c->fail(TIM_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) inner=%lld rows=%lld pitch=%lld box=(%d,%d)", static_cast<int>(r),
                       inner, rows, pitch_bytes, box_inner, box_rows);
    return TIM_OK;
}
// 16-bit plane tile map of the two-plane residual stream: box (32 columns = 64 bytes, 32 rows), 64-byte swizzle
int make_tmap_plane(tim_ctx* c, CUtensorMap* tm, const void* base, CUtensorMapDataType dt, long long inner, long long rows, long long pitch_bytes) {
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (pitch_bytes & 15) != 0)
        return c->fail(TIM_ERR_INVALID, "TMA map: base / pitch must be 16-byte aligned");
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(rows)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(pitch_bytes)};
    cuuint32_t box[2] = {32, 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = c->encode(tm, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return c->fail(TIM_ERR_CUDA, "cuTensorMapEncodeTiled(plane) failed (%d) inner=%lld rows=%lld", static_cast<int>(r), inner, rows);
    return TIM_OK;
}
// 3-D map (inner elements, rows of a group, groups) with explicit pitches; 16-bit elements, box (64, box_rows, 1), 128-byte swizzle.
// Rows / groups outside the tensor are zero-filled on load and skipped on store (the attention kernel relies on both).
int make_tmap_3d16(tim_ctx* c, CUtensorMap* tm, const void* base, long long inner, long long rows, long long groups,
                   long long row_pitch_bytes, long long group_pitch_bytes, int box_rows) {
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (row_pitch_bytes & 15) != 0 || (group_pitch_bytes & 15) != 0)
        return c->fail(TIM_ERR_INVALID, "TMA map: base / pitches must be 16-byte aligned");
    if (inner <= 0 || rows <= 0 || groups <= 0 || box_rows <= 0 || box_rows > 256)
        return c->fail(TIM_ERR_INVALID, "TMA map: empty tensor or bad box");
    const CUtensorMapDataType dt = c->cfg.compute_dtype == TIM_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(groups)};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(row_pitch_bytes), static_cast<cuuint64_t>(group_pitch_bytes)};
    cuuint32_t box[3] = {64, static_cast<cuuint32_t>(box_rows), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = c->encode(tm, dt, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return c->fail(TIM_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) inner=%lld rows=%lld groups=%lld box_rows=%d", static_cast<int>(r),
                       inner, rows, groups, box_rows);
    return TIM_OK;
}
inline CUtensorMapDataType op_dtype(const tim_ctx* c) {
    return c->cfg.compute_dtype == TIM_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
}

int make_tmap_w(tim_ctx* c, LinearW& w) {
    const CUtensorMapDataType dt = c->cfg.compute_dtype == TIM_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    if (w.K % 8) return c->fail(TIM_ERR_INVALID, "tcgen05 path needs K %% 8 == 0 (weight K=%d)", w.K);
    w.block_n = umma_block_n(w.N);
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(w.K), static_cast<cuuint64_t>(w.N)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(w.K) * 2};
    cuuint32_t box[2] = {64, static_cast<cuuint32_t>(w.block_n)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = c->encode(&w.tmB, dt, 2, w.w, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return c->fail(TIM_ERR_CUDA, "cuTensorMapEncodeTiled(weight %dx%d) failed (%d)", w.N, w.K, static_cast<int>(r));
    w.has_tmB2 = false;
    if (w.N >= 128) {
        TIM_TRY(make_tmap_2d(c, &w.tmB2, w.w, dt, 2, w.K, w.N, static_cast<long long>(w.K) * 2, 64, 128));
        w.has_tmB2 = true;
    }
    return TIM_OK;
}

int add_linear(tim_ctx* c, const std::string& prefix, LinearW& w, int N, int K, int scale_rows = 0, float scale = 1.0f,
               const char* wname = "weight", const char* bname = "bias") {
    w.N = N; w.K = K; w.scale_rows = scale_rows; w.scale = scale;
    TIM_TRY(dev_alloc(c, &w.w, static_cast<size_t>(N) * K * c->esize));
    TIM_TRY(dev_alloc(c, reinterpret_cast<void**>(&w.bias), static_cast<size_t>(N) * sizeof(float)));
    if (c->cfg.compute_dtype != TIM_FP32) TIM_TRY(make_tmap_w(c, w));
    Slot sw; sw.kind = Slot::LIN_W; sw.lin = &w; sw.shape = {N, K};
    Slot sb; sb.kind = Slot::LIN_B; sb.lin = &w; sb.shape = {N};
    c->slots[prefix + wname] = sw;
    c->slots[prefix + bname] = sb;
    return TIM_OK;
}

int add_vec(tim_ctx* c, const std::string& key, float** dst, std::vector<int64_t> shape) {
    size_t n = 1;
    for (auto s : shape) n *= static_cast<size_t>(s);
    TIM_TRY(dev_alloc(c, reinterpret_cast<void**>(dst), n * sizeof(float)));
    Slot s; s.kind = Slot::VEC; s.vec = *dst; s.shape = shape;
    c->slots[key] = s;
    return TIM_OK;
}

int build_weights(tim_ctx* c) {
    const tim_config& g = c->cfg;
    const int d = c->d, E = c->E, FF = c->FF;
    TIM_TRY(add_vec(c, "time_mlp.0.weight", &c->t0w, {d, 2}));
    TIM_TRY(add_vec(c, "time_mlp.0.bias", &c->t0b, {d}));
    TIM_TRY(add_linear(c, "time_mlp.2.", c->t2, d, d));
    TIM_TRY(add_linear(c, "time_mlp.4.", c->t4, d, d));
    TIM_TRY(add_vec(c, "time_mlp.6.weight", &c->tlg, {d}));
    TIM_TRY(add_vec(c, "time_mlp.6.bias", &c->tlb, {d}));
    const std::string fe = "feature_encoding.";
    const bool recog = g.variant == TIM_RECOGNITION;
    if (g.input_modality == TIM_AUDIO_VISUAL) {
        TIM_TRY(add_vec(c, fe + "visual_modality_encoding", &c->mod_v, {1, 1, E}));
        TIM_TRY(add_vec(c, fe + "audio_modality_encoding", &c->mod_a, {1, 1, E}));
        if (c->vis_data) {
            TIM_TRY(add_vec(c, fe + "visual_action_cls", &c->cls_action, {1, 1, d}));
            if (c->vn_tokens) {
                TIM_TRY(add_vec(c, fe + "visual_verb_cls", &c->cls_verb, {1, 1, d}));
                TIM_TRY(add_vec(c, fe + "visual_noun_cls", &c->cls_noun, {1, 1, d}));
            }
        }
        if (c->aud_data) TIM_TRY(add_vec(c, fe + "audio_action_cls", &c->cls_audio, {1, 1, d}));
    } else if (g.input_modality == TIM_VISUAL) {
        TIM_TRY(add_vec(c, fe + (recog ? "action_cls" : "visual_action_cls"), &c->cls_action, {1, 1, d}));
        if (c->vn_tokens) {
            TIM_TRY(add_vec(c, fe + "verb_cls", &c->cls_verb, {1, 1, d}));
            TIM_TRY(add_vec(c, fe + "noun_cls", &c->cls_noun, {1, 1, d}));
        }
    } else {
        TIM_TRY(add_vec(c, fe + (recog ? "action_cls" : "audio_action_cls"), &c->cls_audio, {1, 1, d}));
    }
    if (c->Fv) {
        TIM_TRY(add_linear(c, fe + "visual_embedder.1.", c->emb_v, d, g.vis_dim));
        TIM_TRY(add_vec(c, fe + "visual_embedder.3.weight", &c->lnv_g, {d}));
        TIM_TRY(add_vec(c, fe + "visual_embedder.3.bias", &c->lnv_b, {d}));
    }
    if (c->Fa) {
        TIM_TRY(add_linear(c, fe + "audio_embedder.1.", c->emb_a, d, g.aud_dim));
        TIM_TRY(add_vec(c, fe + "audio_embedder.3.weight", &c->lna_g, {d}));
        TIM_TRY(add_vec(c, fe + "audio_embedder.3.bias", &c->lna_b, {d}));
    }
    if (g.n_verb) TIM_TRY(add_linear(c, "cls_head.fc_visual_verb.", c->h_verb, g.n_verb, E));
    if (g.n_noun) TIM_TRY(add_linear(c, "cls_head.fc_visual_noun.", c->h_noun, g.n_noun, E));
    if (g.n_action) TIM_TRY(add_linear(c, "cls_head.fc_visual_action.", c->h_action, g.n_action, E));
    if (g.n_audio) TIM_TRY(add_linear(c, "cls_head.fc_audio_action.", c->h_audio, g.n_audio, E));
    if (g.variant == TIM_DETECTION) {
        auto add_reg = [&](const std::string& p, RegHead& r) -> int {
            TIM_TRY(add_linear(c, p + "0.", r.l0, E / 2, E));
            TIM_TRY(add_linear(c, p + "2.", r.l2, E / 2, E / 2));
            TIM_TRY(add_vec(c, p + "4.weight", &r.w4, {2, E / 2}));
            TIM_TRY(add_vec(c, p + "4.bias", &r.b4, {2}));
            return TIM_OK;
        };
        if (c->vis_data) TIM_TRY(add_reg("reg_head.fc_visual_action.", c->reg_v));
        if (c->aud_data) TIM_TRY(add_reg("reg_head.fc_audio_action.", c->reg_a));
    }
    c->layers.resize(c->L);
    const std::string enc = recog ? "transformer_encoder" : "backbone";
    const float qscale = static_cast<float>(std::pow(static_cast<double>(c->hd), -0.5) * 1.4426950408889634074);
    for (int l = 0; l < c->L; ++l) {
        Layer& ly = c->layers[l];
        const std::string p = enc + ".layers." + std::to_string(l) + ".";
        TIM_TRY(add_linear(c, p + "self_attn.", ly.in_proj, 3 * E, E, E, qscale, "in_proj_weight", "in_proj_bias"));
        TIM_TRY(add_linear(c, p + "self_attn.out_proj.", ly.out_proj, E, E));
        TIM_TRY(add_vec(c, p + "norm1.weight", &ly.n1g, {E}));
        TIM_TRY(add_vec(c, p + "norm1.bias", &ly.n1b, {E}));
        TIM_TRY(add_linear(c, p + "linear1.", ly.lin1, FF, E));
        TIM_TRY(add_linear(c, p + "linear2.", ly.lin2, E, FF));
        TIM_TRY(add_vec(c, p + "norm2.weight", &ly.n2g, {E}));
        TIM_TRY(add_vec(c, p + "norm2.bias", &ly.n2b, {E}));
    }
    // folded LayerNorm: needs every encoder GEMM on the CTA-pair kernel (shape conditions are independent of the batch)
    c->fold_ln = g.compute_dtype != TIM_FP32 && c->gemm_version >= 2 && c->L > 0 && E <= 2048 && umma2_supported(1, 3 * E, E) &&
                 umma2_supported(1, E, E) && umma2_supported(1, FF, E) && umma2_supported(1, E, FF);
    if (const char* fv = std::getenv("TIM_B200_FOLD")) if (std::atoi(fv) == 0) c->fold_ln = false;
    if (const char* pv = std::getenv("TIM_B200_PLANES")) c->planes = std::atoi(pv) != 0;
    if (c->fold_ln) {
        TIM_TRY(dev_alloc(c, reinterpret_cast<void**>(&c->fold_alarm_dev), sizeof(int)));
        if (cudaMemset(c->fold_alarm_dev, 0, sizeof(int)) != cudaSuccess ||
            cudaHostAlloc(reinterpret_cast<void**>(const_cast<int**>(&c->fold_alarm_host)), sizeof(int), cudaHostAllocDefault) != cudaSuccess)
            return c->fail(TIM_ERR_NOMEM, "fold alarm allocation failed");
        *c->fold_alarm_host = 0;
        auto make_foldable = [&](LinearW& w) -> int {
            w.foldable = true;
            TIM_TRY(dev_alloc(c, reinterpret_cast<void**>(&w.w32), static_cast<size_t>(w.N) * w.K * sizeof(float)));
            TIM_TRY(dev_alloc(c, reinterpret_cast<void**>(&w.b32), static_cast<size_t>(w.N) * sizeof(float)));
            TIM_TRY(dev_alloc(c, &w.wf, static_cast<size_t>(w.N) * w.K * c->esize));
            TIM_TRY(dev_alloc(c, reinterpret_cast<void**>(&w.cs), static_cast<size_t>(w.N) * sizeof(float)));
            TIM_TRY(dev_alloc(c, reinterpret_cast<void**>(&w.bwf), static_cast<size_t>(w.N) * sizeof(float)));
            return make_tmap_2d(c, &w.tmB2f, w.wf, op_dtype(c), 2, w.K, w.N, static_cast<long long>(w.K) * 2, 64, 128);
        };
        for (int l = 0; l < c->L; ++l) {
            if (l > 0) TIM_TRY(make_foldable(c->layers[l].in_proj));
            TIM_TRY(make_foldable(c->layers[l].lin1));
        }
    }
    return TIM_OK;
}

int ensure_ws(tim_ctx* c, size_t bytes) {
    if (bytes <= c->ws_bytes) return TIM_OK;
    if (c->ws) { cudaDeviceSynchronize(); cudaFree(c->ws); c->ws = nullptr; c->ws_bytes = 0; }
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&c->ws), bytes);
    if (e != cudaSuccess) return c->fail(TIM_ERR_NOMEM, "workspace cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    c->ws_bytes = bytes;
    return TIM_OK;
}

struct Arena {
    uint8_t* base; size_t off = 0;
    template <typename P> void take(P** p, size_t bytes) {
        *p = base ? reinterpret_cast<P*>(base + off) : nullptr;
        off += align_up(bytes ? bytes : 16, 256);
    }
};

// ------------------------------------------------------------------------------------------------------------------
// one linear layer through the selected compute path
// ------------------------------------------------------------------------------------------------------------------
// LayerNorm feeding a linear layer in the un-folded flow: A (16-bit, [rows, K]) = LayerNorm(in) (+ row statistics), a kernel of
// its own before the GEMM. (Running it as a prologue inside the CTA-pair GEMM was built and measured slower: DESIGN.md §5.)
struct LnBefore {
    const float* in;      // fp32 [rows, K]
    const float* gamma;
    const float* beta;
    float2* stats;        // [rows]
};

template <typename T>
int run_linear(tim_ctx* c, const void* A, int lda, const LinearW& w, RowMap rm, Epilogue ep, cudaStream_t s,
               const LnBefore* ln = nullptr) {
    if (lda != w.K) return c->fail(TIM_ERR_INVALID, "linear: lda %d != K %d", lda, w.K);
    if (rm.G <= 0 || rm.R <= 0) return TIM_OK;
    if (!ep.bias) ep.bias = w.bias;
    if (ln) {
        if constexpr (std::is_same<T, float>::value) {
            return c->fail(TIM_ERR_INVALID, "linear: LayerNorm prologue is a 16-bit path feature");
        } else {
            LAUNCH_C(c, 2, 0.0, s, launch_layernorm<T>(ln->in, w.K, ln->gamma, ln->beta, nullptr, 0, static_cast<T*>(const_cast<void*>(A)), w.K,
                                                             rm.R, w.K, s, ln->stats));
        }
    }
    if constexpr (std::is_same<T, float>::value) {
        if (ep.rstats) return c->fail(TIM_ERR_INVALID, "linear: LayerNorm-on-read residual is a 16-bit path feature");
        if (ep.out_act || ep.dact_of) return c->fail(TIM_ERR_INVALID, "linear: fused activation outputs are a 16-bit path feature");
        ep.out_fp32 = 1;
        LAUNCH_C(c, 0, 2.0 * rm.G * rm.R * w.N * w.K, s,
                 launch_linear_simt(static_cast<const float*>(A), lda, static_cast<const float*>(w.w), w.N, w.K, rm, ep, s));
    } else {
        const bool plain = rm.G == 1 && rm.box_g == 1 && rm.a_row_off == 0 && rm.out_row_off == 0;
        const int esz_out = ep.out_fp32 ? 4 : 2;
        if (c->gemm_version >= 2 && plain && w.has_tmB2 && umma2_supported(rm.R, w.N, w.K) && (ep.out_fp32 || !ep.resid) &&
            (static_cast<size_t>(ep.ldo) * esz_out) % 16 == 0 && (!ep.resid || (static_cast<size_t>(ep.ldr) * 4) % 16 == 0) &&
            (reinterpret_cast<uintptr_t>(ep.out) & 15) == 0 && (reinterpret_cast<uintptr_t>(ep.resid) & 15) == 0) {
            Umma2Params q;
            std::memset(&q, 0, sizeof(q));
            const int M = rm.R;
            TIM_TRY(make_tmap_2d(c, &q.tmA, A, op_dtype(c), 2, w.K, M, static_cast<long long>(lda) * 2, 64, 128));
            q.tmB = w.tmB2;
            if (ep.out_fp32)
                TIM_TRY(make_tmap_2d(c, &q.tmOut, ep.out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, w.N, M, static_cast<long long>(ep.ldo) * 4, 32, 32));
            else
                TIM_TRY(make_tmap_2d(c, &q.tmOut, ep.out, op_dtype(c), 2, w.N, M, static_cast<long long>(ep.ldo) * 2, 64, 32));
            if (ep.resid)
                TIM_TRY(make_tmap_2d(c, &q.tmRes, ep.resid, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, w.N, M, static_cast<long long>(ep.ldr) * 4, 32, 32));
            q.bias = ep.bias; q.M = M; q.N = w.N; q.K = w.K;
            q.rstats = ep.rstats; q.rgamma = ep.rgamma; q.rbeta = ep.rbeta;
            int mode = ep.out_fp32 ? (ep.resid ? (ep.rstats ? 3 : 2) : 1) : 0;
            if (ep.out_act || ep.dact_of) {         // training-leg fusions (modes 8 / 9): 16-bit output, no residual, one of the two
                if (mode != 0 || (ep.out_act && ep.dact_of) || (ep.out_act && ep.act != ACT_GELU) || (ep.dact_of && ep.act != ACT_NONE) ||
                    ((reinterpret_cast<uintptr_t>(ep.out_act) | reinterpret_cast<uintptr_t>(ep.dact_of)) & 15))
                    return c->fail(TIM_ERR_INVALID, "linear: unsupported combination for a fused activation output");
                if (ep.out_act) { TIM_TRY(make_tmap_2d(c, &q.tmOut16, ep.out_act, op_dtype(c), 2, w.N, M, static_cast<long long>(ep.ldo) * 2, 64, 32)); mode = 8; }
                else { TIM_TRY(make_tmap_2d(c, &q.tmRes, ep.dact_of, op_dtype(c), 2, w.N, M, static_cast<long long>(ep.ldo) * 2, 64, 32)); mode = 9; }
            }
            LAUNCH_C(c, 0, 2.0 * M * w.N * w.K, s, launch_linear_umma2<T>(q, mode, ep.act, c->num_sms, s));
            return TIM_OK;
        }
        if (ep.out_act || ep.dact_of) return c->fail(TIM_ERR_INVALID, "linear: fused activation outputs need the CTA-pair kernel (check fused_act_ok first)");
        UmmaParams p;
        std::memset(&p, 0, sizeof(p));
        TIM_TRY(make_tmap(c, &p.tmA, A, w.K, rm.a_group_rows, rm.G, rm.box_r, rm.box_g));
        p.tmB = w.tmB;
        p.N = w.N; p.K = w.K; p.rm = rm; p.ep = ep;
        LAUNCH_C(c, 0, 2.0 * rm.G * rm.R * w.N * w.K, s, launch_linear_umma<T>(p, w.block_n, c->num_sms, s));
    }
    return TIM_OK;
}

// attention over the two-stream qkv buffer: tcgen05 kernel where it applies (head_dim 64 / 128 / 192, Ft <= 128), else the
// warp-MMA kernel. `ap` caches the TMA maps for (qkv, out, B, Qt): they are identical for every layer of one forward.
template <typename T>
int prepare_attention(tim_ctx* c, AttnUmmaParams* ap, bool* use_umma, const T* qkv, T* out, int B, int Ft, int Qt) {
    *use_umma = c->attn_version >= 2 && attention_umma_supported(Ft, c->hd);
    if (!*use_umma) return TIM_OK;
    std::memset(ap, 0, sizeof(*ap));
    const long long E = c->E, ld = 3LL * c->E;
    const int Fp = (Ft + 15) & ~15;
    TIM_TRY(make_tmap_3d16(c, &ap->tmKV, qkv, ld, Ft, B, ld * 2, ld * 2 * Ft, Fp));
    TIM_TRY(make_tmap_3d16(c, &ap->tmQf, qkv, ld, Ft, B, ld * 2, ld * 2 * Ft, 128));
    TIM_TRY(make_tmap_3d16(c, &ap->tmOf, out, E, Ft, B, E * 2, E * 2 * Ft, 32));
    if (Qt > 0) {
        TIM_TRY(make_tmap_3d16(c, &ap->tmQq, qkv + static_cast<size_t>(B) * Ft * ld, ld, Qt, B, ld * 2, ld * 2 * Qt, 128));
        TIM_TRY(make_tmap_3d16(c, &ap->tmOq, out + static_cast<size_t>(B) * Ft * E, E, Qt, B, E * 2, E * 2 * Qt, 32));
    } else {
        ap->tmQq = ap->tmQf; ap->tmOq = ap->tmOf;
    }
    ap->qkv = qkv; ap->B = B; ap->Ft = Ft; ap->Qt = Qt; ap->H = c->H;
    return TIM_OK;
}

// tcgen05 attention forward: the decoupled-pipeline kernel (attention_umma4.cu) where it applies (head_dim 64 / 128 and the staging
// tile fits next to two stages), else attention_umma.cu; TIM_B200_ATTN=2 forces the latter
template <typename T>
cudaError_t launch_attention_tc(const tim_ctx* c, const AttnUmmaParams& ap, cudaStream_t s) {
    return c->attn_version >= 4 && attention_umma4_supported(ap.Ft, c->hd) ? launch_attention_umma4<T>(ap, c->hd, c->num_sms, s)
                                                                          : launch_attention_umma<T>(ap, c->hd, c->num_sms, s);
}

// whether run_linear sends a plain [M, w.N] 16-bit-output linear (row stride ldo elements) to the CTA-pair kernel: the precondition of the
// fused activation outputs (Epilogue::out_act / dact_of)
inline bool fused_act_ok(const tim_ctx* c, const LinearW& w, int M, int ldo) {
    return c->gemm_version >= 2 && w.has_tmB2 && umma2_supported(M, w.N, w.K) && (static_cast<size_t>(ldo) * 2) % 16 == 0;
}

inline Epilogue epi(void* out, int ldo, bool out_fp32, int act = ACT_NONE, const float* resid = nullptr, int ldr = 0) {
    Epilogue e;
    e.bias = nullptr; e.resid = resid; e.ldr = ldr; e.out = out; e.ldo = ldo; e.out_fp32 = out_fp32 ? 1 : 0; e.act = act;
    e.rstats = nullptr; e.rgamma = nullptr; e.rbeta = nullptr;
    e.out_act = nullptr; e.dact_of = nullptr;
    return e;
}

// rows of one query-token group [off, off+Q) inside each clip's Qt query rows
inline RowMap group_rows(int B, int Qt, int off, int Q) {
    RowMap rm;
    rm.G = B; rm.R = Q;
    if (Q <= 128) { rm.box_r = Q; rm.box_g = 128 / Q; if (rm.box_g > B) rm.box_g = B; if (rm.box_g > 256) rm.box_g = 256; }
    else { rm.box_r = 128; rm.box_g = 1; }
    rm.a_group_rows = Qt; rm.a_row_off = off;
    rm.out_group_rows = Q; rm.out_row_off = 0;
    return rm;
}

// (re)build the gamma-folded weight copies after any parameter changed (12 small kernels for 6 layers)
template <typename T>
int refold_weights(tim_ctx* c, cudaStream_t s) {
    for (int l = 0; l < c->L; ++l) {
        Layer& ly = c->layers[l];
        if (l > 0) {
            LinearW& w = ly.in_proj;
            LAUNCH(c, launch_fold_ln<T>(w.w32, w.b32, c->layers[l - 1].n2g, c->layers[l - 1].n2b, static_cast<T*>(w.wf), w.cs, w.bwf, w.N, w.K,
                                        w.scale_rows, w.scale, s));
        }
        LinearW& w1 = ly.lin1;
        LAUNCH(c, launch_fold_ln<T>(w1.w32, w1.b32, ly.n1g, ly.n1b, static_cast<T*>(w1.wf), w1.cs, w1.bwf, w1.N, w1.K, w1.scale_rows, w1.scale, s));
    }
    c->fold_dirty = false;
    return TIM_OK;
}

// One launch of the CTA-pair kernel in one of the folded-LayerNorm modes (gemm_umma2.cu):
//   mode 5 (producer): out32 [M,N] fp32 = A W^T + bias + R(resid), out16 = its 16-bit copy, opart = its partial row sums;
//                      R = LayerNorm-on-read with (rstats, rgamma, rbeta) or the identity when rstats == nullptr
//   mode 6 (consumer): out16 [M,N] = act(rstd * (A Wf^T - mean * cs) + bwf) with (mean, rstd) = rstats[row]
//   mode 7 (producer, two-plane residual stream): (out16, out_lo) = planes of  A W^T + bias + R(res_hi + res_lo); opart as mode 5
template <typename T>
int run_fold_gemm(tim_ctx* c, int mode, int act, const void* A, int M, int N, int K, const CUtensorMap& tmB, const float* bias,
                  float* out32, void* out16, const float* resid, const float2* rstats, const float* rgamma, const float* rbeta,
                  float2* opart, const float* cs, cudaStream_t s, const void* res_hi = nullptr, const void* res_lo = nullptr,
                  void* out_lo = nullptr) {
    Umma2Params q;
    std::memset(&q, 0, sizeof(q));
    TIM_TRY(make_tmap_2d(c, &q.tmA, A, op_dtype(c), 2, K, M, static_cast<long long>(K) * 2, 64, 128));
    q.tmB = tmB;
    if (mode == 7) {
        const long long pitch = static_cast<long long>(N) * 2;
        TIM_TRY(make_tmap_plane(c, &q.tmOut, out16, op_dtype(c), N, M, pitch));
        TIM_TRY(make_tmap_plane(c, &q.tmOutLo, out_lo, op_dtype(c), N, M, pitch));
        TIM_TRY(make_tmap_plane(c, &q.tmRes, res_hi, op_dtype(c), N, M, pitch));
        TIM_TRY(make_tmap_plane(c, &q.tmResLo, res_lo, op_dtype(c), N, M, pitch));
    } else if (mode == 5) {
        TIM_TRY(make_tmap_2d(c, &q.tmOut, out32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, N, M, static_cast<long long>(N) * 4, 32, 32));
        TIM_TRY(make_tmap_2d(c, &q.tmRes, resid, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, N, M, static_cast<long long>(N) * 4, 32, 32));
        TIM_TRY(make_tmap_2d(c, &q.tmOut16, out16, op_dtype(c), 2, N, M, static_cast<long long>(N) * 2, 64, 32));
    } else {
        TIM_TRY(make_tmap_2d(c, &q.tmOut, out16, op_dtype(c), 2, N, M, static_cast<long long>(N) * 2, 64, 32));
    }
    q.bias = bias; q.M = M; q.N = N; q.K = K;
    q.rstats = rstats; q.rgamma = rgamma; q.rbeta = rbeta; q.opart = opart; q.cs = cs;
    LAUNCH_C(c, mode == 6 ? 5 : 6, 2.0 * M * N * K, s, launch_linear_umma2<T>(q, mode, act, c->num_sms, s));
    return TIM_OK;
}
static_assert(sizeof(Umma2Params) <= 4000, "kernel parameter space");

template <typename T>
int time_mlp_impl(tim_ctx* c, const float* times, float* out, int B, int T_, cudaStream_t s, uint8_t* ws_base, size_t* ws_need) {
    const int M = B * T_, d = c->d;
    Arena a{ws_base};
    T *t1, *t2; float* t3;
    a.take(&t1, static_cast<size_t>(M) * d * sizeof(T));
    a.take(&t2, static_cast<size_t>(M) * d * sizeof(T));
    a.take(&t3, static_cast<size_t>(M) * d * sizeof(float));
    if (ws_need) { *ws_need = a.off; return TIM_OK; }
    constexpr bool f32 = std::is_same<T, float>::value;
    LAUNCH(c, launch_time_l1<T>(times, c->t0w, c->t0b, t1, M, d, s));
    TIM_TRY(run_linear<T>(c, t1, d, c->t2, plain_rows(M), epi(t2, d, f32, ACT_RELU), s));
    TIM_TRY(run_linear<T>(c, t2, d, c->t4, plain_rows(M), epi(t3, d, true, ACT_RELU), s));
    LAUNCH(c, launch_layernorm<T>(t3, d, c->tlg, c->tlb, out, d, static_cast<T*>(nullptr), 0, M, d, s));
    return TIM_OK;
}

struct QueryPlan {
    int Qv = 0, Qa = 0, Qt = 0;
    int n_groups = 0;
    TokenGroup groups[4];
    int off_verb = -1, off_noun = -1, off_action = -1, off_audio = -1;
};

int plan_queries(tim_ctx* c, int T_, int Qv, int Qa, QueryPlan* qp) {
    const tim_config& g = c->cfg;
    qp->Qv = c->vis_data ? Qv : 0;
    qp->Qa = c->aud_data ? Qa : 0;
    if (qp->Qv < 0 || qp->Qa < 0) return c->fail(TIM_ERR_INVALID, "negative query count");
    int off = 0;
    auto add = [&](const float* cls, const float* mod, int te_off, int count, int* where) {
        TokenGroup& tg = qp->groups[qp->n_groups++];
        tg.cls = cls; tg.mod = mod; tg.te_off = te_off; tg.count = count;
        *where = off; off += count;
    };
    const bool av = g.input_modality == TIM_AUDIO_VISUAL;
    if (qp->Qv > 0) {
        const int te_off = c->Ft;
        const float* mod = av ? c->mod_v : nullptr;
        if (c->vn_tokens) {
            add(c->cls_verb, mod, te_off, qp->Qv, &qp->off_verb);
            add(c->cls_noun, mod, te_off, qp->Qv, &qp->off_noun);
        }
        add(c->cls_action, mod, te_off, qp->Qv, &qp->off_action);
        if (te_off + qp->Qv > T_) return c->fail(TIM_ERR_INVALID, "time encodings too short: T=%d < %d", T_, te_off + qp->Qv);
    }
    if (qp->Qa > 0) {
        const int te_off = av ? T_ - qp->Qa : c->Ft;      // encodings.py:241 uses query_time_encoding[:, -num_a_queries:]
        if (te_off < c->Ft || te_off + qp->Qa > T_) return c->fail(TIM_ERR_INVALID, "time encodings too short for %d audio queries (T=%d)", qp->Qa, T_);
        add(c->cls_audio, av ? c->mod_a : nullptr, te_off, qp->Qa, &qp->off_audio);
    }
    qp->Qt = off;
    if (T_ < c->Ft) return c->fail(TIM_ERR_INVALID, "T=%d smaller than the %d feature tokens", T_, c->Ft);
    return TIM_OK;
}

template <typename T>
int encoder_impl(tim_ctx* c, const float* vis, const float* aud, const float* te, int B, int T_, int Qv, int Qa,
                 const tim_outputs* o, cudaStream_t s, uint8_t* ws_base, size_t* ws_need, const tim_feature_bank* fb = nullptr,
                 bool ws_indexed = false, bool in16 = false) {
    constexpr bool f32 = std::is_same<T, float>::value;
    const tim_config& g = c->cfg;
    const int d = c->d, E = c->E, FF = c->FF;
    QueryPlan qp;
    TIM_TRY(plan_queries(c, T_, Qv, Qa, &qp));
    const int Ft = c->Ft, Qt = qp.Qt;
    const size_t Mf = static_cast<size_t>(B) * Ft, Mq = static_cast<size_t>(B) * Qt, M = Mf + Mq;
    if (M > 0x7fffffffull) return c->fail(TIM_ERR_INVALID, "too many token rows (%zu)", M);

    Arena a{ws_base};
    float *x32, *z, *embv, *emba; T *x16, *qkv, *att, *hid, *vis16, *aud16, *r1, *r2;
    a.take(&x32, M * E * sizeof(float));
    a.take(&z, M * E * sizeof(float));
    a.take(&x16, f32 ? 0 : M * E * sizeof(T));
    a.take(&qkv, M * 3 * E * sizeof(T));
    a.take(&att, M * E * sizeof(T));
    a.take(&hid, M * FF * sizeof(T));
    a.take(&embv, static_cast<size_t>(B) * c->Fv * d * sizeof(float));
    a.take(&emba, static_cast<size_t>(B) * c->Fa * d * sizeof(float));
    // dense copies of the inputs in the compute dtype: the 16-bit cast of vis / aud, or (any dtype) the rows gathered from a bank
    const bool indexed = fb != nullptr || ws_indexed;
    a.take(&vis16, (f32 && !indexed) ? 0 : static_cast<size_t>(B) * c->Fv * g.vis_dim * sizeof(T));
    a.take(&aud16, (f32 && !indexed) ? 0 : static_cast<size_t>(B) * c->Fa * g.aud_dim * sizeof(T));
    const size_t regrows = static_cast<size_t>(B) * (qp.Qv > qp.Qa ? qp.Qv : qp.Qa);
    a.take(&r1, g.variant == TIM_DETECTION ? regrows * (E / 2) * sizeof(T) : 0);
    a.take(&r2, g.variant == TIM_DETECTION ? regrows * (E / 2) * sizeof(T) : 0);
    // two-plane residual stream of the folded flow: (x16, xlo) = tokens / z2, (zhi, zlo) = z1
    T *xlo, *zhi, *zlo;
    const bool planes_ws = !f32 && c->fold_ln && c->planes;
    a.take(&xlo, planes_ws ? M * E * sizeof(T) : 0);
    a.take(&zhi, planes_ws ? M * E * sizeof(T) : 0);
    a.take(&zlo, planes_ws ? M * E * sizeof(T) : 0);
    float2 *stats, *part;
    a.take(&stats, f32 ? 0 : M * sizeof(float2));
    const size_t nparts = 2 * static_cast<size_t>((E + 255) / 256);       // per-tile partial row sums of the folded-LayerNorm GEMMs
    a.take(&part, c->fold_ln ? nparts * M * sizeof(float2) : 0);
    if (ws_need) { *ws_need = a.off; return TIM_OK; }

    // ---- embedders: Linear -> GELU (epilogue) -> pre-LN rows; LN happens while assembling tokens ----
    const int Mv = B * c->Fv, Ma = B * c->Fa;
    if (c->Fv) {
        const void* A = vis;
        if (fb) {
            if (!fb->vis_bank || !fb->vis_rows) return c->fail(TIM_ERR_INVALID, "visual feature bank / row indices are NULL");
            LAUNCH(c, launch_gather_rows<T>(fb->vis_bank, fb->bank_dtype, fb->vis_bank_rows, reinterpret_cast<const long long*>(fb->vis_rows), vis16, Mv,
                                            g.vis_dim, s, c->idx_alarm_dev));
            A = vis16;
        } else {
            if (!vis) return c->fail(TIM_ERR_INVALID, "visual input is NULL");
            // in16: the caller's rows already are in the operand type (16-bit host feature bank of tim_forward_host_ex): no cast pass
            if constexpr (!f32) { if (!in16) { LAUNCH(c, launch_cast<T>(vis, vis16, Mv, g.vis_dim, 0, 1.0f, s)); A = vis16; } }
        }
        TIM_TRY(run_linear<T>(c, A, g.vis_dim, c->emb_v, plain_rows(Mv), epi(embv, d, true, ACT_GELU), s));
    }
    if (c->Fa) {
        const void* A = aud;
        if (fb) {
            if (!fb->aud_bank || !fb->aud_rows) return c->fail(TIM_ERR_INVALID, "audio feature bank / row indices are NULL");
            LAUNCH(c, launch_gather_rows<T>(fb->aud_bank, fb->bank_dtype, fb->aud_bank_rows, reinterpret_cast<const long long*>(fb->aud_rows), aud16, Ma,
                                            g.aud_dim, s, c->idx_alarm_dev));
            A = aud16;
        } else {
            if (!aud) return c->fail(TIM_ERR_INVALID, "audio input is NULL");
            if constexpr (!f32) { if (!in16) { LAUNCH(c, launch_cast<T>(aud, aud16, Ma, g.aud_dim, 0, 1.0f, s)); A = aud16; } }
        }
        TIM_TRY(run_linear<T>(c, A, g.aud_dim, c->emb_a, plain_rows(Ma), epi(emba, d, true, ACT_GELU), s));
    }
    // ---- token assembly into the two-stream buffer ----
    AssembleParams ap;
    std::memset(&ap, 0, sizeof(ap));
    ap.B = B; ap.d = d; ap.T = T_; ap.Fv = c->Fv; ap.Fa = c->Fa;
    ap.emb_v = embv; ap.emb_a = emba;
    ap.ln_v_g = c->lnv_g; ap.ln_v_b = c->lnv_b; ap.ln_a_g = c->lna_g; ap.ln_a_b = c->lna_b;
    ap.mod_v = g.input_modality == TIM_AUDIO_VISUAL ? c->mod_v : nullptr;
    ap.mod_a = g.input_modality == TIM_AUDIO_VISUAL ? c->mod_a : nullptr;
    ap.te = te; ap.n_groups = qp.n_groups;
    for (int i = 0; i < qp.n_groups; ++i) ap.groups[i] = qp.groups[i];
    ap.Qt = Qt; ap.x32 = x32; ap.x16 = f32 ? nullptr : x16;
    bool folded = false;
    if constexpr (!f32) {
        if (c->fold_ln && !c->fold_tripped && *c->fold_alarm_host) {
            c->fold_tripped = true;
            std::fprintf(stderr, "[tim_b200] rows with |mean| > 8 std seen in the residual stream: LayerNorm folding is switched off "
                                 "for this context (un-folded LayerNorm-on-read flow from now on)\n");
        }
        folded = c->fold_ln && !c->fold_tripped;
    }
    const bool planes = folded && planes_ws;
    ap.xlo = planes ? xlo : nullptr;
    if (planes && c->L > 0) ap.x32 = nullptr;       // the planes flow never reads the fp32 tokens: 4 of the 8 bytes per element stay unwritten
    LAUNCH_C(c, 3, 0.0, s, launch_assemble<T>(ap, s));

    // ---- encoder layers (post-LN): x = LN1(x + out_proj(attn(in_proj(x)))); x = LN2(x + W2 gelu(W1 x)) ----
    const int Mi = static_cast<int>(M);
    // algorithmic (mask-aware) attention FLOPs per layer: 4E(Ft^2 + Qt(Ft+1)) per clip
    const double attn_flops = 4.0 * E * (static_cast<double>(Ft) * Ft + static_cast<double>(Qt) * (Ft + 1)) * B;
    const void* xin = f32 ? static_cast<const void*>(x32) : static_cast<const void*>(x16);
    T* x16o = f32 ? nullptr : x16;
    AttnUmmaParams attn_p;
    bool attn_umma = false;
    if constexpr (!f32) TIM_TRY(prepare_attention<T>(c, &attn_p, &attn_umma, qkv, att, B, Ft, Qt));
    if constexpr (!f32) {
        if (folded && c->fold_dirty) TIM_TRY(refold_weights<T>(c, s));
    }
    for (int l = 0; l < c->L; ++l) {
        Layer& ly = c->layers[l];
        // ---- in_proj ----
        if constexpr (f32) {
            TIM_TRY(run_linear<T>(c, xin, E, ly.in_proj, plain_rows(Mi), epi(qkv, 3 * E, f32), s));
        } else if (folded) {
            // x16 holds the tokens (layer 0) or the 16-bit copy of the previous layer's pre-norm2 rows z2, whose LayerNorm
            // lives in this GEMM: gamma in the weight, (mean, rstd) and beta in the epilogue
            if (l == 0) TIM_TRY(run_linear<T>(c, xin, E, ly.in_proj, plain_rows(Mi), epi(qkv, 3 * E, false), s));
            else TIM_TRY(run_fold_gemm<T>(c, 6, ACT_NONE, x16, Mi, 3 * E, E, ly.in_proj.tmB2f, ly.in_proj.bwf, nullptr, qkv, nullptr, stats,
                                          nullptr, nullptr, nullptr, ly.in_proj.cs, s));
        } else {
            // un-folded 16-bit path: the fp32 LayerNorm output is never materialised either. x32 holds the tokens (layer 0)
            // or the pre-norm2 rows z2 of the previous layer, z the pre-norm1 rows z1; a LayerNorm kernel writes only the
            // 16-bit operand copy + (mean, rstd) per row, and the GEMM that needs LN(.) as its residual normalises on read.
            LnBefore lp{x32, l > 0 ? c->layers[l - 1].n2g : nullptr, l > 0 ? c->layers[l - 1].n2b : nullptr, stats};
            TIM_TRY(run_linear<T>(c, xin, E, ly.in_proj, plain_rows(Mi), epi(qkv, 3 * E, f32), s, l > 0 ? &lp : nullptr));
        }
        // ---- attention ----
        if constexpr (f32) {
            LAUNCH_C(c, 1, attn_flops, s, launch_attention_simt(reinterpret_cast<const float*>(qkv), reinterpret_cast<float*>(att), B, Ft, Qt, c->H, c->hd, s));
        } else if (attn_umma) {
            LAUNCH_C(c, 1, attn_flops, s, launch_attention_tc<T>(c, attn_p, s));
        } else {
            LAUNCH_C(c, 1, attn_flops, s, launch_attention_mma<T>(qkv, att, B, Ft, Qt, c->H, c->hd, s));
        }
        // ---- out_proj + residual (+ norm1), FFN + residual (+ norm2) ----
        if constexpr (f32) {
            TIM_TRY(run_linear<T>(c, att, E, ly.out_proj, plain_rows(Mi), epi(z, E, true, ACT_NONE, x32, E), s));
            LAUNCH_C(c, 2, 0.0, s, launch_layernorm<T>(z, E, ly.n1g, ly.n1b, x32, E, x16o, E, Mi, E, s));
            TIM_TRY(run_linear<T>(c, xin, E, ly.lin1, plain_rows(Mi), epi(hid, FF, f32, ACT_GELU), s));
            TIM_TRY(run_linear<T>(c, hid, FF, ly.lin2, plain_rows(Mi), epi(z, E, true, ACT_NONE, x32, E), s));
            LAUNCH_C(c, 2, 0.0, s, launch_layernorm<T>(z, E, ly.n2g, ly.n2b, x32, E, x16o, E, Mi, E, s));
        } else if (planes) {
            // the folded flow on the two-plane residual stream (mode 7): (x16, xlo) hold the tokens / z2 of the previous layer,
            // (zhi, zlo) receive z1; the hi planes are the A operands of the consumer GEMMs. Same statistics protocol as below.
            TIM_TRY(run_fold_gemm<T>(c, 7, ACT_NONE, att, Mi, E, E, ly.out_proj.tmB2, ly.out_proj.bias, nullptr, zhi, nullptr, l > 0 ? stats : nullptr,
                                     l > 0 ? c->layers[l - 1].n2g : nullptr, l > 0 ? c->layers[l - 1].n2b : nullptr, part, nullptr, s, x16, xlo, zlo));
            LAUNCH_C(c, 2, 0.0, s, launch_row_stats_finalize(part, static_cast<int>(nparts), E, stats, Mi, 64.0f, c->fold_alarm_dev, s));
            TIM_TRY(run_fold_gemm<T>(c, 6, ACT_GELU, zhi, Mi, FF, E, ly.lin1.tmB2f, ly.lin1.bwf, nullptr, hid, nullptr, stats, nullptr, nullptr,
                                     nullptr, ly.lin1.cs, s));
            TIM_TRY(run_fold_gemm<T>(c, 7, ACT_NONE, hid, Mi, E, FF, ly.lin2.tmB2, ly.lin2.bias, nullptr, x16, nullptr, stats, ly.n1g, ly.n1b, part, nullptr, s,
                                     zhi, zlo, xlo));
            if (l < c->L - 1) LAUNCH_C(c, 2, 0.0, s, launch_row_stats_finalize(part, static_cast<int>(nparts), E, stats, Mi, 64.0f, c->fold_alarm_dev, s));
            // the last layer's norm2 feeds the heads: over the query rows only, in place on the hi plane (each row is read whole first)
            if (l == c->L - 1 && Mq)
                LAUNCH_C(c, 2, 0.0, s, launch_layernorm_planes<T>(x16 + Mf * E, xlo + Mf * E, E, ly.n2g, ly.n2b, nullptr, 0, x16 + Mf * E, E,
                                                                  static_cast<int>(Mq), E, s));
        } else if (folded) {
            // `stats` holds (mean, rstd) of the rows whose LayerNorm is pending: z2 of the previous layer here ...
            // z1 = att Wo^T + bo + LN2_prev(z2_prev)   -> z (fp32), x16 (16-bit copy), partial sums
            TIM_TRY(run_fold_gemm<T>(c, 5, ACT_NONE, att, Mi, E, E, ly.out_proj.tmB2, ly.out_proj.bias, z, x16, x32, l > 0 ? stats : nullptr,
                                     l > 0 ? c->layers[l - 1].n2g : nullptr, l > 0 ? c->layers[l - 1].n2b : nullptr, part, nullptr, s));
            LAUNCH_C(c, 2, 0.0, s, launch_row_stats_finalize(part, static_cast<int>(nparts), E, stats, Mi, 64.0f, c->fold_alarm_dev, s));   // ... z1 from here on
            // hid = gelu(LN1(z1) W1^T + b1) with norm1 folded
            TIM_TRY(run_fold_gemm<T>(c, 6, ACT_GELU, x16, Mi, FF, E, ly.lin1.tmB2f, ly.lin1.bwf, nullptr, hid, nullptr, stats, nullptr, nullptr,
                                     nullptr, ly.lin1.cs, s));
            // z2 = hid W2^T + b2 + LN1(z1)             -> x32 (fp32), x16 (16-bit copy), partial sums
            TIM_TRY(run_fold_gemm<T>(c, 5, ACT_NONE, hid, Mi, E, FF, ly.lin2.tmB2, ly.lin2.bias, x32, x16, z, stats, ly.n1g, ly.n1b, part, nullptr, s));
            if (l < c->L - 1) LAUNCH_C(c, 2, 0.0, s, launch_row_stats_finalize(part, static_cast<int>(nparts), E, stats, Mi, 64.0f, c->fold_alarm_dev, s));   // z2
            // the last layer's norm2 feeds the heads: the only LayerNorm kernel of the stack, and only over the query rows
            // (the heads read nothing else; the feature rows' fp32 LayerNorm goes straight to `feats` below)
            if (l == c->L - 1 && Mq)
                LAUNCH_C(c, 2, 0.0, s, launch_layernorm<T>(x32 + Mf * E, E, ly.n2g, ly.n2b, nullptr, 0, x16o + Mf * E, E, static_cast<int>(Mq), E, s));
        } else {
            Epilogue e1 = epi(z, E, true, ACT_NONE, x32, E);
            if (l > 0) { e1.rstats = stats; e1.rgamma = c->layers[l - 1].n2g; e1.rbeta = c->layers[l - 1].n2b; }
            TIM_TRY(run_linear<T>(c, att, E, ly.out_proj, plain_rows(Mi), e1, s));
            LnBefore lp1{z, ly.n1g, ly.n1b, stats};
            TIM_TRY(run_linear<T>(c, xin, E, ly.lin1, plain_rows(Mi), epi(hid, FF, f32, ACT_GELU), s, &lp1));
            Epilogue e2 = epi(x32, E, true, ACT_NONE, z, E);
            e2.rstats = stats; e2.rgamma = ly.n1g; e2.rbeta = ly.n1b;
            TIM_TRY(run_linear<T>(c, hid, FF, ly.lin2, plain_rows(Mi), e2, s));
            // the last layer's norm2 feeds the heads: a kernel of its own, over the query rows only
            if (l == c->L - 1 && Mq)
                LAUNCH_C(c, 2, 0.0, s, launch_layernorm<T>(x32 + Mf * E, E, ly.n2g, ly.n2b, nullptr, 0, x16o + Mf * E, E, static_cast<int>(Mq), E, s));
        }
    }
    // fp32 features returned to the caller (tim.py:172, x[:, :num_feats]): LayerNorm of the last layer's z2 feature rows
    const float* feats_src = x32;
    bool feats_done = false;
    if constexpr (!f32) {
        if (c->L > 0 && o->feats && Mf) {
            Layer& ly = c->layers[c->L - 1];
            if (planes) LAUNCH_C(c, 2, 0.0, s, launch_layernorm_planes<T>(x16, xlo, E, ly.n2g, ly.n2b, o->feats, E, static_cast<T*>(nullptr), 0, static_cast<int>(Mf), E, s));
            else LAUNCH_C(c, 2, 0.0, s, launch_layernorm<T>(x32, E, ly.n2g, ly.n2b, o->feats, E, static_cast<T*>(nullptr), 0, static_cast<int>(Mf), E, s));
            feats_done = true;
        }
    }

    // ---- heads read slices of the query stream ----
    const uint8_t* qbase = reinterpret_cast<const uint8_t*>(xin) + Mf * E * sizeof(T);
    auto cls = [&](const LinearW& w, int off, int Q, float* out) -> int {
        if (!w.N || Q <= 0) return TIM_OK;
        if (!out) return c->fail(TIM_ERR_INVALID, "output pointer for a %d-class head is NULL", w.N);
        return run_linear<T>(c, qbase, E, w, group_rows(B, Qt, off, Q), epi(out, w.N, true), s);
    };
    auto reg = [&](RegHead& r, int off, int Q, float* out) -> int {
        if (Q <= 0) return TIM_OK;
        if (!out) return c->fail(TIM_ERR_INVALID, "regression output pointer is NULL");
        RowMap rm = group_rows(B, Qt, off, Q);
        TIM_TRY(run_linear<T>(c, qbase, E, r.l0, rm, epi(r1, E / 2, f32, ACT_RELU), s));
        TIM_TRY(run_linear<T>(c, r1, E / 2, r.l2, plain_rows(B * Q), epi(r2, E / 2, f32, ACT_RELU), s));
        LAUNCH(c, launch_reg_final<T>(r2, E / 2, r.w4, r.b4, out, B * Q, E / 2, s));
        return TIM_OK;
    };
    if (g.variant == TIM_RECOGNITION) {
        if (qp.Qv > 0) {
            if (g.n_verb && qp.off_verb >= 0) TIM_TRY(cls(c->h_verb, qp.off_verb, qp.Qv, o->verb));
            if (g.n_noun && qp.off_noun >= 0) TIM_TRY(cls(c->h_noun, qp.off_noun, qp.Qv, o->noun));
            TIM_TRY(cls(c->h_action, qp.off_action, qp.Qv, o->action));
        }
        if (qp.Qa > 0) TIM_TRY(cls(c->h_audio, qp.off_audio, qp.Qa, o->audio));
    } else {
        if (qp.Qv > 0) {
            TIM_TRY(cls(c->h_verb, qp.off_action, qp.Qv, o->verb));
            TIM_TRY(cls(c->h_noun, qp.off_action, qp.Qv, o->noun));
            TIM_TRY(cls(c->h_action, qp.off_action, qp.Qv, o->action));
            TIM_TRY(reg(c->reg_v, qp.off_action, qp.Qv, o->reg_visual));
        }
        if (qp.Qa > 0) {
            TIM_TRY(cls(c->h_audio, qp.off_audio, qp.Qa, o->audio));
            TIM_TRY(reg(c->reg_a, qp.off_audio, qp.Qa, o->reg_audio));
        }
    }
    if (o->feats && Mf && !feats_done) CU_OK(c, cudaMemcpyAsync(o->feats, feats_src, Mf * E * sizeof(float), cudaMemcpyDeviceToDevice, s));
    c->last_folded = folded;
    if (folded) {
        CU_OK(c, cudaMemcpyAsync(const_cast<int*>(c->fold_alarm_host), c->fold_alarm_dev, sizeof(int), cudaMemcpyDeviceToHost, s));
        if (!c->fold_event) CU_OK(c, cudaEventCreateWithFlags(&c->fold_event, cudaEventDisableTiming));
        CU_OK(c, cudaEventRecord(c->fold_event, s));
    }
    return TIM_OK;
}

int check_ready(tim_ctx* c) {
    for (auto& kv : c->slots)
        if (!kv.second.set) return c->fail(TIM_ERR_WEIGHTS, "weight '%s' has not been set", kv.first.c_str());
    return TIM_OK;
}

template <typename F>
int dispatch_dtype(tim_ctx* c, F&& f) {
    switch (c->cfg.compute_dtype) {
        case TIM_FP32: return f(float{});
        case TIM_BF16: return f(__nv_bfloat16{});
        case TIM_FP16: return f(__half{});
    }
    return c->fail(TIM_ERR_INVALID, "bad compute_dtype %d", c->cfg.compute_dtype);
}

int time_mlp_ws(tim_ctx* c, int B, int T_, size_t* need) {
    return dispatch_dtype(c, [&](auto tag) { return time_mlp_impl<decltype(tag)>(c, nullptr, nullptr, B, T_, nullptr, nullptr, need); });
}
int encoder_ws(tim_ctx* c, int B, int T_, int Qv, int Qa, size_t* need, bool indexed = false) {
    tim_outputs o; std::memset(&o, 0, sizeof(o));
    return dispatch_dtype(c, [&](auto tag) {
        return encoder_impl<decltype(tag)>(c, nullptr, nullptr, nullptr, B, T_, Qv, Qa, &o, nullptr, nullptr, need, nullptr, indexed);
    });
}

}  // namespace

#include "train.inl"

// ======================================================================================================================
// extern "C"
// ======================================================================================================================
extern "C" {

int tim_abi_version(void) { return TIM_ABI_VERSION; }

const char* tim_last_error(const tim_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int tim_seq_len(const tim_config* g, int Qv, int Qa) {
    if (!g) return -1;
    const int Ft = g->num_feats * (g->input_modality == TIM_AUDIO_VISUAL ? 2 : 1);
    const bool vis = g->data_modality != TIM_AUDIO, aud = g->data_modality != TIM_VISUAL;
    const bool vn = g->variant == TIM_RECOGNITION && g->include_verb_noun && vis;
    int q = 0;
    if (vis && Qv > 0) q += Qv * (vn ? 3 : 1);
    if (aud && Qa > 0) q += Qa;
    return Ft + q;
}

int tim_create(tim_ctx** out, const tim_config* cfg, int device) {
    if (!out || !cfg) { g_create_error = "tim_create: NULL argument"; return TIM_ERR_INVALID; }
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0) {
        g_create_error = std::string("tim_create: no CUDA device (") + cudaGetErrorString(e) + "); libtim_b200 has no CPU fallback";
        return TIM_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= ndev) { g_create_error = "tim_create: bad device index"; return TIM_ERR_INVALID; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) {
        g_create_error = "tim_create: device is not sm_100 (B200); libtim_b200 is built for sm_100a only and has no fallback";
        return TIM_ERR_NO_DEVICE;
    }
    tim_ctx* c = new tim_ctx();
    auto bail = [&](int code) { g_create_error = c->err; tim_destroy(c); return code; };
    c->cfg = *cfg;
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    const tim_config& g = c->cfg;
    if (g.d_model <= 0 || g.nhead <= 0 || g.num_layers < 0 || g.ff_dim <= 0 || g.num_feats <= 0 || (2 * g.d_model) % g.nhead)
        return bail(c->fail(TIM_ERR_INVALID, "bad model dims"));
    if (g.compute_dtype < TIM_FP32 || g.compute_dtype > TIM_FP16) return bail(c->fail(TIM_ERR_INVALID, "bad compute_dtype"));
    if (g.input_modality != TIM_AUDIO_VISUAL && g.data_modality != g.input_modality)
        return bail(c->fail(TIM_ERR_INVALID, "uni-modal input requires data_modality == input_modality"));
    c->d = g.d_model; c->E = 2 * g.d_model; c->FF = g.ff_dim; c->H = g.nhead; c->hd = c->E / g.nhead; c->L = g.num_layers;
    c->F = g.num_feats;
    c->Fv = g.input_modality != TIM_AUDIO ? g.num_feats : 0;
    c->Fa = g.input_modality != TIM_VISUAL ? g.num_feats : 0;
    c->Ft = c->Fv + c->Fa;
    c->vis_data = g.data_modality != TIM_AUDIO;
    c->aud_data = g.data_modality != TIM_VISUAL;
    c->vn_tokens = g.variant == TIM_RECOGNITION && g.include_verb_noun && c->vis_data;
    c->esize = g.compute_dtype == TIM_FP32 ? 4 : 2;
    if (const char* gv = std::getenv("TIM_B200_GEMM")) c->gemm_version = std::atoi(gv) == 1 ? 1 : 2;
    if (const char* av = std::getenv("TIM_B200_ATTN")) { const int v = std::atoi(av); c->attn_version = v < 1 ? 1 : (v > 4 ? 4 : v); }
    if (const char* wv = std::getenv("TIM_B200_WGRAD_SPLITS")) c->wgrad_splits = std::atoi(wv);
    if (const char* tf = std::getenv("TIM_B200_TRAIN_FUSE")) c->train_fuse = std::atoi(tf) != 0;
    if (c->d % 4) return bail(c->fail(TIM_ERR_INVALID, "d_model must be a multiple of 4"));
    if (c->vis_data && !g.n_action) return bail(c->fail(TIM_ERR_INVALID, "visual data modality needs n_action > 0"));
    if (c->aud_data && !g.n_audio) return bail(c->fail(TIM_ERR_INVALID, "audio data modality needs n_audio > 0"));
    if (c->vn_tokens && (!g.n_verb || !g.n_noun)) return bail(c->fail(TIM_ERR_INVALID, "include_verb_noun needs n_verb and n_noun"));
    if (g.compute_dtype != TIM_FP32) {
        const int hd = c->hd;
        if (!(hd == 16 || hd == 32 || hd == 64 || hd == 128 || hd == 192))
            return bail(c->fail(TIM_ERR_INVALID, "16-bit attention supports head_dim in {16,32,64,128,192}, got %d", hd));
        if (c->Ft > 128) return bail(c->fail(TIM_ERR_INVALID, "16-bit attention supports at most 128 feature tokens per clip, got %d", c->Ft));
        if (c->d % 8 || g.vis_dim % 8 || g.aud_dim % 8 || c->FF % 8)
            return bail(c->fail(TIM_ERR_INVALID, "tcgen05 path needs all reduction dims to be multiples of 8"));
    } else if (attention_simt_smem(c->Ft, c->hd) > 227 * 1024) {
        return bail(c->fail(TIM_ERR_INVALID, "fp32 attention: %d feature tokens x head_dim %d exceed shared memory", c->Ft, c->hd));
    }
    if (cudaSetDevice(device) != cudaSuccess) return bail(c->fail(TIM_ERR_CUDA, "cudaSetDevice failed"));
    if (g.compute_dtype != TIM_FP32) { int r = get_encode_fn(c); if (r) return bail(r); }
    int r = build_weights(c);
    if (r) return bail(r);
    *out = c;
    return TIM_OK;
}

void tim_destroy(tim_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (void* p : c->allocs) cudaFree(p);
    if (c->ws) cudaFree(c->ws);
    if (c->fold_alarm_host) cudaFreeHost(const_cast<int*>(c->fold_alarm_host));
    if (c->idx_alarm_host) cudaFreeHost(const_cast<int*>(c->idx_alarm_host));
    if (c->idx_event) cudaEventDestroy(c->idx_event);
    if (c->fold_event) cudaEventDestroy(c->fold_event);
    for (cudaEvent_t e : c->host_events) cudaEventDestroy(e);
    if (c->train) {
        if (c->train->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->train->nccl_comm);
        if (c->train->mem) cudaFree(c->train->mem);
        if (c->train->tmem) cudaFree(c->train->tmem);
        delete c->train;
    }
    if (c->io) cudaFree(c->io);
    for (auto& r : c->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto e : c->ev_pool) cudaEventDestroy(e);
    if (c->s_h2d) cudaStreamDestroy(c->s_h2d);
    if (c->s_comp) cudaStreamDestroy(c->s_comp);
    if (c->s_d2h) cudaStreamDestroy(c->s_d2h);
    delete c;
}

int tim_set_weight(tim_ctx* c, const char* key, const float* data, const int64_t* shape, int ndim, void* stream) {
    if (!c || !key) return TIM_ERR_INVALID;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (!std::strncmp(key, "drloc_mlp.", 10) || !std::strncmp(key, "pool.", 5)) return TIM_OK;   // off the hot path
    auto it = c->slots.find(key);
    if (it == c->slots.end()) return c->fail(TIM_ERR_WEIGHTS, "unknown state_dict key '%s' for this configuration", key);
    Slot& sl = it->second;
    if (!data || !shape) return c->fail(TIM_ERR_INVALID, "NULL data/shape for '%s'", key);
    bool ok = static_cast<size_t>(ndim) == sl.shape.size();
    for (int i = 0; ok && i < ndim; ++i) ok = shape[i] == sl.shape[i];
    if (!ok) return c->fail(TIM_ERR_WEIGHTS, "shape mismatch for '%s'", key);
    CU_OK(c, cudaSetDevice(c->device));
    if (sl.kind == Slot::VEC) {
        size_t n = 1;
        for (auto v : sl.shape) n *= static_cast<size_t>(v);
        CU_OK(c, cudaMemcpyAsync(sl.vec, data, n * sizeof(float), cudaMemcpyDeviceToDevice, s));
    } else if (sl.kind == Slot::LIN_B) {
        LinearW& w = *sl.lin;
        if (w.scale_rows % 4) return c->fail(TIM_ERR_INVALID, "q rows (%d) must be a multiple of 4", w.scale_rows);
        LAUNCH(c, launch_scale_copy(data, w.bias, w.N, 1, w.scale_rows, w.scale, s));
        if (w.foldable) CU_OK(c, cudaMemcpyAsync(w.b32, data, static_cast<size_t>(w.N) * sizeof(float), cudaMemcpyDeviceToDevice, s));
    } else {
        LinearW& w = *sl.lin;
        if ((static_cast<size_t>(w.scale_rows) * w.K) % 4) return c->fail(TIM_ERR_INVALID, "scaled prefix not float4 aligned");
        switch (c->cfg.compute_dtype) {
            case TIM_FP32: LAUNCH(c, launch_scale_copy(data, static_cast<float*>(w.w), w.N, w.K, w.scale_rows, w.scale, s)); break;
            case TIM_BF16: LAUNCH(c, launch_cast<__nv_bfloat16>(data, static_cast<__nv_bfloat16*>(w.w), w.N, w.K, w.scale_rows, w.scale, s)); break;
            case TIM_FP16: LAUNCH(c, launch_cast<__half>(data, static_cast<__half*>(w.w), w.N, w.K, w.scale_rows, w.scale, s)); break;
        }
        if (w.foldable) CU_OK(c, cudaMemcpyAsync(w.w32, data, static_cast<size_t>(w.N) * w.K * sizeof(float), cudaMemcpyDeviceToDevice, s));
        if (w.wt) {
            switch (c->cfg.compute_dtype) {
                case TIM_FP32: LAUNCH(c, launch_transpose_pack<float>(data, static_cast<float*>(w.wt), w.N, w.K, w.Np, s)); break;
                case TIM_BF16: LAUNCH(c, launch_transpose_pack<__nv_bfloat16>(data, static_cast<__nv_bfloat16*>(w.wt), w.N, w.K, w.Np, s)); break;
                case TIM_FP16: LAUNCH(c, launch_transpose_pack<__half>(data, static_cast<__half*>(w.wt), w.N, w.K, w.Np, s)); break;
            }
            w.wt_set = true;
        }
    }
    c->fold_dirty = true;
    sl.set = true;
    return TIM_OK;
}

int tim_weights_missing(const tim_ctx* c, char* buf, size_t buflen) {
    if (!c) return TIM_ERR_INVALID;
    int n = 0;
    std::string names;
    for (auto& kv : c->slots)
        if (!kv.second.set) { ++n; names += kv.first; names += '\n'; }
    if (buf && buflen) { std::strncpy(buf, names.c_str(), buflen - 1); buf[buflen - 1] = 0; }
    return n;
}

int tim_time_mlp_fwd(tim_ctx* c, const float* times, float* out, int B, int T_, void* stream) {
    if (!c) return TIM_ERR_INVALID;
    if (!times || !out || B <= 0 || T_ <= 0) return c->fail(TIM_ERR_INVALID, "tim_time_mlp_fwd: bad arguments");
    TIM_TRY(check_ready(c));
    CU_OK(c, cudaSetDevice(c->device));
    size_t need = 0;
    TIM_TRY(time_mlp_ws(c, B, T_, &need));
    TIM_TRY(ensure_ws(c, need));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    return dispatch_dtype(c, [&](auto tag) { return time_mlp_impl<decltype(tag)>(c, times, out, B, T_, s, c->ws, nullptr); });
}

int tim_encoder_fwd(tim_ctx* c, const float* vis, const float* aud, const float* te, int B, int T_, int Qv, int Qa,
                    const tim_outputs* outs, void* stream) {
    if (!c) return TIM_ERR_INVALID;
    if (!te || !outs || B <= 0 || T_ <= 0) return c->fail(TIM_ERR_INVALID, "tim_encoder_fwd: bad arguments");
    TIM_TRY(check_ready(c));
    CU_OK(c, cudaSetDevice(c->device));
    size_t need = 0;
    TIM_TRY(encoder_ws(c, B, T_, Qv, Qa, &need));
    TIM_TRY(ensure_ws(c, need));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    return dispatch_dtype(c, [&](auto tag) { return encoder_impl<decltype(tag)>(c, vis, aud, te, B, T_, Qv, Qa, outs, s, c->ws, nullptr); });
}

int tim_encoder_fwd_indexed(tim_ctx* c, const tim_feature_bank* fb, const float* te, int B, int T_, int Qv, int Qa,
                            const tim_outputs* outs, void* stream) {
    if (!c) return TIM_ERR_INVALID;
    if (!fb || !te || !outs || B <= 0 || T_ <= 0) return c->fail(TIM_ERR_INVALID, "tim_encoder_fwd_indexed: bad arguments");
    if (fb->bank_dtype < TIM_FP32 || fb->bank_dtype > TIM_FP16) return c->fail(TIM_ERR_INVALID, "tim_encoder_fwd_indexed: bad bank_dtype");
    TIM_TRY(check_ready(c));
    CU_OK(c, cudaSetDevice(c->device));
    // a row index outside the bank (the reference's host-side indexing would raise IndexError): counted by the gather kernel, the count
    // reaches pinned host memory behind this forward. The verdict is asynchronous like everything on this entry point: a later call on
    // the context fails if an earlier forward saw bad indices, tim_index_check() blocks for the verdict on the last one.
    if (!c->idx_alarm_dev) {
        TIM_TRY(dev_alloc(c, reinterpret_cast<void**>(&c->idx_alarm_dev), sizeof(int)));
        if (cudaMemset(c->idx_alarm_dev, 0, sizeof(int)) != cudaSuccess ||
            cudaHostAlloc(reinterpret_cast<void**>(const_cast<int**>(&c->idx_alarm_host)), sizeof(int), cudaHostAllocDefault) != cudaSuccess ||
            cudaEventCreateWithFlags(&c->idx_event, cudaEventDisableTiming) != cudaSuccess)
            return c->fail(TIM_ERR_NOMEM, "index alarm allocation failed");
        *c->idx_alarm_host = 0;
    }
    if (*c->idx_alarm_host) {
        const int n = *c->idx_alarm_host;
        *c->idx_alarm_host = 0;
        cudaMemsetAsync(c->idx_alarm_dev, 0, sizeof(int), static_cast<cudaStream_t>(stream));
        return c->fail(TIM_ERR_INVALID, "tim_encoder_fwd_indexed: %d feature-bank row indices of an earlier call were out of range (those rows were read as zeros)", n);
    }
    size_t need = 0;
    TIM_TRY(encoder_ws(c, B, T_, Qv, Qa, &need, true));
    TIM_TRY(ensure_ws(c, need));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int r = dispatch_dtype(c, [&](auto tag) {
        return encoder_impl<decltype(tag)>(c, nullptr, nullptr, te, B, T_, Qv, Qa, outs, s, c->ws, nullptr, fb, true);
    });
    if (r != TIM_OK) return r;
    CU_OK(c, cudaMemcpyAsync(const_cast<int*>(c->idx_alarm_host), c->idx_alarm_dev, sizeof(int), cudaMemcpyDeviceToHost, s));
    CU_OK(c, cudaEventRecord(c->idx_event, s));
    return TIM_OK;
}

// Blocks until the index check of the last tim_encoder_fwd_indexed is known: TIM_OK, or TIM_ERR_INVALID (tim_last_error says how many row
// indices were outside their bank; the outputs of that forward are to be discarded). Clears the count.
int tim_index_check(tim_ctx* c) {
    if (!c) return TIM_ERR_INVALID;
    if (!c->idx_event) return TIM_OK;
    CU_OK(c, cudaSetDevice(c->device));
    CU_OK(c, cudaEventSynchronize(c->idx_event));
    const int n = *c->idx_alarm_host;
    if (!n) return TIM_OK;
    *c->idx_alarm_host = 0;
    CU_OK(c, cudaMemset(c->idx_alarm_dev, 0, sizeof(int)));
    return c->fail(TIM_ERR_INVALID, "tim_encoder_fwd_indexed: %d feature-bank row indices were out of range (those rows were read as zeros)", n);
}

// Chunk sizes of tim_forward_host (pure host arithmetic; tim_host_chunk_schedule exposes it to the CPU tests). cpc = target clips
// per chunk, 1 <= cpc <= B. Every chunk is <= cpc and > 0, the sizes add up to B.
static std::vector<int> chunk_schedule(int B, int cpc, int rows_per_clip, int E, int num_sms, bool sixteen_bit) {
    int unit_tiles = 0;
    if (sixteen_bit && E > 0 && E % 256 == 0 && rows_per_clip > 0 && num_sms > 0) {
        int a = num_sms, b = E / 256;
        while (b) { const int t = a % b; a = b; b = t; }
        unit_tiles = num_sms / a;
    }
    auto unit_clips = [&](int k) { return static_cast<int>(static_cast<long long>(k) * unit_tiles * 256 / rows_per_clip); };
    std::vector<int> head, tail, out;
    int left = B, body = cpc;
    // the largest k with unit_clips(k) <= cpc (unit_clips(k) can exceed k * unit_clips(1): floor(k * u) >= k * floor(u))
    const int body_units = (unit_tiles > 0 && unit_clips(1) >= 8)
        ? static_cast<int>(((static_cast<long long>(cpc) + 1) * rows_per_clip - 1) / (static_cast<long long>(unit_tiles) * 256)) : 0;
    if (body_units >= 1) {
        body = unit_clips(body_units);
        if (B >= 3 * body && body_units >= 4) {
            head = {unit_clips(body_units / 4), unit_clips(body_units / 2)};
            tail = {head[1], head[0]};
        }
    } else if (B >= 4 * cpc && cpc >= 8) {
        head = {cpc / 4, cpc / 2};
        tail = {cpc / 2, cpc / 4};
    }
    for (int n : head) left -= 2 * n;
    for (int n : head) out.push_back(n);
    // the remainder that does not fill a body chunk goes early (behind the head), not into the D2H tail
    if (left % body) { out.push_back(left % body); left -= left % body; }
    while (left > 0) { out.push_back(body); left -= body; }
    for (int n : tail) out.push_back(n);
    return out;
}

int tim_host_chunk_schedule(int B, int clips_per_chunk, int rows_per_clip, int E, int num_sms, int sixteen_bit, int* out, int max_out) {
    if (B <= 0 || rows_per_clip <= 0 || E <= 0 || num_sms <= 0 || (max_out > 0 && !out)) return TIM_ERR_INVALID;
    int cpc = clips_per_chunk;
    if (cpc <= 0 || cpc > B) cpc = B;
    const std::vector<int> v = chunk_schedule(B, cpc, rows_per_clip, E, num_sms, sixteen_bit != 0);
    for (size_t i = 0; i < v.size() && static_cast<int>(i) < max_out; ++i) out[i] = v[i];
    return static_cast<int>(v.size());
}

int tim_forward_host_ex(tim_ctx* c, const void* vis, const void* aud, const float* times, int B, int T_, int Qv, int Qa,
                        const tim_outputs* ho, int cpc, int in_dtype, int out_dtype, void* stream, uint64_t* h2d_bytes, uint64_t* d2h_bytes) {
    if (!c) return TIM_ERR_INVALID;
    if (!times || !ho || B <= 0 || T_ <= 0) return c->fail(TIM_ERR_INVALID, "tim_forward_host: bad arguments");
    const tim_config& g = c->cfg;
    const bool in16 = in_dtype != TIM_FP32;
    if (in16 && in_dtype != g.compute_dtype) return c->fail(TIM_ERR_INVALID, "tim_forward_host_ex: 16-bit host features must be in the context's compute dtype");
    if (out_dtype != TIM_FP32 && out_dtype != TIM_FP16) return c->fail(TIM_ERR_INVALID, "tim_forward_host_ex: out_dtype must be TIM_FP32 or TIM_FP16");
    const bool out16 = out_dtype == TIM_FP16;
    TIM_TRY(check_ready(c));
    CU_OK(c, cudaSetDevice(c->device));
    if (cpc <= 0 || cpc > B) cpc = B;
    QueryPlan qp;
    TIM_TRY(plan_queries(c, T_, Qv, Qa, &qp));
    if (!c->s_h2d) {
        CU_OK(c, cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
        CU_OK(c, cudaStreamCreateWithFlags(&c->s_comp, cudaStreamNonBlocking));
        CU_OK(c, cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
    }
    // per-clip sizes (elements)
    const size_t isz = in16 ? 2 : 4, osz = out16 ? 2 : 4;
    const size_t n_vis = static_cast<size_t>(c->Fv) * g.vis_dim, n_aud = static_cast<size_t>(c->Fa) * g.aud_dim;
    const size_t n_times = static_cast<size_t>(T_) * 2, n_te = static_cast<size_t>(T_) * c->d;
    const size_t n_verb = ho->verb ? static_cast<size_t>(qp.Qv) * g.n_verb : 0, n_noun = ho->noun ? static_cast<size_t>(qp.Qv) * g.n_noun : 0;
    const size_t n_act = ho->action ? static_cast<size_t>(qp.Qv) * g.n_action : 0, n_au = ho->audio ? static_cast<size_t>(qp.Qa) * g.n_audio : 0;
    const size_t n_rv = ho->reg_visual ? static_cast<size_t>(qp.Qv) * 2 : 0, n_ra = ho->reg_audio ? static_cast<size_t>(qp.Qa) * 2 : 0;
    const size_t n_feats = ho->feats ? static_cast<size_t>(c->Ft) * c->E : 0;
    const size_t n_out[7] = {n_verb, n_noun, n_act, n_au, n_rv, n_ra, n_feats};
    // two staging sets (double buffering across chunks)
    struct Set { uint8_t *vis, *aud; float *times, *te; float* out[7]; __half* out16[7]; };
    Set st[2];
    for (int pass = 0; pass < 2; ++pass) {
        Arena a{pass ? c->io : nullptr};
        for (int k = 0; k < 2; ++k) {
            a.take(&st[k].vis, cpc * n_vis * isz); a.take(&st[k].aud, cpc * n_aud * isz); a.take(&st[k].times, cpc * n_times * 4);
            a.take(&st[k].te, cpc * n_te * 4);
            for (int j = 0; j < 7; ++j) { a.take(&st[k].out[j], cpc * n_out[j] * 4); a.take(&st[k].out16[j], out16 ? cpc * n_out[j] * 2 : 0); }
        }
        if (!pass && a.off > c->io_bytes) {
            if (c->io) { cudaDeviceSynchronize(); cudaFree(c->io); c->io = nullptr; c->io_bytes = 0; }
            cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&c->io), a.off);
            if (e != cudaSuccess) return c->fail(TIM_ERR_NOMEM, "staging cudaMalloc(%zu) failed: %s", a.off, cudaGetErrorString(e));
            c->io_bytes = a.off;
        }
    }
    size_t need_t = 0, need_e = 0;
    TIM_TRY(time_mlp_ws(c, cpc, T_, &need_t));
    TIM_TRY(encoder_ws(c, cpc, T_, Qv, Qa, &need_e));
    TIM_TRY(ensure_ws(c, need_t > need_e ? need_t : need_e));

    // chunk schedule: full chunks of about cpc clips with tapered ends (a quarter, a half ... a half, a quarter) so that the
    // un-overlapped H2D of the first chunk and D2H of the last chunk are short. On the 16-bit path chunk sizes are aligned to
    // whole WAVES of GEMM tiles (see chunk_schedule).
    std::vector<int> chunk_b0, chunk_nb = chunk_schedule(B, cpc, c->Ft + qp.Qt, c->E, c->num_sms, g.compute_dtype != TIM_FP32);
    {
        int b0 = 0;
        for (int n : chunk_nb) { chunk_b0.push_back(b0); b0 += n; }
    }
    const int nchunks = static_cast<int>(chunk_nb.size());
    // events live in the context and are reused call after call (3 per chunk + 1 for the caller's stream)
    while (c->host_events.size() < 3 * static_cast<size_t>(nchunks) + 1) {
        cudaEvent_t e = nullptr;
        CU_OK(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->host_events.push_back(e);
    }
    cudaEvent_t* ev_in = c->host_events.data();
    cudaEvent_t* ev_comp = ev_in + nchunks;
    cudaEvent_t* ev_out = ev_comp + nchunks;
    cudaEvent_t ev_caller = c->host_events[3 * static_cast<size_t>(nchunks)];
    // Work the caller enqueued earlier on ITS stream (tim_set_weight packing, a previous device-path forward) must be complete
    // before the context's own streams read weights / reuse the workspace: an event on that stream, waited for by the H2D and
    // compute streams. No device-wide synchronisation: other streams and contexts of the device keep running.
    CU_OK(c, cudaEventRecord(ev_caller, static_cast<cudaStream_t>(stream)));
    CU_OK(c, cudaStreamWaitEvent(c->s_h2d, ev_caller, 0));
    CU_OK(c, cudaStreamWaitEvent(c->s_comp, ev_caller, 0));

    const bool was_folded_ctx = c->fold_ln && !c->fold_tripped;
    uint64_t up = 0, down = 0;
    int rc = TIM_OK;
    cudaError_t ce = cudaSuccess;
#define HOST_CU(expr) do { if (ce == cudaSuccess) ce = (expr); } while (0)
    auto h2d = [&](void* dst, const void* src, size_t bytes) {
        if (!bytes) return;
        up += bytes;
        HOST_CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->s_h2d));
    };
    void* const hdst[7] = {ho->verb, ho->noun, ho->action, ho->audio, ho->reg_visual, ho->reg_audio, ho->feats};
    for (int i = 0; i < nchunks && rc == TIM_OK && ce == cudaSuccess; ++i) {
        const size_t b0 = static_cast<size_t>(chunk_b0[i]);
        const int nb = chunk_nb[i];
        Set& s = st[i & 1];
        // inputs of chunk i may overwrite set (i&1) once chunk i-2 has been computed
        if (i >= 2) HOST_CU(cudaStreamWaitEvent(c->s_h2d, ev_comp[i - 2], 0));
        if (c->Fv) h2d(s.vis, static_cast<const uint8_t*>(vis) + b0 * n_vis * isz, nb * n_vis * isz);
        if (c->Fa) h2d(s.aud, static_cast<const uint8_t*>(aud) + b0 * n_aud * isz, nb * n_aud * isz);
        h2d(s.times, times + b0 * n_times, nb * n_times * 4);
        HOST_CU(cudaEventRecord(ev_in[i], c->s_h2d));
        // compute: needs its inputs, and its output set free (chunk i-2 copied out)
        HOST_CU(cudaStreamWaitEvent(c->s_comp, ev_in[i], 0));
        if (i >= 2) HOST_CU(cudaStreamWaitEvent(c->s_comp, ev_out[i - 2], 0));
        if (ce != cudaSuccess) break;
        rc = dispatch_dtype(c, [&](auto tag) { return time_mlp_impl<decltype(tag)>(c, s.times, s.te, nb, T_, c->s_comp, c->ws, nullptr); });
        if (rc != TIM_OK) break;
        tim_outputs dout;
        dout.verb = n_verb ? s.out[0] : nullptr; dout.noun = n_noun ? s.out[1] : nullptr; dout.action = n_act ? s.out[2] : nullptr;
        dout.audio = n_au ? s.out[3] : nullptr; dout.reg_visual = n_rv ? s.out[4] : nullptr; dout.reg_audio = n_ra ? s.out[5] : nullptr;
        dout.feats = n_feats ? s.out[6] : nullptr;
        rc = dispatch_dtype(c, [&](auto tag) {
            return encoder_impl<decltype(tag)>(c, reinterpret_cast<const float*>(s.vis), reinterpret_cast<const float*>(s.aud), s.te, nb, T_, Qv, Qa, &dout,
                                               c->s_comp, c->ws, nullptr, nullptr, false, in16);
        });
        if (rc != TIM_OK) break;
        if (out16)
            for (int j = 0; j < 7; ++j)
                if (n_out[j]) { HOST_CU(launch_cast<__half>(s.out[j], s.out16[j], nb * n_out[j], 1, 0, 1.0f, c->s_comp)); c->launches++; }
        HOST_CU(cudaEventRecord(ev_comp[i], c->s_comp));
        HOST_CU(cudaStreamWaitEvent(c->s_d2h, ev_comp[i], 0));
        for (int j = 0; j < 7; ++j) {
            if (!n_out[j]) continue;
            const size_t bytes = nb * n_out[j] * osz;
            down += bytes;
            HOST_CU(cudaMemcpyAsync(static_cast<uint8_t*>(hdst[j]) + b0 * n_out[j] * osz, out16 ? static_cast<const void*>(s.out16[j]) : static_cast<const void*>(s.out[j]),
                                    bytes, cudaMemcpyDeviceToHost, c->s_d2h));
        }
        HOST_CU(cudaEventRecord(ev_out[i], c->s_d2h));
    }
#undef HOST_CU
    // always settle the three streams before returning - also on an error path: the async copies target caller-owned host buffers
    cudaError_t e1 = cudaStreamSynchronize(c->s_h2d), e2 = cudaStreamSynchronize(c->s_comp), e3 = cudaStreamSynchronize(c->s_d2h);
    if (rc != TIM_OK) return rc;
    if (ce != cudaSuccess) return c->fail(TIM_ERR_CUDA, "tim_forward_host: %s", cudaGetErrorString(ce));
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess)
        return c->fail(TIM_ERR_CUDA, "tim_forward_host: stream sync failed: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3)));
    // precision guard of the folded LayerNorm flow, applied WITHIN this call: if a chunk raised the alarm (rows with |mean| > 8 std),
    // the results already in the caller's buffers came from the folded path - switch the context to the un-folded flow and redo
    if (was_folded_ctx && *c->fold_alarm_host) {
        c->fold_tripped = true;
        std::fprintf(stderr, "[tim_b200] rows with |mean| > 8 std seen in the residual stream: LayerNorm folding is switched off for this context; "
                             "this call is recomputed with the un-folded flow\n");
        return tim_forward_host_ex(c, vis, aud, times, B, T_, Qv, Qa, ho, cpc, in_dtype, out_dtype, stream, h2d_bytes, d2h_bytes);
    }
    if (h2d_bytes) *h2d_bytes = up;
    if (d2h_bytes) *d2h_bytes = down;
    return TIM_OK;
}

int tim_forward_host(tim_ctx* c, const float* vis, const float* aud, const float* times, int B, int T_, int Qv, int Qa,
                     const tim_outputs* ho, int cpc, uint64_t* h2d_bytes, uint64_t* d2h_bytes) {
    if (!c) return TIM_ERR_INVALID;
    // no caller stream in this signature: settle the device first (a blocking API can afford it)
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    return tim_forward_host_ex(c, vis, aud, times, B, T_, Qv, Qa, ho, cpc, TIM_FP32, TIM_FP32, nullptr, h2d_bytes, d2h_bytes);
}

// Blocks until the most recent encoder forward of this context has finished its precision check. Returns 1 if that forward ran
// with the LayerNorms folded AND saw rows that break the folded path's error bound (its outputs should be recomputed: the context
// has switched to the un-folded flow, so calling the forward again does that), 0 otherwise.
int tim_fold_check(tim_ctx* c) {
    if (!c) return TIM_ERR_INVALID;
    if (!c->fold_ln || c->fold_tripped || !c->last_folded || !c->fold_event) return 0;
    CU_OK(c, cudaSetDevice(c->device));
    CU_OK(c, cudaEventSynchronize(c->fold_event));
    if (*c->fold_alarm_host) {
        c->fold_tripped = true;
        std::fprintf(stderr, "[tim_b200] rows with |mean| > 8 std seen in the residual stream: LayerNorm folding is switched off for this context\n");
        return 1;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------- training leg
int tim_train_enable(tim_ctx* c) {
    if (!c) return TIM_ERR_INVALID;
    if (c->train && c->train->enabled) return TIM_OK;
    CU_OK(c, cudaSetDevice(c->device));
    if (!c->train) c->train = new tim_train_state();
    tim_train_state& tr = *c->train;
    const bool f32 = c->cfg.compute_dtype == TIM_FP32;
    int maxn = 8;
    for (auto& kv : c->slots) {
        Slot& sl = kv.second;
        if (sl.kind == Slot::VEC) { tr.names[sl.vec] = kv.first; continue; }
        LinearW& w = *sl.lin;
        if (sl.kind == Slot::LIN_B) { tr.names[w.bias] = kv.first; continue; }
        tr.names[w.w] = kv.first;
        if (w.N > maxn) maxn = w.N;
        if (w.K > maxn) maxn = w.K;
        if (w.wt) continue;
        w.Np = f32 ? w.N : ((w.N + 7) & ~7);
        TIM_TRY(dev_alloc(c, &w.wt, static_cast<size_t>(w.K) * w.Np * c->esize));
        CU_OK(c, cudaMemset(w.wt, 0, static_cast<size_t>(w.K) * w.Np * c->esize));
        w.wt_set = false;
        if (!f32) {
            LinearW v;
            v.N = w.K; v.K = w.Np; v.w = w.wt;
            TIM_TRY(make_tmap_w(c, v));
            w.tmBt = v.tmB; w.tmBt2 = v.tmB2; w.has_tmBt2 = v.has_tmB2; w.block_n_t = v.block_n;
        }
    }
    TIM_TRY(dev_alloc(c, reinterpret_cast<void**>(&tr.zero_bias), static_cast<size_t>(maxn) * sizeof(float)));
    CU_OK(c, cudaMemset(tr.zero_bias, 0, static_cast<size_t>(maxn) * sizeof(float)));
    tr.zero_bias_n = maxn;
    tr.enabled = true;
    return TIM_OK;
}

int tim_bind_grad(tim_ctx* c, const char* key, float* dst) {
    if (!c || !key) return TIM_ERR_INVALID;
    if (!c->train || !c->train->enabled) return c->fail(TIM_ERR_INVALID, "tim_bind_grad: call tim_train_enable first");
    if (!std::strncmp(key, "drloc_mlp.", 10) || !std::strncmp(key, "pool.", 5)) return TIM_OK;
    auto it = c->slots.find(key);
    if (it == c->slots.end()) return c->fail(TIM_ERR_WEIGHTS, "unknown state_dict key '%s' for this configuration", key);
    if (dst && (reinterpret_cast<uintptr_t>(dst) & 15)) return c->fail(TIM_ERR_INVALID, "gradient destination of '%s' is not 16-byte aligned", key);
    Slot& sl = it->second;
    const void* p = sl.kind == Slot::VEC ? static_cast<const void*>(sl.vec) : (sl.kind == Slot::LIN_W ? sl.lin->w : static_cast<const void*>(sl.lin->bias));
    c->train->grads[p] = dst;
    return TIM_OK;
}

int tim_set_dropout(tim_ctx* c, float p_feat, float p_seq, float p_enc, uint64_t seed) {
    if (!c) return TIM_ERR_INVALID;
    if (!c->train || !c->train->enabled) return c->fail(TIM_ERR_INVALID, "tim_set_dropout: call tim_train_enable first");
    if (!(p_feat >= 0.0f && p_feat < 1.0f && p_seq >= 0.0f && p_seq < 1.0f && p_enc >= 0.0f && p_enc < 1.0f))
        return c->fail(TIM_ERR_INVALID, "tim_set_dropout: probabilities must lie in [0, 1)");
    c->train->p_feat = p_feat; c->train->p_seq = p_seq; c->train->p_enc = p_enc;
    c->train->drop_seed = static_cast<uint32_t>(seed ^ (seed >> 32));
    return TIM_OK;
}

int tim_time_mlp_fwd_train(tim_ctx* c, const float* times, float* out, int B, int T_, void* stream) {
    if (!c) return TIM_ERR_INVALID;
    if (!times || !out || B <= 0 || T_ <= 0) return c->fail(TIM_ERR_INVALID, "tim_time_mlp_fwd_train: bad arguments");
    TIM_TRY(train_ready(c));
    CU_OK(c, cudaSetDevice(c->device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    return dispatch_dtype(c, [&](auto tag) { return time_mlp_train_fwd<decltype(tag)>(c, times, out, B, T_, s); });
}

int tim_time_mlp_bwd(tim_ctx* c, const float* d_out, void* stream) {
    if (!c) return TIM_ERR_INVALID;
    if (!d_out) return c->fail(TIM_ERR_INVALID, "tim_time_mlp_bwd: bad arguments");
    TIM_TRY(train_ready(c));
    CU_OK(c, cudaSetDevice(c->device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    return dispatch_dtype(c, [&](auto tag) { return time_mlp_train_bwd<decltype(tag)>(c, d_out, s); });
}

int tim_encoder_fwd_train(tim_ctx* c, const float* vis, const float* aud, const float* te, int B, int T_, int Qv, int Qa,
                          const tim_outputs* outs, void* stream) {
    if (!c) return TIM_ERR_INVALID;
    if (!te || !outs || B <= 0 || T_ <= 0) return c->fail(TIM_ERR_INVALID, "tim_encoder_fwd_train: bad arguments");
    TIM_TRY(train_ready(c));
    CU_OK(c, cudaSetDevice(c->device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    return dispatch_dtype(c, [&](auto tag) { return encoder_train_fwd<decltype(tag)>(c, vis, aud, te, B, T_, Qv, Qa, outs, s); });
}

int tim_encoder_bwd(tim_ctx* c, const tim_outputs* grad_outs, float* d_time_enc, void* stream) {
    if (!c) return TIM_ERR_INVALID;
    if (!grad_outs) return c->fail(TIM_ERR_INVALID, "tim_encoder_bwd: bad arguments");
    TIM_TRY(train_ready(c));
    CU_OK(c, cudaSetDevice(c->device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    return dispatch_dtype(c, [&](auto tag) { return encoder_train_bwd<decltype(tag)>(c, grad_outs, d_time_enc, s); });
}

size_t tim_train_tape_bytes(const tim_ctx* c) { return (c && c->train) ? c->train->mem_bytes + c->train->tmem_bytes : 0; }

int tim_comm_unique_id(void* out128) {
    if (!out128) return TIM_ERR_INVALID;
    if (const char* e = load_nccl()) { g_create_error = e; return TIM_ERR_INVALID; }
    const int r = g_nccl.GetUniqueId(out128);
    if (r != 0) { g_create_error = std::string("ncclGetUniqueId failed: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"); return TIM_ERR_CUDA; }
    return TIM_OK;
}

int tim_comm_init(tim_ctx* c, const void* id128, int rank, int world) {
    if (!c || !id128 || world < 1 || rank < 0 || rank >= world) return TIM_ERR_INVALID;
    if (!c->train) c->train = new tim_train_state();
    if (const char* e = load_nccl()) return c->fail(TIM_ERR_INVALID, "%s", e);
    CU_OK(c, cudaSetDevice(c->device));
    if (c->train->nccl_comm) { g_nccl.CommDestroy(c->train->nccl_comm); c->train->nccl_comm = nullptr; }
    ncclUniqueIdBlob id;
    std::memcpy(&id, id128, sizeof(id));
    const int r = g_nccl.CommInitRank(&c->train->nccl_comm, world, id, rank);
    if (r != 0) return c->fail(TIM_ERR_CUDA, "ncclCommInitRank failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    c->train->world = world;
    return TIM_OK;
}

int tim_allreduce_grads(tim_ctx* c, float* buf, size_t n, void* stream) {
    if (!c) return TIM_ERR_INVALID;
    if (!buf || !n) return c->fail(TIM_ERR_INVALID, "tim_allreduce_grads: bad arguments");
    if (!c->train || !c->train->nccl_comm) return c->fail(TIM_ERR_INVALID, "tim_allreduce_grads: no communicator (tim_comm_init)");
    CU_OK(c, cudaSetDevice(c->device));
    // one collective over the flat buffer, averaged over the ranks (what DDP does bucket by bucket, models/build.py:58-63)
    const int r = g_nccl.AllReduce(buf, buf, n, /*ncclFloat32*/ 7, /*ncclAvg*/ 4, c->train->nccl_comm, static_cast<cudaStream_t>(stream));
    if (r != 0) return c->fail(TIM_ERR_CUDA, "ncclAllReduce failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    c->launches++;
    return TIM_OK;
}

int tim_profile_begin(tim_ctx* c) {
    if (!c) return TIM_ERR_INVALID;
    for (auto& r : c->prof) { c->ev_pool.push_back(r.a); c->ev_pool.push_back(r.b); }
    c->prof.clear();
    c->profiling = true;
    return TIM_OK;
}

int tim_profile_end(tim_ctx* c, double* ms, double* flops, uint64_t* count, int n_classes) {
    if (!c) return TIM_ERR_INVALID;
    c->profiling = false;
    CU_OK(c, cudaSetDevice(c->device));
    CU_OK(c, cudaDeviceSynchronize());
    for (int i = 0; i < n_classes; ++i) { if (ms) ms[i] = 0; if (flops) flops[i] = 0; if (count) count[i] = 0; }
    for (auto& r : c->prof) {
        float t = 0.0f;
        CU_OK(c, cudaEventElapsedTime(&t, r.a, r.b));
        if (r.cls >= 0 && r.cls < n_classes) {
            if (ms) ms[r.cls] += t;
            if (flops) flops[r.cls] += r.flops;
            if (count) count[r.cls] += 1;
        }
        c->ev_pool.push_back(r.a); c->ev_pool.push_back(r.b);
    }
    c->prof.clear();
    return TIM_OK;
}

size_t tim_workspace_bytes(const tim_ctx* c) { return c ? c->ws_bytes + c->io_bytes : 0; }
uint64_t tim_launch_count(const tim_ctx* c) { return c ? c->launches : 0; }
int tim_fold_active(const tim_ctx* c) {
    if (!c || !c->fold_ln || c->fold_tripped) return 0;
    return (c->fold_alarm_host && *c->fold_alarm_host) ? 0 : 1;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------- single-kernel test hooks
namespace {
struct TmpCtx {           // minimal context for the test hooks (no weights)
    tim_ctx c;
    std::vector<void*> tmp;
    ~TmpCtx() { for (void* p : tmp) cudaFree(p); }
    template <typename P> cudaError_t alloc(P** p, size_t bytes) {
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), bytes ? bytes : 16);
        if (e == cudaSuccess) tmp.push_back(*p);
        return e;
    }
};
}  // namespace

extern "C" {

int tim_test_linear(int dtype, const float* A, const float* W, const float* bias, const float* resid, float* out, int M, int N,
                    int K, int act, void* stream) {
    TmpCtx t;
    tim_ctx* c = &t.c;
    auto fin = [&](int r) { g_create_error = c->err; return r; };
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess || prop.major != 10) {
        g_create_error = "tim_test_linear: no sm_100 device";
        return TIM_ERR_NO_DEVICE;
    }
    c->num_sms = prop.multiProcessorCount; c->device = dev;
    c->cfg.compute_dtype = dtype;
    c->esize = dtype == TIM_FP32 ? 4 : 2;
    if (const char* gv = std::getenv("TIM_B200_GEMM")) c->gemm_version = std::atoi(gv) == 1 ? 1 : 2;
    LinearW w;
    w.N = N; w.K = K; w.bias = const_cast<float*>(bias);
    Epilogue ep = epi(out, N, true, act, resid, N);
    float* zero_bias = nullptr;
    if (!bias) {
        if (t.alloc(&zero_bias, N * sizeof(float)) != cudaSuccess) return fin(c->fail(TIM_ERR_NOMEM, "alloc"));
        cudaMemsetAsync(zero_bias, 0, N * sizeof(float), s);
        w.bias = zero_bias;
    }
    int r;
    if (dtype == TIM_FP32) {
        w.w = const_cast<float*>(W);
        r = run_linear<float>(c, A, K, w, plain_rows(M), ep, s);
    } else {
        r = get_encode_fn(c);
        if (r) return fin(r);
        void *a16 = nullptr, *w16 = nullptr;
        if (t.alloc(&a16, static_cast<size_t>(M) * K * 2) != cudaSuccess || t.alloc(&w16, static_cast<size_t>(N) * K * 2) != cudaSuccess)
            return fin(c->fail(TIM_ERR_NOMEM, "alloc"));
        w.w = w16;
        if (dtype == TIM_BF16) {
            launch_cast<__nv_bfloat16>(A, static_cast<__nv_bfloat16*>(a16), M, K, 0, 1.0f, s);
            launch_cast<__nv_bfloat16>(W, static_cast<__nv_bfloat16*>(w16), N, K, 0, 1.0f, s);
        } else {
            launch_cast<__half>(A, static_cast<__half*>(a16), M, K, 0, 1.0f, s);
            launch_cast<__half>(W, static_cast<__half*>(w16), N, K, 0, 1.0f, s);
        }
        r = make_tmap_w(c, w);
        if (r) return fin(r);
        r = dtype == TIM_BF16 ? run_linear<__nv_bfloat16>(c, a16, K, w, plain_rows(M), ep, s)
                              : run_linear<__half>(c, a16, K, w, plain_rows(M), ep, s);
    }
    if (r) return fin(r);
    cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return fin(c->fail(TIM_ERR_CUDA, "tim_test_linear: %s", cudaGetErrorString(e)));
    return TIM_OK;
}

int tim_bench_linear(int dtype, const void* A16, const void* W16, const float* bias, const float* resid, void* out, int M, int N, int K,
                     int act, int out_fp32, int version, int iters, float* ms_per_iter) {
    TmpCtx t;
    tim_ctx* c = &t.c;
    auto fin = [&](int r) { g_create_error = c->err; return r; };
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess || prop.major != 10) {
        g_create_error = "tim_bench_linear: no sm_100 device";
        return TIM_ERR_NO_DEVICE;
    }
    if (dtype != TIM_BF16 && dtype != TIM_FP16) return fin(c->fail(TIM_ERR_INVALID, "tim_bench_linear: 16-bit dtypes only"));
    c->num_sms = prop.multiProcessorCount; c->device = dev;
    c->cfg.compute_dtype = dtype;
    c->esize = 2;
    c->gemm_version = version == 1 ? 1 : 2;
    int r = get_encode_fn(c);
    if (r) return fin(r);
    LinearW w;
    w.N = N; w.K = K; w.bias = const_cast<float*>(bias); w.w = const_cast<void*>(W16);
    r = make_tmap_w(c, w);
    if (r) return fin(r);
    Epilogue ep = epi(out, N, out_fp32 != 0, act, resid, N);
    cudaStream_t s = nullptr;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < iters + 2 && r == TIM_OK; ++i) {
        if (i == 2) cudaEventRecord(e0, s);
        r = dtype == TIM_BF16 ? run_linear<__nv_bfloat16>(c, A16, K, w, plain_rows(M), ep, s) : run_linear<__half>(c, A16, K, w, plain_rows(M), ep, s);
    }
    cudaEventRecord(e1, s);
    cudaError_t e = cudaEventSynchronize(e1);
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (r) return fin(r);
    if (e != cudaSuccess) return fin(c->fail(TIM_ERR_CUDA, "tim_bench_linear: %s", cudaGetErrorString(e)));
    if (ms_per_iter) *ms_per_iter = ms / (iters > 0 ? iters : 1);
    return TIM_OK;
}

int tim_debug_role_prof(unsigned long long* dev_buf) {
    set_attention_role_prof(dev_buf);
    return TIM_OK;
}

int tim_bench_attention(int dtype, const void* qkv16, void* out16, int B, int Ft, int Qt, int H, int hd, int version, int iters,
                        float* ms_per_iter) {
    TmpCtx t;
    tim_ctx* c = &t.c;
    auto fin = [&](int r) { g_create_error = c->err; return r; };
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess || prop.major != 10) {
        g_create_error = "tim_bench_attention: no sm_100 device";
        return TIM_ERR_NO_DEVICE;
    }
    if (dtype != TIM_BF16 && dtype != TIM_FP16) return fin(c->fail(TIM_ERR_INVALID, "tim_bench_attention: 16-bit dtypes only"));
    c->num_sms = prop.multiProcessorCount; c->device = dev; c->cfg.compute_dtype = dtype; c->esize = 2;
    c->H = H; c->hd = hd; c->E = H * hd; c->attn_version = version < 1 ? 1 : (version > 4 ? 4 : version);
    int r = get_encode_fn(c);
    if (r) return fin(r);
    cudaStream_t s = nullptr;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaError_t e = cudaSuccess;
    auto run = [&](auto tag) -> int {
        using TT = decltype(tag);
        AttnUmmaParams ap;
        bool use_umma = false;
        TIM_TRY(prepare_attention<TT>(c, &ap, &use_umma, static_cast<const TT*>(qkv16), static_cast<TT*>(out16), B, Ft, Qt));
        for (int i = 0; i < iters + 2 && e == cudaSuccess; ++i) {
            if (i == 2) cudaEventRecord(e0, s);
            e = use_umma ? launch_attention_tc<TT>(c, ap, s)
                         : launch_attention_mma<TT>(static_cast<const TT*>(qkv16), static_cast<TT*>(out16), B, Ft, Qt, H, hd, s);
        }
        return TIM_OK;
    };
    r = dtype == TIM_BF16 ? run(__nv_bfloat16{}) : run(__half{});
    cudaEventRecord(e1, s);
    cudaError_t es = cudaEventSynchronize(e1);
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (r) return fin(r);
    if (e != cudaSuccess || es != cudaSuccess) return fin(c->fail(TIM_ERR_CUDA, "tim_bench_attention: %s", cudaGetErrorString(e != cudaSuccess ? e : es)));
    if (ms_per_iter) *ms_per_iter = ms / (iters > 0 ? iters : 1);
    return TIM_OK;
}

int tim_test_attention(int dtype, const float* qkv, float* out, int B, int Ft, int Qt, int H, int hd, void* stream) {
    TmpCtx t;
    tim_ctx* c = &t.c;
    auto fin = [&](int r) { g_create_error = c->err; return r; };
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t M = static_cast<size_t>(B) * (Ft + Qt), E = static_cast<size_t>(H) * hd;
    cudaError_t e = cudaSuccess;
    const float* scaled = qkv;          // the caller pre-scales q by hd^-0.5 * log2(e), as the packed in_proj weights do
    if (dtype == TIM_FP32) {
        e = launch_attention_simt(scaled, out, B, Ft, Qt, H, hd, s);
    } else {
        void *q16 = nullptr, *o16 = nullptr;
        int dev = 0;
        cudaDeviceProp prop;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess || prop.major != 10) {
            g_create_error = "tim_test_attention: no sm_100 device";
            return TIM_ERR_NO_DEVICE;
        }
        c->num_sms = prop.multiProcessorCount; c->device = dev; c->cfg.compute_dtype = dtype; c->esize = 2;
        c->H = H; c->hd = hd; c->E = static_cast<int>(E);
        if (const char* av = std::getenv("TIM_B200_ATTN")) { const int v = std::atoi(av); c->attn_version = v < 1 ? 1 : (v > 4 ? 4 : v); }
        int rc = get_encode_fn(c);
        if (rc) return fin(rc);
        // same dispatch as the forward: tcgen05 kernel where the shape allows, warp-MMA kernel otherwise
        auto test_attn = [&](auto* qp, auto* op) -> cudaError_t {
            using TT = std::remove_pointer_t<decltype(qp)>;
            AttnUmmaParams ap;
            bool use_umma = false;
            if (prepare_attention<TT>(c, &ap, &use_umma, qp, op, B, Ft, Qt) != TIM_OK) return cudaErrorInvalidValue;
            if (use_umma) return launch_attention_tc<TT>(c, ap, s);
            return launch_attention_mma<TT>(qp, op, B, Ft, Qt, H, hd, s);
        };
        if (t.alloc(&q16, M * 3 * E * 2) != cudaSuccess || t.alloc(&o16, M * E * 2) != cudaSuccess) return fin(c->fail(TIM_ERR_NOMEM, "alloc"));
        if (dtype == TIM_BF16) {
            launch_cast<__nv_bfloat16>(scaled, static_cast<__nv_bfloat16*>(q16), M, 3 * E, 0, 1.0f, s);
            e = test_attn(static_cast<__nv_bfloat16*>(q16), static_cast<__nv_bfloat16*>(o16));
        } else {
            launch_cast<__half>(scaled, static_cast<__half*>(q16), M, 3 * E, 0, 1.0f, s);
            e = test_attn(static_cast<__half*>(q16), static_cast<__half*>(o16));
        }
        if (e == cudaSuccess) {
            // widen back to fp32 with a plain device loop (cast kernels only go fp32 -> T): use cudaMemcpy2D-free host path
            std::vector<uint16_t> h(M * E);
            e = cudaStreamSynchronize(s);
            if (e == cudaSuccess) e = cudaMemcpy(h.data(), o16, M * E * 2, cudaMemcpyDeviceToHost);
            if (e == cudaSuccess) {
                std::vector<float> f(M * E);
                for (size_t i = 0; i < M * E; ++i) {
                    if (dtype == TIM_BF16) { uint32_t u = static_cast<uint32_t>(h[i]) << 16; std::memcpy(&f[i], &u, 4); }
                    else {
                        __half_raw hr; hr.x = h[i];
                        f[i] = __half2float(__half(hr));
                    }
                }
                e = cudaMemcpy(out, f.data(), M * E * 4, cudaMemcpyHostToDevice);
            }
        }
    }
    if (e != cudaSuccess) return fin(c->fail(TIM_ERR_CUDA, "tim_test_attention: %s", cudaGetErrorString(e)));
    e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return fin(c->fail(TIM_ERR_CUDA, "tim_test_attention: %s", cudaGetErrorString(e)));
    return TIM_OK;
}

// dW[N, K] (fp32, accumulated into) += dY[M, N]^T X[M, K] through the selected compute path (16-bit paths cast dY and X on the device
// first; N is padded to a multiple of 8 columns for the TMA pitch). splits <= 0: chosen by the library.
int tim_test_wgrad(int dtype, const float* dY, const float* X, float* dW, int M, int N, int K, int splits, void* stream) {
    TmpCtx t;
    tim_ctx* c = &t.c;
    auto fin = [&](int r) { g_create_error = c->err; return r; };
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess || prop.major != 10) {
        g_create_error = "tim_test_wgrad: no sm_100 device";
        return TIM_ERR_NO_DEVICE;
    }
    c->num_sms = prop.multiProcessorCount; c->device = dev; c->cfg.compute_dtype = dtype; c->esize = dtype == TIM_FP32 ? 4 : 2;
    c->wgrad_splits = splits;
    int r;
    if (dtype == TIM_FP32) {
        r = run_wgrad<float>(c, dY, N, X, K, dW, K, M, N, K, s);
    } else {
        r = get_encode_fn(c);
        if (r) return fin(r);
        const int Np = (N + 7) & ~7, Kp = (K + 7) & ~7;
        void *y16 = nullptr, *x16 = nullptr;
        if (t.alloc(&y16, static_cast<size_t>(M) * Np * 2) != cudaSuccess || t.alloc(&x16, static_cast<size_t>(M) * Kp * 2) != cudaSuccess)
            return fin(c->fail(TIM_ERR_NOMEM, "alloc"));
        if (dtype == TIM_BF16) {
            launch_cast_pad<__nv_bfloat16>(dY, static_cast<__nv_bfloat16*>(y16), M, N, Np, s);
            launch_cast_pad<__nv_bfloat16>(X, static_cast<__nv_bfloat16*>(x16), M, K, Kp, s);
            r = run_wgrad<__nv_bfloat16>(c, y16, Np, x16, Kp, dW, K, M, N, K, s);
        } else {
            launch_cast_pad<__half>(dY, static_cast<__half*>(y16), M, N, Np, s);
            launch_cast_pad<__half>(X, static_cast<__half*>(x16), M, K, Kp, s);
            r = run_wgrad<__half>(c, y16, Np, x16, Kp, dW, K, M, N, K, s);
        }
    }
    if (r) return fin(r);
    cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return fin(c->fail(TIM_ERR_CUDA, "tim_test_wgrad: %s", cudaGetErrorString(e)));
    return TIM_OK;
}

// timing hook (tools/wgrad_bench.py): `iters` back-to-back weight-gradient launches on 16-bit device operands
int tim_bench_wgrad(int dtype, const void* dY16, const void* X16, float* dW, int M, int N, int K, int splits, int iters, float* ms_per_iter) {
    TmpCtx t;
    tim_ctx* c = &t.c;
    auto fin = [&](int r) { g_create_error = c->err; return r; };
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess || prop.major != 10) {
        g_create_error = "tim_bench_wgrad: no sm_100 device";
        return TIM_ERR_NO_DEVICE;
    }
    if (dtype != TIM_BF16 && dtype != TIM_FP16) return fin(c->fail(TIM_ERR_INVALID, "tim_bench_wgrad: 16-bit dtypes only"));
    c->num_sms = prop.multiProcessorCount; c->device = dev; c->cfg.compute_dtype = dtype; c->esize = 2; c->wgrad_splits = splits;
    int r = get_encode_fn(c);
    if (r) return fin(r);
    cudaStream_t s = nullptr;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < iters + 2 && r == TIM_OK; ++i) {
        if (i == 2) cudaEventRecord(e0, s);
        r = dtype == TIM_BF16 ? run_wgrad<__nv_bfloat16>(c, dY16, N, X16, K, dW, K, M, N, K, s) : run_wgrad<__half>(c, dY16, N, X16, K, dW, K, M, N, K, s);
    }
    cudaEventRecord(e1, s);
    cudaError_t e = cudaEventSynchronize(e1);
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (r) return fin(r);
    if (e != cudaSuccess) return fin(c->fail(TIM_ERR_CUDA, "tim_bench_wgrad: %s", cudaGetErrorString(e)));
    if (ms_per_iter) *ms_per_iter = ms / (iters > 0 ? iters : 1);
    return TIM_OK;
}

// attention backward over a two-stream qkv buffer: qkv [(B*Ft + B*Qt), 3*H*hd], dO [.., H*hd] -> dqkv [.., 3*H*hd], all fp32 on the
// device (16-bit paths cast on the device, run the mma kernels and widen the result on the host). qscale: see attention_bwd.cu.
int tim_test_attention_bwd(int dtype, const float* qkv, const float* dO, float* dqkv, int B, int Ft, int Qt, int H, int hd, float qscale,
                           void* stream) {
    TmpCtx t;
    tim_ctx* c = &t.c;
    auto fin = [&](int r) { g_create_error = c->err; return r; };
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t M = static_cast<size_t>(B) * (Ft + Qt), E = static_cast<size_t>(H) * hd;
    cudaError_t e = cudaSuccess;
    if (dtype == TIM_FP32) {
        e = cudaMemsetAsync(dqkv, 0, M * 3 * E * sizeof(float), s);
        if (e == cudaSuccess) e = launch_attention_bwd_simt(qkv, dO, dqkv, B, Ft, Qt, H, hd, qscale, s);
    } else {
        void *q16 = nullptr, *d16 = nullptr, *o16 = nullptr, *st = nullptr;
        if (t.alloc(&q16, M * 3 * E * 2) != cudaSuccess || t.alloc(&d16, M * E * 2) != cudaSuccess || t.alloc(&o16, M * 3 * E * 2) != cudaSuccess ||
            t.alloc(&st, attention_bwd_stats_bytes(B, Ft, Qt, H)) != cudaSuccess)
            return fin(c->fail(TIM_ERR_NOMEM, "alloc"));
        cudaMemsetAsync(o16, 0, M * 3 * E * 2, s);
        int dev = 0;
        cudaDeviceProp prop;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess || prop.major != 10) {
            g_create_error = "tim_test_attention_bwd: no sm_100 device";
            return TIM_ERR_NO_DEVICE;
        }
        c->num_sms = prop.multiProcessorCount; c->device = dev; c->cfg.compute_dtype = dtype; c->esize = 2;
        c->H = H; c->hd = hd; c->E = static_cast<int>(E);
        int rc = get_encode_fn(c);
        if (rc) return fin(rc);
        // same dispatch as the training leg: tcgen05 kernel where the shape allows (TIM_B200_ATTN_BWD=1 forces the warp-MMA kernels)
        if (dtype == TIM_BF16) {
            launch_cast<__nv_bfloat16>(qkv, static_cast<__nv_bfloat16*>(q16), M, 3 * E, 0, 1.0f, s);
            launch_cast<__nv_bfloat16>(dO, static_cast<__nv_bfloat16*>(d16), M, E, 0, 1.0f, s);
            rc = run_attention_bwd<__nv_bfloat16>(c, static_cast<const __nv_bfloat16*>(q16), static_cast<const __nv_bfloat16*>(d16),
                                                  static_cast<__nv_bfloat16*>(o16), st, B, Ft, Qt, qscale, 0.0, s);
        } else {
            launch_cast<__half>(qkv, static_cast<__half*>(q16), M, 3 * E, 0, 1.0f, s);
            launch_cast<__half>(dO, static_cast<__half*>(d16), M, E, 0, 1.0f, s);
            rc = run_attention_bwd<__half>(c, static_cast<const __half*>(q16), static_cast<const __half*>(d16), static_cast<__half*>(o16), st, B, Ft,
                                           Qt, qscale, 0.0, s);
        }
        if (rc) return fin(rc);
        if (e == cudaSuccess) {
            std::vector<uint16_t> h(M * 3 * E);
            e = cudaStreamSynchronize(s);
            if (e == cudaSuccess) e = cudaMemcpy(h.data(), o16, M * 3 * E * 2, cudaMemcpyDeviceToHost);
            if (e == cudaSuccess) {
                std::vector<float> f(M * 3 * E);
                for (size_t i = 0; i < M * 3 * E; ++i) {
                    if (dtype == TIM_BF16) { uint32_t u = static_cast<uint32_t>(h[i]) << 16; std::memcpy(&f[i], &u, 4); }
                    else { __half_raw hr; hr.x = h[i]; f[i] = __half2float(__half(hr)); }
                }
                e = cudaMemcpy(dqkv, f.data(), M * 3 * E * 4, cudaMemcpyHostToDevice);
            }
        }
    }
    if (e != cudaSuccess) return fin(c->fail(TIM_ERR_CUDA, "tim_test_attention_bwd: %s", cudaGetErrorString(e)));
    e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return fin(c->fail(TIM_ERR_CUDA, "tim_test_attention_bwd: %s", cudaGetErrorString(e)));
    return TIM_OK;
}

}  // extern "C"
