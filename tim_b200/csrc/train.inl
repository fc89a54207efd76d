// Training leg of libtim_b200 (included by api.cu): the forward that keeps what the backward needs, the backward schedule, the
// gradient destinations and the one data-path collective (NCCL all-reduce of the flat gradient buffer).
//
// Reference: recognition/scripts/train.py:190-260, 354-366 (autocast forward, loss, GradScaler.scale(loss).backward(), DDP bucketed
// all-reduce from models/build.py:58-63, optimizer step); detection/time_interval_machine/models/tim.py:272-337 (forward_train).
// The reference has no hand-written backward: this file schedules what torch.autograd would run for the graph of tim.py:147-172,
// with dropout p = 0 semantics (the drop-in refuses p > 0, see tim_b200/plugin.py).
//
// Forward (16-bit modes): the un-folded LayerNorm flow of encoder_impl - every LayerNorm writes its 16-bit output (the A operand the
// weight gradient needs) + row statistics, and the GEMM that needs LN(.) as its residual normalises on read. Kept per layer:
// xin (A of in_proj), qkv, att, z1 (fp32, pre-norm1), x1 = LN1(z1), u (pre-GELU), hid = GELU(u), z2 (fp32, pre-norm2): 28 KB per
// token row and layer at E = 1024. fp32 mode keeps the same tensors in fp32.
// Backward per layer: LN2' -> wgrad / dgrad linear2 -> GELU' -> wgrad / dgrad linear1 (+ residual) -> LN1' -> wgrad / dgrad out_proj
// -> attention backward -> wgrad / dgrad in_proj (+ residual). dgrad = the forward CTA-pair GEMM on a transposed weight copy
// (packed by tim_set_weight); wgrad = gemm_wgrad.cu (MN-major tcgen05, split over token rows, TMA reduce-add).

struct LayerTape {
    void *xin = nullptr, *qkv = nullptr, *att = nullptr, *x1 = nullptr, *u = nullptr, *hid = nullptr;
    float *z1 = nullptr, *z2 = nullptr;
    float2 *st1 = nullptr, *st2 = nullptr;
};
struct RegTape { void *r1 = nullptr, *r2 = nullptr; float* y = nullptr; };

struct tim_train_state {
    bool enabled = false;
    float* zero_bias = nullptr;
    int zero_bias_n = 0;
    std::map<const void*, float*> grads;        // device parameter pointer (LinearW::w / ::bias / vector) -> gradient destination
    std::map<const void*, std::string> names;   // for error messages
    // encoder tape
    bool enc_valid = false;
    int B = 0, T = 0, Qv = 0, Qa = 0;
    QueryPlan qp;
    uint8_t* mem = nullptr; size_t mem_bytes = 0;
    float *tok32 = nullptr, *embv_pre = nullptr, *emba_pre = nullptr, *embv_act = nullptr, *emba_act = nullptr;
    void *visT = nullptr, *audT = nullptr, *xh = nullptr;
    std::vector<LayerTape> layers;
    RegTape reg_v, reg_a;
    // time-MLP tape
    bool time_valid = false;
    int tM = 0;
    uint8_t* tmem = nullptr; size_t tmem_bytes = 0;
    float *times = nullptr, *t3 = nullptr;
    void *t1 = nullptr, *t2 = nullptr;
    // dropout of the NEXT training forward (tim_set_dropout); the encoder tape remembers what its forward used
    float p_feat = 0.0f, p_seq = 0.0f, p_enc = 0.0f;
    uint32_t drop_seed = 0;
    float tape_p_seq = 0.0f, tape_p_enc = 0.0f;
    uint32_t tape_seed = 0;
    // NCCL (dlopen'ed, see tim_comm_init)
    void* nccl_comm = nullptr;
    int world = 1;
};

namespace {

struct NcclApi {
    void* h = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, ncclUniqueIdBlob, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;

const char* load_nccl() {
    if (g_nccl.h) return nullptr;
    // the process normally has torch's bundled libnccl.so.2 mapped already; dlopen by soname returns that copy
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return "libnccl.so.2 not found (import torch first, or put NCCL on the library path)";
    g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    g_nccl.AllReduce = reinterpret_cast<decltype(g_nccl.AllReduce)>(dlsym(h, "ncclAllReduce"));
    g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy) return "NCCL symbols missing";
    g_nccl.h = h;
    return nullptr;
}

// ------------------------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------------------------
int grad_dst(tim_ctx* c, const void* param, float** out) {
    auto it = c->train->grads.find(param);
    if (it == c->train->grads.end() || !it->second) {
        auto nm = c->train->names.find(param);
        return c->fail(TIM_ERR_WEIGHTS, "no gradient destination bound for '%s' (tim_bind_grad)", nm != c->train->names.end() ? nm->second.c_str() : "?");
    }
    *out = it->second;
    return TIM_OK;
}

// the transposed copy of a linear layer as the "weight" of the dgrad GEMM: dX[rows, K] = dY[rows, Np] * Wt[K, Np]^T
inline LinearW dgrad_view(const tim_ctx* c, const LinearW& w) {
    LinearW v;
    v.N = w.K; v.K = w.Np; v.w = w.wt; v.bias = c->train->zero_bias;
    v.tmB = w.tmBt; v.tmB2 = w.tmBt2; v.has_tmB2 = w.has_tmBt2; v.block_n = w.block_n_t;
    return v;
}

template <typename T>
int run_wgrad(tim_ctx* c, const void* dY, int ldy, const void* X, int ldx, float* dW, int ldw, int M, int N, int K, cudaStream_t s) {
    if (M <= 0 || N <= 0 || K <= 0) return TIM_OK;
    const double fl = 2.0 * M * static_cast<double>(N) * K;
    if constexpr (std::is_same<T, float>::value) {
        LAUNCH_C(c, 8, fl, s, launch_wgrad_simt<float>(static_cast<const float*>(dY), ldy, static_cast<const float*>(X), ldx, dW, ldw, M, N, K, s));
    } else {
        const bool aligned = ((reinterpret_cast<uintptr_t>(dY) | reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(dW)) & 15) == 0;
        if (aligned && wgrad_umma_supported(M, N, K, ldy, ldx, ldw)) {
            WgradParams q;
            std::memset(&q, 0, sizeof(q));
            TIM_TRY(make_tmap_2d(c, &q.tmA, dY, op_dtype(c), 2, ldy, M, static_cast<long long>(ldy) * 2, 64, 64));
            TIM_TRY(make_tmap_2d(c, &q.tmB, X, op_dtype(c), 2, K, M, static_cast<long long>(ldx) * 2, 64, 64));
            TIM_TRY(make_tmap_2d(c, &q.tmOut, dW, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, K, N, static_cast<long long>(ldw) * 4, 32, 32));
            q.M = M; q.N = N; q.K = K; q.splits = c->wgrad_splits;
            LAUNCH_C(c, 8, fl, s, launch_wgrad_umma<T>(q, c->num_sms, s));
        } else {
            LAUNCH_C(c, 8, fl, s, launch_wgrad_simt<T>(static_cast<const T*>(dY), ldy, static_cast<const T*>(X), ldx, dW, ldw, M, N, K, s));
        }
    }
    return TIM_OK;
}

// attention backward: tcgen05 kernel where it applies (head_dim 64 / 128, Ft <= 128), else the two warp-MMA kernels
template <typename T>
int run_attention_bwd(tim_ctx* c, const T* qkv, const T* dO, T* dqkv, void* stats, int B, int Ft, int Qt, float qscale, double flops, cudaStream_t s,
                      DropSite drop = DropSite()) {
    if (drop.thr && !attention_bwd_umma_supported(Ft, c->hd))
        return c->fail(TIM_ERR_INVALID, "attention dropout in the 16-bit modes needs head_dim 64 or 128 (tcgen05 attention kernels), got %d", c->hd);
    if (attention_bwd_umma_supported(Ft, c->hd)) {
        AttnBwdUmmaParams ap;
        std::memset(&ap, 0, sizeof(ap));
        const long long E = c->E, ld = 3LL * c->E;
        const int Fp = (Ft + 15) & ~15;
        TIM_TRY(make_tmap_3d16(c, &ap.tmKV, qkv, ld, Ft, B, ld * 2, ld * 2 * Ft, Fp));
        TIM_TRY(make_tmap_3d16(c, &ap.tmQf, qkv, ld, Ft, B, ld * 2, ld * 2 * Ft, 128));
        TIM_TRY(make_tmap_3d16(c, &ap.tmDf, dO, E, Ft, B, E * 2, E * 2 * Ft, 128));
        TIM_TRY(make_tmap_3d16(c, &ap.tmGf, dqkv, ld, Ft, B, ld * 2, ld * 2 * Ft, 32));
        if (Qt > 0) {
            TIM_TRY(make_tmap_3d16(c, &ap.tmQq, qkv + static_cast<size_t>(B) * Ft * ld, ld, Qt, B, ld * 2, ld * 2 * Qt, 128));
            TIM_TRY(make_tmap_3d16(c, &ap.tmDq, dO + static_cast<size_t>(B) * Ft * E, E, Qt, B, E * 2, E * 2 * Qt, 128));
            TIM_TRY(make_tmap_3d16(c, &ap.tmGq, dqkv + static_cast<size_t>(B) * Ft * ld, ld, Qt, B, ld * 2, ld * 2 * Qt, 32));
        } else {
            ap.tmQq = ap.tmQf; ap.tmDq = ap.tmDf; ap.tmGq = ap.tmGf;
        }
        ap.qkv = qkv; ap.dO = dO; ap.dqkv = dqkv; ap.B = B; ap.Ft = Ft; ap.Qt = Qt; ap.H = c->H; ap.qscale = qscale; ap.drop = drop;
        if (drop.thr && 1ull * B * c->H * (Ft + Qt) * DROP_ATTN_KW > 0xffffffffull) return c->fail(TIM_ERR_INVALID, "attention dropout: batch too large for the 32-bit element index");
        LAUNCH_C(c, 9, flops, s, launch_attention_bwd_umma<T>(ap, c->hd, c->num_sms, s));
    } else {
        LAUNCH_C(c, 9, flops, s, launch_attention_bwd<T>(qkv, dO, dqkv, stats, B, Ft, Qt, c->H, c->hd, qscale, s));
        c->launches++;      // two kernels per call
    }
    return TIM_OK;
}

// dgrad through the forward GEMM kernels; profiled as class 7
template <typename T>
int run_dgrad(tim_ctx* c, const void* dY, const LinearW& w, int rows, Epilogue ep, cudaStream_t s) {
    if (!w.wt) return c->fail(TIM_ERR_WEIGHTS, "transposed weight copy missing: set the weights again after tim_train_enable()");
    const LinearW v = dgrad_view(c, w);
    c->class_override = 7;
    const int r = run_linear<T>(c, dY, w.Np, v, plain_rows(rows), ep, s);
    c->class_override = -1;
    return r;
}

int ensure_tape(tim_ctx* c, uint8_t** mem, size_t* have, size_t need) {
    if (need <= *have) return TIM_OK;
    if (*mem) { cudaDeviceSynchronize(); cudaFree(*mem); *mem = nullptr; *have = 0; }
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(mem), need);
    if (e != cudaSuccess) return c->fail(TIM_ERR_NOMEM, "activation tape cudaMalloc(%zu) failed: %s", need, cudaGetErrorString(e));
    *have = need;
    return TIM_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// time MLP: forward that keeps the three activations, and its backward (tim.py:66-74)
// ------------------------------------------------------------------------------------------------------------------
template <typename T>
int time_mlp_train_fwd(tim_ctx* c, const float* times, float* out, int B, int T_, cudaStream_t s) {
    constexpr bool f32 = std::is_same<T, float>::value;
    tim_train_state& tr = *c->train;
    const int M = B * T_, d = c->d;
    for (int pass = 0; pass < 2; ++pass) {
        Arena a{pass ? tr.tmem : nullptr};
        a.take(&tr.times, static_cast<size_t>(M) * 2 * sizeof(float));
        a.take(&tr.t1, static_cast<size_t>(M) * d * sizeof(T));
        a.take(&tr.t2, static_cast<size_t>(M) * d * sizeof(T));
        a.take(&tr.t3, static_cast<size_t>(M) * d * sizeof(float));
        if (!pass) TIM_TRY(ensure_tape(c, &tr.tmem, &tr.tmem_bytes, a.off));
    }
    tr.time_valid = false;
    CU_OK(c, cudaMemcpyAsync(tr.times, times, static_cast<size_t>(M) * 2 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    LAUNCH(c, launch_time_l1<T>(tr.times, c->t0w, c->t0b, static_cast<T*>(tr.t1), M, d, s));
    TIM_TRY(run_linear<T>(c, tr.t1, d, c->t2, plain_rows(M), epi(tr.t2, d, f32, ACT_RELU), s));
    TIM_TRY(run_linear<T>(c, tr.t2, d, c->t4, plain_rows(M), epi(tr.t3, d, true, ACT_RELU), s));
    LAUNCH(c, launch_layernorm<T>(tr.t3, d, c->tlg, c->tlb, out, d, static_cast<T*>(nullptr), 0, M, d, s));
    tr.tM = M;
    tr.time_valid = true;
    return TIM_OK;
}

template <typename T>
int time_mlp_train_bwd(tim_ctx* c, const float* d_out, cudaStream_t s) {
    constexpr bool f32 = std::is_same<T, float>::value;
    tim_train_state& tr = *c->train;
    if (!tr.time_valid) return c->fail(TIM_ERR_INVALID, "tim_time_mlp_bwd without a matching tim_time_mlp_fwd_train");
    const int M = tr.tM, d = c->d;
    size_t need = 0;
    float* g32 = nullptr; T *d3 = nullptr, *d2 = nullptr;
    for (int pass = 0; pass < 2; ++pass) {
        Arena a{pass ? c->ws : nullptr};
        a.take(&g32, static_cast<size_t>(M) * d * sizeof(float));
        a.take(&d3, f32 ? 0 : static_cast<size_t>(M) * d * sizeof(T));
        a.take(&d2, static_cast<size_t>(M) * d * sizeof(T));
        if (!pass) { need = a.off; TIM_TRY(ensure_ws(c, need)); }
    }
    float *g_lg, *g_lb, *g_b4, *g_w4, *g_b2, *g_w2, *g_b0, *g_w0;
    TIM_TRY(grad_dst(c, c->tlg, &g_lg)); TIM_TRY(grad_dst(c, c->tlb, &g_lb));
    TIM_TRY(grad_dst(c, c->t4.w, &g_w4)); TIM_TRY(grad_dst(c, c->t4.bias, &g_b4));
    TIM_TRY(grad_dst(c, c->t2.w, &g_w2)); TIM_TRY(grad_dst(c, c->t2.bias, &g_b2));
    TIM_TRY(grad_dst(c, c->t0w, &g_w0)); TIM_TRY(grad_dst(c, c->t0b, &g_b0));
    CU_OK(c, cudaMemcpyAsync(g32, d_out, static_cast<size_t>(M) * d * sizeof(float), cudaMemcpyDeviceToDevice, s));
    LAUNCH(c, launch_ln_bwd<T>(g32, d, tr.t3, d, c->tlg, static_cast<T*>(nullptr), 0, g_lg, g_lb, nullptr, M, d, s));
    const T* d3op;
    if constexpr (f32) {
        LAUNCH(c, (launch_act_bwd<float, float, float>(1, g32, tr.t3, g32, M, d, g_b4, s)));
        d3op = g32;
    } else {
        LAUNCH(c, (launch_act_bwd<float, float, T>(1, g32, tr.t3, d3, M, d, g_b4, s)));
        d3op = d3;
    }
    TIM_TRY(run_wgrad<T>(c, d3op, d, tr.t2, d, g_w4, d, M, d, d, s));
    TIM_TRY(run_dgrad<T>(c, d3op, c->t4, M, epi(d2, d, f32), s));
    LAUNCH(c, (launch_act_bwd<T, T, T>(1, d2, static_cast<const T*>(tr.t2), d2, M, d, g_b2, s)));
    TIM_TRY(run_wgrad<T>(c, d2, d, tr.t1, d, g_w2, d, M, d, d, s));
    // d1 overwrites the d3 buffer (fp32 mode: g32)
    T* d1 = f32 ? reinterpret_cast<T*>(g32) : d3;
    TIM_TRY(run_dgrad<T>(c, d2, c->t2, M, epi(d1, d, f32), s));
    LAUNCH(c, (launch_act_bwd<T, T, T>(1, d1, static_cast<const T*>(tr.t1), d1, M, d, g_b0, s)));
    LAUNCH(c, launch_time_l0_bwd<T>(d1, tr.times, g_w0, M, d, s));
    tr.time_valid = false;
    return TIM_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// encoder: training forward
// ------------------------------------------------------------------------------------------------------------------
template <typename T>
size_t layout_tape(tim_ctx* c, int B, const QueryPlan& qp, uint8_t* base) {
    constexpr bool f32 = std::is_same<T, float>::value;
    tim_train_state& tr = *c->train;
    const tim_config& g = c->cfg;
    const int d = c->d, E = c->E, FF = c->FF;
    const size_t Mf = static_cast<size_t>(B) * c->Ft, Mq = static_cast<size_t>(B) * qp.Qt, M = Mf + Mq;
    Arena a{base};
    a.take(&tr.embv_pre, static_cast<size_t>(B) * c->Fv * d * sizeof(float));
    a.take(&tr.emba_pre, static_cast<size_t>(B) * c->Fa * d * sizeof(float));
    a.take(&tr.embv_act, static_cast<size_t>(B) * c->Fv * d * sizeof(float));
    a.take(&tr.emba_act, static_cast<size_t>(B) * c->Fa * d * sizeof(float));
    a.take(&tr.visT, static_cast<size_t>(B) * c->Fv * g.vis_dim * sizeof(T));
    a.take(&tr.audT, static_cast<size_t>(B) * c->Fa * g.aud_dim * sizeof(T));
    tr.layers.resize(c->L);
    for (int l = 0; l < c->L; ++l) {
        LayerTape& t = tr.layers[l];
        a.take(&t.xin, M * E * sizeof(T));
        a.take(&t.qkv, M * 3 * E * sizeof(T));
        a.take(&t.att, M * E * sizeof(T));
        a.take(&t.z1, M * E * sizeof(float));
        a.take(&t.x1, M * E * sizeof(T));
        a.take(&t.u, M * FF * sizeof(T));
        a.take(&t.hid, M * FF * sizeof(T));
        a.take(&t.z2, M * E * sizeof(float));
        a.take(&t.st1, f32 ? 0 : M * sizeof(float2));
        a.take(&t.st2, f32 ? 0 : M * sizeof(float2));
    }
    // 16-bit: the fp32 tokens (residual of layer 0) live next to their 16-bit copy layers[0].xin; fp32: they ARE layers[0].xin
    a.take(&tr.tok32, f32 ? 0 : M * E * sizeof(float));
    if (f32) tr.tok32 = c->L ? static_cast<float*>(tr.layers[0].xin) : nullptr;
    // heads read LN2 of the last layer: the query rows in the operand type (fp32 mode: all rows, the feature rows are `feats`)
    a.take(&tr.xh, (f32 ? M : Mq) * E * sizeof(T));
    const size_t rv = static_cast<size_t>(B) * qp.Qv, ra = static_cast<size_t>(B) * qp.Qa;
    const bool det = g.variant == TIM_DETECTION;
    a.take(&tr.reg_v.r1, det ? rv * (E / 2) * sizeof(T) : 0);
    a.take(&tr.reg_v.r2, det ? rv * (E / 2) * sizeof(T) : 0);
    a.take(&tr.reg_v.y, det ? rv * 2 * sizeof(float) : 0);
    a.take(&tr.reg_a.r1, det ? ra * (E / 2) * sizeof(T) : 0);
    a.take(&tr.reg_a.r2, det ? ra * (E / 2) * sizeof(T) : 0);
    a.take(&tr.reg_a.y, det ? ra * 2 * sizeof(float) : 0);
    return a.off;
}

template <typename T>
int encoder_train_fwd(tim_ctx* c, const float* vis, const float* aud, const float* te, int B, int T_, int Qv, int Qa, const tim_outputs* o,
                      cudaStream_t s) {
    constexpr bool f32 = std::is_same<T, float>::value;
    tim_train_state& tr = *c->train;
    const tim_config& g = c->cfg;
    const int d = c->d, E = c->E, FF = c->FF;
    if (c->L < 1) return c->fail(TIM_ERR_INVALID, "the training leg needs num_layers >= 1");
    if (E > 2048 || d > 2048) return c->fail(TIM_ERR_INVALID, "the training leg supports 2 * d_model <= 2048");
    QueryPlan qp;
    TIM_TRY(plan_queries(c, T_, Qv, Qa, &qp));
    const int Ft = c->Ft, Qt = qp.Qt;
    const size_t Mf = static_cast<size_t>(B) * Ft, Mq = static_cast<size_t>(B) * Qt, M = Mf + Mq;
    if (M > 0x7fffffffull) return c->fail(TIM_ERR_INVALID, "too many token rows (%zu)", M);
    tr.enc_valid = false;
    const size_t need = layout_tape<T>(c, B, qp, nullptr);
    TIM_TRY(ensure_tape(c, &tr.mem, &tr.mem_bytes, need));
    layout_tape<T>(c, B, qp, tr.mem);
    tr.B = B; tr.T = T_; tr.Qv = Qv; tr.Qa = Qa; tr.qp = qp;

    // ---- embedders: Linear (pre-activation kept) -> GELU -> (LayerNorm inside the assembly) ----
    const int Mv = B * c->Fv, Ma = B * c->Fa;
    // dropout configuration of this forward (all zero unless tim_set_dropout was called)
    const float p_feat = tr.p_feat, p_seq = tr.p_seq, p_enc = tr.p_enc;
    const uint32_t seed = tr.drop_seed;
    tr.tape_p_seq = p_seq; tr.tape_p_enc = p_enc; tr.tape_seed = seed;
    if (p_enc > 0.0f && !f32 && !(c->attn_version >= 2 && attention_umma_supported(c->Ft, c->hd) && attention_bwd_umma_supported(c->Ft, c->hd)))
        return c->fail(TIM_ERR_INVALID, "attention dropout in the 16-bit modes needs head_dim 64 or 128 and the tcgen05 attention kernels (head_dim %d)", c->hd);
    auto embed = [&](const float* x, void* xT, const LinearW& w, int rows, int dim, float* pre, float* act, uint32_t site) -> int {
        if (!x) return c->fail(TIM_ERR_INVALID, "input features are NULL");
        if (p_feat > 0.0f) {       // feat_drop (encodings.py:141,149): fused into the operand cast
            if (dim & 1) return c->fail(TIM_ERR_INVALID, "feature dropout needs an even feature width");
            LAUNCH(c, launch_drop_cast<T>(x, static_cast<T*>(xT), static_cast<size_t>(rows) * dim, make_drop_site(p_feat, seed, site, 0), s));
        } else if constexpr (f32) CU_OK(c, cudaMemcpyAsync(xT, x, static_cast<size_t>(rows) * dim * sizeof(float), cudaMemcpyDeviceToDevice, s));
        else LAUNCH(c, launch_cast<T>(x, static_cast<T*>(xT), rows, dim, 0, 1.0f, s));
        TIM_TRY(run_linear<T>(c, xT, dim, w, plain_rows(rows), epi(pre, d, true, ACT_NONE), s));
        LAUNCH(c, launch_gelu_fwd<float>(pre, act, static_cast<size_t>(rows) * d, s));
        return TIM_OK;
    };
    if (c->Fv) TIM_TRY(embed(vis, tr.visT, c->emb_v, Mv, g.vis_dim, tr.embv_pre, tr.embv_act, DROP_FEAT_VIS));
    if (c->Fa) TIM_TRY(embed(aud, tr.audT, c->emb_a, Ma, g.aud_dim, tr.emba_pre, tr.emba_act, DROP_FEAT_AUD));
    AssembleParams ap;
    std::memset(&ap, 0, sizeof(ap));
    ap.B = B; ap.d = d; ap.T = T_; ap.Fv = c->Fv; ap.Fa = c->Fa;
    ap.emb_v = tr.embv_act; ap.emb_a = tr.emba_act;
    ap.ln_v_g = c->lnv_g; ap.ln_v_b = c->lnv_b; ap.ln_a_g = c->lna_g; ap.ln_a_b = c->lna_b;
    ap.mod_v = g.input_modality == TIM_AUDIO_VISUAL ? c->mod_v : nullptr;
    ap.mod_a = g.input_modality == TIM_AUDIO_VISUAL ? c->mod_a : nullptr;
    ap.te = te; ap.n_groups = qp.n_groups;
    for (int i = 0; i < qp.n_groups; ++i) ap.groups[i] = qp.groups[i];
    ap.Qt = Qt; ap.x32 = tr.tok32; ap.x16 = f32 ? nullptr : tr.layers[0].xin;
    LAUNCH_C(c, 3, 0.0, s, launch_assemble<T>(ap, s));
    if (p_seq > 0.0f)          // seq_drop on the assembled tokens (encodings.py:177, 250), flat index over the two-stream rows
        LAUNCH(c, launch_drop_apply<T>(tr.tok32, f32 ? static_cast<T*>(nullptr) : static_cast<T*>(tr.layers[0].xin), M * E, make_drop_site(p_seq, seed, DROP_SEQ, 0), s));

    const int Mi = static_cast<int>(M);
    const double attn_flops = 4.0 * E * (static_cast<double>(Ft) * Ft + static_cast<double>(Qt) * (Ft + 1)) * B;
    // with dropout1 / dropout2 the sub-layer output is written WITHOUT its residual into a scratch buffer and a row kernel forms
    // z = R(residual) + output o mask (the GEMM epilogues stay the ones the inference forward validated)
    float* sub32 = nullptr;
    if (p_enc > 0.0f) {
        TIM_TRY(ensure_ws(c, M * E * sizeof(float) + 256));
        sub32 = reinterpret_cast<float*>(c->ws);
    }
    for (int l = 0; l < c->L; ++l) {
        Layer& ly = c->layers[l];
        LayerTape& t = tr.layers[l];
        const DropSite d_attn = make_drop_site(p_enc, seed, DROP_ATTN, l), d_sub1 = make_drop_site(p_enc, seed, DROP_SUB1, l);
        const DropSite d_ffn = make_drop_site(p_enc, seed, DROP_FFN, l), d_sub2 = make_drop_site(p_enc, seed, DROP_SUB2, l);
        const Layer* lp = l > 0 ? &c->layers[l - 1] : nullptr;
        const LayerTape* tp = l > 0 ? &tr.layers[l - 1] : nullptr;
        if constexpr (!f32) {
            if (l > 0) LAUNCH_C(c, 2, 0.0, s, launch_layernorm<T>(tp->z2, E, lp->n2g, lp->n2b, nullptr, 0, static_cast<T*>(t.xin), E, Mi, E, s, tp->st2));
        }
        TIM_TRY(run_linear<T>(c, t.xin, E, ly.in_proj, plain_rows(Mi), epi(t.qkv, 3 * E, f32), s));
        if constexpr (f32) {
            LAUNCH_C(c, 1, attn_flops, s, launch_attention_simt(static_cast<const float*>(t.qkv), static_cast<float*>(t.att), B, Ft, Qt, c->H, c->hd, s, d_attn));
        } else {
            AttnUmmaParams attn_p;
            bool attn_umma = false;
            TIM_TRY(prepare_attention<T>(c, &attn_p, &attn_umma, static_cast<const T*>(t.qkv), static_cast<T*>(t.att), B, Ft, Qt));
            attn_p.drop = d_attn;
            if (d_attn.thr && 1ull * B * c->H * (Ft + Qt) * DROP_ATTN_KW > 0xffffffffull) return c->fail(TIM_ERR_INVALID, "attention dropout: batch too large for the 32-bit element index");
            if (attn_umma) LAUNCH_C(c, 1, attn_flops, s, launch_attention_tc<T>(c, attn_p, s));
            else LAUNCH_C(c, 1, attn_flops, s, launch_attention_mma<T>(static_cast<const T*>(t.qkv), static_cast<T*>(t.att), B, Ft, Qt, c->H, c->hd, s));
        }
        if constexpr (f32) {
            if (d_sub1.thr) {
                TIM_TRY(run_linear<T>(c, t.att, E, ly.out_proj, plain_rows(Mi), epi(sub32, E, true), s));
                LAUNCH(c, launch_residual_drop(sub32, static_cast<const float*>(t.xin), nullptr, nullptr, nullptr, t.z1, Mi, E, d_sub1, s));
            } else {
                TIM_TRY(run_linear<T>(c, t.att, E, ly.out_proj, plain_rows(Mi), epi(t.z1, E, true, ACT_NONE, static_cast<const float*>(t.xin), E), s));
            }
            LAUNCH_C(c, 2, 0.0, s, launch_layernorm<T>(t.z1, E, ly.n1g, ly.n1b, static_cast<float*>(t.x1), E, static_cast<T*>(nullptr), 0, Mi, E, s));
            TIM_TRY(run_linear<T>(c, t.x1, E, ly.lin1, plain_rows(Mi), epi(t.u, FF, true, ACT_NONE), s));
            LAUNCH(c, launch_gelu_fwd<T>(static_cast<const T*>(t.u), static_cast<T*>(t.hid), M * FF, s, d_ffn));
            if (d_sub2.thr) {
                TIM_TRY(run_linear<T>(c, t.hid, FF, ly.lin2, plain_rows(Mi), epi(sub32, E, true), s));
                LAUNCH(c, launch_residual_drop(sub32, static_cast<const float*>(t.x1), nullptr, nullptr, nullptr, t.z2, Mi, E, d_sub2, s));
            } else {
                TIM_TRY(run_linear<T>(c, t.hid, FF, ly.lin2, plain_rows(Mi), epi(t.z2, E, true, ACT_NONE, static_cast<const float*>(t.x1), E), s));
            }
            float* nxt = l < c->L - 1 ? static_cast<float*>(tr.layers[l + 1].xin) : static_cast<float*>(tr.xh);
            LAUNCH_C(c, 2, 0.0, s, launch_layernorm<T>(t.z2, E, ly.n2g, ly.n2b, nxt, E, static_cast<T*>(nullptr), 0, Mi, E, s));
        } else {
            if (d_sub1.thr) {
                TIM_TRY(run_linear<T>(c, t.att, E, ly.out_proj, plain_rows(Mi), epi(sub32, E, true), s));
                LAUNCH(c, launch_residual_drop(sub32, l > 0 ? tp->z2 : tr.tok32, l > 0 ? tp->st2 : nullptr, l > 0 ? lp->n2g : nullptr,
                                               l > 0 ? lp->n2b : nullptr, t.z1, Mi, E, d_sub1, s));
            } else {
                Epilogue e1 = epi(t.z1, E, true, ACT_NONE, l > 0 ? tp->z2 : tr.tok32, E);
                if (l > 0) { e1.rstats = tp->st2; e1.rgamma = lp->n2g; e1.rbeta = lp->n2b; }
                TIM_TRY(run_linear<T>(c, t.att, E, ly.out_proj, plain_rows(Mi), e1, s));
            }
            LAUNCH_C(c, 2, 0.0, s, launch_layernorm<T>(t.z1, E, ly.n1g, ly.n1b, nullptr, 0, static_cast<T*>(t.x1), E, Mi, E, s, t.st1));
            if (!d_ffn.thr && c->train_fuse && fused_act_ok(c, ly.lin1, Mi, FF)) {
                // linear1 writes the pre-activation u (kept for the backward) AND hid = GELU(u) from one epilogue (gemm_umma2.cu mode 8)
                Epilogue eu = epi(t.u, FF, false, ACT_GELU);
                eu.out_act = t.hid;
                TIM_TRY(run_linear<T>(c, t.x1, E, ly.lin1, plain_rows(Mi), eu, s));
            } else {
                TIM_TRY(run_linear<T>(c, t.x1, E, ly.lin1, plain_rows(Mi), epi(t.u, FF, false, ACT_NONE), s));
                LAUNCH(c, launch_gelu_fwd<T>(static_cast<const T*>(t.u), static_cast<T*>(t.hid), M * FF, s, d_ffn));
            }
            if (d_sub2.thr) {
                TIM_TRY(run_linear<T>(c, t.hid, FF, ly.lin2, plain_rows(Mi), epi(sub32, E, true), s));
                LAUNCH(c, launch_residual_drop(sub32, t.z1, t.st1, ly.n1g, ly.n1b, t.z2, Mi, E, d_sub2, s));
            } else {
                Epilogue e2 = epi(t.z2, E, true, ACT_NONE, t.z1, E);
                e2.rstats = t.st1; e2.rgamma = ly.n1g; e2.rbeta = ly.n1b;
                TIM_TRY(run_linear<T>(c, t.hid, FF, ly.lin2, plain_rows(Mi), e2, s));
            }
            if (l == c->L - 1 && Mq)
                LAUNCH_C(c, 2, 0.0, s, launch_layernorm<T>(t.z2 + Mf * E, E, ly.n2g, ly.n2b, nullptr, 0, static_cast<T*>(tr.xh), E, static_cast<int>(Mq), E, s));
        }
    }
    // feature rows returned to the caller (tim.py:172)
    if (o->feats && Mf) {
        if constexpr (f32) {
            CU_OK(c, cudaMemcpyAsync(o->feats, tr.xh, Mf * E * sizeof(float), cudaMemcpyDeviceToDevice, s));
        } else {
            Layer& ly = c->layers[c->L - 1];
            LAUNCH_C(c, 2, 0.0, s, launch_layernorm<T>(tr.layers[c->L - 1].z2, E, ly.n2g, ly.n2b, o->feats, E, static_cast<T*>(nullptr), 0, static_cast<int>(Mf), E, s));
        }
    }
    // ---- heads ----
    const uint8_t* qbase = reinterpret_cast<const uint8_t*>(tr.xh) + (f32 ? Mf * E * sizeof(T) : 0);
    auto cls = [&](const LinearW& w, int off, int Q, float* out) -> int {
        if (!w.N || Q <= 0) return TIM_OK;
        if (!out) return c->fail(TIM_ERR_INVALID, "output pointer for a %d-class head is NULL", w.N);
        return run_linear<T>(c, qbase, E, w, group_rows(B, Qt, off, Q), epi(out, w.N, true), s);
    };
    auto reg = [&](RegHead& r, RegTape& rt, int off, int Q, float* out) -> int {
        if (Q <= 0) return TIM_OK;
        if (!out) return c->fail(TIM_ERR_INVALID, "regression output pointer is NULL");
        TIM_TRY(run_linear<T>(c, qbase, E, r.l0, group_rows(B, Qt, off, Q), epi(rt.r1, E / 2, f32, ACT_RELU), s));
        TIM_TRY(run_linear<T>(c, rt.r1, E / 2, r.l2, plain_rows(B * Q), epi(rt.r2, E / 2, f32, ACT_RELU), s));
        LAUNCH(c, launch_reg_final<T>(static_cast<const T*>(rt.r2), E / 2, r.w4, r.b4, out, B * Q, E / 2, s));
        CU_OK(c, cudaMemcpyAsync(rt.y, out, static_cast<size_t>(B) * Q * 2 * sizeof(float), cudaMemcpyDeviceToDevice, s));
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
            TIM_TRY(reg(c->reg_v, tr.reg_v, qp.off_action, qp.Qv, o->reg_visual));
        }
        if (qp.Qa > 0) {
            TIM_TRY(cls(c->h_audio, qp.off_audio, qp.Qa, o->audio));
            TIM_TRY(reg(c->reg_a, tr.reg_a, qp.off_audio, qp.Qa, o->reg_audio));
        }
    }
    tr.enc_valid = true;
    return TIM_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// encoder: backward. go = gradients of the outputs (device fp32, NULL where the loss does not touch an output),
// d_te [B, T, d] receives the gradient w.r.t. the time encodings; parameter gradients accumulate (+=) into the bound destinations.
// ------------------------------------------------------------------------------------------------------------------
template <typename T>
int encoder_train_bwd(tim_ctx* c, const tim_outputs* go, float* d_te, cudaStream_t s) {
    constexpr bool f32 = std::is_same<T, float>::value;
    tim_train_state& tr = *c->train;
    if (!tr.enc_valid) return c->fail(TIM_ERR_INVALID, "tim_encoder_bwd without a matching tim_encoder_fwd_train");
    const tim_config& g = c->cfg;
    const int d = c->d, E = c->E, FF = c->FF, B = tr.B, T_ = tr.T;
    const QueryPlan& qp = tr.qp;
    const int Ft = c->Ft, Qt = qp.Qt;
    const size_t Mf = static_cast<size_t>(B) * Ft, Mq = static_cast<size_t>(B) * Qt, M = Mf + Mq;
    const int Mi = static_cast<int>(M);
    const int Mv = B * c->Fv, Ma = B * c->Fa;

    // ---- workspace ----
    const int qmax = qp.Qv > qp.Qa ? qp.Qv : qp.Qa;
    const size_t hrows = static_cast<size_t>(B) * qmax;
    int cmax = 8;
    for (int n : {g.n_verb, g.n_noun, g.n_action, g.n_audio}) if (n > cmax) cmax = n;
    const int cpmax = (cmax + 7) & ~7;
    const float p_seq = tr.tape_p_seq, p_enc = tr.tape_p_enc;
    const uint32_t seed = tr.tape_seed;
    float *g32, *dxq, *demb_v, *demb_a; T *g16, *dh, *da, *dqkv, *dY16, *xg, *dr1, *dr2, *dpre; void* astats;
    for (int pass = 0; pass < 2; ++pass) {
        Arena a{pass ? c->ws : nullptr};
        a.take(&g32, M * E * sizeof(float));
        // operand copy of dz: the 16-bit copy, or - fp32 mode with dropout1 / dropout2 - the masked fp32 copy (dz o mask differs from dz)
        a.take(&g16, (f32 && p_enc <= 0.0f) ? 0 : M * E * sizeof(T));
        a.take(&dh, M * FF * sizeof(T));
        a.take(&da, M * E * sizeof(T));
        a.take(&dqkv, M * 3 * E * sizeof(T));
        a.take(&astats, f32 ? 0 : attention_bwd_stats_bytes(B, Ft, Qt, c->H));
        a.take(&dY16, f32 ? 0 : hrows * cpmax * sizeof(T));
        a.take(&xg, hrows * E * sizeof(T));
        a.take(&dxq, hrows * E * sizeof(float));
        const bool det = g.variant == TIM_DETECTION;
        a.take(&dr1, det ? hrows * (E / 2) * sizeof(T) : 0);
        a.take(&dr2, det ? hrows * (E / 2) * sizeof(T) : 0);
        a.take(&demb_v, static_cast<size_t>(Mv) * d * sizeof(float));
        a.take(&demb_a, static_cast<size_t>(Ma) * d * sizeof(float));
        a.take(&dpre, f32 ? 0 : static_cast<size_t>(Mv > Ma ? Mv : Ma) * d * sizeof(T));
        if (!pass) TIM_TRY(ensure_ws(c, a.off));
    }
    const bool own_copy = !f32 || p_enc > 0.0f;
    const T* gop = own_copy ? g16 : reinterpret_cast<const T*>(g32);   // dz (o dropout mask) as the GEMM operand
    T* g16w = own_copy ? g16 : nullptr;

    // ---- seed: d L / d LN2(z2_last): feature rows from `feats`, query rows from the heads ----
    CU_OK(c, cudaMemsetAsync(g32, 0, M * E * sizeof(float), s));
    if (go->feats && Mf) CU_OK(c, cudaMemcpyAsync(g32, go->feats, Mf * E * sizeof(float), cudaMemcpyDeviceToDevice, s));
    const T* xq = reinterpret_cast<const T*>(reinterpret_cast<const uint8_t*>(tr.xh) + (f32 ? Mf * E * sizeof(T) : 0));
    float* gq = g32 + Mf * E;

    auto cls_bwd = [&](const LinearW& w, int off, int Q, const float* dlog) -> int {
        if (!w.N || Q <= 0 || !dlog) return TIM_OK;
        const int rows = B * Q, C = w.N;
        float *gw, *gb;
        TIM_TRY(grad_dst(c, w.w, &gw)); TIM_TRY(grad_dst(c, w.bias, &gb));
        LAUNCH(c, launch_colsum<float>(dlog, C, 1, 0, 0, rows, 0, C, gb, s));
        const T* dY; int ldy;
        if constexpr (f32) { dY = dlog; ldy = C; }
        else { LAUNCH(c, launch_cast_pad<T>(dlog, dY16, rows, C, w.Np, s)); dY = dY16; ldy = w.Np; }
        LAUNCH(c, launch_gather_group<T>(xq, xg, B, Qt, off, Q, E, s));
        TIM_TRY(run_wgrad<T>(c, dY, ldy, xg, E, gw, E, rows, C, E, s));
        TIM_TRY(run_dgrad<T>(c, dY, w, rows, epi(dxq, E, true), s));
        LAUNCH(c, launch_scatter_add_group(dxq, gq, B, Qt, off, Q, E, s));
        return TIM_OK;
    };
    auto reg_bwd = [&](RegHead& r, RegTape& rt, int off, int Q, const float* dout) -> int {
        if (Q <= 0 || !dout) return TIM_OK;
        const int rows = B * Q, Eh = E / 2;
        float *gw4, *gb4, *gw2, *gb2, *gw0, *gb0;
        TIM_TRY(grad_dst(c, r.w4, &gw4)); TIM_TRY(grad_dst(c, r.b4, &gb4));
        TIM_TRY(grad_dst(c, r.l2.w, &gw2)); TIM_TRY(grad_dst(c, r.l2.bias, &gb2));
        TIM_TRY(grad_dst(c, r.l0.w, &gw0)); TIM_TRY(grad_dst(c, r.l0.bias, &gb0));
        LAUNCH(c, launch_reg_final_bwd<T>(dout, rt.y, static_cast<const T*>(rt.r2), r.w4, gw4, gb4, dr2, gb2, rows, Eh, s));
        TIM_TRY(run_wgrad<T>(c, dr2, Eh, rt.r1, Eh, gw2, Eh, rows, Eh, Eh, s));
        TIM_TRY(run_dgrad<T>(c, dr2, r.l2, rows, epi(dr1, Eh, f32), s));
        LAUNCH(c, (launch_act_bwd<T, T, T>(1, dr1, static_cast<const T*>(rt.r1), dr1, rows, Eh, gb0, s)));
        LAUNCH(c, launch_gather_group<T>(xq, xg, B, Qt, off, Q, E, s));
        TIM_TRY(run_wgrad<T>(c, dr1, Eh, xg, E, gw0, E, rows, Eh, E, s));
        TIM_TRY(run_dgrad<T>(c, dr1, r.l0, rows, epi(dxq, E, true), s));
        LAUNCH(c, launch_scatter_add_group(dxq, gq, B, Qt, off, Q, E, s));
        return TIM_OK;
    };
    if (g.variant == TIM_RECOGNITION) {
        if (qp.Qv > 0) {
            if (g.n_verb && qp.off_verb >= 0) TIM_TRY(cls_bwd(c->h_verb, qp.off_verb, qp.Qv, go->verb));
            if (g.n_noun && qp.off_noun >= 0) TIM_TRY(cls_bwd(c->h_noun, qp.off_noun, qp.Qv, go->noun));
            TIM_TRY(cls_bwd(c->h_action, qp.off_action, qp.Qv, go->action));
        }
        if (qp.Qa > 0) TIM_TRY(cls_bwd(c->h_audio, qp.off_audio, qp.Qa, go->audio));
    } else {
        if (qp.Qv > 0) {
            TIM_TRY(cls_bwd(c->h_verb, qp.off_action, qp.Qv, go->verb));
            TIM_TRY(cls_bwd(c->h_noun, qp.off_action, qp.Qv, go->noun));
            TIM_TRY(cls_bwd(c->h_action, qp.off_action, qp.Qv, go->action));
            TIM_TRY(reg_bwd(c->reg_v, tr.reg_v, qp.off_action, qp.Qv, go->reg_visual));
        }
        if (qp.Qa > 0) {
            TIM_TRY(cls_bwd(c->h_audio, qp.off_audio, qp.Qa, go->audio));
            TIM_TRY(reg_bwd(c->reg_a, tr.reg_a, qp.off_audio, qp.Qa, go->reg_audio));
        }
    }

    // ---- encoder layers, last to first ----
    const float qscale = static_cast<float>(std::pow(static_cast<double>(c->hd), -0.5));
    const double attn_bwd_flops = 2.5 * 4.0 * E * (static_cast<double>(Ft) * Ft + static_cast<double>(Qt) * (Ft + 1)) * B;
    for (int l = c->L - 1; l >= 0; --l) {
        Layer& ly = c->layers[l];
        LayerTape& t = tr.layers[l];
        float *g_n2g, *g_n2b, *g_w2, *g_b2, *g_w1, *g_b1, *g_n1g, *g_n1b, *g_wo, *g_bo, *g_wi, *g_bi;
        TIM_TRY(grad_dst(c, ly.n2g, &g_n2g)); TIM_TRY(grad_dst(c, ly.n2b, &g_n2b));
        TIM_TRY(grad_dst(c, ly.lin2.w, &g_w2)); TIM_TRY(grad_dst(c, ly.lin2.bias, &g_b2));
        TIM_TRY(grad_dst(c, ly.lin1.w, &g_w1)); TIM_TRY(grad_dst(c, ly.lin1.bias, &g_b1));
        TIM_TRY(grad_dst(c, ly.n1g, &g_n1g)); TIM_TRY(grad_dst(c, ly.n1b, &g_n1b));
        TIM_TRY(grad_dst(c, ly.out_proj.w, &g_wo)); TIM_TRY(grad_dst(c, ly.out_proj.bias, &g_bo));
        TIM_TRY(grad_dst(c, ly.in_proj.w, &g_wi)); TIM_TRY(grad_dst(c, ly.in_proj.bias, &g_bi));
        const DropSite d_attn = make_drop_site(p_enc, seed, DROP_ATTN, l), d_sub1 = make_drop_site(p_enc, seed, DROP_SUB1, l);
        const DropSite d_ffn = make_drop_site(p_enc, seed, DROP_FFN, l), d_sub2 = make_drop_site(p_enc, seed, DROP_SUB2, l);
        // norm2 backward: g32 = d / d LN2(z2) -> dz2 (in place, + operand copy); dz2 (o dropout2's mask) is the gradient of linear2's output
        LAUNCH_C(c, 2, 0.0, s, launch_ln_bwd<T>(g32, E, t.z2, E, ly.n2g, g16w, E, g_n2g, g_n2b, g_b2, Mi, E, s, d_sub2));
        TIM_TRY(run_wgrad<T>(c, gop, E, t.hid, FF, g_w2, FF, Mi, E, FF, s));
        bool fused_dact = false;
        if constexpr (!f32) fused_dact = !d_ffn.thr && c->train_fuse && ly.lin2.has_tmBt2 && fused_act_ok(c, dgrad_view(c, ly.lin2), Mi, FF);
        if (fused_dact) {
            // d / d u = (dz2 W2) o GELU'(u) straight from the dgrad epilogue (gemm_umma2.cu mode 9); the bias gradient is its column sum
            Epilogue ed = epi(dh, FF, false);
            ed.dact_of = t.u;
            TIM_TRY(run_dgrad<T>(c, gop, ly.lin2, Mi, ed, s));
            LAUNCH(c, launch_colsum<T>(static_cast<const T*>(dh), FF, 1, 0, 0, Mi, 0, FF, g_b1, s));
        } else {
            TIM_TRY(run_dgrad<T>(c, gop, ly.lin2, Mi, epi(dh, FF, f32), s));
            LAUNCH(c, (launch_act_bwd<T, T, T>(0, dh, static_cast<const T*>(t.u), dh, Mi, FF, g_b1, s, d_ffn)));
        }
        TIM_TRY(run_wgrad<T>(c, dh, FF, t.x1, E, g_w1, E, Mi, FF, E, s));
        // d / d x1 = du W1 + dz2 (the residual branch), in place in g32
        TIM_TRY(run_dgrad<T>(c, dh, ly.lin1, Mi, epi(g32, E, true, ACT_NONE, g32, E), s));
        LAUNCH_C(c, 2, 0.0, s, launch_ln_bwd<T>(g32, E, t.z1, E, ly.n1g, g16w, E, g_n1g, g_n1b, g_bo, Mi, E, s, d_sub1));
        TIM_TRY(run_wgrad<T>(c, gop, E, t.att, E, g_wo, E, Mi, E, E, s));
        TIM_TRY(run_dgrad<T>(c, gop, ly.out_proj, Mi, epi(da, E, f32), s));
        if constexpr (f32) {
            CU_OK(c, cudaMemsetAsync(dqkv, 0, M * 3 * E * sizeof(float), s));
            LAUNCH_C(c, 9, attn_bwd_flops, s, launch_attention_bwd_simt(static_cast<const float*>(t.qkv), da, dqkv, B, Ft, Qt, c->H, c->hd, qscale, s, d_attn));
        } else {
            TIM_TRY(run_attention_bwd<T>(c, static_cast<const T*>(t.qkv), da, dqkv, astats, B, Ft, Qt, qscale, attn_bwd_flops, s, d_attn));
        }
        LAUNCH(c, launch_colsum<T>(dqkv, 3 * E, 1, 0, 0, Mi, 0, 3 * E, g_bi, s));
        TIM_TRY(run_wgrad<T>(c, dqkv, 3 * E, t.xin, E, g_wi, E, Mi, 3 * E, E, s));
        // d / d x_l = dqkv Win + dz1 (the residual branch), in place in g32: the input of the previous layer's norm2 backward
        TIM_TRY(run_dgrad<T>(c, dqkv, ly.in_proj, Mi, epi(g32, E, true, ACT_NONE, g32, E), s));
    }

    // ---- token assembly backward (encodings.py:181-251) ----
    AssembleBwdParams bp;
    std::memset(&bp, 0, sizeof(bp));
    bp.B = B; bp.d = d; bp.T = T_; bp.Fv = c->Fv; bp.Fa = c->Fa; bp.Qt = Qt;
    bp.dtok = g32; bp.dte = d_te; bp.demb_v = demb_v; bp.demb_a = demb_a;
    bp.n_groups = qp.n_groups;
    for (int i = 0; i < qp.n_groups; ++i) bp.groups[i] = qp.groups[i];
    if (!d_te) return c->fail(TIM_ERR_INVALID, "tim_encoder_bwd: d_time_enc is NULL");
    if (p_seq > 0.0f)          // seq_drop: the gradient w.r.t. the assembled tokens is the gradient w.r.t. the dropped tokens o mask
        LAUNCH(c, launch_drop_apply<T>(g32, static_cast<T*>(nullptr), M * E, make_drop_site(p_seq, seed, DROP_SEQ, 0), s));
    LAUNCH_C(c, 3, 0.0, s, launch_assemble_bwd(bp, s));
    {
        int start = 0;
        for (int i = 0; i < qp.n_groups; ++i) {
            const TokenGroup& tg = qp.groups[i];
            float* gc;
            TIM_TRY(grad_dst(c, tg.cls, &gc));
            LAUNCH(c, launch_colsum<float>(gq, E, B, Qt, start, tg.count, 0, d, gc, s));
            if (tg.mod) {
                float* gm;
                TIM_TRY(grad_dst(c, tg.mod, &gm));
                LAUNCH(c, launch_colsum<float>(gq, E, B, Qt, start, tg.count, 0, E, gm, s));
            }
            start += tg.count;
        }
        if (g.input_modality == TIM_AUDIO_VISUAL) {
            float *gmv, *gma;
            TIM_TRY(grad_dst(c, c->mod_v, &gmv)); TIM_TRY(grad_dst(c, c->mod_a, &gma));
            LAUNCH(c, launch_colsum<float>(g32, E, B, Ft, 0, c->Fv, 0, E, gmv, s));
            LAUNCH(c, launch_colsum<float>(g32, E, B, Ft, c->Fv, c->Fa, 0, E, gma, s));
        }
    }
    auto emb_bwd = [&](const LinearW& w, float* lng, float* lnb, float* demb, const float* act, const float* pre, const void* xT, int rows,
                       int dim) -> int {
        float *g_lg, *g_lb, *gw, *gb;
        TIM_TRY(grad_dst(c, lng, &g_lg)); TIM_TRY(grad_dst(c, lnb, &g_lb));
        TIM_TRY(grad_dst(c, w.w, &gw)); TIM_TRY(grad_dst(c, w.bias, &gb));
        LAUNCH(c, launch_ln_bwd<T>(demb, d, act, d, lng, static_cast<T*>(nullptr), 0, g_lg, g_lb, nullptr, rows, d, s));
        const T* dop;
        if constexpr (f32) { LAUNCH(c, (launch_act_bwd<float, float, float>(0, demb, pre, demb, rows, d, gb, s))); dop = demb; }
        else { LAUNCH(c, (launch_act_bwd<float, float, T>(0, demb, pre, dpre, rows, d, gb, s))); dop = dpre; }
        return run_wgrad<T>(c, dop, d, xT, dim, gw, dim, rows, d, dim, s);
    };
    if (c->Fv) TIM_TRY(emb_bwd(c->emb_v, c->lnv_g, c->lnv_b, demb_v, tr.embv_act, tr.embv_pre, tr.visT, Mv, g.vis_dim));
    if (c->Fa) TIM_TRY(emb_bwd(c->emb_a, c->lna_g, c->lna_b, demb_a, tr.emba_act, tr.emba_pre, tr.audT, Ma, g.aud_dim));
    tr.enc_valid = false;
    return TIM_OK;
}

int train_ready(tim_ctx* c) {
    if (!c->train || !c->train->enabled) return c->fail(TIM_ERR_INVALID, "training is not enabled for this context (tim_train_enable)");
    TIM_TRY(check_ready(c));
    for (auto& kv : c->slots)
        if (kv.second.kind == Slot::LIN_W && !kv.second.lin->wt_set)
            return c->fail(TIM_ERR_WEIGHTS, "weight '%s' must be set again after tim_train_enable() (its transposed copy is missing)", kv.first.c_str());
    return TIM_OK;
}

}  // namespace
