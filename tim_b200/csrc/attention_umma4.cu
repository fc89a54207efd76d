// Mask-aware attention of the TIM encoder layer on tcgen05 / TMEM / TMA - fourth form of the forward kernel: DECOUPLED PIPELINE.
//
// Same math, token layout, tensor maps and reference mapping as attention_umma.cu (nn.MultiheadAttention of
// */models/helpers/transformers.py:102 under the mask of recognition/.../models/tim.py:161-166). What changed and why - the source-level
// ncu capture of attention_umma.cu (profiles/r02i_attention_umma_roles.txt) shows where a tile's 5465 cycles go: the softmax warps
// wait 43 % of the time, the epilogue warps 50 %, the MMA warp sits on q_full and the producer on q_empty. The loop that bounds it is
//     epilogue(g-1) -> stage released -> TMA load(g+1) -> S(g+1) issued -> only then P.V(g) issued -> epilogue(g) ...
// because (1) the MMA warp issues in program order (S(g+1) before P.V(g)) and blocks on the load of tile g+1 while P(g) has long
// been ready, and (2) a stage is released only when the epilogue's TMA store has read the staging tile that aliases the K_q buffer.
// Here:
//   * the MMA warp is EVENT-DRIVEN: it polls "S of the next tile can go" (Q / K_q loaded, TMEM stage drained) and "P.V of the oldest
//     tile can go" (P written) and issues whichever is ready;
//   * the output staging tile is its OWN buffer (one, shared by both stages: each epilogue warp owns a 32-row slab and waits for
//     its previous bulk store to have read it), so a stage's shared memory is free the moment P.V completes: the producer waits on
//     o_full itself and the load of tile g+2 runs under the epilogue of tile g;
//   * the epilogue releases the TMEM stage (t_empty) as soon as the second half of the O tile is in registers; the own-value rows
//     (16-byte loads issued before the waits) are parked in the staging slab in the output layout and each thread overwrites its
//     own row in place (no scratch in the stage buffers);
//   * the number of 16-key chunks is a template parameter for Ft in (96, 112] and (112, 128] (cfg2 / cfg4: 100 keys), which removes
//     the data-dependent branches from the softmax loops.
// head_dim 64 / 128 and K_f + V_f + 2 stages + staging <= 227 KB; everything else stays on attention_umma.cu.
//
// One persistent CTA per SM, 320 threads: warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 softmax, warps 6-9 epilogue.
#include <cstdio>
#include <cstdlib>

#include "kernels.h"
#include "ptx.cuh"

namespace tim {
namespace {

constexpr int A4_THREADS = 320;
constexpr int A4_BM = 128;
constexpr int A4_NST = 2;                // tiles in flight (TMEM: 2 x 256 columns)
constexpr int A4_P_BYTES = 32768;        // P tile (128 rows x up to 128 keys)
constexpr int A4_PF = 1;                 // tiles prefetched into L2 ahead of the smem loads
constexpr int A4_MIN_SMEM = 120 * 1024;  // keeps it at one CTA per SM (each CTA allocates all 512 TMEM columns)
constexpr int A4_MAX_SMEM = 232448;      // 227 KB

template <int HD> struct A4Cfg {
    static constexpr int KBOX = HD / 64;                   // 64-column (128-byte) boxes per head row
    static constexpr int Q_BYTES = KBOX * A4_BM * 128;     // Q tile (buffer A: Q -> P)
    static constexpr int A_BYTES = Q_BYTES > A4_P_BYTES ? Q_BYTES : A4_P_BYTES;
    static constexpr int STAGE_BYTES = A_BYTES + Q_BYTES;  // buffer B: own-key tile K_q
    static constexpr int TMEM_STAGE = 256;                 // [0,128): S, later O; [128,256): S_self
};

template <typename T> struct FmtOf4;
template <> struct FmtOf4<__half> { static constexpr uint32_t v = 0; };
template <> struct FmtOf4<__nv_bfloat16> { static constexpr uint32_t v = 1; };

struct Unit4 { int b, h, t_lo, t_hi; };
__device__ __forceinline__ Unit4 decode_unit4(const AttnUmmaParams& p, int u) {
    Unit4 ui;
    const int item = u / p.chunks, ch = u - item * p.chunks;
    ui.b = item / p.H; ui.h = item - ui.b * p.H;
    ui.t_lo = ch * p.tpu;
    ui.t_hi = min(ui.t_lo + p.tpu, 1 + p.tiles_q);
    return ui;
}

// non-blocking phase test, warp-uniform verdict (lanes may observe the flip at different instants: all of them must have seen it)
__device__ __forceinline__ bool mbar_test_all(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return __all_sync(0xffffffffu, done != 0);
}

// role profile (debug hook, off unless tim_debug_role_prof gave a buffer): cycles one warp of each role spends in its waits
struct ProfClock {
    bool on;
    long long t0;
    __device__ __forceinline__ void begin() { if (on) t0 = clock64(); }
    __device__ __forceinline__ void end(unsigned long long& acc) { if (on) acc += static_cast<unsigned long long>(clock64() - t0); }
};
enum { PR_TOTAL = 0, PR_PROD_KF, PR_PROD_STAGE, PR_PROD_VF, PR_MMA_IDLE, PR_MMA_TOTAL, PR_SM_WAIT_S, PR_SM_TOTAL, PR_EP_WAIT_READ, PR_EP_WAIT_P,
       PR_EP_WAIT_O, PR_EP_TOTAL, PR_TILES, PR_EP_WAIT_TMEM, PR_COUNT };

// the MMA warp's walk over (unit, tile): two of these run independently, one for the S products, one for the P.V products
struct TileIt {
    int u, t, t_lo, t_hi;
    uint32_t un;
    bool valid;
    __device__ __forceinline__ void load(const AttnUmmaParams& p) {
        valid = u < p.num_units;
        if (valid) { const Unit4 ui = decode_unit4(p, u); t_lo = ui.t_lo; t_hi = ui.t_hi; t = t_lo; }
    }
    __device__ __forceinline__ void init(const AttnUmmaParams& p) { u = blockIdx.x; un = 0; load(p); }
    __device__ __forceinline__ void next(const AttnUmmaParams& p) {
        if (++t >= t_hi) { u += gridDim.x; ++un; load(p); }
    }
    __device__ __forceinline__ bool first() const { return t == t_lo; }
    __device__ __forceinline__ bool last() const { return t == t_hi - 1; }
};

// FPC: Ft rounded up to 16 as a compile-time constant (112 / 128), 0 = run-time (p.Fp).
// DROP: attention-probability dropout compiled in (training forward with p > 0 only).
template <typename T, int HD, int FPC, bool DROP>
__global__ void __launch_bounds__(A4_THREADS, 1) attention_umma4_kernel(const __grid_constant__ AttnUmmaParams p) {
    using C = A4Cfg<HD>;
    constexpr int NST = A4_NST, KBOX = C::KBOX;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int Fp = FPC ? FPC : p.Fp;
    const int Ft = p.Ft, Qt = p.Qt;
    const uint32_t box_kv = static_cast<uint32_t>(Fp) * 128u;       // one 64-column box of K_f / V_f
    const uint32_t kv_bytes = KBOX * box_kv;
    const uint32_t kv_pad = (kv_bytes + 1023u) & ~1023u;
    const uint32_t sKF = base, sVF = base + kv_pad;
    const uint32_t stage0 = sVF + kv_pad;
    auto sQ = [&](int st) { return stage0 + static_cast<uint32_t>(st) * C::STAGE_BYTES; };
    auto sP = [&](int st) { return stage0 + static_cast<uint32_t>(st) * C::STAGE_BYTES; };                    // aliases Q
    auto sB = [&](int st) { return stage0 + static_cast<uint32_t>(st) * C::STAGE_BYTES + C::A_BYTES; };
    const uint32_t sOut = stage0 + NST * C::STAGE_BYTES;             // output staging [KBOX][128 rows][128 B], chunk-swizzled
    const uint32_t stat_base = sOut + C::Q_BYTES;                    // float [NST][2][128]: 1/l, p_self/l
    auto stat = [&](int st, int which, int row) { return stat_base + static_cast<uint32_t>(((st * 2 + which) * A4_BM + row) * 4); };
    const uint32_t bar_base = stat_base + NST * 1024;
    auto q_full = [&](int st) { return bar_base + 8u * st; };
    auto t_empty = [&](int st) { return bar_base + 8u * (NST + st); };
    auto s_full = [&](int st) { return bar_base + 8u * (2 * NST + st); };
    auto p_full = [&](int st) { return bar_base + 8u * (3 * NST + st); };
    auto o_full = [&](int st) { return bar_base + 8u * (4 * NST + st); };
    const uint32_t kf_full = bar_base + 8u * (5 * NST), kf_empty = kf_full + 8u, vf_full = kf_full + 16u, vf_empty = kf_full + 24u;
    const uint32_t tmem_slot = kf_full + 32u;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);     // warp-uniform for the compiler
    const int lane = threadIdx.x & 31;
    const int E = p.H * HD;
    const int ld = 3 * E;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tmKV); tma_prefetch_desc(&p.tmQf); tma_prefetch_desc(&p.tmQq);
        tma_prefetch_desc(&p.tmOf); tma_prefetch_desc(&p.tmOq);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int st = 0; st < NST; ++st) {
                mbar_init(q_full(st), 1); mbar_init(t_empty(st), 4); mbar_init(s_full(st), 1);
                mbar_init(p_full(st), 4); mbar_init(o_full(st), 1);
            }
            mbar_init(kf_full, 1); mbar_init(kf_empty, 1); mbar_init(vf_full, 1); mbar_init(vf_empty, 1);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_launch_dependents();          // programmatic dependent launch (ptx.cuh): the set-up above overlapped the previous kernel's tail
    pdl_wait();
    ProfClock pc;
    pc.on = p.prof != nullptr;
    unsigned long long pr[PR_COUNT] = {};
    const long long t_start = pc.on ? clock64() : 0;

    if (warp == 0) {
        // ===================== TMA producer =====================
        // warp-uniform loop state, one elected lane issues the asynchronous instructions (see attention_umma.cu)
        int pf_u = blockIdx.x, pf_t = 0;
        Unit4 pf_ui = {0, 0, 0, 0};
        bool pf_valid = pf_u < p.num_units;
        if (pf_valid) { pf_ui = decode_unit4(p, pf_u); pf_t = pf_ui.t_lo; }
        auto prefetch_next = [&]() {        // L2 prefetch iterator running ahead of the loads
            if (!pf_valid) return;
            const int h0 = pf_ui.h * HD;
            const bool tiles_too = p.pf_mode >= 2;
            if (elect_one()) {
                if (pf_t == pf_ui.t_lo) {
#pragma unroll
                    for (int j = 0; j < KBOX; ++j) {
                        tma_prefetch_l2_3d(&p.tmKV, E + h0 + 64 * j, 0, pf_ui.b);
                        tma_prefetch_l2_3d(&p.tmKV, 2 * E + h0 + 64 * j, 0, pf_ui.b);
                    }
                }
#pragma unroll
                for (int j = 0; j < KBOX; ++j) {
                    if (!tiles_too) break;
                    if (pf_t == 0) {
                        tma_prefetch_l2_3d(&p.tmQf, h0 + 64 * j, 0, pf_ui.b);
                    } else {
                        tma_prefetch_l2_3d(&p.tmQq, h0 + 64 * j, (pf_t - 1) * A4_BM, pf_ui.b);
                        tma_prefetch_l2_3d(&p.tmQq, E + h0 + 64 * j, (pf_t - 1) * A4_BM, pf_ui.b);
                        tma_prefetch_l2_3d(&p.tmQq, 2 * E + h0 + 64 * j, (pf_t - 1) * A4_BM, pf_ui.b);    // own-value rows (epilogue)
                    }
                }
            }
            if (++pf_t >= pf_ui.t_hi) {
                pf_u += gridDim.x;
                pf_valid = pf_u < p.num_units;
                if (pf_valid) { pf_ui = decode_unit4(p, pf_u); pf_t = pf_ui.t_lo; }
            }
        };
        if (p.pf_mode >= 1) for (int i = 0; i < p.pf_tiles; ++i) prefetch_next();
        uint32_t g = 0, un = 0;
        for (int u = blockIdx.x; u < p.num_units; u += gridDim.x, ++un) {
            const Unit4 ui = decode_unit4(p, u);
            pc.begin();
            mbar_wait(kf_empty, (un & 1u) ^ 1u);
            pc.end(pr[PR_PROD_KF]);
            if (elect_one()) {
                mbar_arrive_expect_tx(kf_full, kv_bytes);
#pragma unroll
                for (int j = 0; j < KBOX; ++j) tma_load_3d(sKF + j * box_kv, &p.tmKV, kf_full, E + ui.h * HD + 64 * j, 0, ui.b);
            }
            for (int t = ui.t_lo; t < ui.t_hi; ++t, ++g) {
                const int st = g % NST;
                const uint32_t ph = (g / NST) & 1u;
                if (p.pf_mode >= 1) prefetch_next();
                // the stage's buffers (Q -> P, K_q) are free once P.V of the tile that used them has completed (which implies its S
                // products have): nothing else reads them
                pc.begin();
                mbar_wait(o_full(st), ph ^ 1u);
                pc.end(pr[PR_PROD_STAGE]);
                if (elect_one()) {
                    mbar_arrive_expect_tx(q_full(st), t == 0 ? C::Q_BYTES : 2 * C::Q_BYTES);
#pragma unroll
                    for (int j = 0; j < KBOX; ++j) {
                        if (t == 0) {
                            tma_load_3d(sQ(st) + j * 16384, &p.tmQf, q_full(st), ui.h * HD + 64 * j, 0, ui.b);
                        } else {
                            tma_load_3d(sQ(st) + j * 16384, &p.tmQq, q_full(st), ui.h * HD + 64 * j, (t - 1) * A4_BM, ui.b);
                            tma_load_3d(sB(st) + j * 16384, &p.tmQq, q_full(st), E + ui.h * HD + 64 * j, (t - 1) * A4_BM, ui.b);
                        }
                    }
                }
                if (t == ui.t_lo) {
                    pc.begin();
                    mbar_wait(vf_empty, (un & 1u) ^ 1u);
                    pc.end(pr[PR_PROD_VF]);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(vf_full, kv_bytes);
#pragma unroll
                        for (int j = 0; j < KBOX; ++j)
                            tma_load_3d(sVF + j * box_kv, &p.tmKV, vf_full, 2 * E + ui.h * HD + 64 * j, 0, ui.b);
                    }
                }
            }
        }
        __syncwarp();
        if (pc.on && lane == 0) {
            pr[PR_TOTAL] = static_cast<unsigned long long>(clock64() - t_start);
            for (int i = PR_TOTAL; i <= PR_PROD_VF; ++i) atomicAdd(p.prof + blockIdx.x * 16 + i, pr[i]);
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (event-driven) =====================
        const uint32_t idesc_s = umma_idesc_f16(FmtOf4<T>::v, A4_BM, static_cast<uint32_t>(Fp));
        const uint32_t idesc_self = umma_idesc_f16(FmtOf4<T>::v, A4_BM, A4_BM);
        const uint32_t idesc_o = umma_idesc_f16(FmtOf4<T>::v, A4_BM, HD) | (1u << 16);     // B (= V_f) is MN-major
        const int ksteps_o = Fp / 16;
        TileIt is, ip;
        is.init(p); ip.init(p);
        uint32_t gs = 0, gp = 0, idle = 0;
        while (ip.valid) {
            bool progressed = false;
            if (is.valid) {
                const int st = gs % NST;
                const uint32_t ph = (gs / NST) & 1u;
                // Q (and K_q) of the tile in smem; the TMEM stage drained by the epilogue of the tile that used it; K_f of the unit in smem
                bool ready = mbar_test_all(q_full(st), ph) && mbar_test_all(t_empty(st), ph ^ 1u);
                if (ready && is.first()) ready = mbar_test_all(kf_full, is.un & 1u);
                if (ready) {
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(st * C::TMEM_STAGE);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < HD / 16; ++k) {
                            const uint64_t adesc = umma_desc_sw128(sQ(st) + (k >> 2) * 16384) + 2u * (k & 3);
                            const uint64_t bdesc = umma_desc_sw128(sKF + (k >> 2) * box_kv) + 2u * (k & 3);
                            umma_f16_ss(d_tmem, adesc, bdesc, idesc_s, k != 0 ? 1u : 0u);
                        }
                        if (is.t > 0) {
#pragma unroll
                            for (int k = 0; k < HD / 16; ++k) {
                                const uint64_t adesc = umma_desc_sw128(sQ(st) + (k >> 2) * 16384) + 2u * (k & 3);
                                const uint64_t bdesc = umma_desc_sw128(sB(st) + (k >> 2) * 16384) + 2u * (k & 3);
                                umma_f16_ss(d_tmem + 128, adesc, bdesc, idesc_self, k != 0 ? 1u : 0u);
                            }
                        }
                        umma_commit(s_full(st));
                        if (is.last()) umma_commit(kf_empty);
                    }
                    __syncwarp();
                    is.next(p); ++gs;
                    progressed = true;
                }
            }
            if (gp < gs) {
                const int st = gp % NST;
                const uint32_t ph = (gp / NST) & 1u;
                bool ready = mbar_test_all(p_full(st), ph);
                if (ready && ip.first()) ready = mbar_test_all(vf_full, ip.un & 1u);
                if (ready) {
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(st * C::TMEM_STAGE);      // O overwrites the consumed S
                    if (elect_one()) {
                        for (int k = 0; k < ksteps_o; ++k) {
                            const uint64_t adesc = umma_desc_sw128(sP(st) + (k >> 2) * 16384) + 2u * (k & 3);
                            const uint64_t bdesc = umma_desc_mn_sw128(sVF + k * 2048, box_kv);
                            umma_f16_ss(d_tmem, adesc, bdesc, idesc_o, k != 0 ? 1u : 0u);
                        }
                        umma_commit(o_full(st));
                        if (ip.last()) umma_commit(vf_empty);
                    }
                    __syncwarp();
                    ip.next(p); ++gp;
                    progressed = true;
                }
            }
            if (progressed) { if (idle) pc.end(pr[PR_MMA_IDLE]); idle = 0; }
            else { if (idle == 0) pc.begin(); if (++idle > TIM_SPIN_LIMIT) __trap(); }
        }
        __syncwarp();
        if (pc.on && lane == 0) {
            atomicAdd(p.prof + blockIdx.x * 16 + PR_MMA_IDLE, pr[PR_MMA_IDLE]);
            atomicAdd(p.prof + blockIdx.x * 16 + PR_MMA_TOTAL, static_cast<unsigned long long>(clock64() - t_start));
        }
    } else if (warp < 6) {
        // ===================== softmax (warps 2..5) =====================
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t swz = static_cast<uint32_t>(row & 7);
        uint32_t g = 0;
        for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
            const Unit4 ui = decode_unit4(p, u);
            for (int t = ui.t_lo; t < ui.t_hi; ++t, ++g) {
                const int st = g % NST;
                const uint32_t ph = (g / NST) & 1u;
                const bool qt = t > 0;
                pc.begin();
                mbar_wait(s_full(st), ph);
                pc.end(pr[PR_SM_WAIT_S]);
                tc_fence_after();
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(st * C::TMEM_STAGE);
                float s[128];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    if (c * 16 < Fp) {
                        uint32_t v[16];
                        tmem_ld_32x16(taddr + c * 16, v);
#pragma unroll
                        for (int j = 0; j < 16; ++j) s[c * 16 + j] = __uint_as_float(v[j]);
                    }
                }
                float sself = -INFINITY;
                if (qt) {
                    // own-key score = element (row, row) of S_self: this warp's 32 x 32 diagonal block, lane i keeps column i
                    uint32_t v[32];
                    tmem_ld_32x32(taddr + 128 + quarter * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int o = 16; o >= 1; o >>= 1) {
                        const bool up = (lane & o) != 0;
#pragma unroll
                        for (int j = 0; j < o; ++j) v[j] = up ? v[j + o] : v[j];
                    }
                    sself = __uint_as_float(v[0]);
                } else {
                    tmem_ld_wait();
                }
                float mx[4] = {sself, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    if (c * 16 < Fp) {
                        // only the last chunk holds padded keys (their K_f rows are zero-filled); with FPC the chunk is known
                        if (FPC ? (c == FPC / 16 - 1) : (c * 16 + 16 > Ft)) {
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                if (c * 16 + j >= Ft) s[c * 16 + j] = -INFINITY;
                        }
#pragma unroll
                        for (int j = 0; j < 16; ++j) mx[j & 3] = fmaxf(mx[j & 3], s[c * 16 + j]);
                    }
                }
                const float m = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
                float ls[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                // attention-probability dropout (training forward only): the probabilities that meet V are dropped, the normaliser is
                // not. Element index ((b H + h) S + row in clip) * DROP_ATTN_KW + key, own key at Ft (kernels.h: DropSite)
                const uint32_t drop_pair0 = ((((static_cast<uint32_t>(ui.b) * p.H + ui.h) * static_cast<uint32_t>(Ft + Qt) +
                                              static_cast<uint32_t>((qt ? Ft + (t - 1) * A4_BM : 0) + row)) * static_cast<uint32_t>(DROP_ATTN_KW)) >> 1);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    if (c * 16 < Fp) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float e = ex2_approx(s[c * 16 + j] - m);
                            s[c * 16 + j] = e;
                            ls[j & 3] += e;
                        }
                        if constexpr (DROP) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                float m0, m1;
                                drop_pair(drop_pair0 + c * 8 + j, p.drop.key, p.drop.thr, p.drop.scale, m0, m1);
                                s[c * 16 + 2 * j] *= m0; s[c * 16 + 2 * j + 1] *= m1;
                            }
                        }
#pragma unroll
                        for (int h8 = 0; h8 < 2; ++h8) {
                            const int c8 = c * 2 + h8;
                            uint4 q;
                            q.x = pack2<T>(s[c8 * 8 + 0], s[c8 * 8 + 1]); q.y = pack2<T>(s[c8 * 8 + 2], s[c8 * 8 + 3]);
                            q.z = pack2<T>(s[c8 * 8 + 4], s[c8 * 8 + 5]); q.w = pack2<T>(s[c8 * 8 + 6], s[c8 * 8 + 7]);
                            sts_u128(sP(st) + (c8 >> 3) * 16384 + row * 128 + ((static_cast<uint32_t>(c8 & 7) ^ swz) << 4), q);
                        }
                    }
                }
                const float l = (ls[0] + ls[1]) + (ls[2] + ls[3]);
                const float ps = qt ? ex2_approx(sself - m) : 0.0f;
                const float inv = 1.0f / (l + ps);
                const float m_self = DROP ? drop_one(2u * drop_pair0 + static_cast<uint32_t>(Ft), p.drop.key, p.drop.thr, p.drop.scale) : 1.0f;
                sts_f32(stat(st, 0, row), inv);
                sts_f32(stat(st, 1, row), ps * inv * m_self);
                tc_fence_before();
                fence_proxy_async_smem();          // P (generic-proxy writes) -> visible to the tensor core's operand reads
                __syncwarp();
                if (lane == 0) mbar_arrive(p_full(st));
            }
        }
        if (pc.on && warp == 2 && lane == 0) {
            atomicAdd(p.prof + blockIdx.x * 16 + PR_SM_WAIT_S, pr[PR_SM_WAIT_S]);
            atomicAdd(p.prof + blockIdx.x * 16 + PR_SM_TOTAL, static_cast<unsigned long long>(clock64() - t_start));
            atomicAdd(p.prof + blockIdx.x * 16 + PR_TILES, static_cast<unsigned long long>(g));
        }
    } else {
        // ===================== epilogue (warps 6..9) =====================
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t swz = static_cast<uint32_t>(row & 7);
        const T* qkv = static_cast<const T*>(p.qkv);
        constexpr int HALF32 = HD / 64;                // 32-column TMEM loads per half of the O tile
        uint32_t g = 0;
        for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
            const Unit4 ui = decode_unit4(p, u);
            for (int t = ui.t_lo; t < ui.t_hi; ++t, ++g) {
                const int st = g % NST;
                const uint32_t ph = (g / NST) & 1u;
                const bool qt = t > 0;
                const int row0 = qt ? (t - 1) * A4_BM : 0;
                const int nrows = min(A4_BM, (qt ? Qt : Ft) - row0);
                const int nvalid = nrows - quarter * 32;       // rows of this warp's slab that exist
                const bool vterm = qt && nvalid > 0;
                // own-value rows: coalesced 16-byte loads (8 lanes per row) issued first; their latency runs under the waits, the
                // first TMEM loads and the previous bulk store's read of the slab
                uint4 vreg[8][KBOX];
                if (vterm) {
                    const T* vbase = qkv + (static_cast<size_t>(p.B) * Ft + static_cast<size_t>(ui.b) * Qt + row0 + quarter * 32) * ld + 2 * E + ui.h * HD + (lane & 7) * 8;
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int rl = min(it * 4 + (lane >> 3), nvalid - 1);
#pragma unroll
                        for (int w = 0; w < KBOX; ++w)
                            vreg[it][w] = ldg_nc_u128_pinned(vbase + static_cast<size_t>(rl) * ld + w * 64);
                    }
                }
                pc.begin();
                mbar_wait(p_full(st), ph);                 // softmax statistics of this tile are in smem
                pc.end(pr[PR_EP_WAIT_P]);
                const float inv = lds_f32(stat(st, 0, row));
                const float wself = lds_f32(stat(st, 1, row));
                pc.begin();
                mbar_wait(o_full(st), ph);                 // P.V complete: O in TMEM
                pc.end(pr[PR_EP_WAIT_O]);
                tc_fence_after();
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(st * C::TMEM_STAGE);
                uint32_t o[HALF32][32];
#pragma unroll
                for (int i = 0; i < HALF32; ++i) tmem_ld_32x32(taddr + i * 32, o[i]);
                // this warp's slab of the staging tile is free once its previous bulk store has read it
                pc.begin();
                if (lane == 0) tma_store_wait_read<0>();
                __syncwarp();
                pc.end(pr[PR_EP_WAIT_READ]);
                if (vterm) {
                    // parked in the slab in the OUTPUT layout; after the __syncwarp each thread only touches its own row, which it
                    // overwrites with the finished output
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int rl = it * 4 + (lane >> 3);
#pragma unroll
                        for (int w = 0; w < KBOX; ++w)
                            sts_u128(sOut + w * 16384 + (quarter * 32 + rl) * 128 + ((static_cast<uint32_t>(lane & 7) ^ static_cast<uint32_t>(rl & 7)) << 4),
                                     vreg[it][w]);
                    }
                    __syncwarp();
                }
                pc.begin();
                tmem_ld_wait();
                pc.end(pr[PR_EP_WAIT_TMEM]);
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    if (half == 1) {
#pragma unroll
                        for (int i = 0; i < HALF32; ++i) tmem_ld_32x32(taddr + (HALF32 + i) * 32, o[i]);
                        tmem_ld_wait();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(t_empty(st));   // the next S into this TMEM stage may be issued
                    }
#pragma unroll
                    for (int k16 = 0; k16 < HALF32 * 2; ++k16) {
                        const int c16 = half * HALF32 * 2 + k16;       // 16-column group of the head row
                        uint32_t addr[2];
                        uint4 q4[2];
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            addr[c] = sOut + (c16 >> 2) * 16384 + row * 128 + ((static_cast<uint32_t>((c16 & 3) * 2 + c) ^ swz) << 4);
                            if (vterm) q4[c] = lds_u128(addr[c]);      // both chunks before either store (in place: kept in this order)
                        }
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            const int o0 = (k16 & 1) * 16 + 8 * c;
                            float f[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(o[k16 >> 1][o0 + j]) * inv;
                            if (vterm) {
                                const float2 v0 = unpack2<T>(q4[c].x), v1 = unpack2<T>(q4[c].y), v2 = unpack2<T>(q4[c].z), v3 = unpack2<T>(q4[c].w);
                                f[0] = fmaf(wself, v0.x, f[0]); f[1] = fmaf(wself, v0.y, f[1]);
                                f[2] = fmaf(wself, v1.x, f[2]); f[3] = fmaf(wself, v1.y, f[3]);
                                f[4] = fmaf(wself, v2.x, f[4]); f[5] = fmaf(wself, v2.y, f[5]);
                                f[6] = fmaf(wself, v3.x, f[6]); f[7] = fmaf(wself, v3.y, f[7]);
                            }
                            uint4 q;
                            q.x = pack2<T>(f[0], f[1]); q.y = pack2<T>(f[2], f[3]);
                            q.z = pack2<T>(f[4], f[5]); q.w = pack2<T>(f[6], f[7]);
                            sts_u128(addr[c], q);
                        }
                    }
                }
                fence_proxy_async_smem();                  // staged output -> visible to the TMA store
                __syncwarp();
                if (lane == 0 && nvalid > 0) {
#pragma unroll
                    for (int j = 0; j < KBOX; ++j)
                        tma_store_3d(qt ? &p.tmOq : &p.tmOf, sOut + j * 16384 + quarter * 4096, ui.h * HD + 64 * j, row0 + quarter * 32, ui.b);
                    tma_store_commit();
                }
                __syncwarp();
            }
        }
        if (lane == 0) tma_store_wait<0>();
        __syncwarp();
        if (pc.on && warp == 6 && lane == 0) {
            for (int i = PR_EP_WAIT_READ; i <= PR_EP_WAIT_O; ++i) atomicAdd(p.prof + blockIdx.x * 16 + i, pr[i]);
            atomicAdd(p.prof + blockIdx.x * 16 + PR_EP_WAIT_TMEM, pr[PR_EP_WAIT_TMEM]);
            atomicAdd(p.prof + blockIdx.x * 16 + PR_EP_TOTAL, static_cast<unsigned long long>(clock64() - t_start));
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

template <int HD> size_t smem_for4(int Fp) {
    using C = A4Cfg<HD>;
    const size_t kv_pad = (static_cast<size_t>(C::KBOX) * Fp * 128 + 1023) & ~static_cast<size_t>(1023);
    size_t b = 1024 + 2 * kv_pad + static_cast<size_t>(A4_NST) * (C::STAGE_BYTES + 1024) + C::Q_BYTES + 256;
    return b < static_cast<size_t>(A4_MIN_SMEM) ? A4_MIN_SMEM : b;
}

template <typename T, int HD, int FPC, bool DROP>
cudaError_t launch_one(const AttnUmmaParams& p, int grid, size_t smem, cudaStream_t s) {
    auto kern = attention_umma4_kernel<T, HD, FPC, DROP>;
    static SmemAttrCache cache;
    if (cudaError_t e = ensure_dynamic_smem(kern, smem, cache); e != cudaSuccess) return e;
    return launch_maybe_pdl(kern, static_cast<unsigned>(grid), A4_THREADS, smem, s, pdl_enabled(), p);
}

template <typename T, int HD>
cudaError_t launch_hd4(const AttnUmmaParams& p, int num_sms, cudaStream_t s) {
    const size_t smem = smem_for4<HD>(p.Fp);
    const int grid = p.num_units < num_sms ? p.num_units : num_sms;
    if (p.drop.thr) return launch_one<T, HD, 0, true>(p, grid, smem, s);
    if (p.Fp == 112) return launch_one<T, HD, 112, false>(p, grid, smem, s);
    if (p.Fp == 128) return launch_one<T, HD, 128, false>(p, grid, smem, s);
    return launch_one<T, HD, 0, false>(p, grid, smem, s);
}

unsigned long long* g_role_prof = nullptr;

}  // namespace

void set_attention_role_prof(unsigned long long* dev_buf) { g_role_prof = dev_buf; }

bool attention_umma4_supported(int Ft, int hd) {
    if (Ft < 1 || Ft > 128 || (hd != 64 && hd != 128)) return false;
    const int Fp = (Ft + 15) & ~15;
    return (hd == 64 ? smem_for4<64>(Fp) : smem_for4<128>(Fp)) <= static_cast<size_t>(A4_MAX_SMEM);
}

template <typename T>
cudaError_t launch_attention_umma4(AttnUmmaParams p, int hd, int num_sms, cudaStream_t s) {
    if (!attention_umma4_supported(p.Ft, hd) || p.B <= 0 || p.H <= 0 || p.Qt < 0) return cudaErrorInvalidValue;
    p.Fp = (p.Ft + 15) & ~15;
    p.tiles_q = (p.Qt + A4_BM - 1) / A4_BM;
    p.prof = g_role_prof;
    const long long items = 1LL * p.B * p.H;
    const int tiles_total = 1 + p.tiles_q;
    // split an item's tiles into several work units when the items alone are too few to balance the persistent CTAs
    long long chunks = (16LL * num_sms + items - 1) / items;
    if (chunks < 1) chunks = 1;
    if (chunks > tiles_total) chunks = tiles_total;
    p.tpu = static_cast<int>((tiles_total + chunks - 1) / chunks);
    p.chunks = (tiles_total + p.tpu - 1) / p.tpu;
    const long long units = items * p.chunks;
    if (units > 0x7fffffffLL) return cudaErrorInvalidValue;
    p.num_units = static_cast<int>(units);
    // L2 prefetch policy (TIM_B200_ATTN_PF = "<mode>,<tiles>"): mode 0 off, 1 K_f / V_f of upcoming units only, 2 everything.
    // Measured at HEAD (isolated launches, profiles/README.md r03j): two tiles per unit (cfg2) 271 us with mode 2 against 283 with mode 1;
    // 17 tiles per unit (cfg4) 306 us with mode 1 against 319 with mode 2 - the query tiles of a long unit stream in order anyway.
    p.pf_mode = tiles_total > 4 ? 1 : 2; p.pf_tiles = A4_PF;
    if (const char* e = std::getenv("TIM_B200_ATTN_PF")) {
        int m = 0, t = A4_PF;
        const int n = std::sscanf(e, "%d,%d", &m, &t);
        if (n >= 1) p.pf_mode = m;
        if (n >= 2 && t >= 0 && t <= 16) p.pf_tiles = t;
    }
    return hd == 64 ? launch_hd4<T, 64>(p, num_sms, s) : launch_hd4<T, 128>(p, num_sms, s);
}
template cudaError_t launch_attention_umma4<__half>(AttnUmmaParams, int, int, cudaStream_t);
template cudaError_t launch_attention_umma4<__nv_bfloat16>(AttnUmmaParams, int, int, cudaStream_t);

}  // namespace tim
