// Mask-aware attention of the TIM encoder layer on tcgen05 / TMEM / TMA - third form of the forward kernel: DEEPER PIPELINE.
//
// Same math, token layout and reference mapping as attention_umma.cu (nn.MultiheadAttention of */models/helpers/transformers.py:102
// under the mask of recognition/.../models/tim.py:161-166). What changed and why: attention_umma.cu keeps, per tile in flight, a Q
// tile, an own-key tile K_q (so that the own-key score is the diagonal of a second product S_self = Q K_q^T) and an output staging
// tile - 64 KB of shared memory and 256 TMEM columns per stage, i.e. TWO tiles in flight at head_dim 128 and ONE at 192, against a
// 6 us load -> MMA -> softmax -> MMA -> epilogue chain: 3.2 ms per step in situ, 0.49 of the HBM peak, tensor pipe 79 % idle
// (VERDICT r01 "what's weak" 2). Here the own-key / own-value terms are element-wise work of the thread that owns the row (its
// q row from the Q tile, its k / v rows straight from global memory, 16-byte loads issued under the previous phase), so a stage is ONE
// 128-row buffer that is Q, then P, then the output staging tile, and 128 TMEM columns (S, overwritten by O):
//   head_dim 64 / 128 : 4 tiles in flight (K_f / V_f 56 KB + 4 x 33 KB);  head_dim 192 : 2 (was 1).
// HBM traffic is unchanged (qkv read once, out written once).
//
// MEASURED (profiles/r02g_attn_bench.txt, same visit, isolated launches): correct (equal to the warp-MMA kernel to 7e-6, 116 forward
// tests green with it as the default) but SLOWER than attention_umma.cu - cfg2 538 us against 344 us, cfg4 766 against 404, head_dim
// 192 717 against 689 - and the stage count does not matter (2 / 3 / 4 stages: 510 / 544 / 541 us). So the forward kernel was never
// pipeline-depth-bound: its limit is the throughput of the four softmax warps (one thread per row, one warp per scheduler), and
// moving the own-key dot product from the idle tensor core (the S_self trick) onto those warps costs more than the extra stages
// give. NOT the default (TIM_B200_ATTN=3 selects it); kept as the measured A/B partner and for what it says about where to go next:
// more threads per row in the softmax role, not more tiles in flight.
//
// One persistent CTA per SM, 320 threads: warp 0 TMA producer (+ L2 prefetch of the next tiles and of the own k / v rows),
// warp 1 MMA issuer (S(t+1) is issued before O(t)), warps 2-5 softmax (one thread per row), warps 6-9 epilogue.
#include <cstdio>
#include <cstdlib>

#include "kernels.h"
#include "ptx.cuh"

namespace tim {
namespace {

constexpr int A3_THREADS = 320;
constexpr int A3_BM = 128;
constexpr int A3_NST_MAX = 4;
constexpr int A3_MIN_SMEM = 120 * 1024;  // keeps it at one CTA per SM (each CTA allocates all 512 TMEM columns)

template <typename T> struct FmtOf3;
template <> struct FmtOf3<__half> { static constexpr uint32_t v = 0; };
template <> struct FmtOf3<__nv_bfloat16> { static constexpr uint32_t v = 1; };

struct Unit3 { int b, h, t_lo, t_hi; };
__device__ __forceinline__ Unit3 decode_unit3(const AttnUmmaParams& p, int u) {
    Unit3 ui;
    const int item = u / p.chunks, ch = u - item * p.chunks;
    ui.b = item / p.H; ui.h = item - ui.b * p.H;
    ui.t_lo = ch * p.tpu;
    ui.t_hi = min(ui.t_lo + p.tpu, 1 + p.tiles_q);
    return ui;
}

// p.pf_mode is reused as the number of stages (2 .. 4), p.pf_tiles as the L2 prefetch distance in tiles
template <typename T, int HD>
__global__ void __launch_bounds__(A3_THREADS, 1) attention_umma3_kernel(const __grid_constant__ AttnUmmaParams p) {
    constexpr int KBOX = HD / 64;
    constexpr int Q_BYTES = KBOX * A3_BM * 128;
    constexpr int TMEM_STAGE = HD <= 128 ? 128 : 256;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int Fp = p.Fp, Ft = p.Ft, Qt = p.Qt;
    const int NST = p.pf_mode;
    const uint32_t box_kv = static_cast<uint32_t>(Fp) * 128u;
    const uint32_t kv_bytes = KBOX * box_kv;
    const uint32_t kv_pad = (kv_bytes + 1023u) & ~1023u;
    const uint32_t p_bytes = static_cast<uint32_t>((Fp + 63) / 64) * 16384u;
    const uint32_t stage_bytes = p_bytes > static_cast<uint32_t>(Q_BYTES) ? p_bytes : static_cast<uint32_t>(Q_BYTES);
    const uint32_t sKF = base, sVF = base + kv_pad;
    const uint32_t stage0 = sVF + kv_pad;
    auto sA = [&](int st) { return stage0 + static_cast<uint32_t>(st) * stage_bytes; };      // Q -> P -> output staging
    const uint32_t stat_base = stage0 + static_cast<uint32_t>(NST) * stage_bytes;            // float [NST][2][128]: 1/l, p_self/l
    auto stat = [&](int st, int which, int row) { return stat_base + static_cast<uint32_t>(((st * 2 + which) * A3_BM + row) * 4); };
    const uint32_t bar_base = stat_base + A3_NST_MAX * 1024;
    auto q_full = [&](int st) { return bar_base + 8u * st; };
    auto q_empty = [&](int st) { return bar_base + 8u * (A3_NST_MAX + st); };
    auto s_full = [&](int st) { return bar_base + 8u * (2 * A3_NST_MAX + st); };
    auto p_full = [&](int st) { return bar_base + 8u * (3 * A3_NST_MAX + st); };
    auto o_full = [&](int st) { return bar_base + 8u * (4 * A3_NST_MAX + st); };
    const uint32_t kf_full = bar_base + 8u * (5 * A3_NST_MAX), kf_empty = kf_full + 8u, vf_full = kf_full + 16u, vf_empty = kf_full + 24u;
    const uint32_t tmem_slot = kf_full + 32u;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const int E = p.H * HD;
    const int ld = 3 * E;
    const T* qkv = static_cast<const T*>(p.qkv);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tmKV); tma_prefetch_desc(&p.tmQf); tma_prefetch_desc(&p.tmQq);
        tma_prefetch_desc(&p.tmOf); tma_prefetch_desc(&p.tmOq);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int st = 0; st < NST; ++st) {
                mbar_init(q_full(st), 1); mbar_init(q_empty(st), 4); mbar_init(s_full(st), 1);
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

    if (warp == 0) {
        // ===================== TMA producer =====================
        // L2 prefetch iterator running p.pf_tiles tiles ahead of the loads (Q tile + the own k / v rows the softmax / epilogue threads read)
        int pf_u = blockIdx.x, pf_t = 0;
        Unit3 pf_ui = {0, 0, 0, 0};
        bool pf_valid = pf_u < p.num_units;
        if (pf_valid) { pf_ui = decode_unit3(p, pf_u); pf_t = pf_ui.t_lo; }
        auto prefetch_next = [&]() {
            if (!pf_valid) return;
            const int h0 = pf_ui.h * HD;
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
                    if (pf_t == 0) {
                        tma_prefetch_l2_3d(&p.tmQf, h0 + 64 * j, 0, pf_ui.b);
                    } else {
                        tma_prefetch_l2_3d(&p.tmQq, h0 + 64 * j, (pf_t - 1) * A3_BM, pf_ui.b);
                        tma_prefetch_l2_3d(&p.tmQq, E + h0 + 64 * j, (pf_t - 1) * A3_BM, pf_ui.b);        // own-key rows (softmax threads)
                        tma_prefetch_l2_3d(&p.tmQq, 2 * E + h0 + 64 * j, (pf_t - 1) * A3_BM, pf_ui.b);    // own-value rows (epilogue threads)
                    }
                }
            }
            if (++pf_t >= pf_ui.t_hi) {
                pf_u += gridDim.x;
                pf_valid = pf_u < p.num_units;
                if (pf_valid) { pf_ui = decode_unit3(p, pf_u); pf_t = pf_ui.t_lo; }
            }
        };
        for (int i = 0; i < p.pf_tiles; ++i) prefetch_next();
        uint32_t g = 0, un = 0;
        for (int u = blockIdx.x; u < p.num_units; u += gridDim.x, ++un) {
            const Unit3 ui = decode_unit3(p, u);
            mbar_wait(kf_empty, (un & 1u) ^ 1u);
            if (elect_one()) {
                mbar_arrive_expect_tx(kf_full, kv_bytes);
#pragma unroll
                for (int j = 0; j < KBOX; ++j) tma_load_3d(sKF + j * box_kv, &p.tmKV, kf_full, E + ui.h * HD + 64 * j, 0, ui.b);
            }
            for (int t = ui.t_lo; t < ui.t_hi; ++t, ++g) {
                const int st = g % NST;
                const uint32_t ph = (g / NST) & 1u;
                if (p.pf_tiles > 0) prefetch_next();
                mbar_wait(q_empty(st), ph ^ 1u);
                if (elect_one()) {
                    mbar_arrive_expect_tx(q_full(st), Q_BYTES);
#pragma unroll
                    for (int j = 0; j < KBOX; ++j) {
                        if (t == 0) tma_load_3d(sA(st) + j * 16384, &p.tmQf, q_full(st), ui.h * HD + 64 * j, 0, ui.b);
                        else tma_load_3d(sA(st) + j * 16384, &p.tmQq, q_full(st), ui.h * HD + 64 * j, (t - 1) * A3_BM, ui.b);
                    }
                }
                if (t == ui.t_lo) {
                    // V_f after the unit's first Q tile: the MMA warp issues S of that tile before the last P.V of the previous unit,
                    // which is what releases the V_f buffer
                    mbar_wait(vf_empty, (un & 1u) ^ 1u);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(vf_full, kv_bytes);
#pragma unroll
                        for (int j = 0; j < KBOX; ++j) tma_load_3d(sVF + j * box_kv, &p.tmKV, vf_full, 2 * E + ui.h * HD + 64 * j, 0, ui.b);
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc_s = umma_idesc_f16(FmtOf3<T>::v, A3_BM, static_cast<uint32_t>(Fp));
        const uint32_t idesc_o = umma_idesc_f16(FmtOf3<T>::v, A3_BM, HD) | (1u << 16);     // B (= V_f) is MN-major
        const int ksteps_o = Fp / 16;
        auto issue_pv = [&](int st, uint32_t ph, uint32_t un, bool first, bool last) {
            if (first) mbar_wait(vf_full, un & 1u);
            mbar_wait(p_full(st), ph);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(st * TMEM_STAGE);      // O overwrites the consumed S
            if (elect_one()) {
                for (int k = 0; k < ksteps_o; ++k) {
                    const uint64_t adesc = umma_desc_sw128(sA(st) + (k >> 2) * 16384) + 2u * (k & 3);
                    const uint64_t bdesc = umma_desc_mn_sw128(sVF + k * 2048, box_kv);
                    umma_f16_ss(d_tmem, adesc, bdesc, idesc_o, k != 0 ? 1u : 0u);
                }
                umma_commit(o_full(st));
                if (last) umma_commit(vf_empty);
            }
        };
        uint32_t g = 0, un = 0;
        bool have_prev = false, prev_first = false, prev_last = false;
        int prev_st = 0; uint32_t prev_ph = 0, prev_un = 0;
        for (int u = blockIdx.x; u < p.num_units; u += gridDim.x, ++un) {
            const Unit3 ui = decode_unit3(p, u);
            mbar_wait(kf_full, un & 1u);
            for (int t = ui.t_lo; t < ui.t_hi; ++t, ++g) {
                const int st = g % NST;
                const uint32_t ph = (g / NST) & 1u;
                mbar_wait(q_full(st), ph);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(st * TMEM_STAGE);
                const bool first = t == ui.t_lo, last = t == ui.t_hi - 1;
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < HD / 16; ++k) {
                        const uint64_t adesc = umma_desc_sw128(sA(st) + (k >> 2) * 16384) + 2u * (k & 3);
                        const uint64_t bdesc = umma_desc_sw128(sKF + (k >> 2) * box_kv) + 2u * (k & 3);
                        umma_f16_ss(d_tmem, adesc, bdesc, idesc_s, k != 0 ? 1u : 0u);
                    }
                    umma_commit(s_full(st));
                    if (last) umma_commit(kf_empty);
                }
                if (NST == 1) {
                    issue_pv(st, ph, un, first, last);       // a single stage cannot hold S(t+1) next to P(t)
                } else {
                    if (have_prev) issue_pv(prev_st, prev_ph, prev_un, prev_first, prev_last);
                    have_prev = true; prev_st = st; prev_ph = ph; prev_un = un; prev_first = first; prev_last = last;
                }
            }
        }
        if (NST > 1 && have_prev) issue_pv(prev_st, prev_ph, prev_un, prev_first, prev_last);
        __syncwarp();
    } else if (warp < 6) {
        // ===================== softmax (warps 2..5): one thread per row =====================
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t swz = static_cast<uint32_t>(row & 7);
        uint32_t g = 0;
        for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
            const Unit3 ui = decode_unit3(p, u);
            for (int t = ui.t_lo; t < ui.t_hi; ++t, ++g) {
                const int st = g % NST;
                const uint32_t ph = (g / NST) & 1u;
                const bool qt = t > 0;
                float sself = -INFINITY;
                if (qt) {
                    // own-key score q_r . k_r: k row straight from global memory (L2: prefetched by the producer), q row from the Q tile
                    const int row0 = (t - 1) * A3_BM;
                    const int rl = min(row, Qt - row0 - 1);          // padded rows read a valid row; their result is never stored
                    const T* krow = qkv + (static_cast<size_t>(p.B) * Ft + static_cast<size_t>(ui.b) * Qt + row0 + rl) * ld + E + ui.h * HD;
                    uint4 kreg[HD / 8];
#pragma unroll
                    for (int c = 0; c < HD / 8; ++c) kreg[c] = __ldg(reinterpret_cast<const uint4*>(krow + 8 * c));
                    mbar_wait(q_full(st), ph);                       // the Q tile of this stage has landed
                    float acc = 0.0f;
#pragma unroll
                    for (int c = 0; c < HD / 8; ++c) {
                        const uint4 qq = lds_u128(sA(st) + static_cast<uint32_t>(c >> 3) * 16384 + row * 128 + ((static_cast<uint32_t>(c & 7) ^ swz) << 4));
                        const uint32_t qw[4] = {qq.x, qq.y, qq.z, qq.w}, kw[4] = {kreg[c].x, kreg[c].y, kreg[c].z, kreg[c].w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float2 qf = unpack2<T>(qw[j]), kf = unpack2<T>(kw[j]);
                            acc = fmaf(qf.x, kf.x, fmaf(qf.y, kf.y, acc));
                        }
                    }
                    sself = acc;
                }
                mbar_wait(s_full(st), ph);
                tc_fence_after();
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(st * TMEM_STAGE);
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
                tmem_ld_wait();
                float mx[4] = {sself, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    if (c * 16 < Fp) {
                        if (c * 16 + 16 > Ft) {          // only the last chunk holds padded keys (their K_f rows are zero-filled)
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
                // P overwrites this thread's OWN row of the Q tile (same 128-byte rows, same swizzle): the S product has retired and
                // nobody else reads that row
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    if (c * 16 < Fp) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float e = ex2_approx(s[c * 16 + j] - m);
                            s[c * 16 + j] = e;
                            ls[j & 3] += e;
                        }
#pragma unroll
                        for (int h8 = 0; h8 < 2; ++h8) {
                            const int c8 = c * 2 + h8;
                            uint4 q;
                            q.x = pack2<T>(s[c8 * 8 + 0], s[c8 * 8 + 1]); q.y = pack2<T>(s[c8 * 8 + 2], s[c8 * 8 + 3]);
                            q.z = pack2<T>(s[c8 * 8 + 4], s[c8 * 8 + 5]); q.w = pack2<T>(s[c8 * 8 + 6], s[c8 * 8 + 7]);
                            sts_u128(sA(st) + (c8 >> 3) * 16384 + row * 128 + ((static_cast<uint32_t>(c8 & 7) ^ swz) << 4), q);
                        }
                    }
                }
                const float l = (ls[0] + ls[1]) + (ls[2] + ls[3]);
                const float ps = qt ? ex2_approx(sself - m) : 0.0f;
                const float inv = 1.0f / (l + ps);
                sts_f32(stat(st, 0, row), inv);
                sts_f32(stat(st, 1, row), ps * inv);
                tc_fence_before();
                fence_proxy_async_smem();          // P (generic-proxy writes) -> visible to the tensor core's operand reads
                __syncwarp();
                if (lane == 0) mbar_arrive(p_full(st));
            }
        }
    } else {
        // ===================== epilogue (warps 6..9) =====================
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t swz = static_cast<uint32_t>(row & 7);
        uint32_t g = 0;
        for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
            const Unit3 ui = decode_unit3(p, u);
            for (int t = ui.t_lo; t < ui.t_hi; ++t, ++g) {
                const int st = g % NST;
                const uint32_t ph = (g / NST) & 1u;
                const bool qt = t > 0;
                const int row0 = qt ? (t - 1) * A3_BM : 0;
                const int nrows = min(A3_BM, (qt ? Qt : Ft) - row0);
                const int nvalid = nrows - quarter * 32;       // rows of this warp's slab that exist
                // own-value row of a query row: straight from global memory into registers, under the tile's products
                uint4 vreg[HD / 8];
                if (qt) {
                    const int rl = min(row, nrows - 1);
                    const T* vrow = qkv + (static_cast<size_t>(p.B) * Ft + static_cast<size_t>(ui.b) * Qt + row0 + rl) * ld + 2 * E + ui.h * HD;
#pragma unroll
                    for (int c = 0; c < HD / 8; ++c) vreg[c] = __ldg(reinterpret_cast<const uint4*>(vrow + 8 * c));
                }
                mbar_wait(p_full(st), ph);                 // softmax statistics of this tile are in smem
                const float inv = lds_f32(stat(st, 0, row));
                const float wself = lds_f32(stat(st, 1, row));
                mbar_wait(o_full(st), ph);                 // P.V complete: O in TMEM, the P tile is no longer read by the tensor core
                tc_fence_after();
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(st * TMEM_STAGE);
#pragma unroll
                for (int c32 = 0; c32 < HD / 32; ++c32) {
                    uint32_t v[32];
                    tmem_ld_32x32(taddr + c32 * 32, v);
                    tmem_ld_wait();
                    float f[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) * inv;
                    if (qt) {
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const uint4 q4 = vreg[c32 * 4 + c];
                            const float2 v0 = unpack2<T>(q4.x), v1 = unpack2<T>(q4.y), v2 = unpack2<T>(q4.z), v3 = unpack2<T>(q4.w);
                            f[8 * c + 0] = fmaf(wself, v0.x, f[8 * c + 0]); f[8 * c + 1] = fmaf(wself, v0.y, f[8 * c + 1]);
                            f[8 * c + 2] = fmaf(wself, v1.x, f[8 * c + 2]); f[8 * c + 3] = fmaf(wself, v1.y, f[8 * c + 3]);
                            f[8 * c + 4] = fmaf(wself, v2.x, f[8 * c + 4]); f[8 * c + 5] = fmaf(wself, v2.y, f[8 * c + 5]);
                            f[8 * c + 6] = fmaf(wself, v3.x, f[8 * c + 6]); f[8 * c + 7] = fmaf(wself, v3.y, f[8 * c + 7]);
                        }
                    }
                    // staged into this thread's own row of the stage buffer (the P tile is dead): box c32 / 2, 16-byte chunks (c32 & 1) * 4 + c
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint4 q;
                        q.x = pack2<T>(f[8 * c], f[8 * c + 1]); q.y = pack2<T>(f[8 * c + 2], f[8 * c + 3]);
                        q.z = pack2<T>(f[8 * c + 4], f[8 * c + 5]); q.w = pack2<T>(f[8 * c + 6], f[8 * c + 7]);
                        sts_u128(sA(st) + (c32 >> 1) * 16384 + row * 128 + ((static_cast<uint32_t>((c32 & 1) * 4 + c) ^ swz) << 4), q);
                    }
                }
                tc_fence_before();
                fence_proxy_async_smem();                  // staged output -> visible to the TMA store
                __syncwarp();
                if (lane == 0) {
                    if (nvalid > 0) {
#pragma unroll
                        for (int j = 0; j < KBOX; ++j)
                            tma_store_3d(qt ? &p.tmOq : &p.tmOf, sA(st) + j * 16384 + quarter * 4096, ui.h * HD + 64 * j, row0 + quarter * 32, ui.b);
                        tma_store_commit();
                        tma_store_wait_read<0>();          // the staging rows are the next Q tile of this stage
                    }
                    mbar_arrive(q_empty(st));
                }
                __syncwarp();
            }
        }
        if (lane == 0) tma_store_wait<0>();
        __syncwarp();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

template <int HD> size_t smem3_for(int Fp, int nst) {
    const size_t kv = ((static_cast<size_t>(HD / 64) * Fp * 128) + 1023) & ~static_cast<size_t>(1023);
    const size_t pb = static_cast<size_t>((Fp + 63) / 64) * 16384, qb = static_cast<size_t>(HD / 64) * A3_BM * 128;
    const size_t b = 1024 + 2 * kv + nst * (pb > qb ? pb : qb) + A3_NST_MAX * 1024 + 256;
    return b < static_cast<size_t>(A3_MIN_SMEM) ? A3_MIN_SMEM : b;
}

template <typename T, int HD>
cudaError_t launch3_hd(AttnUmmaParams& p, int num_sms, cudaStream_t s) {
    int nst = HD <= 128 ? 4 : 2;                     // TMEM: 128 columns per stage (256 at head_dim 192)
    while (nst > 1 && smem3_for<HD>(p.Fp, nst) > 227 * 1024) --nst;
    if (smem3_for<HD>(p.Fp, nst) > 227 * 1024) return cudaErrorInvalidValue;
    if (const char* e = std::getenv("TIM_B200_ATTN_STAGES")) { const int v = std::atoi(e); if (v >= 1 && v < nst) nst = v; }
    p.pf_mode = nst;
    const size_t smem = smem3_for<HD>(p.Fp, nst);
    auto kern = attention_umma3_kernel<T, HD>;
    static SmemAttrCache cache;
    if (cudaError_t e = ensure_dynamic_smem(kern, smem, cache); e != cudaSuccess) return e;
    const int grid = p.num_units < num_sms ? p.num_units : num_sms;
    kern<<<grid, A3_THREADS, smem, s>>>(p);
    return cudaGetLastError();
}

}  // namespace

// same parameter block and tensor maps as launch_attention_umma (attention_umma.cu); tmOf / tmOq boxes are (64, 32, 1)
template <typename T>
cudaError_t launch_attention_umma3(AttnUmmaParams p, int hd, int num_sms, cudaStream_t s) {
    if (!attention_umma_supported(p.Ft, hd) || p.B <= 0 || p.H <= 0 || p.Qt < 0) return cudaErrorInvalidValue;
    p.Fp = (p.Ft + 15) & ~15;
    p.tiles_q = (p.Qt + A3_BM - 1) / A3_BM;
    const long long items = 1LL * p.B * p.H;
    const int tiles_total = 1 + p.tiles_q;
    long long chunks = (16LL * num_sms + items - 1) / items;
    if (chunks < 1) chunks = 1;
    if (chunks > tiles_total) chunks = tiles_total;
    p.tpu = static_cast<int>((tiles_total + chunks - 1) / chunks);
    p.chunks = (tiles_total + p.tpu - 1) / p.tpu;
    const long long units = items * p.chunks;
    if (units > 0x7fffffffLL) return cudaErrorInvalidValue;
    p.num_units = static_cast<int>(units);
    p.pf_tiles = 2;
    if (const char* e = std::getenv("TIM_B200_ATTN_PF")) {
        int m = 0, t = 2;
        const int n = std::sscanf(e, "%d,%d", &m, &t);
        if (n >= 2 && t >= 0 && t <= 16) p.pf_tiles = t;
    }
    switch (hd) {
        case 64: return launch3_hd<T, 64>(p, num_sms, s);
        case 128: return launch3_hd<T, 128>(p, num_sms, s);
        case 192: return launch3_hd<T, 192>(p, num_sms, s);
        default: return cudaErrorInvalidValue;
    }
}
template cudaError_t launch_attention_umma3<__half>(AttnUmmaParams, int, int, cudaStream_t);
template cudaError_t launch_attention_umma3<__nv_bfloat16>(AttnUmmaParams, int, int, cudaStream_t);

}  // namespace tim
