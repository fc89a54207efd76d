// Mask-aware attention of the TIM encoder layer on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
// Reference: nn.MultiheadAttention called from */models/helpers/transformers.py:102 with the boolean mask of
// recognition/.../models/tim.py:161-166 (detection/.../tim.py:384-389): mask[i, j] = (j >= num_feats) && (i != j), i.e.
//   feature rows : softmax over the Ft feature keys of the clip
//   query rows   : softmax over the Ft feature keys + the row's own key
// (token layout, q pre-scaling and the validated decomposition are described in attention.cu).
//
// Work unit = (clip b, head h, a run of 128-row tiles): tile 0 is the feature rows of the clip, tiles 1.. its query rows.
// Per unit K_f / V_f ([Ft, hd]) are staged once; per tile:
//   S = Q K_f^T          tcgen05.mma  M=128, N=Fp, K=hd   (Q, K_f K-major in 128B-swizzled smem, loaded by TMA)
//   S_self = Q K_q^T     query tiles only: the tile's own 128 key rows as a second B operand (N=128); only the diagonal
//                        (q_r . k_r) is used - the tensor core is idle anyway and this keeps the dot products off the
//                        CUDA cores and the k rows on the TMA path
//   softmax              one thread per row, the whole row (<= 128 scores) in registers straight from TMEM, exp2
//   O = P V_f            tcgen05.mma  M=128, N=hd, K=Fp    (P written to smem by the softmax warps, V_f MN-major)
//   out = O / l + (p_self / l) v_self    epilogue: TMEM -> registers -> swizzled smem -> TMA store (row-clipped by the map)
// HBM traffic is the algorithmic minimum: qkv is read once (6 KB per row), out written once (2 KB per row).
//
// One persistent CTA per SM, 320 threads, mbarrier pipelines between the roles:
//   warp 0    : TMA producer (K_f, V_f per unit; Q tile per stage)
//   warp 1    : MMA issuer; software-pipelined so that S(t+1) is issued before O(t) (two TMEM / smem stages, hd <= 128)
//   warps 2-5 : softmax (TMEM lane quarter = warp % 4)
//   warps 6-9 : epilogue (same quarters); softmax of tile t+1 overlaps the epilogue of tile t
#include <cstdio>
#include <cstdlib>

#include "kernels.h"
#include "ptx.cuh"

namespace tim {
namespace {

constexpr int AU_THREADS = 320;
constexpr int AU_BM = 128;
constexpr int AU_P_BYTES = 32768;        // P tile (128 rows x up to 128 keys); reused as the epilogue's v_self scratch
constexpr int AU_PF = 1;                 // tiles prefetched into L2 ahead of the smem loads (measured: 1 > 2 > 0 > 4)
constexpr int AU_MIN_SMEM = 120 * 1024;  // keeps it at one CTA per SM (each CTA allocates all 512 TMEM columns)

template <int HD> struct AUCfg {
    static constexpr int NST = HD <= 128 ? 2 : 1;          // tiles in flight
    static constexpr int KBOX = HD / 64;                   // 64-column (128-byte) boxes per head row
    static constexpr int Q_BYTES = KBOX * AU_BM * 128;     // Q tile (buffer A: Q -> P -> v_self scratch)
    static constexpr int A_BYTES = Q_BYTES > AU_P_BYTES ? Q_BYTES : AU_P_BYTES;
    static constexpr int STAGE_BYTES = A_BYTES + Q_BYTES;  // buffer B: own-key tile K_q -> output staging tile
    static constexpr int TMEM_STAGE = 256;                 // [0,128): S, later O (O may spill into [128,192)); [128,256): S_self
};

template <typename T> struct FmtOfA;
template <> struct FmtOfA<__half> { static constexpr uint32_t v = 0; };
template <> struct FmtOfA<__nv_bfloat16> { static constexpr uint32_t v = 1; };

struct UnitInfo { int b, h, t_lo, t_hi; };
__device__ __forceinline__ UnitInfo decode_unit(const AttnUmmaParams& p, int u) {
    UnitInfo ui;
    const int item = u / p.chunks, ch = u - item * p.chunks;
    ui.b = item / p.H; ui.h = item - ui.b * p.H;
    ui.t_lo = ch * p.tpu;
    ui.t_hi = min(ui.t_lo + p.tpu, 1 + p.tiles_q);
    return ui;
}

// DROP: attention-probability dropout compiled in (training forward with p > 0 only: the inference kernel carries none of it)
template <typename T, int HD, bool DROP>
__global__ void __launch_bounds__(AU_THREADS, 1) attention_umma_kernel(const __grid_constant__ AttnUmmaParams p) {
    using C = AUCfg<HD>;
    constexpr int NST = C::NST, KBOX = C::KBOX;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int Fp = p.Fp, Ft = p.Ft, Qt = p.Qt;
    const uint32_t box_kv = static_cast<uint32_t>(Fp) * 128u;       // one 64-column box of K_f / V_f
    const uint32_t kv_bytes = KBOX * box_kv;
    const uint32_t sKF = base, sVF = base + kv_bytes;
    const uint32_t stage0 = sVF + kv_bytes;
    auto sQ = [&](int st) { return stage0 + static_cast<uint32_t>(st) * C::STAGE_BYTES; };
    auto sP = [&](int st) { return stage0 + static_cast<uint32_t>(st) * C::STAGE_BYTES; };                    // aliases Q
    auto sB = [&](int st) { return stage0 + static_cast<uint32_t>(st) * C::STAGE_BYTES + C::A_BYTES; };
    const uint32_t stat_base = stage0 + NST * C::STAGE_BYTES;        // float [NST][2][128]: 1/l, p_self/l
    auto stat = [&](int st, int which, int row) { return stat_base + static_cast<uint32_t>(((st * 2 + which) * AU_BM + row) * 4); };
    const uint32_t bar_base = stat_base + NST * 1024;
    auto q_full = [&](int st) { return bar_base + 8u * st; };
    auto q_empty = [&](int st) { return bar_base + 8u * (NST + st); };
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
        // producer and MMA warps run warp-uniformly (all lanes carry the loop state and wait on the barriers); one elected
        // lane issues the asynchronous instructions. A loop under `if (lane == 0)` makes the compiler re-derive uniform
        // operands per instruction and starves the issue slot the warp shares with the softmax / epilogue warps.
        {
            // L2 prefetch iterator running AU_PF tiles ahead of the loads: with two smem stages the loads in flight
            // (<= 2 x 64 KB per SM) cannot cover the HBM latency-bandwidth product; the prefetches can.
            int pf_u = blockIdx.x, pf_t = 0;
            UnitInfo pf_ui = {0, 0, 0, 0};
            bool pf_valid = pf_u < p.num_units;
            if (pf_valid) { pf_ui = decode_unit(p, pf_u); pf_t = pf_ui.t_lo; }
            auto prefetch_next = [&]() {
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
                            tma_prefetch_l2_3d(&p.tmQq, h0 + 64 * j, (pf_t - 1) * AU_BM, pf_ui.b);
                            tma_prefetch_l2_3d(&p.tmQq, E + h0 + 64 * j, (pf_t - 1) * AU_BM, pf_ui.b);
                            tma_prefetch_l2_3d(&p.tmQq, 2 * E + h0 + 64 * j, (pf_t - 1) * AU_BM, pf_ui.b);    // own-value rows (epilogue)
                        }
                    }
                }
                if (++pf_t >= pf_ui.t_hi) {
                    pf_u += gridDim.x;
                    pf_valid = pf_u < p.num_units;
                    if (pf_valid) { pf_ui = decode_unit(p, pf_u); pf_t = pf_ui.t_lo; }
                }
            };
            if (p.pf_mode >= 1) for (int i = 0; i < p.pf_tiles; ++i) prefetch_next();
            uint32_t g = 0, un = 0;
            for (int u = blockIdx.x; u < p.num_units; u += gridDim.x, ++un) {
                const UnitInfo ui = decode_unit(p, u);
                mbar_wait(kf_empty, (un & 1u) ^ 1u);
                if (elect_one()) {
                    mbar_arrive_expect_tx(kf_full, kv_bytes);
#pragma unroll
                    for (int j = 0; j < KBOX; ++j) tma_load_3d(sKF + j * box_kv, &p.tmKV, kf_full, E + ui.h * HD + 64 * j, 0, ui.b);
                }
                for (int t = ui.t_lo; t < ui.t_hi; ++t, ++g) {
                    const int st = g % NST;
                    const uint32_t ph = (g / NST) & 1u;
                    if (p.pf_mode >= 1) prefetch_next();
                    mbar_wait(q_empty(st), ph ^ 1u);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(q_full(st), t == 0 ? C::Q_BYTES : 2 * C::Q_BYTES);
#pragma unroll
                        for (int j = 0; j < KBOX; ++j) {
                            if (t == 0) {
                                tma_load_3d(sQ(st) + j * 16384, &p.tmQf, q_full(st), ui.h * HD + 64 * j, 0, ui.b);
                            } else {
                                tma_load_3d(sQ(st) + j * 16384, &p.tmQq, q_full(st), ui.h * HD + 64 * j, (t - 1) * AU_BM, ui.b);
                                tma_load_3d(sB(st) + j * 16384, &p.tmQq, q_full(st), E + ui.h * HD + 64 * j, (t - 1) * AU_BM, ui.b);
                            }
                        }
                    }
                    if (t == ui.t_lo) {
                        // V_f after the unit's first Q tile: the MMA warp issues S of that tile before the last P.V of
                        // the previous unit, which is what releases the V_f buffer
                        mbar_wait(vf_empty, (un & 1u) ^ 1u);
                        if (elect_one()) {
                            mbar_arrive_expect_tx(vf_full, kv_bytes);
#pragma unroll
                            for (int j = 0; j < KBOX; ++j)
                                tma_load_3d(sVF + j * box_kv, &p.tmKV, vf_full, 2 * E + ui.h * HD + 64 * j, 0, ui.b);
                        }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        {
            const uint32_t idesc_s = umma_idesc_f16(FmtOfA<T>::v, AU_BM, static_cast<uint32_t>(Fp));
            const uint32_t idesc_self = umma_idesc_f16(FmtOfA<T>::v, AU_BM, AU_BM);
            const uint32_t idesc_o = umma_idesc_f16(FmtOfA<T>::v, AU_BM, HD) | (1u << 16);     // B (= V_f) is MN-major
            const int ksteps_o = Fp / 16;
            auto issue_pv = [&](int st, uint32_t ph, uint32_t un, bool first, bool last) {
                if (first) mbar_wait(vf_full, un & 1u);
                mbar_wait(p_full(st), ph);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(st * C::TMEM_STAGE);      // O overwrites the consumed S
                if (elect_one()) {
                    for (int k = 0; k < ksteps_o; ++k) {
                        const uint64_t adesc = umma_desc_sw128(sP(st) + (k >> 2) * 16384) + 2u * (k & 3);
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
                const UnitInfo ui = decode_unit(p, u);
                mbar_wait(kf_full, un & 1u);
                for (int t = ui.t_lo; t < ui.t_hi; ++t, ++g) {
                    const int st = g % NST;
                    const uint32_t ph = (g / NST) & 1u;
                    mbar_wait(q_full(st), ph);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(st * C::TMEM_STAGE);
                    const bool first = t == ui.t_lo, last = t == ui.t_hi - 1;
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < HD / 16; ++k) {
                            const uint64_t adesc = umma_desc_sw128(sQ(st) + (k >> 2) * 16384) + 2u * (k & 3);
                            const uint64_t bdesc = umma_desc_sw128(sKF + (k >> 2) * box_kv) + 2u * (k & 3);
                            umma_f16_ss(d_tmem, adesc, bdesc, idesc_s, k != 0 ? 1u : 0u);
                        }
                        if (t > 0) {
#pragma unroll
                            for (int k = 0; k < HD / 16; ++k) {
                                const uint64_t adesc = umma_desc_sw128(sQ(st) + (k >> 2) * 16384) + 2u * (k & 3);
                                const uint64_t bdesc = umma_desc_sw128(sB(st) + (k >> 2) * 16384) + 2u * (k & 3);
                                umma_f16_ss(d_tmem + 128, adesc, bdesc, idesc_self, k != 0 ? 1u : 0u);
                            }
                        }
                        umma_commit(s_full(st));
                        if (last) umma_commit(kf_empty);
                    }
                    if (NST == 1) {
                        issue_pv(st, ph, un, first, last);
                    } else {
                        if (have_prev) issue_pv(prev_st, prev_ph, prev_un, prev_first, prev_last);
                        have_prev = true; prev_st = st; prev_ph = ph; prev_un = un; prev_first = first; prev_last = last;
                    }
                }
            }
            if (NST > 1 && have_prev) issue_pv(prev_st, prev_ph, prev_un, prev_first, prev_last);
        }
        __syncwarp();
    } else if (warp < 6) {
        // ===================== softmax (warps 2..5) =====================
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t swz = static_cast<uint32_t>(row & 7);
        uint32_t g = 0;
        for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
            const UnitInfo ui = decode_unit(p, u);
            for (int t = ui.t_lo; t < ui.t_hi; ++t, ++g) {
                const int st = g % NST;
                const uint32_t ph = (g / NST) & 1u;
                const bool qt = t > 0;
                mbar_wait(s_full(st), ph);
                tc_fence_after();
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(st * C::TMEM_STAGE);
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
                }
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
                // four independent max / sum chains: one warp per scheduler runs this, so dependent-issue latency is what it costs
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
                // attention-probability dropout (training forward only): the probabilities that meet V are dropped, the normaliser is
                // not. Element index ((b H + h) S + row in clip) * DROP_ATTN_KW + key, own key at Ft (kernels.h: DropSite)
                const uint32_t drop_pair0 = ((((static_cast<uint32_t>(ui.b) * p.H + ui.h) * static_cast<uint32_t>(Ft + Qt) +
                                              static_cast<uint32_t>((qt ? Ft + (t - 1) * AU_BM : 0) + row)) * static_cast<uint32_t>(DROP_ATTN_KW)) >> 1);
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
    } else {
        // ===================== epilogue (warps 6..9) =====================
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t swz = static_cast<uint32_t>(row & 7);
        const T* qkv = static_cast<const T*>(p.qkv);
        uint32_t g = 0;
        for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
            const UnitInfo ui = decode_unit(p, u);
            for (int t = ui.t_lo; t < ui.t_hi; ++t, ++g) {
                const int st = g % NST;
                const uint32_t ph = (g / NST) & 1u;
                const bool qt = t > 0;
                const int row0 = qt ? (t - 1) * AU_BM : 0;
                const int nrows = min(AU_BM, (qt ? Qt : Ft) - row0);
                const int nvalid = nrows - quarter * 32;       // rows of this warp's slab that exist
                const bool vterm = qt && nvalid > 0;
                uint4 vreg[8][KBOX];                           // own-value rows, same lane layout as the softmax warps' k rows
                if (vterm) {
                    const T* vbase = qkv + (static_cast<size_t>(p.B) * Ft + static_cast<size_t>(ui.b) * Qt + row0 + quarter * 32) * ld + 2 * E + ui.h * HD + (lane & 7) * 8;
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int rl = min(it * 4 + (lane >> 3), nvalid - 1);
#pragma unroll
                        for (int w = 0; w < KBOX; ++w)
                            vreg[it][w] = __ldg(reinterpret_cast<const uint4*>(vbase + static_cast<size_t>(rl) * ld + w * 64));
                    }
                }
                mbar_wait(p_full(st), ph);                 // softmax statistics of this tile are in smem
                const float inv = lds_f32(stat(st, 0, row));
                const float wself = lds_f32(stat(st, 1, row));
                mbar_wait(o_full(st), ph);                 // P.V complete: O in TMEM, Q / P smem no longer read by the MMA
                tc_fence_after();
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(st * C::TMEM_STAGE);
                const uint32_t scr = sP(st) + static_cast<uint32_t>(warp - 6) * 8192;    // [2 boxes][32 rows][128 B], chunk-swizzled
#pragma unroll
                for (int jb = 0; jb < KBOX; jb += 2) {
                    const int nb = (KBOX - jb) < 2 ? (KBOX - jb) : 2;
                    if (vterm) {
                        // own-value rows: registers (coalesced global loads) -> scratch -> one row per thread
                        __syncwarp();
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            const int rl = it * 4 + (lane >> 3);
#pragma unroll
                            for (int w = 0; w < 2; ++w)
                                if (w < nb) sts_u128(scr + w * 4096 + rl * 128 + ((static_cast<uint32_t>(lane & 7) ^ static_cast<uint32_t>(rl & 7)) << 4),
                                                     vreg[it][jb + w < KBOX ? jb + w : 0]);
                        }
                        __syncwarp();
                    }
                    // TMEM loads of one 64-column box at a time (one wait per box), then the arithmetic
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        if (cc < nb * 2) {
                            const int c32 = jb * 2 + cc;
                            uint32_t v[2][32];
                            if ((cc & 1) == 0) {
                                tmem_ld_32x32(taddr + c32 * 32, v[0]);
                                tmem_ld_32x32(taddr + c32 * 32 + 32, v[1]);
                                tmem_ld_wait();
                            }
                            float f[32];
#pragma unroll
                            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[cc & 1][j]) * inv;
                            if (vterm) {
#pragma unroll
                                for (int c = 0; c < 4; ++c) {
                                    const uint4 q4 = lds_u128(scr + (cc >> 1) * 4096 + lane * 128 + ((static_cast<uint32_t>((cc & 1) * 4 + c) ^ static_cast<uint32_t>(lane & 7)) << 4));
                                    const float2 v0 = unpack2<T>(q4.x), v1 = unpack2<T>(q4.y), v2 = unpack2<T>(q4.z), v3 = unpack2<T>(q4.w);
                                    f[8 * c + 0] = fmaf(wself, v0.x, f[8 * c + 0]); f[8 * c + 1] = fmaf(wself, v0.y, f[8 * c + 1]);
                                    f[8 * c + 2] = fmaf(wself, v1.x, f[8 * c + 2]); f[8 * c + 3] = fmaf(wself, v1.y, f[8 * c + 3]);
                                    f[8 * c + 4] = fmaf(wself, v2.x, f[8 * c + 4]); f[8 * c + 5] = fmaf(wself, v2.y, f[8 * c + 5]);
                                    f[8 * c + 6] = fmaf(wself, v3.x, f[8 * c + 6]); f[8 * c + 7] = fmaf(wself, v3.y, f[8 * c + 7]);
                                }
                            }
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                uint4 q;
                                q.x = pack2<T>(f[8 * c], f[8 * c + 1]); q.y = pack2<T>(f[8 * c + 2], f[8 * c + 3]);
                                q.z = pack2<T>(f[8 * c + 4], f[8 * c + 5]); q.w = pack2<T>(f[8 * c + 6], f[8 * c + 7]);
                                sts_u128(sB(st) + (c32 >> 1) * 16384 + row * 128 + ((static_cast<uint32_t>((c32 & 1) * 4 + c) ^ swz) << 4), q);
                            }
                        }
                    }
                }
                tc_fence_before();
                fence_proxy_async_smem();                  // staged output -> visible to the TMA store
                __syncwarp();
                if (lane == 0) {
                    if (nvalid > 0) {
#pragma unroll
                        for (int j = 0; j < KBOX; ++j)
                            tma_store_3d(qt ? &p.tmOq : &p.tmOf, sB(st) + j * 16384 + quarter * 4096, ui.h * HD + 64 * j, row0 + quarter * 32, ui.b);
                        tma_store_commit();
                        tma_store_wait_read<0>();          // the staging rows may be overwritten by the next K_q tile
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

template <int HD> size_t smem_for(int Fp) {
    using C = AUCfg<HD>;
    size_t b = 1024 + 2 * static_cast<size_t>(C::KBOX) * Fp * 128 + static_cast<size_t>(C::NST) * (C::STAGE_BYTES + 1024) + 256;
    return b < static_cast<size_t>(AU_MIN_SMEM) ? AU_MIN_SMEM : b;
}

template <typename T, int HD>
cudaError_t launch_hd(const AttnUmmaParams& p, int num_sms, cudaStream_t s) {
    const size_t smem = smem_for<HD>(p.Fp);
    const int grid = p.num_units < num_sms ? p.num_units : num_sms;
    if (p.drop.thr) {
        auto kern = attention_umma_kernel<T, HD, true>;
        static SmemAttrCache cache;
        if (cudaError_t e = ensure_dynamic_smem(kern, smem, cache); e != cudaSuccess) return e;
        kern<<<grid, AU_THREADS, smem, s>>>(p);
    } else {
        auto kern = attention_umma_kernel<T, HD, false>;
        static SmemAttrCache cache;
        if (cudaError_t e = ensure_dynamic_smem(kern, smem, cache); e != cudaSuccess) return e;
        kern<<<grid, AU_THREADS, smem, s>>>(p);
    }
    return cudaGetLastError();
}

}  // namespace

bool attention_umma_supported(int Ft, int hd) { return Ft >= 1 && Ft <= 128 && (hd == 64 || hd == 128 || hd == 192); }

size_t attention_umma_smem(int Ft, int hd) {
    const int Fp = (Ft + 15) & ~15;
    return hd == 64 ? smem_for<64>(Fp) : (hd == 128 ? smem_for<128>(Fp) : smem_for<192>(Fp));
}

template <typename T>
cudaError_t launch_attention_umma(AttnUmmaParams p, int hd, int num_sms, cudaStream_t s) {
    if (!attention_umma_supported(p.Ft, hd) || p.B <= 0 || p.H <= 0 || p.Qt < 0) return cudaErrorInvalidValue;
    p.Fp = (p.Ft + 15) & ~15;
    p.tiles_q = (p.Qt + AU_BM - 1) / AU_BM;
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
    // L2 prefetch policy (TIM_B200_ATTN_PF = "<mode>,<tiles>"): mode 0 off, 1 K_f / V_f of upcoming units only, 2 everything
    p.pf_mode = 2; p.pf_tiles = AU_PF;
    if (const char* e = std::getenv("TIM_B200_ATTN_PF")) {
        int m = 0, t = AU_PF;
        const int n = std::sscanf(e, "%d,%d", &m, &t);
        if (n >= 1) p.pf_mode = m;
        if (n >= 2 && t >= 0 && t <= 16) p.pf_tiles = t;
    }
    switch (hd) {
        case 64: return launch_hd<T, 64>(p, num_sms, s);
        case 128: return launch_hd<T, 128>(p, num_sms, s);
        case 192: return launch_hd<T, 192>(p, num_sms, s);
        default: return cudaErrorInvalidValue;
    }
}
template cudaError_t launch_attention_umma<__half>(AttnUmmaParams, int, int, cudaStream_t);
template cudaError_t launch_attention_umma<__nv_bfloat16>(AttnUmmaParams, int, int, cudaStream_t);

}  // namespace tim
