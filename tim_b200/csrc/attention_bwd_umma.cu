// Backward of the mask-aware attention core on the 5th-generation tensor cores (tcgen05 + TMEM + TMA) - the training leg's
// counterpart of attention_umma.cu. Math, scaling conventions and what it replaces in the reference: see attention_bwd.cu (the
// warp-MMA version, which stays for head_dim 16 / 32 / 192 and as the A/B partner: TIM_B200_ATTN_BWD=1).
//
// Work unit = (clip b, head h). K_f / V_f ([Ft, hd]) are staged once per unit; the unit walks its 128-row tiles (tile 0 = the
// clip's feature rows, tiles 1.. = its query rows). Per tile, five products on the tensor cores:
//   S   = Q K_f^T          M=128 rows, N=Fp keys, K=hd     A, B K-major                         -> TMEM [0, Fp)
//   dP  = dO V_f^T         same shapes                                                          -> TMEM [128, 128 + Fp)
//   softmax warps (one thread per row): P, D = sum P dP (+ own key), dS = P o (dP - D); P and dS go to shared memory ONCE,
//   in the K-major 128-byte-swizzled tile layout - which, read with the other descriptor, is also the MN-major layout:
//   dQ  = dS K_f           M=128 rows, N=hd, K=Fp          A = dS K-major, B = K_f MN-major     -> TMEM [0, hd)   (S is consumed)
//   dV_f += P^T dO         M=128 keys, N=hd, K=128 rows    A = P MN-major, B = dO MN-major      -> TMEM [256, 256 + hd), accumulates over the unit's tiles
//   dK_f += dS^T Q         same                            A = dS MN-major, B = Q MN-major      -> TMEM [384, 384 + hd)
// The own-key / own-value terms of query rows are element-wise per row. Their global traffic runs in the COALESCED pattern - 8 lanes per
// 128-byte row segment, 4 rows per instruction - and is turned to / from the one-thread-per-row view by shuffles (dot products of the
// softmax warps, per-row scale factors of the own-key / own-value gradients) or through a staging slab (the epilogue parks the own-key
// rows in its dQ slab, finishes each row in place and leaves the slab to a bulk store). The r02e form did all of it with one thread
// per row, 32 different lines per load / store instruction (profiles/r02l_attention_bwd_roles.txt).
// No atomics, no P / dS in global memory; HBM traffic is the algorithmic 14 KB per token row (+ the own k / v rows once more).
//
// One persistent CTA per SM, 320 threads: warp 0 TMA producer, warp 1 MMA issuer (and TMEM allocation), warps 2-5 softmax
// (TMEM lane quarter = warp % 4; S, D and dS passes only), warps 6-9 epilogue: dQ per tile, dK_f / dV_f per unit, and ALL the own-row
// work of query tiles - the dot products s_self / dP_self of tile g (handed to the softmax warps through shared memory + a_full, needed
// only once the softmax over the feature keys is done) and, one tile late, the own-key / own-value gradients. The softmax warps are the
// serial resource of a tile; the epilogue warps idled 72 % of the time (profiles/r02l). Single-stage per tile (the 512 TMEM
// columns and 227 KB of shared memory are both full at hd = 128): tiles of one SM run back to back, overlap comes from the
// roles (the loads of tile t+1 start as soon as the products of tile t have retired, under its epilogue).
#include <cstdlib>

#include "kernels.h"
#include "ptx.cuh"

namespace tim {
namespace {

constexpr int BU_THREADS = 320;
constexpr int BU_BM = 128;
constexpr float kLn2u = 0.69314718055994530942f;

template <typename T> struct FmtOfB;
template <> struct FmtOfB<__half> { static constexpr uint32_t v = 0; };
template <> struct FmtOfB<__nv_bfloat16> { static constexpr uint32_t v = 1; };

template <typename T, int HD>
__global__ void __launch_bounds__(BU_THREADS, 1) attention_bwd_umma_kernel(const __grid_constant__ AttnBwdUmmaParams p) {
    constexpr int KBOX = HD / 64;
    constexpr int TILE_BYTES = KBOX * BU_BM * 128;          // Q / dO tile
    constexpr int PS_BYTES = 2 * BU_BM * 128;               // P / dS tile: 128 rows x 128 keys
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int Fp = p.Fp, Ft = p.Ft, Qt = p.Qt;
    const uint32_t box_kv = static_cast<uint32_t>(Fp) * 128u;
    const uint32_t kv_bytes = KBOX * box_kv;
    const uint32_t kv_pad = (kv_bytes + 1023u) & ~1023u;
    const uint32_t sKF = base, sVF = base + kv_pad;
    const uint32_t sQ = sVF + kv_pad, sDO = sQ + TILE_BYTES, sP = sDO + TILE_BYTES, sDS = sP + PS_BYTES;
    const uint32_t sG = sDS + PS_BYTES;                     // dQ staging tile [KBOX][128 rows][128 B], chunk-swizzled (bulk-store source)
    const uint32_t stat_base = sG + TILE_BYTES;             // float [2][128], one exchange slot per row: epilogue -> softmax (s_self, dP_self), then softmax -> epilogue (dS_self, p_self)
    const uint32_t bar_base = stat_base + 1024;
    const uint32_t kv_full = bar_base, kv_empty = bar_base + 8, q_full = bar_base + 16, s_full = bar_base + 24, p_full = bar_base + 32,
                   o_full = bar_base + 40, dq_done = bar_base + 48, acc_empty = bar_base + 56, a_full = bar_base + 64, tmem_slot = bar_base + 72;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const int E = p.H * HD;
    const size_t ld = 3 * static_cast<size_t>(E);
    const int tiles = 1 + p.tiles_q;
    const T* qkv = static_cast<const T*>(p.qkv);
    const T* dOp = static_cast<const T*>(p.dO);
    T* dqkv = static_cast<T*>(p.dqkv);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tmKV); tma_prefetch_desc(&p.tmQf); tma_prefetch_desc(&p.tmQq);
        tma_prefetch_desc(&p.tmDf); tma_prefetch_desc(&p.tmDq); tma_prefetch_desc(&p.tmGf); tma_prefetch_desc(&p.tmGq);
    }
    if (warp == 1) {
        if (lane == 0) {
            mbar_init(kv_full, 1); mbar_init(kv_empty, 1); mbar_init(q_full, 1); mbar_init(s_full, 1); mbar_init(p_full, 4);
            mbar_init(o_full, 1); mbar_init(dq_done, 4); mbar_init(acc_empty, 4); mbar_init(a_full, 4);
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
        uint32_t g = 0, un = 0;
        for (int u = blockIdx.x; u < p.num_units; u += gridDim.x, ++un) {
            const int b = u / p.H, h = u - b * p.H;
            mbar_wait(kv_empty, (un & 1u) ^ 1u);
            if (elect_one()) {
                mbar_arrive_expect_tx(kv_full, 2 * kv_bytes);
#pragma unroll
                for (int j = 0; j < KBOX; ++j) {
                    tma_load_3d(sKF + j * box_kv, &p.tmKV, kv_full, E + h * HD + 64 * j, 0, b);
                    tma_load_3d(sVF + j * box_kv, &p.tmKV, kv_full, 2 * E + h * HD + 64 * j, 0, b);
                }
            }
            for (int t = 0; t < tiles; ++t, ++g) {
                if (g > 0) mbar_wait(o_full, (g - 1) & 1u);          // the previous tile's products have retired: Q / dO tiles are free
                if (elect_one()) {
                    mbar_arrive_expect_tx(q_full, 2 * TILE_BYTES);
#pragma unroll
                    for (int j = 0; j < KBOX; ++j) {
                        if (t == 0) {
                            tma_load_3d(sQ + j * 16384, &p.tmQf, q_full, h * HD + 64 * j, 0, b);
                            tma_load_3d(sDO + j * 16384, &p.tmDf, q_full, h * HD + 64 * j, 0, b);
                        } else {
                            tma_load_3d(sQ + j * 16384, &p.tmQq, q_full, h * HD + 64 * j, (t - 1) * BU_BM, b);
                            tma_load_3d(sDO + j * 16384, &p.tmDq, q_full, h * HD + 64 * j, (t - 1) * BU_BM, b);
                        }
                    }
                    // the tile buffers are single: the NEXT tile's loads cannot start before this tile's products retire, so pull its
                    // rows into L2 now - the shared-memory load then costs an L2 round trip instead of an HBM one on the critical path
                    const int nu = u + static_cast<int>(gridDim.x);
                    if (t + 1 < tiles) {
#pragma unroll
                        for (int j = 0; j < KBOX; ++j) {
                            tma_prefetch_l2_3d(&p.tmQq, h * HD + 64 * j, t * BU_BM, b);
                            tma_prefetch_l2_3d(&p.tmDq, h * HD + 64 * j, t * BU_BM, b);
                        }
                    } else if (nu < p.num_units) {
                        const int nb = nu / p.H, nh = nu - nb * p.H;
#pragma unroll
                        for (int j = 0; j < KBOX; ++j) {
                            tma_prefetch_l2_3d(&p.tmKV, E + nh * HD + 64 * j, 0, nb);
                            tma_prefetch_l2_3d(&p.tmKV, 2 * E + nh * HD + 64 * j, 0, nb);
                            tma_prefetch_l2_3d(&p.tmQf, nh * HD + 64 * j, 0, nb);
                            tma_prefetch_l2_3d(&p.tmDf, nh * HD + 64 * j, 0, nb);
                        }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc_s = umma_idesc_f16(FmtOfB<T>::v, BU_BM, static_cast<uint32_t>(Fp));
        const uint32_t idesc_dq = umma_idesc_f16(FmtOfB<T>::v, BU_BM, HD) | (1u << 16);                 // B = K_f MN-major
        const uint32_t idesc_kv = umma_idesc_f16(FmtOfB<T>::v, BU_BM, HD) | (1u << 15) | (1u << 16);   // A and B MN-major
        const int ksteps_f = Fp / 16;
        uint32_t g = 0, un = 0;
        for (int u = blockIdx.x; u < p.num_units; u += gridDim.x, ++un) {
            mbar_wait(kv_full, un & 1u);
            for (int t = 0; t < tiles; ++t, ++g) {
                mbar_wait(q_full, g & 1u);
                if (g > 0) mbar_wait(dq_done, (g - 1) & 1u);        // dQ of the previous tile has been read out of [0, hd)
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < HD / 16; ++k) {
                        const uint64_t bq = umma_desc_sw128(sQ + (k >> 2) * 16384) + 2u * (k & 3);
                        const uint64_t bd = umma_desc_sw128(sDO + (k >> 2) * 16384) + 2u * (k & 3);
                        const uint64_t bk = umma_desc_sw128(sKF + (k >> 2) * box_kv) + 2u * (k & 3);
                        const uint64_t bv = umma_desc_sw128(sVF + (k >> 2) * box_kv) + 2u * (k & 3);
                        umma_f16_ss(tmem_base, bq, bk, idesc_s, k != 0 ? 1u : 0u);             // S
                        umma_f16_ss(tmem_base + 128, bd, bv, idesc_s, k != 0 ? 1u : 0u);       // dP
                    }
                    umma_commit(s_full);
                }
                mbar_wait(p_full, g & 1u);                           // P and dS are in shared memory, S and dP consumed
                if (t == 0 && un > 0) mbar_wait(acc_empty, (un - 1) & 1u);     // dK_f / dV_f of the previous unit have been read out
                tc_fence_after();
                if (elect_one()) {
                    for (int k = 0; k < ksteps_f; ++k) {             // dQ = dS K_f
                        const uint64_t a = umma_desc_sw128(sDS + (k >> 2) * 16384) + 2u * (k & 3);
                        const uint64_t bdesc = umma_desc_mn_sw128(sKF + k * 2048, box_kv);
                        umma_f16_ss(tmem_base, a, bdesc, idesc_dq, k != 0 ? 1u : 0u);
                    }
#pragma unroll
                    for (int k = 0; k < BU_BM / 16; ++k) {           // dV_f += P^T dO,  dK_f += dS^T Q  (k runs over the tile's rows)
                        const uint32_t acc = (t != 0 || k != 0) ? 1u : 0u;
                        umma_f16_ss(tmem_base + 256, umma_desc_mn_sw128(sP + k * 2048, 16384), umma_desc_mn_sw128(sDO + k * 2048, 16384), idesc_kv, acc);
                        umma_f16_ss(tmem_base + 384, umma_desc_mn_sw128(sDS + k * 2048, 16384), umma_desc_mn_sw128(sQ + k * 2048, 16384), idesc_kv, acc);
                    }
                    umma_commit(o_full);
                    if (t == tiles - 1) umma_commit(kv_empty);
                }
            }
        }
        __syncwarp();
    } else if (warp < 6) {
        // ===================== softmax / dS (warps 2..5): one thread per row =====================
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t swz = static_cast<uint32_t>(row & 7);
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        uint32_t g = 0, aq = 0;                                      // aq: query tiles so far (phase of a_full)
        for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
            const int b = u / p.H, h = u - b * p.H;
            for (int t = 0; t < tiles; ++t, ++g) {
                const bool qt = t > 0;
                const int row0 = qt ? (t - 1) * BU_BM : 0;
                const int nrows = min(BU_BM, (qt ? Qt : Ft) - row0);
                const bool valid = row < nrows;
                if (g > 0) mbar_wait(o_full, (g - 1) & 1u);          // P / dS tiles of the previous tile are no longer read by the tensor core
                mbar_wait(s_full, g & 1u);
                tc_fence_after();
                // ---- P over the feature keys (whole row in registers): max, exp2, sum ----
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
                float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    if (c * 16 < Fp) {
                        if (c * 16 + 16 > Ft) {
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                if (c * 16 + j >= Ft) s[c * 16 + j] = -INFINITY;
                        }
#pragma unroll
                        for (int j = 0; j < 16; ++j) mx[j & 3] = fmaxf(mx[j & 3], s[c * 16 + j]);
                    }
                }
                const float mf = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
                // the three passes below run on packed fp32 pairs (FADD2 / FMUL2 / FFMA2, two elements per issue slot): the softmax warps
                // are the serial resource of a tile (profiles/r02q_attention_bwd_roles.txt) and their instruction count is its critical path
                float2 lsA = make_float2(0.0f, 0.0f), lsB = make_float2(0.0f, 0.0f);
                const float2 nmf = make_float2(-mf, -mf);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    if (c * 16 < Fp) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            const float2 a = ex2_approx2(__fadd2_rn(make_float2(s[c * 16 + j], s[c * 16 + j + 1]), nmf));
                            const float2 b2 = ex2_approx2(__fadd2_rn(make_float2(s[c * 16 + j + 2], s[c * 16 + j + 3]), nmf));
                            s[c * 16 + j] = a.x; s[c * 16 + j + 1] = a.y; s[c * 16 + j + 2] = b2.x; s[c * 16 + j + 3] = b2.y;
                            lsA = __fadd2_rn(lsA, a); lsB = __fadd2_rn(lsB, b2);
                        }
                    }
                }
                // ---- the own key joins: s_self and dP_self come from the epilogue warps (computed under the products and the passes above) ----
                float sself = -INFINITY, dps = 0.0f;
                if (qt) {
                    mbar_wait(a_full, aq & 1u);
                    ++aq;
                    sself = lds_f32(stat_base + static_cast<uint32_t>(row * 4));
                    dps = lds_f32(stat_base + static_cast<uint32_t>(512 + row * 4));
                }
                const float m = fmaxf(mf, sself);
                const float rf = ex2_approx(mf - m);                 // the feature-key terms were taken against mf
                float ps = qt ? ex2_approx(sself - m) : 0.0f;
                const float inv = valid ? 1.0f / (((lsA.x + lsA.y) + (lsB.x + lsB.y)) * rf + ps) : 0.0f;      // padded rows: P = dS = 0
                const float invf = inv * rf;
                ps *= inv;
                // dropout of the probabilities in the forward (kernels.h: DropSite): the gradient w.r.t. P is dP o mask, dV_f sees the
                // dropped P, D and dS use the un-dropped P
                const uint32_t drop_pair0 = ((((static_cast<uint32_t>(b) * p.H + h) * static_cast<uint32_t>(Ft + Qt) +
                                              static_cast<uint32_t>((qt ? Ft + row0 : 0) + row)) * static_cast<uint32_t>(DROP_ATTN_KW)) >> 1);
                const float m_self = p.drop.thr ? drop_one(2u * drop_pair0 + static_cast<uint32_t>(Ft), p.drop.key, p.drop.thr, p.drop.scale) : 1.0f;
                dps *= m_self;
                // ---- D = sum_j P_j dP_j (+ own key): first pass over dP ----
                float2 dA = make_float2(0.0f, 0.0f), dB = make_float2(0.0f, 0.0f);
                const float2 invf2 = make_float2(invf, invf);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    if (c * 16 < Fp) {
                        uint32_t v[16];
                        tmem_ld_32x16(taddr + 128 + c * 16, v);
                        tmem_ld_wait();
                        if (p.drop.thr) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                float m0, m1;
                                drop_pair(drop_pair0 + c * 8 + j, p.drop.key, p.drop.thr, p.drop.scale, m0, m1);
                                v[2 * j] = __float_as_uint(__uint_as_float(v[2 * j]) * m0); v[2 * j + 1] = __float_as_uint(__uint_as_float(v[2 * j + 1]) * m1);
                            }
                        }
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            const float2 pa = __fmul2_rn(make_float2(s[c * 16 + j], s[c * 16 + j + 1]), invf2);
                            const float2 pb = __fmul2_rn(make_float2(s[c * 16 + j + 2], s[c * 16 + j + 3]), invf2);
                            s[c * 16 + j] = pa.x; s[c * 16 + j + 1] = pa.y; s[c * 16 + j + 2] = pb.x; s[c * 16 + j + 3] = pb.y;
                            dA = __ffma2_rn(pa, make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), dA);
                            dB = __ffma2_rn(pb, make_float2(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])), dB);
                        }
                    }
                }
                const float D = (dA.x + dA.y) + (dB.x + dB.y) + ps * dps;
                const float2 nD = make_float2(-D, -D);
                const float dss = ps * (dps - D);
                // ---- second pass: dS = P o (dP - D); P and dS -> shared memory (K-major, 128-byte swizzle; 64-key blocks 16 KB apart) ----
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    if (c * 16 < Fp) {
                        uint32_t v[16];
                        tmem_ld_32x16(taddr + 128 + c * 16, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int h8 = 0; h8 < 2; ++h8) {
                            const int c8 = c * 2 + h8;
                            uint4 qp, qd;
                            uint32_t* wp = reinterpret_cast<uint32_t*>(&qp);
                            uint32_t* wd = reinterpret_cast<uint32_t*>(&qd);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float2 pp = make_float2(s[c8 * 8 + 2 * j], s[c8 * 8 + 2 * j + 1]);
                                float2 vv = make_float2(__uint_as_float(v[h8 * 8 + 2 * j]), __uint_as_float(v[h8 * 8 + 2 * j + 1]));
                                float2 pm = pp;
                                if (p.drop.thr) {
                                    float2 mm;
                                    drop_pair(drop_pair0 + c8 * 4 + j, p.drop.key, p.drop.thr, p.drop.scale, mm.x, mm.y);
                                    vv = __fmul2_rn(vv, mm); pm = __fmul2_rn(pp, mm);
                                }
                                wp[j] = pack2<T>(pm.x, pm.y);
                                const float2 ds = __fmul2_rn(pp, __fadd2_rn(vv, nD));
                                wd[j] = pack2<T>(ds.x, ds.y);
                            }
                            const uint32_t off = static_cast<uint32_t>(c8 >> 3) * 16384 + row * 128 + ((static_cast<uint32_t>(c8 & 7) ^ swz) << 4);
                            sts_u128(sP + off, qp);
                            sts_u128(sDS + off, qd);
                        }
                    }
                }
                // back through the same exchange slots (this thread read them above): dS_self and the dropped own probability, for the
                // epilogue warps' dQ correction and own-key / own-value gradients
                sts_f32(stat_base + static_cast<uint32_t>(row * 4), dss);
                sts_f32(stat_base + static_cast<uint32_t>(512 + row * 4), ps * m_self);
                tc_fence_before();
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(p_full);
            }
        }
    } else {
        // ===================== epilogue (warps 6..9): own-row work, dQ per tile, dK_f / dV_f per unit =====================
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t swz = static_cast<uint32_t>(row & 7);
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        // own-key / own-value gradients of the PREVIOUS query tile (dk_own = ln2 dS_self q~, dv_own = p_self dO): deferred by one tile
        // so that the dot products of the next tile, which its softmax is waiting for, go first. Coalesced (8 lanes per 128-byte row
        // segment); q~ and dO rows come from global memory (an L2 hit), the per-row factors from the row's thread by shuffle.
        bool pend = false;
        size_t pend_slab = 0;
        int pend_nv = 0, pend_h = 0;
        float pend_kk = 0.0f, pend_pv = 0.0f;
        auto own_grads = [&]() {
            if (!pend) return;
            pend = false;
            T* oslab = dqkv + pend_slab * ld + pend_h * HD + (lane & 7) * 8;
            const T* qslab = qkv + pend_slab * ld + pend_h * HD + (lane & 7) * 8;
            const T* dslab = dOp + pend_slab * static_cast<size_t>(E) + pend_h * HD + (lane & 7) * 8;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint4 qown[4][KBOX], down[4][KBOX];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int rl = min((half * 4 + i) * 4 + (lane >> 3), pend_nv - 1);
#pragma unroll
                    for (int w = 0; w < KBOX; ++w) {
                        qown[i][w] = ldg_nc_u128_pinned(qslab + static_cast<size_t>(rl) * ld + 64 * w);
                        down[i][w] = ldg_nc_u128_pinned(dslab + static_cast<size_t>(rl) * E + 64 * w);
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int rl = (half * 4 + i) * 4 + (lane >> 3);
                    const float kk_r = __shfl_sync(0xffffffffu, pend_kk, rl), pv_r = __shfl_sync(0xffffffffu, pend_pv, rl);
                    if (rl < pend_nv) {
#pragma unroll
                        for (int w = 0; w < KBOX; ++w) {
                            const uint4 qq = qown[i][w], dd = down[i][w];
                            const uint32_t qw[4] = {qq.x, qq.y, qq.z, qq.w}, dw[4] = {dd.x, dd.y, dd.z, dd.w};
                            uint4 ok, ov;
                            uint32_t* wk = reinterpret_cast<uint32_t*>(&ok);
                            uint32_t* wv = reinterpret_cast<uint32_t*>(&ov);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float2 qf = unpack2<T>(qw[j]), df = unpack2<T>(dw[j]);
                                wk[j] = pack2<T>(kk_r * qf.x, kk_r * qf.y);
                                wv[j] = pack2<T>(pv_r * df.x, pv_r * df.y);
                            }
                            *reinterpret_cast<uint4*>(oslab + static_cast<size_t>(rl) * ld + E + 64 * w) = ok;
                            *reinterpret_cast<uint4*>(oslab + static_cast<size_t>(rl) * ld + 2 * E + 64 * w) = ov;
                        }
                    }
                }
            }
        };
        // s_self = q . k_own, dP_self = dO . v_own of a query tile's rows, ENTIRELY from global memory (the q / dO rows are an L2 hit: the
        // producer prefetched them), so that it can run one tile ahead, under the softmax of the tile before. Coalesced: 8 lanes read
        // one 128-byte segment of a row (4 rows per instruction) and take their part of the dot products; the 8 partial sums are folded
        // by shuffles and handed to the lane that owns the row. Four batches of 2 row groups, each batch in flight together.
        auto own_dots = [&](size_t slab_row, int h, int nv, float& sself, float& dps) {
            sself = -INFINITY; dps = 0.0f;
            if (nv <= 0) return;
            const T* qs = qkv + slab_row * ld + h * HD + (lane & 7) * 8;
            const T* ds = dOp + slab_row * static_cast<size_t>(E) + h * HD + (lane & 7) * 8;
#pragma unroll
            for (int bt = 0; bt < 4; ++bt) {
                uint4 kown[2][KBOX], vown[2][KBOX], qown[2][KBOX], down[2][KBOX];
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const size_t r = static_cast<size_t>(min((bt * 2 + i) * 4 + (lane >> 3), nv - 1));
#pragma unroll
                    for (int w = 0; w < KBOX; ++w) {
                        qown[i][w] = ldg_nc_u128_pinned(qs + r * ld + 64 * w);
                        kown[i][w] = ldg_nc_u128_pinned(qs + r * ld + E + 64 * w);
                        vown[i][w] = ldg_nc_u128_pinned(qs + r * ld + 2 * E + 64 * w);
                        down[i][w] = ldg_nc_u128_pinned(ds + r * E + 64 * w);
                    }
                }
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int it = bt * 2 + i;
                    float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
                    for (int w = 0; w < KBOX; ++w) {
                        const uint4 kq = kown[i][w], vq = vown[i][w], qq = qown[i][w], dd = down[i][w];
                        const uint32_t kw[4] = {kq.x, kq.y, kq.z, kq.w}, vw[4] = {vq.x, vq.y, vq.z, vq.w};
                        const uint32_t qw[4] = {qq.x, qq.y, qq.z, qq.w}, dw[4] = {dd.x, dd.y, dd.z, dd.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float2 kf = unpack2<T>(kw[j]), vf = unpack2<T>(vw[j]), qf = unpack2<T>(qw[j]), df = unpack2<T>(dw[j]);
                            a0 = fmaf(qf.x, kf.x, fmaf(qf.y, kf.y, a0));
                            a1 = fmaf(df.x, vf.x, fmaf(df.y, vf.y, a1));
                        }
                    }
#pragma unroll
                    for (int o = 1; o <= 4; o <<= 1) {
                        a0 += __shfl_xor_sync(0xffffffffu, a0, o);
                        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
                    }
                    const float t0 = __shfl_sync(0xffffffffu, a0, (lane & 3) * 8), t1 = __shfl_sync(0xffffffffu, a1, (lane & 3) * 8);
                    if (it == (lane >> 2)) { sself = t0; dps = t1; }
                }
            }
            if (lane >= nv) { sself = -INFINITY; dps = 0.0f; }
        };
        uint32_t g = 0;
        for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
            const int b = u / p.H, h = u - b * p.H;
            for (int t = 0; t < tiles; ++t, ++g) {
                const bool qt = t > 0;
                const int row0 = qt ? (t - 1) * BU_BM : 0;
                const int nrows = min(BU_BM, (qt ? Qt : Ft) - row0);
                const int nv = nrows - quarter * 32;                 // rows of this warp's 32-row slab that exist
                const bool kterm = qt && nv > 0;
                const size_t slab_row = qt ? static_cast<size_t>(p.B) * Ft + static_cast<size_t>(b) * Qt + row0 + quarter * 32 : 0;
                // own-key rows (query tiles), coalesced, parked in this warp's slab of the dQ staging tile in the OUTPUT layout once the
                // previous bulk store has read it; after the __syncwarp each thread only touches its own row, which it overwrites
                // with the finished dQ row
                uint4 kreg[8][KBOX];
                if (kterm) {
                    const T* kslab = qkv + slab_row * ld + E + h * HD + (lane & 7) * 8;
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int rl = min(it * 4 + (lane >> 3), nv - 1);
#pragma unroll
                        for (int w = 0; w < KBOX; ++w) kreg[it][w] = ldg_nc_u128_pinned(kslab + static_cast<size_t>(rl) * ld + 64 * w);
                    }
                }
                if (lane == 0) tma_store_wait_read<0>();
                __syncwarp();
                if (kterm) {
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int rl = it * 4 + (lane >> 3);
#pragma unroll
                        for (int w = 0; w < KBOX; ++w)
                            sts_u128(sG + w * 16384 + (quarter * 32 + rl) * 128 + ((static_cast<uint32_t>(lane & 7) ^ static_cast<uint32_t>(rl & 7)) << 4), kreg[it][w]);
                    }
                    __syncwarp();
                }
                own_grads();                                         // of the previous query tile, under this tile's softmax
                // the dot products of the NEXT tile (same unit: a query tile), also under this tile's softmax; published below
                const bool next_q = t + 1 < tiles;
                float n_sself = -INFINITY, n_dps = 0.0f;
                if (next_q) {
                    const int n_row0 = t * BU_BM;
                    const int n_nv = min(BU_BM, Qt - n_row0) - quarter * 32;
                    own_dots(static_cast<size_t>(p.B) * Ft + static_cast<size_t>(b) * Qt + n_row0 + quarter * 32, h, n_nv, n_sself, n_dps);
                }
                mbar_wait(p_full, g & 1u);                           // the softmax warps' per-row results are in shared memory
                const float dss = qt ? lds_f32(stat_base + static_cast<uint32_t>(row * 4)) : 0.0f;
                const float pself = qt ? lds_f32(stat_base + static_cast<uint32_t>(512 + row * 4)) : 0.0f;
                if (next_q) {
                    // the exchange slots are free again (this thread has just read them): the next tile's dot products go in
                    sts_f32(stat_base + static_cast<uint32_t>(row * 4), n_sself);
                    sts_f32(stat_base + static_cast<uint32_t>(512 + row * 4), n_dps);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(a_full);
                }
                mbar_wait(o_full, g & 1u);
                tc_fence_after();
                // dQ: each thread finishes its own row in place in the slab (dQ + dS_self k_own, scaled), a bulk store takes the slab
#pragma unroll
                for (int c64 = 0; c64 < HD / 64; ++c64) {
                    uint32_t v[2][32];
                    tmem_ld_32x32(taddr + c64 * 64, v[0]);
                    tmem_ld_32x32(taddr + c64 * 64 + 32, v[1]);
                    tmem_ld_wait();
                    if (c64 == HD / 64 - 1) {
                        // dQ has left TMEM columns [0, hd): the next tile's S may be issued (the unit-end read-out below is of other columns)
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(dq_done);
                    }
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        const int c = c64 * 4 + cc;
                        const uint32_t a0 = sG + (c >> 2) * 16384 + row * 128 + ((static_cast<uint32_t>((c & 3) * 2) ^ swz) << 4);
                        const uint32_t a1 = sG + (c >> 2) * 16384 + row * 128 + ((static_cast<uint32_t>((c & 3) * 2 + 1) ^ swz) << 4);
                        uint4 k0 = make_uint4(0u, 0u, 0u, 0u), k1 = k0;
                        if (kterm) { k0 = lds_u128(a0); k1 = lds_u128(a1); }
                        float f[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[cc >> 1][(cc & 1) * 16 + j]);
                        if (kterm) {
                            const uint32_t kw[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float2 kf = unpack2<T>(kw[j]);
                                f[2 * j] = fmaf(dss, kf.x, f[2 * j]); f[2 * j + 1] = fmaf(dss, kf.y, f[2 * j + 1]);
                            }
                        }
                        uint4 o0, o1;
                        o0.x = pack2<T>(f[0] * p.qscale, f[1] * p.qscale); o0.y = pack2<T>(f[2] * p.qscale, f[3] * p.qscale);
                        o0.z = pack2<T>(f[4] * p.qscale, f[5] * p.qscale); o0.w = pack2<T>(f[6] * p.qscale, f[7] * p.qscale);
                        o1.x = pack2<T>(f[8] * p.qscale, f[9] * p.qscale); o1.y = pack2<T>(f[10] * p.qscale, f[11] * p.qscale);
                        o1.z = pack2<T>(f[12] * p.qscale, f[13] * p.qscale); o1.w = pack2<T>(f[14] * p.qscale, f[15] * p.qscale);
                        sts_u128(a0, o0);
                        sts_u128(a1, o1);
                    }
                }
                fence_proxy_async_smem();                            // staged dQ rows -> visible to the bulk store
                __syncwarp();
                if (lane == 0 && nv > 0) {
#pragma unroll
                    for (int j = 0; j < KBOX; ++j)
                        tma_store_3d(qt ? &p.tmGq : &p.tmGf, sG + j * 16384 + quarter * 4096, h * HD + 64 * j, row0 + quarter * 32, b);
                    tma_store_commit();
                }
                if (kterm) {
                    pend = true; pend_slab = slab_row; pend_nv = nv; pend_h = h;
                    pend_kk = kLn2u * dss; pend_pv = pself;
                }
                if (t == tiles - 1) {
                    // the unit's accumulators are complete: this thread's TMEM lane is feature key `row`. dK_f, then dV_f, through the
                    // slab (each thread its own row) and a bulk store into the feature rows' k / v columns
                    const int nk = Ft - quarter * 32;                // feature keys of this warp's lane quarter that exist
#pragma unroll
                    for (int which = 0; which < 2; ++which) {        // 0: dK_f (TMEM [384, 384 + hd), x ln 2), 1: dV_f (TMEM [256, 256 + hd))
                        const float sc = which == 0 ? kLn2u : 1.0f;
                        if (lane == 0) tma_store_wait_read<0>();
                        __syncwarp();
#pragma unroll
                        for (int c64 = 0; c64 < HD / 64; ++c64) {
                            uint32_t v[2][32];
                            tmem_ld_32x32(taddr + (which == 0 ? 384 : 256) + c64 * 64, v[0]);
                            tmem_ld_32x32(taddr + (which == 0 ? 384 : 256) + c64 * 64 + 32, v[1]);
                            tmem_ld_wait();
#pragma unroll
                            for (int c8 = 0; c8 < 8; ++c8) {
                                const uint32_t* f = &v[c8 >> 2][(c8 & 3) * 8];
                                uint4 o;
                                o.x = pack2<T>(sc * __uint_as_float(f[0]), sc * __uint_as_float(f[1])); o.y = pack2<T>(sc * __uint_as_float(f[2]), sc * __uint_as_float(f[3]));
                                o.z = pack2<T>(sc * __uint_as_float(f[4]), sc * __uint_as_float(f[5])); o.w = pack2<T>(sc * __uint_as_float(f[6]), sc * __uint_as_float(f[7]));
                                sts_u128(sG + c64 * 16384 + row * 128 + ((static_cast<uint32_t>(c8) ^ swz) << 4), o);
                            }
                        }
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0 && nk > 0) {
#pragma unroll
                            for (int j = 0; j < KBOX; ++j)
                                tma_store_3d(&p.tmGf, sG + j * 16384 + quarter * 4096, (which == 0 ? E : 2 * E) + h * HD + 64 * j, quarter * 32, b);
                            tma_store_commit();
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0 && t == tiles - 1) mbar_arrive(acc_empty);
            }
        }
        own_grads();
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

template <int HD> size_t bwd_smem_for(int Fp) {
    const size_t kv = ((static_cast<size_t>(HD / 64) * Fp * 128) + 1023) & ~static_cast<size_t>(1023);
    return 1024 + 2 * kv + 3 * static_cast<size_t>(HD / 64) * BU_BM * 128 + 2 * 2 * BU_BM * 128 + 1024 + 256;
}

template <typename T, int HD>
cudaError_t launch_bwd_umma_hd(const AttnBwdUmmaParams& p, int num_sms, cudaStream_t s) {
    const size_t smem = bwd_smem_for<HD>(p.Fp);
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    auto kern = attention_bwd_umma_kernel<T, HD>;
    static SmemAttrCache cache;
    if (cudaError_t e = ensure_dynamic_smem(kern, smem > 120 * 1024 ? smem : 120 * 1024, cache); e != cudaSuccess) return e;
    const int grid = p.num_units < num_sms ? p.num_units : num_sms;
    kern<<<grid, BU_THREADS, smem > 120 * 1024 ? smem : 120 * 1024, s>>>(p);      // >= 120 KB: one CTA per SM (each allocates all of TMEM)
    return cudaGetLastError();
}

}  // namespace

bool attention_bwd_umma_supported(int Ft, int hd) {
    if (const char* e = std::getenv("TIM_B200_ATTN_BWD")) if (std::atoi(e) == 1) return false;
    return Ft >= 1 && Ft <= 128 && (hd == 64 || hd == 128);
}

template <typename T>
cudaError_t launch_attention_bwd_umma(AttnBwdUmmaParams p, int hd, int num_sms, cudaStream_t s) {
    if (!attention_bwd_umma_supported(p.Ft, hd) || p.B <= 0 || p.H <= 0 || p.Qt < 0) return cudaErrorInvalidValue;
    p.Fp = (p.Ft + 15) & ~15;
    p.tiles_q = (p.Qt + BU_BM - 1) / BU_BM;
    const long long units = 1LL * p.B * p.H;
    if (units > 0x7fffffffLL) return cudaErrorInvalidValue;
    p.num_units = static_cast<int>(units);
    return hd == 64 ? launch_bwd_umma_hd<T, 64>(p, num_sms, s) : launch_bwd_umma_hd<T, 128>(p, num_sms, s);
}
template cudaError_t launch_attention_bwd_umma<__half>(AttnBwdUmmaParams, int, int, cudaStream_t);
template cudaError_t launch_attention_bwd_umma<__nv_bfloat16>(AttnBwdUmmaParams, int, int, cudaStream_t);

}  // namespace tim
