// Backward of the mask-aware attention core of the TIM encoder layer (training leg).
//
// What it replaces: the autograd nodes of F.multi_head_attention_forward (baddbmm / softmax / bmm and the [B*H, S, S] mask) that
// torch records for nn.MultiheadAttention in */models/helpers/transformers.py:102 under the mask of recognition/.../models/tim.py:
// 161-166, when recognition/scripts/train.py:354-366 calls backward(). As in the forward (attention.cu), the mask is implied by
// the token index: feature rows attend to the Ft feature keys of their clip, query rows to those plus their own key - so only the
// feature keys / values receive gradient from other rows, and a query row's own key / value only from itself.
//
// Inputs: qkv [M, 3E] as the forward saw it (q columns carry hd^-0.5 * log2 e, softmax = exp2), dO [M, E]. Output dqkv [M, 3E]:
//   dq = qscale * (dS K_f + dS_self k_own)         natural-log dS = P o (dP - D),  D_i = sum_j P_ij dP_ij  (incl. the own key)
//   dk = ln2 * dS^T q~ ,  dv = P^T dO              (feature rows: summed over all rows of the clip; query rows: own terms only)
// qscale = hd^-0.5 makes dq the gradient w.r.t. the UN-scaled in_proj output, which is what the weight / input gradients of in_proj
// (packed without the folded factor) need; qscale = ln2 gives the gradient w.r.t. the stored q~ (kernel-boundary oracle).
//
// 16-bit path, two kernels on warp-level mma.sync m16n8k16 (attention FLOPs are ~2 % of the step; the tcgen05 budget went to the GEMMs):
//   attn_bwd_dq_kernel   row-parallel like the forward warp-MMA kernel: S, softmax, dP, D, dS, dQ, the own-key / own-value terms,
//                        and per (row, head) the softmax statistics (max, 1 / l, D) for the second kernel
//   attn_bwd_dkv_kernel  one CTA per (clip, head), warp w owns 16 feature keys: recomputes S^T = K_f Q^T and dP^T = V_f dO^T over
//                        32-row chunks of the clip's rows (so the C fragments ARE the A fragments of the next product) and
//                        accumulates dK_f = dS^T Q, dV_f = P^T dO in registers - no atomics, no P / dS tensors in memory.
// fp32 path: one CUDA-core kernel, one CTA per (clip, head).
#include "kernels.h"
#include "ptx.cuh"

namespace tim {
namespace {

constexpr float kLn2 = 0.69314718055994530942f;
constexpr int AB_BM = 64;
constexpr int AB_WARPS = 4;
constexpr int AB_MAXKEYS = 128;

template <int HD> struct SwzB {
    static constexpr int CHUNKS = HD / 8;
    static constexpr int MASK = (CHUNKS % 8 == 0) ? 7 : ((CHUNKS % 4 == 0) ? 3 : ((CHUNKS % 2 == 0) ? 1 : 0));
    __device__ static __forceinline__ int off(int row, int chunk) { return row * (HD * 2) + ((chunk ^ (row & MASK)) << 4); }
};

// ------------------------------------------------------------------------------------------------------------------
// kernel 1: dQ (+ own-key / own-value gradients of query rows, + softmax statistics)
// ------------------------------------------------------------------------------------------------------------------
template <typename T, int HD>
__global__ void __launch_bounds__(AB_WARPS * 32, 2) attn_bwd_dq_kernel(const T* __restrict__ qkv, const T* __restrict__ dO, T* __restrict__ dqkv,
                                                                       float4* __restrict__ stats, int B, int Ft, int Qt, int H, int tiles_f,
                                                                       int tiles_q, int tiles_per_cta, float qscale) {
    constexpr int CH = HD / 8;
    constexpr int KSTEPS = HD / 16;
    constexpr int NT_MAX = AB_MAXKEYS / 8;
    using SW = SwzB<HD>;
    extern __shared__ __align__(128) uint8_t smem[];
    const int Fp = (Ft + 15) & ~15;
    uint8_t* sK = smem;                                  // [Fp][HD]
    uint8_t* sV = sK + Fp * HD * 2;                      // [Fp][HD]
    uint8_t* sQ = sV + Fp * HD * 2;                      // [64][HD]  (reused as the dQ staging tile)
    uint8_t* sD = sQ + AB_BM * HD * 2;                   // [64][HD]  dO tile
    // The own key / value rows of a query tile are NOT staged: they are only used element-wise, in exactly the (row, column) pattern
    // of this thread's A / C fragments, so each thread reads its 4-byte pieces straight from global memory (every 32-byte sector
    // is consumed whole by the four lanes of a quad). That keeps the CTA at 2 K/V + 2 tile buffers: two CTAs per SM instead of one
    // (the first version staged them: 120 KB, 4 warps per SM, 5.9 % warps active, profiles/r02b_ncu_full_attnbwd.csv).

    const int item = blockIdx.x;
    const int b = item / H, h = item - b * H;
    const int tile_lo = blockIdx.y * tiles_per_cta;
    const int tile_hi = min(tile_lo + tiles_per_cta, tiles_f + tiles_q);
    const size_t E = static_cast<size_t>(H) * HD;
    const size_t ld = 3 * E;
    const size_t Mtot = static_cast<size_t>(B) * (Ft + Qt);
    const size_t feat_base = static_cast<size_t>(b) * Ft;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int i = tid; i < Fp * CH; i += AB_WARPS * 32) {
        const int r = i / CH, c = i - r * CH;
        if (r < Ft) {
            const T* src = qkv + (feat_base + r) * ld + h * HD + c * 8;
            cp_async_16(smem_u32(sK + SW::off(r, c)), src + E);
            cp_async_16(smem_u32(sV + SW::off(r, c)), src + 2 * E);
        } else {
            *reinterpret_cast<uint4*>(sK + SW::off(r, c)) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(sV + SW::off(r, c)) = make_uint4(0, 0, 0, 0);
        }
    }

    for (int tile = tile_lo; tile < tile_hi; ++tile) {
        const bool qtile = tile >= tiles_f;
        const int t = qtile ? tile - tiles_f : tile;
        const int rows_in_stream = qtile ? Qt : Ft;
        const int row0 = t * AB_BM;
        const int nrows = min(AB_BM, rows_in_stream - row0);
        const size_t tile_base = qtile ? static_cast<size_t>(B) * Ft + static_cast<size_t>(b) * Qt + row0 : feat_base + row0;

        if (tile != tile_lo) __syncthreads();
        for (int i = tid; i < AB_BM * CH; i += AB_WARPS * 32) {
            const int r = i / CH, c = i - r * CH;
            if (r < nrows) {
                const T* src = qkv + (tile_base + r) * ld + h * HD + c * 8;
                cp_async_16(smem_u32(sQ + SW::off(r, c)), src);
                cp_async_16(smem_u32(sD + SW::off(r, c)), dO + (tile_base + r) * E + h * HD + c * 8);
            } else {
                *reinterpret_cast<uint4*>(sQ + SW::off(r, c)) = make_uint4(0, 0, 0, 0);
                *reinterpret_cast<uint4*>(sD + SW::off(r, c)) = make_uint4(0, 0, 0, 0);
            }
        }
        cp_async_commit();
        cp_async_wait_all();
        __syncthreads();

        const int wr0 = warp * 16;
        if (wr0 < nrows) {
            const int g = lane >> 2, tq = lane & 3;
            const int nt = Fp / 8;
            const int lm = lane >> 3, lr = lane & 7;
            const int r_lo = wr0 + g, r_hi = wr0 + g + 8;
            // this thread's own-key / own-value pieces: rows r_lo / r_hi (clamped inside the tile; results of padded rows are unused)
            const T* own_lo = qkv + (tile_base + min(r_lo, nrows - 1)) * ld + E + h * HD + tq * 2;
            const T* own_hi = qkv + (tile_base + min(r_hi, nrows - 1)) * ld + E + h * HD + tq * 2;

            // ---- S = Q K_f^T and dP = dO V_f^T (+ the own-key score q.k_own and dP_self = dO.v_own of query tiles) ----
            float sc[NT_MAX][4], dp[NT_MAX][4];
#pragma unroll
            for (int j = 0; j < NT_MAX; ++j) {
                sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.0f;
                dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.0f;
            }
            float self0 = 0.0f, self1 = 0.0f, dps0 = 0.0f, dps1 = 0.0f;
#pragma unroll
            for (int kk = 0; kk < KSTEPS; ++kk) {
                uint32_t a[4], ad[4];
                const int arow = wr0 + (lm & 1) * 8 + lr, achunk = kk * 2 + (lm >> 1);
                ldmatrix_x4(smem_u32(sQ + SW::off(arow, achunk)), a[0], a[1], a[2], a[3]);
                ldmatrix_x4(smem_u32(sD + SW::off(arow, achunk)), ad[0], ad[1], ad[2], ad[3]);
                if (qtile) {
                    // same element pattern as the A fragments: [0] (r_lo, 16kk + 2tq), [1] (r_hi, ..), [2] (r_lo, 16kk + 8 + 2tq), [3] (r_hi, ..)
                    uint32_t kq[4], vq[4];
                    kq[0] = __ldg(reinterpret_cast<const unsigned int*>(own_lo + 16 * kk));
                    kq[1] = __ldg(reinterpret_cast<const unsigned int*>(own_hi + 16 * kk));
                    kq[2] = __ldg(reinterpret_cast<const unsigned int*>(own_lo + 16 * kk + 8));
                    kq[3] = __ldg(reinterpret_cast<const unsigned int*>(own_hi + 16 * kk + 8));
                    vq[0] = __ldg(reinterpret_cast<const unsigned int*>(own_lo + E + 16 * kk));
                    vq[1] = __ldg(reinterpret_cast<const unsigned int*>(own_hi + E + 16 * kk));
                    vq[2] = __ldg(reinterpret_cast<const unsigned int*>(own_lo + E + 16 * kk + 8));
                    vq[3] = __ldg(reinterpret_cast<const unsigned int*>(own_hi + E + 16 * kk + 8));
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 qa = unpack2<T>(a[i]), ka = unpack2<T>(kq[i]);
                        const float2 da = unpack2<T>(ad[i]), va = unpack2<T>(vq[i]);
                        const float d1 = qa.x * ka.x + qa.y * ka.y, d2 = da.x * va.x + da.y * va.y;
                        if (i & 1) { self1 += d1; dps1 += d2; } else { self0 += d1; dps0 += d2; }
                    }
                }
#pragma unroll
                for (int j = 0; j < NT_MAX; j += 2) {
                    if (j < nt) {
                        uint32_t b0, b1, b2, b3;
                        const int krow = 8 * (j + (lm >> 1)) + lr, kchunk = kk * 2 + (lm & 1);
                        ldmatrix_x4(smem_u32(sK + SW::off(krow, kchunk)), b0, b1, b2, b3);
                        MmaSync<T>::run(sc[j], a, b0, b1);
                        MmaSync<T>::run(sc[j + 1], a, b2, b3);
                        ldmatrix_x4(smem_u32(sV + SW::off(krow, kchunk)), b0, b1, b2, b3);
                        MmaSync<T>::run(dp[j], ad, b0, b1);
                        MmaSync<T>::run(dp[j + 1], ad, b2, b3);
                    }
                }
            }
            // ---- softmax statistics (log2 domain) ----
            float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
            for (int j = 0; j < NT_MAX; ++j) {
                if (j < nt) {
                    const int key = 8 * j + 2 * tq;
                    if (key >= Ft) { sc[j][0] = -INFINITY; sc[j][2] = -INFINITY; }
                    if (key + 1 >= Ft) { sc[j][1] = -INFINITY; sc[j][3] = -INFINITY; }
                    m0 = fmaxf(m0, fmaxf(sc[j][0], sc[j][1]));
                    m1 = fmaxf(m1, fmaxf(sc[j][2], sc[j][3]));
                }
            }
            auto quad_sum = [](float v) { v += __shfl_xor_sync(0xffffffffu, v, 1); v += __shfl_xor_sync(0xffffffffu, v, 2); return v; };
            if (qtile) {
                self0 = quad_sum(self0); self1 = quad_sum(self1);
                dps0 = quad_sum(dps0); dps1 = quad_sum(dps1);
            } else {
                self0 = self1 = -INFINITY;
                dps0 = dps1 = 0.0f;
            }
            m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
            m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
            m0 = fmaxf(m0, self0); m1 = fmaxf(m1, self1);
            float l0 = 0.0f, l1 = 0.0f;
#pragma unroll
            for (int j = 0; j < NT_MAX; ++j) {
                if (j < nt) {
                    sc[j][0] = ex2_approx(sc[j][0] - m0); sc[j][1] = ex2_approx(sc[j][1] - m0);
                    sc[j][2] = ex2_approx(sc[j][2] - m1); sc[j][3] = ex2_approx(sc[j][3] - m1);
                    l0 += sc[j][0] + sc[j][1]; l1 += sc[j][2] + sc[j][3];
                }
            }
            l0 = quad_sum(l0); l1 = quad_sum(l1);
            float ps0 = qtile ? ex2_approx(self0 - m0) : 0.0f, ps1 = qtile ? ex2_approx(self1 - m1) : 0.0f;
            const float inv0 = 1.0f / (l0 + ps0), inv1 = 1.0f / (l1 + ps1);
            ps0 *= inv0; ps1 *= inv1;
            // ---- D = sum_j P_ij dP_ij (+ own key), dS = P o (dP - D) ----
            float D0 = 0.0f, D1 = 0.0f;
#pragma unroll
            for (int j = 0; j < NT_MAX; ++j) {
                if (j < nt) {
                    sc[j][0] *= inv0; sc[j][1] *= inv0; sc[j][2] *= inv1; sc[j][3] *= inv1;
                    D0 = fmaf(sc[j][0], dp[j][0], fmaf(sc[j][1], dp[j][1], D0));
                    D1 = fmaf(sc[j][2], dp[j][2], fmaf(sc[j][3], dp[j][3], D1));
                }
            }
            D0 = quad_sum(D0) + ps0 * dps0; D1 = quad_sum(D1) + ps1 * dps1;
            uint32_t pa[NT_MAX / 2][4];                   // dS as A fragments
#pragma unroll
            for (int j = 0; j < NT_MAX; ++j) {
                if (j < nt) {
                    pa[j >> 1][(j & 1) * 2 + 0] = pack2<T>(sc[j][0] * (dp[j][0] - D0), sc[j][1] * (dp[j][1] - D0));
                    pa[j >> 1][(j & 1) * 2 + 1] = pack2<T>(sc[j][2] * (dp[j][2] - D1), sc[j][3] * (dp[j][3] - D1));
                }
            }
            const float dss0 = ps0 * (dps0 - D0), dss1 = ps1 * (dps1 - D1);      // own-key dS of query rows
            if (tq == 0) {
                if (r_lo < nrows) stats[static_cast<size_t>(h) * Mtot + tile_base + r_lo] = make_float4(m0, inv0, D0, 0.0f);
                if (r_hi < nrows) stats[static_cast<size_t>(h) * Mtot + tile_base + r_hi] = make_float4(m1, inv1, D1, 0.0f);
            }

            // ---- own-key / own-value gradients of query rows: dk_own = ln2 dss q~, dv_own = p_self dO (written straight out) ----
            if (qtile) {
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    const float2 qlo = unpack2<T>(*reinterpret_cast<const uint32_t*>(sQ + SW::off(r_lo, c) + tq * 4));
                    const float2 qhi = unpack2<T>(*reinterpret_cast<const uint32_t*>(sQ + SW::off(r_hi, c) + tq * 4));
                    const float2 dlo = unpack2<T>(*reinterpret_cast<const uint32_t*>(sD + SW::off(r_lo, c) + tq * 4));
                    const float2 dhi = unpack2<T>(*reinterpret_cast<const uint32_t*>(sD + SW::off(r_hi, c) + tq * 4));
                    const size_t col = static_cast<size_t>(h) * HD + c * 8 + tq * 2;
                    if (r_lo < nrows) {
                        T* o = dqkv + (tile_base + r_lo) * ld + col;
                        *reinterpret_cast<uint32_t*>(o + E) = pack2<T>(kLn2 * dss0 * qlo.x, kLn2 * dss0 * qlo.y);
                        *reinterpret_cast<uint32_t*>(o + 2 * E) = pack2<T>(ps0 * dlo.x, ps0 * dlo.y);
                    }
                    if (r_hi < nrows) {
                        T* o = dqkv + (tile_base + r_hi) * ld + col;
                        *reinterpret_cast<uint32_t*>(o + E) = pack2<T>(kLn2 * dss1 * qhi.x, kLn2 * dss1 * qhi.y);
                        *reinterpret_cast<uint32_t*>(o + 2 * E) = pack2<T>(ps1 * dhi.x, ps1 * dhi.y);
                    }
                }
            }

            // ---- dQ = qscale (dS K_f + dss k_own), staged into this warp's rows of sQ ----
            __syncwarp();
#pragma unroll
            for (int jn = 0; jn < HD / 8; jn += 2) {
                float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int k2 = 0; k2 < NT_MAX / 2; ++k2) {
                    if (2 * k2 < nt) {
                        uint32_t b0, b1, b2, b3;
                        const int vrow = 16 * k2 + (lm & 1) * 8 + lr, vchunk = jn + (lm >> 1);
                        ldmatrix_x4_trans(smem_u32(sK + SW::off(vrow, vchunk)), b0, b1, b2, b3);
                        MmaSync<T>::run(o0, pa[k2], b0, b1);
                        MmaSync<T>::run(o1, pa[k2], b2, b3);
                    }
                }
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    float* o = half ? o1 : o0;
                    const int chunk = jn + half;
                    if (qtile) {
                        const float2 klo = unpack2<T>(__ldg(reinterpret_cast<const unsigned int*>(own_lo + 8 * chunk)));     // L1 / L2 hit: read above
                        const float2 khi = unpack2<T>(__ldg(reinterpret_cast<const unsigned int*>(own_hi + 8 * chunk)));
                        o[0] += dss0 * klo.x; o[1] += dss0 * klo.y;
                        o[2] += dss1 * khi.x; o[3] += dss1 * khi.y;
                    }
                    *reinterpret_cast<uint32_t*>(sQ + SW::off(r_lo, chunk) + tq * 4) = pack2<T>(o[0] * qscale, o[1] * qscale);
                    *reinterpret_cast<uint32_t*>(sQ + SW::off(r_hi, chunk) + tq * 4) = pack2<T>(o[2] * qscale, o[3] * qscale);
                }
            }
            __syncwarp();
            for (int i = lane; i < 16 * CH; i += 32) {
                const int r = wr0 + i / CH, c = i % CH;
                if (r < nrows) {
                    const uint4 v = *reinterpret_cast<const uint4*>(sQ + SW::off(r, c));
                    *reinterpret_cast<uint4*>(dqkv + (tile_base + r) * ld + h * HD + c * 8) = v;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// kernel 2: dK_f, dV_f. One CTA per (clip, head), 8 warps, warp w owns feature keys [16 w, 16 w + 16).
// ------------------------------------------------------------------------------------------------------------------
constexpr int KV_WARPS = 8;
constexpr int KV_RC = 32;                  // rows per chunk

template <typename T, int HD, bool DO_K, bool DO_V>
__device__ __forceinline__ void dkv_pass(const T* __restrict__ qkv, const T* __restrict__ dO, T* __restrict__ dqkv, const float4* __restrict__ stats,
                                         uint8_t* sK, uint8_t* sV, uint8_t* sQc, uint8_t* sDc, float4* sSt, int B, int Ft, int Qt, int H,
                                         int b, int h, int Fp) {
    constexpr int CH = HD / 8;
    constexpr int KSTEPS = HD / 16;
    constexpr int NJ = KV_RC / 8;              // n-tiles of the S^T / dP^T fragments (rows of the chunk)
    using SW = SwzB<HD>;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, tq = lane & 3, lm = lane >> 3, lr = lane & 7;
    const size_t E = static_cast<size_t>(H) * HD, ld = 3 * E;
    const size_t Mtot = static_cast<size_t>(B) * (Ft + Qt);
    const bool active = warp * 16 < Fp;          // warp-uniform
    const int key0 = warp * 16;

    float accK[DO_K ? HD / 8 : 1][4], accV[DO_V ? HD / 8 : 1][4];
#pragma unroll
    for (int j = 0; j < (DO_K ? HD / 8 : 1); ++j) accK[j][0] = accK[j][1] = accK[j][2] = accK[j][3] = 0.0f;
#pragma unroll
    for (int j = 0; j < (DO_V ? HD / 8 : 1); ++j) accV[j][0] = accV[j][1] = accV[j][2] = accV[j][3] = 0.0f;

    const int chunks_f = (Ft + KV_RC - 1) / KV_RC, chunks_q = (Qt + KV_RC - 1) / KV_RC;
    for (int ch = 0; ch < chunks_f + chunks_q; ++ch) {
        const bool qs = ch >= chunks_f;
        const int row0 = (qs ? ch - chunks_f : ch) * KV_RC;
        const int nrows = min(KV_RC, (qs ? Qt : Ft) - row0);
        const size_t base = qs ? static_cast<size_t>(B) * Ft + static_cast<size_t>(b) * Qt + row0 : static_cast<size_t>(b) * Ft + row0;
        __syncthreads();                            // every warp is done with the previous chunk's tiles
        for (int i = tid; i < KV_RC * CH; i += KV_WARPS * 32) {
            const int r = i / CH, c = i - r * CH;
            if (r < nrows) {
                cp_async_16(smem_u32(sQc + SW::off(r, c)), qkv + (base + r) * ld + h * HD + c * 8);
                cp_async_16(smem_u32(sDc + SW::off(r, c)), dO + (base + r) * E + h * HD + c * 8);
            } else {
                *reinterpret_cast<uint4*>(sQc + SW::off(r, c)) = make_uint4(0, 0, 0, 0);
                *reinterpret_cast<uint4*>(sDc + SW::off(r, c)) = make_uint4(0, 0, 0, 0);
            }
        }
        if (tid < KV_RC) sSt[tid] = tid < nrows ? stats[static_cast<size_t>(h) * Mtot + base + tid] : make_float4(0.f, 0.f, 0.f, 0.f);
        cp_async_commit();
        cp_async_wait_all();
        __syncthreads();
        if (!active) continue;

        // ---- S^T = K_f Q^T  and  dP^T = V_f dO^T  for this warp's 16 keys x the chunk's 32 rows ----
        float st[NJ][4], dpt[NJ][4];
#pragma unroll
        for (int j = 0; j < NJ; ++j) { st[j][0] = st[j][1] = st[j][2] = st[j][3] = 0.0f; dpt[j][0] = dpt[j][1] = dpt[j][2] = dpt[j][3] = 0.0f; }
#pragma unroll
        for (int kk = 0; kk < KSTEPS; ++kk) {
            uint32_t ak[4], av[4];
            const int arow = key0 + (lm & 1) * 8 + lr, achunk = kk * 2 + (lm >> 1);
            ldmatrix_x4(smem_u32(sK + SW::off(arow, achunk)), ak[0], ak[1], ak[2], ak[3]);
            if (DO_K) ldmatrix_x4(smem_u32(sV + SW::off(arow, achunk)), av[0], av[1], av[2], av[3]);
#pragma unroll
            for (int j = 0; j < NJ; j += 2) {
                uint32_t b0, b1, b2, b3;
                const int brow = 8 * (j + (lm >> 1)) + lr, bchunk = kk * 2 + (lm & 1);
                ldmatrix_x4(smem_u32(sQc + SW::off(brow, bchunk)), b0, b1, b2, b3);
                MmaSync<T>::run(st[j], ak, b0, b1);
                MmaSync<T>::run(st[j + 1], ak, b2, b3);
                if (DO_K) {
                    ldmatrix_x4(smem_u32(sDc + SW::off(brow, bchunk)), b0, b1, b2, b3);
                    MmaSync<T>::run(dpt[j], av, b0, b1);
                    MmaSync<T>::run(dpt[j + 1], av, b2, b3);
                }
            }
        }
        // ---- P^T = exp2(S^T - m_row) / l_row (0 for padded keys / rows), dS^T = P^T o (dP^T - D_row) ----
        uint32_t pT[NJ / 2][4], dsT[NJ / 2][4];        // A fragments: k = chunk rows
        const bool klo_ok = key0 + g < Ft, khi_ok = key0 + g + 8 < Ft;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const float4 s0 = sSt[8 * j + 2 * tq], s1 = sSt[8 * j + 2 * tq + 1];     // (max, 1 / l, D) of the two rows this thread holds
            const float p00 = klo_ok ? ex2_approx(st[j][0] - s0.x) * s0.y : 0.0f, p01 = klo_ok ? ex2_approx(st[j][1] - s1.x) * s1.y : 0.0f;
            const float p10 = khi_ok ? ex2_approx(st[j][2] - s0.x) * s0.y : 0.0f, p11 = khi_ok ? ex2_approx(st[j][3] - s1.x) * s1.y : 0.0f;
            if (DO_V) {
                pT[j >> 1][(j & 1) * 2 + 0] = pack2<T>(p00, p01);
                pT[j >> 1][(j & 1) * 2 + 1] = pack2<T>(p10, p11);
            }
            if (DO_K) {
                dsT[j >> 1][(j & 1) * 2 + 0] = pack2<T>(p00 * (dpt[j][0] - s0.z), p01 * (dpt[j][1] - s1.z));
                dsT[j >> 1][(j & 1) * 2 + 1] = pack2<T>(p10 * (dpt[j][2] - s0.z), p11 * (dpt[j][3] - s1.z));
            }
        }
        // ---- dV_f += P^T dO,  dK_f += dS^T Q~   (B operands: the chunk tiles read transposed) ----
#pragma unroll
        for (int jn = 0; jn < HD / 8; jn += 2) {
#pragma unroll
            for (int k2 = 0; k2 < NJ / 2; ++k2) {
                uint32_t b0, b1, b2, b3;
                const int vrow = 16 * k2 + (lm & 1) * 8 + lr, vchunk = jn + (lm >> 1);
                if (DO_V) {
                    ldmatrix_x4_trans(smem_u32(sDc + SW::off(vrow, vchunk)), b0, b1, b2, b3);
                    MmaSync<T>::run(accV[DO_V ? jn : 0], pT[k2], b0, b1);
                    MmaSync<T>::run(accV[DO_V ? jn + 1 : 0], pT[k2], b2, b3);
                }
                if (DO_K) {
                    ldmatrix_x4_trans(smem_u32(sQc + SW::off(vrow, vchunk)), b0, b1, b2, b3);
                    MmaSync<T>::run(accK[DO_K ? jn : 0], dsT[k2], b0, b1);
                    MmaSync<T>::run(accK[DO_K ? jn + 1 : 0], dsT[k2], b2, b3);
                }
            }
        }
    }
    if (!active) return;        // (no barrier follows inside this function; the next pass starts with its own __syncthreads)
    // ---- write this warp's 16 key rows of dK_f (x ln2) / dV_f ----
    const size_t r_lo = static_cast<size_t>(b) * Ft + key0 + g, r_hi = r_lo + 8;
    const bool lo_ok = key0 + g < Ft, hi_ok = key0 + g + 8 < Ft;
#pragma unroll
    for (int jn = 0; jn < HD / 8; ++jn) {
        const size_t col = static_cast<size_t>(h) * HD + jn * 8 + tq * 2;
        if (DO_K) {
            if (lo_ok) *reinterpret_cast<uint32_t*>(dqkv + r_lo * ld + E + col) = pack2<T>(kLn2 * accK[DO_K ? jn : 0][0], kLn2 * accK[DO_K ? jn : 0][1]);
            if (hi_ok) *reinterpret_cast<uint32_t*>(dqkv + r_hi * ld + E + col) = pack2<T>(kLn2 * accK[DO_K ? jn : 0][2], kLn2 * accK[DO_K ? jn : 0][3]);
        }
        if (DO_V) {
            if (lo_ok) *reinterpret_cast<uint32_t*>(dqkv + r_lo * ld + 2 * E + col) = pack2<T>(accV[DO_V ? jn : 0][0], accV[DO_V ? jn : 0][1]);
            if (hi_ok) *reinterpret_cast<uint32_t*>(dqkv + r_hi * ld + 2 * E + col) = pack2<T>(accV[DO_V ? jn : 0][2], accV[DO_V ? jn : 0][3]);
        }
    }
}

template <typename T, int HD>
__global__ void __launch_bounds__(KV_WARPS * 32, 2) attn_bwd_dkv_kernel(const T* __restrict__ qkv, const T* __restrict__ dO, T* __restrict__ dqkv,
                                                                        const float4* __restrict__ stats, int B, int Ft, int Qt, int H) {
    constexpr int CH = HD / 8;
    using SW = SwzB<HD>;
    extern __shared__ __align__(128) uint8_t smem[];
    const int Fp = (Ft + 15) & ~15;
    uint8_t* sK = smem;
    uint8_t* sV = sK + Fp * HD * 2;
    uint8_t* sQc = sV + Fp * HD * 2;
    uint8_t* sDc = sQc + KV_RC * HD * 2;
    float4* sSt = reinterpret_cast<float4*>(sDc + KV_RC * HD * 2);
    const int item = blockIdx.x;
    const int b = item / H, h = item - b * H;
    const size_t E = static_cast<size_t>(H) * HD, ld = 3 * E;
    for (int i = threadIdx.x; i < Fp * CH; i += KV_WARPS * 32) {
        const int r = i / CH, c = i - r * CH;
        if (r < Ft) {
            const T* src = qkv + (static_cast<size_t>(b) * Ft + r) * ld + h * HD + c * 8;
            cp_async_16(smem_u32(sK + SW::off(r, c)), src + E);
            cp_async_16(smem_u32(sV + SW::off(r, c)), src + 2 * E);
        } else {
            *reinterpret_cast<uint4*>(sK + SW::off(r, c)) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(sV + SW::off(r, c)) = make_uint4(0, 0, 0, 0);
        }
    }
    // (the first chunk's cp.async group waits for these as well)
    // dV_f and dK_f are produced by DIFFERENT CTAs (blockIdx.y): with both accumulators in one thread the kernel needed 255
    // registers and ran one CTA (8 warps) per SM at 30 % tensor-pipe (profiles/r02b_ncu_full_attnbwd.csv); one accumulator set fits
    // 128 registers, two CTAs per SM, and the two passes of a (clip, head) run concurrently. Cost: S^T is computed twice.
    if (blockIdx.y == 0) dkv_pass<T, HD, false, true>(qkv, dO, dqkv, stats, sK, sV, sQc, sDc, sSt, B, Ft, Qt, H, b, h, Fp);
    else dkv_pass<T, HD, true, false>(qkv, dO, dqkv, stats, sK, sV, sQc, sDc, sSt, B, Ft, Qt, H, b, h, Fp);
}

template <typename T, int HD>
cudaError_t launch_bwd_hd(const T* qkv, const T* dO, T* dqkv, float4* stats, int B, int Ft, int Qt, int H, float qscale, cudaStream_t s) {
    const int Fp = (Ft + 15) & ~15;
    const long long items = 1LL * B * H;
    if (items <= 0) return cudaSuccess;
    if (items > 0x7fffffffLL) return cudaErrorInvalidValue;
    {
        const size_t smem = static_cast<size_t>(2 * Fp + 2 * AB_BM) * HD * 2;
        if (smem > 227 * 1024) return cudaErrorInvalidValue;
        auto kern = attn_bwd_dq_kernel<T, HD>;
        static SmemAttrCache cache;
        if (cudaError_t e = ensure_dynamic_smem(kern, smem, cache); e != cudaSuccess) return e;
        const int tiles_f = (Ft + AB_BM - 1) / AB_BM, tiles_q = (Qt + AB_BM - 1) / AB_BM;
        const int tiles = tiles_f + tiles_q;
        int nsplit = static_cast<int>((148LL * 2 * 4 + items - 1) / items);
        if (nsplit < 1) nsplit = 1;
        if (nsplit > tiles) nsplit = tiles;
        const int tiles_per_cta = (tiles + nsplit - 1) / nsplit;
        nsplit = (tiles + tiles_per_cta - 1) / tiles_per_cta;
        dim3 grid(static_cast<unsigned>(items), nsplit);
        kern<<<grid, AB_WARPS * 32, smem, s>>>(qkv, dO, dqkv, stats, B, Ft, Qt, H, tiles_f, tiles_q, tiles_per_cta, qscale);
        if (cudaError_t e = cudaGetLastError(); e != cudaSuccess) return e;
    }
    {
        const size_t smem = static_cast<size_t>(2 * Fp + 2 * KV_RC) * HD * 2 + KV_RC * sizeof(float4);
        auto kern = attn_bwd_dkv_kernel<T, HD>;
        static SmemAttrCache cache;
        if (cudaError_t e = ensure_dynamic_smem(kern, smem, cache); e != cudaSuccess) return e;
        kern<<<dim3(static_cast<unsigned>(items), 2), KV_WARPS * 32, smem, s>>>(qkv, dO, dqkv, stats, B, Ft, Qt, H);
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------------
// fp32 path: one CTA per (clip, head), one warp per row; dK_f / dV_f accumulate with atomics into the (pre-zeroed) output.
// ------------------------------------------------------------------------------------------------------------------
constexpr int SB_WARPS = 4;

__global__ void __launch_bounds__(SB_WARPS * 32) attn_bwd_simt_kernel(const float* __restrict__ qkv, const float* __restrict__ dO,
                                                                      float* __restrict__ dqkv, int B, int Ft, int Qt, int H, int hd, float qscale,
                                                                      DropSite drop) {
    extern __shared__ float smf[];
    float* sK = smf;                                         // [Ft][hd + 1]
    float* sV = sK + static_cast<size_t>(Ft) * (hd + 1);     // [Ft][hd + 1]
    float* sq = sV + static_cast<size_t>(Ft) * (hd + 1);     // [SB_WARPS][hd]
    float* sd = sq + SB_WARPS * hd;                          // [SB_WARPS][hd]
    float* sp = sd + SB_WARPS * hd;                          // [SB_WARPS][Ft]  P
    float* ss = sp + SB_WARPS * Ft;                          // [SB_WARPS][Ft]  dS
    const int b = blockIdx.x / H, h = blockIdx.x - b * H;
    const size_t E = static_cast<size_t>(H) * hd, ld = 3 * E;
    const size_t feat_base = static_cast<size_t>(b) * Ft;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < Ft * hd; i += SB_WARPS * 32) {
        const int r = i / hd, c = i - r * hd;
        const float* src = qkv + (feat_base + r) * ld + h * hd + c;
        sK[r * (hd + 1) + c] = src[E];
        sV[r * (hd + 1) + c] = src[2 * E];
    }
    __syncthreads();
    float* myq = sq + warp * hd; float* myd = sd + warp * hd;
    float* myp = sp + warp * Ft; float* mys = ss + warp * Ft;
    for (int r = warp; r < Ft + Qt; r += SB_WARPS) {
        const bool qrow = r >= Ft;
        const size_t row = qrow ? static_cast<size_t>(B) * Ft + static_cast<size_t>(b) * Qt + (r - Ft) : feat_base + r;
        const float* qp = qkv + row * ld + h * hd;
        const float* dp_ = dO + row * E + h * hd;
        float selfdot = 0.0f, dps = 0.0f;
        for (int c = lane; c < hd; c += 32) {
            const float qv = qp[c], dv = dp_[c];
            myq[c] = qv; myd[c] = dv;
            if (qrow) { selfdot = fmaf(qv, qp[E + c], selfdot); dps = fmaf(dv, qp[2 * E + c], dps); }
        }
        __syncwarp();
        const float s_self = qrow ? warp_sum(selfdot) : -INFINITY;
        dps = qrow ? warp_sum(dps) : 0.0f;
        float mx = s_self;
        for (int j = lane; j < Ft; j += 32) {
            const float* kr = sK + j * (hd + 1);
            const float* vr = sV + j * (hd + 1);
            float a = 0.0f, d = 0.0f;
            for (int c = 0; c < hd; ++c) { a = fmaf(myq[c], kr[c], a); d = fmaf(myd[c], vr[c], d); }
            myp[j] = a; mys[j] = d;
            mx = fmaxf(mx, a);
        }
        mx = warp_max(mx);
        float sum = 0.0f;
        for (int j = lane; j < Ft; j += 32) { const float p = exp2f(myp[j] - mx); myp[j] = p; sum += p; }
        sum = warp_sum(sum);
        float p_self = qrow ? exp2f(s_self - mx) : 0.0f;
        const float inv = 1.0f / (sum + p_self);
        p_self *= inv;
        // dropout of the probabilities in the forward: the gradient w.r.t. P is dP o mask, and dV sees the dropped P
        const uint32_t rbase = ((static_cast<uint32_t>(b) * H + h) * static_cast<uint32_t>(Ft + Qt) + static_cast<uint32_t>(r)) * static_cast<uint32_t>(DROP_ATTN_KW);
        float m_self = 1.0f;
        if (drop.thr) {
            for (int j = lane; j < Ft; j += 32) mys[j] *= drop_one(rbase + j, drop.key, drop.thr, drop.scale);
            m_self = drop_one(rbase + Ft, drop.key, drop.thr, drop.scale);
            dps *= m_self;
        }
        float D = 0.0f;
        for (int j = lane; j < Ft; j += 32) { const float p = myp[j] * inv; myp[j] = p; D = fmaf(p, mys[j], D); }
        D = warp_sum(D) + p_self * dps;
        for (int j = lane; j < Ft; j += 32) mys[j] = myp[j] * (mys[j] - D);
        const float dss = p_self * (dps - D);
        if (drop.thr) {                                  // from here on myp is the DROPPED probability (what multiplied V in the forward)
            for (int j = lane; j < Ft; j += 32) myp[j] *= drop_one(rbase + j, drop.key, drop.thr, drop.scale);
        }
        const float p_self_d = p_self * m_self;
        __syncwarp();
        for (int c = lane; c < hd; c += 32) {
            float acc = qrow ? dss * qp[E + c] : 0.0f;
            for (int j = 0; j < Ft; ++j) acc = fmaf(mys[j], sK[j * (hd + 1) + c], acc);
            dqkv[row * ld + h * hd + c] = acc * qscale;
            if (qrow) {
                dqkv[row * ld + E + h * hd + c] = kLn2 * dss * myq[c];
                dqkv[row * ld + 2 * E + h * hd + c] = p_self_d * myd[c];
            }
            const float qv = myq[c] * kLn2, dv = myd[c];
            for (int j = 0; j < Ft; ++j) {
                atomicAdd(dqkv + (feat_base + j) * ld + E + h * hd + c, mys[j] * qv);
                atomicAdd(dqkv + (feat_base + j) * ld + 2 * E + h * hd + c, myp[j] * dv);
            }
        }
        __syncwarp();
    }
}

}  // namespace

size_t attention_bwd_stats_bytes(int B, int Ft, int Qt, int H) {
    return static_cast<size_t>(B) * (Ft + Qt) * H * sizeof(float4);
}

template <typename T>
cudaError_t launch_attention_bwd(const T* qkv, const T* dO, T* dqkv, void* stats, int B, int Ft, int Qt, int H, int hd, float qscale,
                                 cudaStream_t s) {
    if (Ft > AB_MAXKEYS || Ft <= 0) return cudaErrorInvalidValue;
    float4* st = static_cast<float4*>(stats);
    switch (hd) {
        case 16: return launch_bwd_hd<T, 16>(qkv, dO, dqkv, st, B, Ft, Qt, H, qscale, s);
        case 32: return launch_bwd_hd<T, 32>(qkv, dO, dqkv, st, B, Ft, Qt, H, qscale, s);
        case 64: return launch_bwd_hd<T, 64>(qkv, dO, dqkv, st, B, Ft, Qt, H, qscale, s);
        case 128: return launch_bwd_hd<T, 128>(qkv, dO, dqkv, st, B, Ft, Qt, H, qscale, s);
        case 192: return launch_bwd_hd<T, 192>(qkv, dO, dqkv, st, B, Ft, Qt, H, qscale, s);
        default: return cudaErrorInvalidValue;
    }
}
template cudaError_t launch_attention_bwd<__half>(const __half*, const __half*, __half*, void*, int, int, int, int, int, float, cudaStream_t);
template cudaError_t launch_attention_bwd<__nv_bfloat16>(const __nv_bfloat16*, const __nv_bfloat16*, __nv_bfloat16*, void*, int, int, int, int, int, float, cudaStream_t);

size_t attention_bwd_simt_smem(int Ft, int hd) {
    return (static_cast<size_t>(2) * Ft * (hd + 1) + static_cast<size_t>(2) * SB_WARPS * hd + static_cast<size_t>(2) * SB_WARPS * Ft) * sizeof(float);
}

// dqkv must be zero on entry for the feature rows' k / v columns (the caller memsets the whole buffer)
cudaError_t launch_attention_bwd_simt(const float* qkv, const float* dO, float* dqkv, int B, int Ft, int Qt, int H, int hd, float qscale,
                                      cudaStream_t s, DropSite drop) {
    if (drop.thr && (Ft + 1 > DROP_ATTN_KW || 1ull * B * H * (Ft + Qt) * DROP_ATTN_KW > 0xffffffffull)) return cudaErrorInvalidValue;
    const size_t smem = attention_bwd_simt_smem(Ft, hd);
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    static SmemAttrCache cache;
    if (cudaError_t e = ensure_dynamic_smem(attn_bwd_simt_kernel, smem, cache); e != cudaSuccess) return e;
    const long long items = 1LL * B * H;
    if (items <= 0) return cudaSuccess;
    attn_bwd_simt_kernel<<<static_cast<unsigned>(items), SB_WARPS * 32, smem, s>>>(qkv, dO, dqkv, B, Ft, Qt, H, hd, qscale, drop);
    return cudaGetLastError();
}

}  // namespace tim
