// Mask-aware attention core of the TIM encoder layer.
//
// Reference: nn.MultiheadAttention called from */models/helpers/transformers.py:102 with the boolean mask built in
// recognition/.../models/tim.py:161-166 (detection/.../tim.py:384-389): mask[i, j] = (j >= num_feats) && (i != j).
// The reference materialises a dense [B*H, S, S] score tensor plus the mask; here the mask is implied by the token
// index and never exists in memory:
//   feature rows : softmax over the Ft feature keys of the clip            (Ft x Ft, dense)
//   query rows   : softmax over the Ft feature keys + the row's own key    (Ft + 1 keys)
// Validated decomposition: SURVEY.md §2.1 "Validated decomposition for K6".
//
// Token layout (two streams, batch-major): rows [0, B*Ft) are feature tokens (clip b at b*Ft), rows [B*Ft, B*Ft + B*Qt)
// are query tokens (clip b at B*Ft + b*Qt). qkv row = [q (E) | k (E) | v (E)], head h at columns h*hd.
// q is pre-scaled by hd^-0.5 * log2(e) (folded into the packed in_proj weight/bias), so softmax uses exp2.
#include "kernels.h"
#include "ptx.cuh"

namespace tim {
namespace {

// ------------------------------------------------------------------------------------------------------------------
// 16-bit path: warp-level mma.sync m16n8k16, all Ft (<= 128) keys in one pass, scores and O in registers.
// CTA = 4 warps x 16 rows = one 64-row tile of one (clip, head); feature tiles and query tiles never mix.
// ------------------------------------------------------------------------------------------------------------------
constexpr int AT_BM = 64;
constexpr int AT_WARPS = 4;
constexpr int AT_MAXKEYS = 128;

template <int HD> struct Swz {
    static constexpr int CHUNKS = HD / 8;                                   // 16-byte chunks per row
    static constexpr int MASK = (CHUNKS % 8 == 0) ? 7 : ((CHUNKS % 4 == 0) ? 3 : ((CHUNKS % 2 == 0) ? 1 : 0));
    __device__ static __forceinline__ int off(int row, int chunk) {          // byte offset of a 16-byte chunk
        return row * (HD * 2) + ((chunk ^ (row & MASK)) << 4);
    }
};

// One CTA per (clip, head[, tile range]): K_f / V_f are staged once and stay resident while the CTA walks over its 64-row
// tiles (feature tiles first, then query tiles). Shared memory is sized to the padded key count so that two CTAs fit on
// an SM (104 KB each at Ft = 100, hd = 128): one CTA's cp.async phase overlaps the other's MMA / softmax phase.
template <typename T, int HD>
__global__ void __launch_bounds__(AT_WARPS * 32, 2) attention_mma_kernel(const T* __restrict__ qkv, T* __restrict__ out,
                                                                         int B, int Ft, int Qt, int H, int tiles_f, int tiles_q,
                                                                         int tiles_per_cta) {
    constexpr int CH = HD / 8;
    constexpr int KSTEPS = HD / 16;
    constexpr int NT_MAX = AT_MAXKEYS / 8;
    using SW = Swz<HD>;
    extern __shared__ __align__(128) uint8_t smem[];
    const int Fp = (Ft + 15) & ~15;                       // keys padded to the MMA k granularity
    uint8_t* sK = smem;                                  // [Fp][HD]
    uint8_t* sV = sK + Fp * HD * 2;                      // [Fp][HD]
    uint8_t* sQ = sV + Fp * HD * 2;                      // [64][HD]  (reused as the output staging tile)
    uint8_t* sKq = sQ + AT_BM * HD * 2;                  // [64][HD]  own keys of a query tile
    uint8_t* sVq = sKq + AT_BM * HD * 2;                 // [64][HD]  own values of a query tile

    const int item = blockIdx.x;                          // (clip, head)
    const int b = item / H, h = item - b * H;
    const int tile_lo = blockIdx.y * tiles_per_cta;
    const int tile_hi = min(tile_lo + tiles_per_cta, tiles_f + tiles_q);
    const size_t E = static_cast<size_t>(H) * HD;
    const size_t ld = 3 * E;
    const size_t feat_base = static_cast<size_t>(b) * Ft;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- stage K_f, V_f (+ zero padding rows) once ----
    for (int i = tid; i < Fp * CH; i += AT_WARPS * 32) {
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
        const int row0 = t * AT_BM;                       // first row of the tile inside its stream
        const int nrows = min(AT_BM, rows_in_stream - row0);
        const size_t tile_base = qtile ? static_cast<size_t>(B) * Ft + static_cast<size_t>(b) * Qt + row0 : feat_base + row0;

        if (tile != tile_lo) __syncthreads();             // everyone is done with the previous tile's sQ / sKq / sVq
        // ---- stage the Q tile and, for query tiles, the tile's own K / V rows ----
        for (int i = tid; i < AT_BM * CH; i += AT_WARPS * 32) {
            const int r = i / CH, c = i - r * CH;
            if (r < nrows) {
                const T* src = qkv + (tile_base + r) * ld + h * HD + c * 8;
                cp_async_16(smem_u32(sQ + SW::off(r, c)), src);
                if (qtile) {
                    cp_async_16(smem_u32(sKq + SW::off(r, c)), src + E);
                    cp_async_16(smem_u32(sVq + SW::off(r, c)), src + 2 * E);
                }
            } else {
                *reinterpret_cast<uint4*>(sQ + SW::off(r, c)) = make_uint4(0, 0, 0, 0);
                if (qtile) {
                    *reinterpret_cast<uint4*>(sKq + SW::off(r, c)) = make_uint4(0, 0, 0, 0);
                    *reinterpret_cast<uint4*>(sVq + SW::off(r, c)) = make_uint4(0, 0, 0, 0);
                }
            }
        }
        cp_async_commit();
        cp_async_wait_all();
        __syncthreads();

        const int wr0 = warp * 16;                        // this warp's rows inside the tile
        if (wr0 < nrows) {                                // warp-uniform
            const int g = lane >> 2, tq = lane & 3;
            const int nt = Fp / 8;                        // score n-tiles in use (even)
            const int lm = lane >> 3, lr = lane & 7;      // ldmatrix: matrix index / row inside it

            // ---- S = Q K_f^T  (and the self score q.k_self for query tiles) ----
            float sc[NT_MAX][4];
#pragma unroll
            for (int j = 0; j < NT_MAX; ++j) { sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.0f; }
            float self0 = 0.0f, self1 = 0.0f;             // rows g and g+8
#pragma unroll
            for (int kk = 0; kk < KSTEPS; ++kk) {
                uint32_t a[4];
                const int arow = wr0 + (lm & 1) * 8 + lr, achunk = kk * 2 + (lm >> 1);
                ldmatrix_x4(smem_u32(sQ + SW::off(arow, achunk)), a[0], a[1], a[2], a[3]);
                if (qtile) {
                    uint32_t kq[4];
                    ldmatrix_x4(smem_u32(sKq + SW::off(arow, achunk)), kq[0], kq[1], kq[2], kq[3]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 qa = unpack2<T>(a[i]), ka = unpack2<T>(kq[i]);
                        const float d = qa.x * ka.x + qa.y * ka.y;
                        if (i & 1) self1 += d; else self0 += d;
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
                    }
                }
            }
            // ---- softmax over (feature keys [+ self]) in the log2 domain ----
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
            if (qtile) {
                self0 += __shfl_xor_sync(0xffffffffu, self0, 1); self0 += __shfl_xor_sync(0xffffffffu, self0, 2);
                self1 += __shfl_xor_sync(0xffffffffu, self1, 1); self1 += __shfl_xor_sync(0xffffffffu, self1, 2);
            } else {
                self0 = self1 = -INFINITY;
            }
            m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
            m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
            m0 = fmaxf(m0, self0); m1 = fmaxf(m1, self1);
            float l0 = 0.0f, l1 = 0.0f;
            uint32_t pa[NT_MAX / 2][4];                   // P as A fragments for the PV MMAs
#pragma unroll
            for (int j = 0; j < NT_MAX; ++j) {
                if (j < nt) {
                    const float p0 = exp2f(sc[j][0] - m0), p1 = exp2f(sc[j][1] - m0);
                    const float p2 = exp2f(sc[j][2] - m1), p3 = exp2f(sc[j][3] - m1);
                    l0 += p0 + p1; l1 += p2 + p3;
                    pa[j >> 1][(j & 1) * 2 + 0] = pack2<T>(p0, p1);
                    pa[j >> 1][(j & 1) * 2 + 1] = pack2<T>(p2, p3);
                }
            }
            l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
            l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
            const float ps0 = qtile ? exp2f(self0 - m0) : 0.0f;
            const float ps1 = qtile ? exp2f(self1 - m1) : 0.0f;
            const float inv0 = 1.0f / (l0 + ps0), inv1 = 1.0f / (l1 + ps1);

            // ---- O = P V_f (+ p_self * v_self), normalise, stage into this warp's rows of sQ ----
            __syncwarp();                                 // all lanes of the warp are done reading their sQ rows
#pragma unroll
            for (int jn = 0; jn < HD / 8; jn += 2) {
                float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int k2 = 0; k2 < NT_MAX / 2; ++k2) {
                    if (2 * k2 < nt) {
                        uint32_t b0, b1, b2, b3;
                        const int vrow = 16 * k2 + (lm & 1) * 8 + lr, vchunk = jn + (lm >> 1);
                        ldmatrix_x4_trans(smem_u32(sV + SW::off(vrow, vchunk)), b0, b1, b2, b3);
                        MmaSync<T>::run(o0, pa[k2], b0, b1);
                        MmaSync<T>::run(o1, pa[k2], b2, b3);
                    }
                }
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    float* o = half ? o1 : o0;
                    const int chunk = jn + half;
                    const int r_lo = wr0 + g, r_hi = wr0 + g + 8;
                    if (qtile) {
                        const float2 vlo = unpack2<T>(*reinterpret_cast<const uint32_t*>(sVq + SW::off(r_lo, chunk) + tq * 4));
                        const float2 vhi = unpack2<T>(*reinterpret_cast<const uint32_t*>(sVq + SW::off(r_hi, chunk) + tq * 4));
                        o[0] += ps0 * vlo.x; o[1] += ps0 * vlo.y;
                        o[2] += ps1 * vhi.x; o[3] += ps1 * vhi.y;
                    }
                    *reinterpret_cast<uint32_t*>(sQ + SW::off(r_lo, chunk) + tq * 4) = pack2<T>(o[0] * inv0, o[1] * inv0);
                    *reinterpret_cast<uint32_t*>(sQ + SW::off(r_hi, chunk) + tq * 4) = pack2<T>(o[2] * inv1, o[3] * inv1);
                }
            }
            __syncwarp();
            // ---- coalesced 16-byte stores of this warp's 16 rows ----
            for (int i = lane; i < 16 * CH; i += 32) {
                const int r = wr0 + i / CH, c = i % CH;
                if (r < nrows) {
                    const uint4 v = *reinterpret_cast<const uint4*>(sQ + SW::off(r, c));
                    *reinterpret_cast<uint4*>(out + (tile_base + r) * E + h * HD + c * 8) = v;
                }
            }
        }
    }
}

template <typename T, int HD>
cudaError_t launch_attn_hd(const T* qkv, T* out, int B, int Ft, int Qt, int H, cudaStream_t s) {
    const int Fp = (Ft + 15) & ~15;
    const size_t smem = static_cast<size_t>(2 * Fp + 3 * AT_BM) * HD * 2;
    auto kern = attention_mma_kernel<T, HD>;
    static SmemAttrCache cache;
    if (cudaError_t e = ensure_dynamic_smem(kern, smem, cache); e != cudaSuccess) return e;
    const int tiles_f = (Ft + AT_BM - 1) / AT_BM, tiles_q = (Qt + AT_BM - 1) / AT_BM;
    const int tiles = tiles_f + tiles_q;
    // split an item's tiles over several CTAs only when (clip, head) items alone cannot fill the machine
    const long long items = 1LL * B * H;
    if (items > 0x7fffffffLL) return cudaErrorInvalidValue;
    int nsplit = static_cast<int>((148LL * 2 * 4 + items - 1) / items);
    if (nsplit < 1) nsplit = 1;
    if (nsplit > tiles) nsplit = tiles;
    const int tiles_per_cta = (tiles + nsplit - 1) / nsplit;
    nsplit = (tiles + tiles_per_cta - 1) / tiles_per_cta;
    dim3 grid(static_cast<unsigned>(items), nsplit);
    kern<<<grid, AT_WARPS * 32, smem, s>>>(qkv, out, B, Ft, Qt, H, tiles_f, tiles_q, tiles_per_cta);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------------
// fp32 path (compute_dtype = fp32 parity mode): one warp per row, keys across lanes, K_f / V_f in shared memory.
// ------------------------------------------------------------------------------------------------------------------
constexpr int AS_WARPS = 4;
constexpr int AS_ROWS = 16;          // rows per CTA

__global__ void __launch_bounds__(AS_WARPS * 32) attention_simt_kernel(const float* __restrict__ qkv, float* __restrict__ out,
                                                                       int B, int Ft, int Qt, int H, int hd, int tiles_f, DropSite drop) {
    extern __shared__ float smf[];
    float* sK = smf;                               // [Ft][hd + 1]
    float* sV = sK + static_cast<size_t>(Ft) * (hd + 1);   // [Ft][hd]
    float* sq = sV + static_cast<size_t>(Ft) * hd;          // [AS_WARPS][hd]
    float* sp = sq + AS_WARPS * hd;                          // [AS_WARPS][Ft]
    const int b = blockIdx.z, h = blockIdx.y;
    const bool qtile = static_cast<int>(blockIdx.x) >= tiles_f;
    const int t = qtile ? blockIdx.x - tiles_f : blockIdx.x;
    const int rows_in_stream = qtile ? Qt : Ft;
    const int row0 = t * AS_ROWS;
    const int nrows = min(AS_ROWS, rows_in_stream - row0);
    const size_t E = static_cast<size_t>(H) * hd, ld = 3 * E;
    const size_t feat_base = static_cast<size_t>(b) * Ft;
    const size_t tile_base = qtile ? static_cast<size_t>(B) * Ft + static_cast<size_t>(b) * Qt + row0 : feat_base + row0;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int i = tid; i < Ft * hd; i += AS_WARPS * 32) {
        const int r = i / hd, c = i - r * hd;
        const float* src = qkv + (feat_base + r) * ld + h * hd + c;
        sK[r * (hd + 1) + c] = src[E];
        sV[r * hd + c] = src[2 * E];
    }
    __syncthreads();
    float* myq = sq + warp * hd;
    float* myp = sp + warp * Ft;
    for (int r = warp; r < nrows; r += AS_WARPS) {
        const float* qrow = qkv + (tile_base + r) * ld + h * hd;
        float selfdot = 0.0f;
        for (int c = lane; c < hd; c += 32) {
            const float qv = qrow[c];
            myq[c] = qv;
            if (qtile) selfdot += qv * qrow[E + c];
        }
        __syncwarp();
        float s_self = qtile ? warp_sum(selfdot) : -INFINITY;
        float mx = s_self;
        for (int j = lane; j < Ft; j += 32) {
            const float* kr = sK + j * (hd + 1);
            float acc = 0.0f;
            for (int c = 0; c < hd; ++c) acc = fmaf(myq[c], kr[c], acc);
            myp[j] = acc;
            mx = fmaxf(mx, acc);
        }
        mx = warp_max(mx);
        float sum = 0.0f;
        for (int j = lane; j < Ft; j += 32) {
            const float p = exp2f(myp[j] - mx);
            myp[j] = p;
            sum += p;
        }
        sum = warp_sum(sum);
        float p_self = qtile ? exp2f(s_self - mx) : 0.0f;
        const float inv = 1.0f / (sum + p_self);
        if (drop.thr) {
            // attention-probability dropout (training forward, nn.MultiheadAttention(dropout=p)): the normalised probabilities
            // are dropped, the normaliser is not. Element index: ((b H + h) S + row in clip) * DROP_ATTN_KW + key (own key: Ft)
            const uint32_t rbase = ((static_cast<uint32_t>(b) * H + h) * static_cast<uint32_t>(Ft + Qt) + static_cast<uint32_t>((qtile ? Ft : 0) + row0 + r)) *
                                   static_cast<uint32_t>(DROP_ATTN_KW);
            for (int j = lane; j < Ft; j += 32) myp[j] *= drop_one(rbase + j, drop.key, drop.thr, drop.scale);
            p_self *= drop_one(rbase + Ft, drop.key, drop.thr, drop.scale);
        }
        __syncwarp();
        for (int c = lane; c < hd; c += 32) {
            float acc = qtile ? p_self * qrow[2 * E + c] : 0.0f;
            for (int j = 0; j < Ft; ++j) acc = fmaf(myp[j], sV[j * hd + c], acc);
            out[(tile_base + r) * E + h * hd + c] = acc * inv;
        }
        __syncwarp();
    }
}

}  // namespace

template <typename T>
cudaError_t launch_attention_mma(const T* qkv, T* out, int B, int Ft, int Qt, int H, int hd, cudaStream_t s) {
    if (Ft > AT_MAXKEYS || Ft <= 0) return cudaErrorInvalidValue;
    switch (hd) {
        case 16: return launch_attn_hd<T, 16>(qkv, out, B, Ft, Qt, H, s);
        case 32: return launch_attn_hd<T, 32>(qkv, out, B, Ft, Qt, H, s);
        case 64: return launch_attn_hd<T, 64>(qkv, out, B, Ft, Qt, H, s);
        case 128: return launch_attn_hd<T, 128>(qkv, out, B, Ft, Qt, H, s);
        case 192: return launch_attn_hd<T, 192>(qkv, out, B, Ft, Qt, H, s);
        default: return cudaErrorInvalidValue;
    }
}
template cudaError_t launch_attention_mma<__half>(const __half*, __half*, int, int, int, int, int, cudaStream_t);
template cudaError_t launch_attention_mma<__nv_bfloat16>(const __nv_bfloat16*, __nv_bfloat16*, int, int, int, int, int, cudaStream_t);

size_t attention_simt_smem(int Ft, int hd) {
    return (static_cast<size_t>(Ft) * (hd + 1) + static_cast<size_t>(Ft) * hd + AS_WARPS * hd + AS_WARPS * Ft) * sizeof(float);
}

cudaError_t launch_attention_simt(const float* qkv, float* out, int B, int Ft, int Qt, int H, int hd, cudaStream_t s, DropSite drop) {
    if (drop.thr && (Ft + 1 > DROP_ATTN_KW || 1ull * B * H * (Ft + Qt) * DROP_ATTN_KW > 0xffffffffull)) return cudaErrorInvalidValue;
    const size_t smem = attention_simt_smem(Ft, hd);
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    static SmemAttrCache cache;
    if (cudaError_t e = ensure_dynamic_smem(attention_simt_kernel, smem, cache); e != cudaSuccess) return e;
    const int tiles_f = (Ft + AS_ROWS - 1) / AS_ROWS, tiles_q = (Qt + AS_ROWS - 1) / AS_ROWS;
    dim3 grid(tiles_f + tiles_q, H, B);
    attention_simt_kernel<<<grid, AS_WARPS * 32, smem, s>>>(qkv, out, B, Ft, Qt, H, hd, tiles_f, drop);
    return cudaGetLastError();
}

}  // namespace tim
