// Weight-gradient GEMM of the TIM training leg on the 5th-generation tensor cores:
//     dW[N, K] += dY[M, N]^T * X[M, K]        16-bit operands, fp32 accumulate in TMEM, fp32 reduce-add into the gradient buffer
// What it replaces: the `grad_weight = grad_output.t() @ input` addmm that torch.autograd runs for every nn.Linear of the model
// (recognition/.../models/helpers/transformers.py:75-111, encodings.py:140-153, tim.py:66-74, head.py) when
// recognition/scripts/train.py:354-366 calls backward(). No reference kernel exists (library calls).
//
// Both operands are contracted over their ROWS (the token index m), i.e. each is "MN-major" for the tensor core: a TMA box of
// 64 token rows x 64 columns lands in shared memory as 64 rows of 128 bytes (128-byte swizzle), which is exactly the canonical
// MN-major SWIZZLE_128B layout (8 k-rows per 1024-byte atom, 64-column blocks LBO apart). No transposed copy of dY or X is ever
// made - the activations are read once in the layout the forward / dgrad GEMMs wrote them.
//
// Work item = (n-tile, k-tile, split of the token rows); CTA pair (cta_group::2): 256 x 256 output tile, each CTA stages its 128
// dY columns and 128 X columns per 64-row step (32 KB), the leader issues tcgen05.mma for both, two TMEM accumulators so the
// epilogue of one item overlaps the main loop of the next. The epilogue adds the fp32 tile into dW with TMA reduce
// (cp.reduce.async.bulk.tensor .add.f32), which also makes the split over token rows free of a second pass.
// Persistent; items are ordered split-major so that the clusters running concurrently read the same token rows (L2 reuse).
#include "kernels.h"
#include "ptx.cuh"

namespace tim {
namespace {

constexpr int WG_BM = 128;                 // dW rows (n) per CTA, 256 per pair
constexpr int WG_BN = 256;                 // dW columns (k) per pair tile, 128 staged per CTA
constexpr int WG_BK = 64;                  // token rows per pipeline stage
constexpr int WG_UK = 16;
constexpr int WG_STAGES = 5;
constexpr int WG_A_BYTES = WG_BK * WG_BM * 2;           // 16 KB: two boxes of 64 rows x 64 columns
constexpr int WG_B_BYTES = WG_BK * (WG_BN / 2) * 2;     // 16 KB
constexpr int WG_STAGE_BYTES = WG_A_BYTES + WG_B_BYTES;
constexpr int WG_BOX_BYTES = WG_BK * 128;               // one 64-column box = 8 KB
constexpr int WG_EPI_WARPS = 8;
constexpr int WG_THREADS = 64 + WG_EPI_WARPS * 32;
constexpr int WG_EPI_BUF = 32 * 128;
constexpr int WG_SMEM = WG_STAGES * WG_STAGE_BYTES + WG_EPI_WARPS * 2 * WG_EPI_BUF + 512 + 1024;
constexpr int WG_TMEM_COLS = 512;

template <typename T> struct FmtOfW;
template <> struct FmtOfW<__half> { static constexpr uint32_t v = 0; };
template <> struct FmtOfW<__nv_bfloat16> { static constexpr uint32_t v = 1; };

__device__ __forceinline__ void tma_reduce_add_2d(const void* tmap, uint32_t src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(src), "r"(c0), "r"(c1) : "memory");
}

template <typename T>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(WG_THREADS, 1) wgrad_umma2_kernel(const __grid_constant__ WgradParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t smem_a = smem_base;
    const uint32_t smem_b = smem_base + WG_STAGES * WG_A_BYTES;
    const uint32_t epi_base = smem_base + WG_STAGES * WG_STAGE_BYTES;
    const uint32_t bar_base = epi_base + WG_EPI_WARPS * 2 * WG_EPI_BUF;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (WG_STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * WG_STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * WG_STAGES + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * WG_STAGES + 4);
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int tiles = p.tiles_n * p.tiles_k;
    const int num_items = tiles * p.splits;
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    const int kb_total = (p.M + WG_BK - 1) / WG_BK;

    if (warp == 0 && lane == 0) { tma_prefetch_desc(&p.tmA); tma_prefetch_desc(&p.tmB); tma_prefetch_desc(&p.tmOut); }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < WG_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
            for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 2 * WG_EPI_WARPS); }
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc_cg2(tmem_slot, WG_TMEM_COLS);
        tmem_relinquish_cg2();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    // item -> (tile, split): split-major, so concurrently running clusters walk the same token rows
    auto kb_range = [&](int split, int& kb0, int& kb1) {
        kb0 = split * p.kb_per_split;
        kb1 = min(kb0 + p.kb_per_split, kb_total);
    };

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        int stage = 0; uint32_t phase = 0;
        for (int item = cluster_id; item < num_items; item += num_clusters) {
            const int split = item / tiles, tile = item - split * tiles;
            const int tn = tile / p.tiles_k, tk = tile - tn * p.tiles_k;
            const int n0 = tn * (2 * WG_BM) + static_cast<int>(rank) * WG_BM;          // this CTA's dY columns
            const int k0 = tk * WG_BN + static_cast<int>(rank) * (WG_BN / 2);          // this CTA's X columns
            int kb0, kb1;
            kb_range(split, kb0, kb1);
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(empty_bar(stage), phase ^ 1u);
                const uint32_t fb = mapa_shared(full_bar(stage), 0);
                if (elect_one()) {
                    if (leader) mbar_arrive_expect_tx(full_bar(stage), 2 * WG_STAGE_BYTES);
                    const uint32_t sa = smem_a + stage * WG_A_BYTES, sb = smem_b + stage * WG_B_BYTES;
                    tma_load_2d_cg2(sa, &p.tmA, fb, n0, kb * WG_BK);
                    tma_load_2d_cg2(sa + WG_BOX_BYTES, &p.tmA, fb, n0 + 64, kb * WG_BK);
                    tma_load_2d_cg2(sb, &p.tmB, fb, k0, kb * WG_BK);
                    tma_load_2d_cg2(sb + WG_BOX_BYTES, &p.tmB, fb, k0 + 64, kb * WG_BK);
                }
                if (++stage == WG_STAGES) { stage = 0; phase ^= 1u; }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader) {
            // both operands MN-major: a_major (bit 15) and b_major (bit 16) set
            constexpr uint32_t idesc = umma_idesc_f16(FmtOfW<T>::v, 2 * WG_BM, WG_BN) | (1u << 15) | (1u << 16);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int item = cluster_id; item < num_items; item += num_clusters) {
                const int split = item / tiles;
                int kb0, kb1;
                kb_range(split, kb0, kb1);
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * WG_BN);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t sa = smem_a + stage * WG_A_BYTES, sb = smem_b + stage * WG_B_BYTES;
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < WG_BK / WG_UK; ++k) {
                            // 16 token rows = two 8-row swizzle atoms = 2048 bytes further down each 64-column box
                            const uint64_t adesc = umma_desc_mn_sw128(sa + k * 2048, WG_BOX_BYTES);
                            const uint64_t bdesc = umma_desc_mn_sw128(sb + k * 2048, WG_BOX_BYTES);
                            umma_f16_ss_cg2(d_tmem, adesc, bdesc, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
                        }
                        umma_commit_cg2(empty_bar(stage), 3);
                    }
                    if (++stage == WG_STAGES) { stage = 0; phase ^= 1u; }
                }
                if (elect_one()) umma_commit_cg2(tfull_bar(acc), 3);
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue (warps 2..9, both CTAs): TMEM -> registers -> swizzled smem -> TMA reduce-add =====
        const int e = warp - 2;
        const int quarter = warp & 3;
        const int half = e >> 2;
        const uint32_t buf0 = epi_base + static_cast<uint32_t>(e * 2) * WG_EPI_BUF;
        const uint32_t my_row = static_cast<uint32_t>(lane) * 128u;
        const uint32_t swz = static_cast<uint32_t>(lane & 7);
        const uint32_t tempty_leader0 = mapa_shared(tempty_bar(0), 0), tempty_leader1 = mapa_shared(tempty_bar(1), 0);
        constexpr int BLOCKS = WG_BN / 32;
        int acc = 0; uint32_t acc_phase = 0;
        uint32_t nbuf = 0;
        for (int item = cluster_id; item < num_items; item += num_clusters) {
            const int split = item / tiles, tile = item - split * tiles;
            const int tn = tile / p.tiles_k, tk = tile - tn * p.tiles_k;
            const int row0 = tn * (2 * WG_BM) + static_cast<int>(rank) * WG_BM + quarter * 32;     // dW row (n) of lane 0
            const int k0 = tk * WG_BN;
            const bool rows_live = row0 < p.N;
            int kb0, kb1;
            kb_range(split, kb0, kb1);
            const bool has_work = kb1 > kb0;                  // an empty split never issues an MMA: nothing to add
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * WG_BN);
            if (rows_live && has_work) {
                for (int i = 0; i < BLOCKS / 2; ++i) {
                    const int cb = half + 2 * i;
                    const int col0 = k0 + cb * 32;
                    if (col0 >= p.K) break;
                    const uint32_t buf = buf0 + (nbuf & 1u) * WG_EPI_BUF;
                    if (lane == 0) tma_store_wait_read<1>();      // the reduce that last used this buffer has read it
                    __syncwarp();
                    uint32_t v[32];
                    tmem_ld_32x32(t_addr + static_cast<uint32_t>(cb * 32), v);
                    tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const uint32_t addr = buf + my_row + ((static_cast<uint32_t>(c) ^ swz) << 4);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                                     ::"r"(addr), "r"(v[4 * c]), "r"(v[4 * c + 1]), "r"(v[4 * c + 2]), "r"(v[4 * c + 3]) : "memory");
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_reduce_add_2d(&p.tmOut, buf, col0, row0);
                        tma_store_commit();
                    }
                    ++nbuf;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(acc ? tempty_leader1 : tempty_leader0);
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
        if (lane == 0) tma_store_wait<0>();
        __syncwarp();
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_cg2(tmem_base, WG_TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// fp32 CUDA-core version (compute_dtype = fp32 parity mode, and shapes the tensor-core kernel does not take)
// ---------------------------------------------------------------------------------------------------------------------
constexpr int WS_T = 64, WS_MK = 16;

template <typename TA>
__global__ void __launch_bounds__(256) wgrad_simt_kernel(const TA* __restrict__ dY, int ldy, const TA* __restrict__ X, int ldx,
                                                         float* __restrict__ dW, int ldw, int M, int N, int K, int rows_per_split) {
    __shared__ float Ys[WS_MK][WS_T + 4];
    __shared__ float Xs[WS_MK][WS_T + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int k0 = blockIdx.x * WS_T, n0 = blockIdx.y * WS_T;
    const int m_lo = blockIdx.z * rows_per_split, m_hi = min(M, m_lo + rows_per_split);
    const int lm = tid >> 4, lc = (tid & 15) * 4;           // loader: row lm of the 16-row slab, 4 consecutive columns
    float acc[4][4] = {};
    for (int m0 = m_lo; m0 < m_hi; m0 += WS_MK) {
        const int m = m0 + lm;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + lc + j, k = k0 + lc + j;
            Ys[lm][lc + j] = (m < m_hi && n < N) ? to_float<TA>(dY[static_cast<size_t>(m) * ldy + n]) : 0.0f;
            Xs[lm][lc + j] = (m < m_hi && k < K) ? to_float<TA>(X[static_cast<size_t>(m) * ldx + k]) : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int mm = 0; mm < WS_MK; ++mm) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = Ys[mm][ty * 4 + i]; b[i] = Xs[mm][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int n = n0 + ty * 4 + i;
        if (n >= N) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + tx * 4 + j;
            if (k < K) atomicAdd(dW + static_cast<size_t>(n) * ldw + k, acc[i][j]);
        }
    }
}

}  // namespace

bool wgrad_umma_supported(int M, int N, int K, int ldy, int ldx, int ldw) {
    return M > 0 && N > 0 && K >= 64 && (ldy % 8) == 0 && (ldx % 8) == 0 && (ldw % 4) == 0 && (K % 4) == 0;
}

// number of splits over the token rows: fill whole waves of CTA pairs, keep >= 8 row blocks per split
int wgrad_pick_splits(int M, int N, int K, int num_sms) {
    const int tiles = ((N + 2 * WG_BM - 1) / (2 * WG_BM)) * ((K + WG_BN - 1) / WG_BN);
    const int clusters = num_sms / 2 > 0 ? num_sms / 2 : 1;
    const int kb_total = (M + WG_BK - 1) / WG_BK;
    int best = 1;
    double best_eff = 0.0;
    for (int s = 1; s <= 64; ++s) {
        if (s > 1 && kb_total / s < 8) break;
        const int items = tiles * s;
        const int waves = (items + clusters - 1) / clusters;
        const double eff = static_cast<double>(items) / (static_cast<double>(waves) * clusters);
        if (eff > best_eff + 1e-9) { best_eff = eff; best = s; }
    }
    return best;
}

template <typename T>
cudaError_t launch_wgrad_umma(WgradParams p, int num_sms, cudaStream_t s) {
    p.tiles_n = (p.N + 2 * WG_BM - 1) / (2 * WG_BM);
    p.tiles_k = (p.K + WG_BN - 1) / WG_BN;
    if (p.splits <= 0) p.splits = wgrad_pick_splits(p.M, p.N, p.K, num_sms);
    const int kb_total = (p.M + WG_BK - 1) / WG_BK;
    p.kb_per_split = (kb_total + p.splits - 1) / p.splits;
    p.splits = (kb_total + p.kb_per_split - 1) / p.kb_per_split;       // no empty split
    const int items = p.tiles_n * p.tiles_k * p.splits;
    if (items <= 0) return cudaSuccess;
    auto kern = wgrad_umma2_kernel<T>;
    static SmemAttrCache cache;
    if (cudaError_t e = ensure_dynamic_smem(kern, WG_SMEM, cache); e != cudaSuccess) return e;
    int clusters = num_sms / 2;
    if (clusters > items) clusters = items;
    kern<<<2 * clusters, WG_THREADS, WG_SMEM, s>>>(p);
    return cudaGetLastError();
}
template cudaError_t launch_wgrad_umma<__half>(WgradParams, int, cudaStream_t);
template cudaError_t launch_wgrad_umma<__nv_bfloat16>(WgradParams, int, cudaStream_t);

template <typename TA>
cudaError_t launch_wgrad_simt(const TA* dY, int ldy, const TA* X, int ldx, float* dW, int ldw, int M, int N, int K, cudaStream_t s) {
    if (M <= 0 || N <= 0 || K <= 0) return cudaSuccess;
    int splits = 1;
    const long long tiles = static_cast<long long>((N + WS_T - 1) / WS_T) * ((K + WS_T - 1) / WS_T);
    while (tiles * splits < 592 && M / (splits * 2) >= 256) splits *= 2;
    if (splits > 65535) splits = 65535;
    int rps = (M + splits - 1) / splits;
    rps = (rps + WS_MK - 1) / WS_MK * WS_MK;
    splits = (M + rps - 1) / rps;
    dim3 grid((K + WS_T - 1) / WS_T, (N + WS_T - 1) / WS_T, splits);
    if (grid.y > 65535) return cudaErrorInvalidValue;
    wgrad_simt_kernel<TA><<<grid, 256, 0, s>>>(dY, ldy, X, ldx, dW, ldw, M, N, K, rps);
    return cudaGetLastError();
}
template cudaError_t launch_wgrad_simt<float>(const float*, int, const float*, int, float*, int, int, int, int, cudaStream_t);
template cudaError_t launch_wgrad_simt<__half>(const __half*, int, const __half*, int, float*, int, int, int, int, cudaStream_t);
template cudaError_t launch_wgrad_simt<__nv_bfloat16>(const __nv_bfloat16*, int, const __nv_bfloat16*, int, float*, int, int, int, int, cudaStream_t);

}  // namespace tim
