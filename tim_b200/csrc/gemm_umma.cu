// tcgen05 GEMM for every dense contraction of the TIM forward (in_proj, out_proj, linear1, linear2, embedders,
// time-MLP layers 2/3, CLS and regression heads):   C[rows, N] = A[rows, K] * W[N, K]^T
//
// Replaces the torch addmm calls under  */models/helpers/transformers.py:102-109,  encodings.py:140-153,
// tim.py:66-74, head.py (reference is library calls only; there is no reference kernel).
//
// Structure (one persistent CTA per SM, 320 threads):
//   warp 0   : TMA producer  - A tile (128 rows x 64 k) and W tile (BLOCK_N x 64 k) per stage, 128B swizzle
//   warp 1   : MMA issuer    - one elected lane issues tcgen05.mma 128 x BLOCK_N x 16, fp32 accumulators in TMEM
//   warps 2-9: epilogue      - tcgen05.ld 32 lanes x 32 columns, bias / ReLU / erf-GELU / residual, vector stores; the two
//                              warps of a TMEM lane quarter take half of the tile's columns each (one epilogue warp per
//                              scheduler left the tcgen05.ld / bias-load / store latencies exposed: 20 % issue utilisation)
// Three mbarrier pipelines: smem full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue; two accumulator
// stages so the epilogue of tile i overlaps the main loop of tile i+1), and a static persistent tile schedule
// (tile = blockIdx.x + i * gridDim.x, N fastest so CTAs running concurrently share the A tile through L2).
#include "kernels.h"
#include "ptx.cuh"

namespace tim {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;                    // 64 x 2 B = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int NUM_THREADS = 320;                // TMA warp, MMA warp, 8 epilogue warps
constexpr int EPI_WARPS = 8;
constexpr int SMEM_BUDGET = 184 * 1024;          // pipeline stages (3 x 48 KB at BLOCK_N = 256); the rest holds the epilogue scratch

template <int BLOCK_N> struct Cfg {
    static constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int STAGES = SMEM_BUDGET / STAGE_BYTES;
    static constexpr int TMEM_COLS = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;   // two accumulator stages (power of two)
    // per-epilogue-warp scratch for the coalesced fp32 stores: a 32 x 33 float transpose tile + the 32 output-row offsets
    static constexpr int XPOSE_WARP_BYTES = 32 * 33 * 4 + 32 * 8;
    static constexpr int XPOSE_BYTES = EPI_WARPS * XPOSE_WARP_BYTES;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + XPOSE_BYTES;
    static_assert(SMEM_BYTES <= 227 * 1024, "exceeds the per-CTA shared memory of sm_100");
    static_assert((2 * STAGES + 4) * 8 + 4 <= 256, "barrier area too small");
};

template <typename T> struct FmtOf;
template <> struct FmtOf<__half> { static constexpr uint32_t v = 0; };
template <> struct FmtOf<__nv_bfloat16> { static constexpr uint32_t v = 1; };

template <typename T, int BLOCK_N>
__global__ void __launch_bounds__(NUM_THREADS, 1) linear_umma_kernel(const __grid_constant__ UmmaParams p) {
    using C = Cfg<BLOCK_N>;
    constexpr int STAGES = C::STAGES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t smem_a = smem_base;
    const uint32_t smem_b = smem_base + STAGES * A_STAGE_BYTES;
    const uint32_t bar_base = smem_base + STAGES * C::STAGE_BYTES;      // 8-byte aligned
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + STAGES * C::STAGE_BYTES + 8 * (2 * STAGES + 4));

    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);     // warp-uniform for the compiler
    const int lane = threadIdx.x & 31;
    const int num_kb = (p.K + BLOCK_K - 1) / BLOCK_K;
    const int num_tiles = p.tiles_m * p.tiles_n;
    const uint32_t a_bytes = static_cast<uint32_t>(p.rm.box_r) * p.rm.box_g * (BLOCK_K * 2);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tmA);
        tma_prefetch_desc(&p.tmB);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
            for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), EPI_WARPS); }
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, C::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    // producer and MMA warps: warp-uniform loops, elect_one() only around the asynchronous-issue instructions (see gemm_umma2.cu)
    if (warp == 0) {
        // ===================== TMA producer =====================
        int stage = 0; uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int tm = tile / p.tiles_n, tn = tile % p.tiles_n;
            const int g0 = (tm / p.tiles_r) * p.rm.box_g;
            const int r0 = (tm % p.tiles_r) * p.rm.box_r;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(empty_bar(stage), phase ^ 1u);
                if (elect_one()) {
                    mbar_arrive_expect_tx(full_bar(stage), a_bytes + C::B_STAGE_BYTES);
                    tma_load_3d(smem_a + stage * A_STAGE_BYTES, &p.tmA, full_bar(stage), kb * BLOCK_K, p.rm.a_row_off + r0, g0);
                    tma_load_2d(smem_b + stage * C::B_STAGE_BYTES, &p.tmB, full_bar(stage), kb * BLOCK_K, tn * BLOCK_N);
                }
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = umma_idesc_f16(FmtOf<T>::v, BLOCK_M, BLOCK_N);
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BLOCK_N);
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                const uint64_t adesc = umma_desc_sw128(smem_a + stage * A_STAGE_BYTES);
                const uint64_t bdesc = umma_desc_sw128(smem_b + stage * C::B_STAGE_BYTES);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        // advance 16 elements (32 B) along K inside the swizzle atom: +2 in the (addr >> 4) field
                        umma_f16_ss(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(empty_bar(stage));          // frees the smem stage once these MMAs retire
                }
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
            if (elect_one()) umma_commit(tfull_bar(acc));   // accumulator complete -> epilogue
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
        __syncwarp();
    } else {
        // ===================== epilogue (warps 2..9) =====================
        const int quarter = warp & 3;                        // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;                    // which half of the tile's columns
        const int row_in_tile = quarter * 32 + lane;
        const uint32_t xpose = bar_base + 256u + static_cast<uint32_t>(warp - 2) * C::XPOSE_WARP_BYTES;
        const uint32_t xrows = xpose + 32 * 33 * 4;          // 32 x int64: output element offset of each row of the quarter, -1 = no row
        const Epilogue& ep = p.ep;
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int tm = tile / p.tiles_n, tn = tile % p.tiles_n;
            const int g0 = (tm / p.tiles_r) * p.rm.box_g;
            const int r0 = (tm % p.tiles_r) * p.rm.box_r;
            const int gi = row_in_tile / p.rm.box_r, ri = row_in_tile - gi * p.rm.box_r;
            const bool row_ok = (gi < p.rm.box_g) && (g0 + gi < p.rm.G) && (r0 + ri < p.rm.R);
            const size_t orow = static_cast<size_t>(g0 + gi) * p.rm.out_group_rows + p.rm.out_row_off + r0 + ri;
            const int n0 = tn * BLOCK_N;

            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * BLOCK_N);
            const bool xp_tile = ep.out_fp32 && !ep.resid;   // the transposed-store path may be taken in this tile
            if (xp_tile) {
                __syncwarp();
                sts_u64(xrows + static_cast<uint32_t>(lane) * 8u, row_ok ? static_cast<unsigned long long>(orow) * static_cast<unsigned long long>(ep.ldo) : ~0ull);
                __syncwarp();
            }
#pragma unroll 1
            for (int c0 = half * (BLOCK_N / 2); c0 < (half + 1) * (BLOCK_N / 2); c0 += 32) {
                if (n0 + c0 >= p.N) break;                   // warp-uniform
                uint32_t v[32];
                tmem_ld_32x32(t_addr + c0, v);
                const int n = n0 + c0;
                const bool full = (n + 32 <= p.N);
                // bias of the chunk, fetched while the TMEM load is in flight
                float bia[32];
                if (ep.bias) {
                    if (full) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + n + j));
                            bia[j] = b4.x; bia[j + 1] = b4.y; bia[j + 2] = b4.z; bia[j + 3] = b4.w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) bia[j] = (n + j < p.N) ? __ldg(ep.bias + n + j) : 0.0f;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) bia[j] = 0.0f;
                }
                tmem_ld_wait();
                // fp32 rows whose pitch is not 16-byte aligned (e.g. the 3806-class action head) or a ragged last chunk: transpose
                // through smem so that every store instruction writes 32 consecutive floats of ONE row (warp-uniform choice)
                const bool xp = ep.out_fp32 && !ep.resid && (!full || (ep.ldo & 3) != 0);
                if (!row_ok && !xp) continue;
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) + bia[j];
                if (ep.act == ACT_RELU) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.0f);
                } else if (ep.act == ACT_GELU) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = gelu_erf(f[j]);
                }
                if (ep.resid) {
                    const float* rp = ep.resid + orow * ep.ldr + n;
                    if (ep.rstats) {                         // LayerNorm-on-read of the residual row
                        const float2 st = ep.rstats[orow];
                        const float nmr = -st.x * st.y;
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (n + j < p.N) f[j] += fmaf(fmaf(rp[j], st.y, nmr), __ldg(ep.rgamma + n + j), __ldg(ep.rbeta + n + j));
                    } else if (full && (ep.ldr & 3) == 0) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 r4 = *reinterpret_cast<const float4*>(rp + j);
                            f[j] += r4.x; f[j + 1] += r4.y; f[j + 2] += r4.z; f[j + 3] += r4.w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) if (n + j < p.N) f[j] += rp[j];
                    }
                }
                if (ep.out_fp32) {
                    float* op = reinterpret_cast<float*>(ep.out) + orow * ep.ldo + n;
                    if (xp) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) sts_f32(xpose + static_cast<uint32_t>((lane * 33 + j) * 4), f[j]);
                        __syncwarp();
                        float* obase = reinterpret_cast<float*>(ep.out) + n + lane;
                        const bool col_ok = n + lane < p.N;
#pragma unroll 8
                        for (int r = 0; r < 32; ++r) {
                            const unsigned long long ro = lds_u64(xrows + static_cast<uint32_t>(r) * 8u);      // broadcast read
                            const float val = lds_f32(xpose + static_cast<uint32_t>((r * 33 + lane) * 4));
                            if (ro != ~0ull && col_ok) obase[ro] = val;
                        }
                        __syncwarp();
                    } else if (full && (ep.ldo & 3) == 0) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            *reinterpret_cast<float4*>(op + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) if (n + j < p.N) op[j] = f[j];
                    }
                } else {
                    T* op = reinterpret_cast<T*>(ep.out) + orow * ep.ldo + n;
                    if (full && (ep.ldo & 7) == 0) {
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            uint4 q;
                            q.x = pack2<T>(f[j], f[j + 1]); q.y = pack2<T>(f[j + 2], f[j + 3]);
                            q.z = pack2<T>(f[j + 4], f[j + 5]); q.w = pack2<T>(f[j + 6], f[j + 7]);
                            *reinterpret_cast<uint4*>(op + j) = q;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) if (n + j < p.N) op[j] = from_float<T>(f[j]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

template <typename T, int BLOCK_N>
cudaError_t launch_one(const UmmaParams& p, int num_sms, cudaStream_t s) {
    using C = Cfg<BLOCK_N>;
    auto kern = linear_umma_kernel<T, BLOCK_N>;
    static SmemAttrCache cache;
    if (cudaError_t e = ensure_dynamic_smem(kern, C::SMEM_BYTES, cache); e != cudaSuccess) return e;
    const int tiles = p.tiles_m * p.tiles_n;
    if (tiles <= 0) return cudaSuccess;
    const int grid = tiles < num_sms ? tiles : num_sms;
    kern<<<grid, NUM_THREADS, C::SMEM_BYTES, s>>>(p);
    return cudaGetLastError();
}

}  // namespace

int umma_block_n(int N) { return N > 128 ? 256 : (N > 64 ? 128 : 64); }

template <typename T>
cudaError_t launch_linear_umma(UmmaParams p, int block_n, int num_sms, cudaStream_t s) {
    p.tiles_r = (p.rm.R + p.rm.box_r - 1) / p.rm.box_r;
    const int tiles_g = (p.rm.G + p.rm.box_g - 1) / p.rm.box_g;
    p.tiles_m = p.tiles_r * tiles_g;
    p.tiles_n = (p.N + block_n - 1) / block_n;
    switch (block_n) {
        case 256: return launch_one<T, 256>(p, num_sms, s);
        case 128: return launch_one<T, 128>(p, num_sms, s);
        case 64: return launch_one<T, 64>(p, num_sms, s);
        default: return cudaErrorInvalidValue;
    }
}

template cudaError_t launch_linear_umma<__half>(UmmaParams, int, int, cudaStream_t);
template cudaError_t launch_linear_umma<__nv_bfloat16>(UmmaParams, int, int, cudaStream_t);

}  // namespace tim
