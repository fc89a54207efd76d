// CTA-pair tcgen05 GEMM (cta_group::2) for the big dense contractions of the TIM encoder layers
// (in_proj, out_proj, linear1, linear2; also the embedders / time-MLP / regression-head layers when N % 64 == 0):
//     C[M, N] = A[M, K] * W[N, K]^T      16-bit operands, fp32 accumulate in TMEM
// Replaces the torch addmm calls under */models/helpers/transformers.py:102-109 (no reference kernel exists).
//
// Why a second kernel: the single-CTA 128 x 256 tile of gemm_umma.cu pulls 48 KB from L2 per 64-wide k-step
// (96 B/clk/SM), measured at 59 % tensor-pipe for in_proj and 26-45 % when the epilogue does uncoalesced fp32
// traffic (profiles/r01a_*). Here two CTAs of a cluster share one 256 x 256 tile: each stages only its own 128 A rows
// and 128 W rows (32 KB per k-step, 64 B/clk/SM), the leader issues tcgen05.mma.cta_group::2 for both, and the
// epilogue goes through swizzled shared memory and TMA (bulk tensor stores, TMA loads for the fp32 residual) so that
// global traffic is full 128-byte lines.
//
// Per CTA (320 threads):
//   warp 0    : TMA producer (own A / W halves; completion bytes are signalled on the LEADER's full barrier)
//   warp 1    : leader: MMA issuer (256 x 256 x 16 per instruction); both: TMEM allocation (512 columns = 2 accumulators)
//   warps 2-9 : epilogue; warp w owns TMEM lanes 32*(w%4).. and every other column block of the tile
// Pipelines: smem full/empty ring (5 stages; 4 in mode 5), TMEM full/empty (2 accumulators), per-warp TMA store / residual-load
// double buffers. Persistent: cluster c walks tiles c, c + #clusters, ... (N fastest, so A tiles are shared through L2).
//
// Epilogue modes (template MODE): 0 16-bit out; 1 fp32 out; 2 fp32 out + fp32 residual; 3 as 2 with LayerNorm applied to the
// residual rows on read; 5 / 6 the folded-LayerNorm producer / consumer pair that removes the LayerNorm kernels of the encoder
// stack (transformers.py:105,109) - described in front of the kernel. Measured per-launch numbers: profiles/README.md.
#include <cstdlib>

#include "kernels.h"
#include "ptx.cuh"

namespace tim {
namespace {

constexpr int BM = 128;                 // rows per CTA (256 per pair)
constexpr int BN = 256;                 // columns per pair tile; each CTA stages BN/2 weight rows
constexpr int BK = 64;                  // one 128-byte swizzle row of 16-bit elements
constexpr int UK = 16;
constexpr int STAGES_MAX = 5;
constexpr int A_BYTES = BM * BK * 2;            // 16 KB
constexpr int B_BYTES = (BN / 2) * BK * 2;      // 16 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int EPI_WARPS = 8;
constexpr int THREADS = 64 + EPI_WARPS * 32;
constexpr int EPI_BUF = 32 * 128;               // 32 rows x 128 B
constexpr int EPI_BUFS = 2;                     // per warp
constexpr int BAR_BYTES = 512;
constexpr int SMEM_BYTES = STAGES_MAX * STAGE_BYTES + EPI_WARPS * EPI_BUFS * EPI_BUF + BAR_BYTES + 1024;
constexpr int TMEM_COLS = 512;

template <typename T> struct FmtOf2;
template <> struct FmtOf2<__half> { static constexpr uint32_t v = 0; };
template <> struct FmtOf2<__nv_bfloat16> { static constexpr uint32_t v = 1; };

// activations on PAIRS of values: the epilogue arithmetic runs on the packed fp32 instructions (FADD2 / FFMA2 / FMUL2, ptx.cuh)
template <int ACT> __device__ __forceinline__ float2 apply_act2(float2 v) {
    if (ACT == ACT_RELU) return make_float2(fmaxf(v.x, 0.0f), fmaxf(v.y, 0.0f));
    if (ACT == ACT_GELU) return gelu_erf_fast2(v);
    return v;
}

// MODE 0: 16-bit output, 1: fp32 output, 2: fp32 output + fp32 residual (added after the activation),
// 3: as 2 with LayerNorm applied to the residual rows on the fly (row statistics in p.rstats, affine in p.rgamma / p.rbeta)
// Folded-LayerNorm pair (the LayerNorm between two GEMMs costs no kernel and no extra pass over memory):
// 5: producer - z = acc + bias + resid (resid raw when p.rstats == nullptr, else LayerNorm-on-read as in mode 3), written as
//    fp32 (tmOut) AND as the 16-bit operand copy (tmOut16), plus per-row partial sums (sum z, sum z^2) of the 128 columns this
//    thread owns in the tile: p.opart[(2 tn + half) * M + row]  (row_stats_finalize_kernel turns them into (mean, rstd));
// 6: consumer - A = the 16-bit copy of z, W pre-multiplied by gamma: out = T(act(rstd * (acc - mean * cs[n]) + bw[n])) with
//    (mean, rstd) = p.rstats[row], cs[n] = sum_k W'[n,k], bw[n] = sum_k beta[k] W[n,k] + bias[n] (passed as p.bias),
// Training-leg fusions: 8: two 16-bit outputs, T(acc + bias) -> tmOut (the pre-activation the backward needs) and T(act(that)) -> tmOut16
//    (linear1 + GELU without the separate activation pass); 9: 16-bit out = T((acc + bias) * gelu'(u)) with the 16-bit pre-activation
//    tile u read through tmRes (TMA, same box as the output, overwritten in place) - linear2's dgrad with the GELU backward inside.
// 7: producer on a TWO-PLANE residual stream: z is kept as hi = T(z), lo = T(z - hi) (two 16-bit planes, ~22 significant bits with
//    fp16; tools/residual_split_study.py: 5.9e-7 end to end) instead of fp32 + a 16-bit copy. The hi plane IS the A operand of the
//    consumer GEMM, so a producer writes 4 bytes per element where mode 5 writes 6, reads the residual as two 16-bit tiles (TMA, box
//    32 x 32, 64-byte swizzle) and overwrites them in shared memory with the output planes before the TMA stores. Per warp that is
//    two 4 KB slots, the same as the plain modes - so mode 7 keeps all 5 pipeline stages where mode 5 gave one up.
template <typename T, int MODE, int ACT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1) linear_umma2_kernel(const __grid_constant__ Umma2Params p) {
    // MODE 5 trades one pipeline stage for a third epilogue buffer per warp (the 16-bit copy); same total
    constexpr int STAGES = MODE == 5 ? STAGES_MAX - 1 : STAGES_MAX;
    constexpr int EPI_BUFS_W = MODE == 5 ? 3 : 2;
    static_assert(STAGES * STAGE_BYTES + EPI_WARPS * EPI_BUFS_W * EPI_BUF <= STAGES_MAX * STAGE_BYTES + EPI_WARPS * EPI_BUFS * EPI_BUF, "smem");
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t smem_a = smem_base;
    const uint32_t smem_b = smem_base + STAGES * A_BYTES;
    const uint32_t epi_base = smem_base + STAGES * STAGE_BYTES;
    const uint32_t bar_base = epi_base + EPI_WARPS * EPI_BUFS_W * EPI_BUF;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
    auto epild_bar = [&](int w, int b) { return bar_base + 8u * (2 * STAGES + 4 + w * EPI_BUFS + b); };     // residual loads: 2 per warp
    const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4 + EPI_WARPS * EPI_BUFS);
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);     // warp-uniform for the compiler
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int num_kb = (p.K + BK - 1) / BK;
    const int num_tiles = p.tiles_m * p.tiles_n;
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tmA);
        tma_prefetch_desc(&p.tmB);
        tma_prefetch_desc(&p.tmOut);
        if (MODE == 2 || MODE == 3 || MODE == 5 || MODE == 7 || MODE == 9) tma_prefetch_desc(&p.tmRes);
        if (MODE == 5 || MODE == 8) tma_prefetch_desc(&p.tmOut16);
        if (MODE == 7) { tma_prefetch_desc(&p.tmResLo); tma_prefetch_desc(&p.tmOutLo); }
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
            for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 2 * EPI_WARPS); }
            for (int w = 0; w < EPI_WARPS; ++w)
                for (int b = 0; b < EPI_BUFS; ++b) mbar_init(epild_bar(w, b), 1);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc_cg2(tmem_slot, TMEM_COLS);
        tmem_relinquish_cg2();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    // everything above touched only this CTA's shared memory / TMEM and the kernel parameters: under a programmatic dependent launch it
    // ran while the previous kernel of the stream was draining. From here on global memory is read and written.
    pdl_launch_dependents();
    pdl_wait();

    // Producer and MMA warps run their loops WARP-UNIFORMLY (all 32 lanes wait on the barriers and carry the loop state) and
    // only the asynchronous-issue instructions sit under elect_one(): the descriptors, coordinates and barrier addresses are
    // then provably uniform and live in uniform registers, instead of being moved there (R2UR + ELECT) in front of every
    // UTCHMMA / UTMALDG of a single divergent thread - the MMA issue loop shares its scheduler with two busy epilogue warps
    // and its instruction count is what keeps the tensor pipe fed.
    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        int stage = 0; uint32_t phase = 0;
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
            const int tm = tile / p.tiles_n, tn = tile % p.tiles_n;
            const int m0 = tm * (2 * BM) + static_cast<int>(rank) * BM;
            const int n0 = tn * BN + static_cast<int>(rank) * (BN / 2);
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(empty_bar(stage), phase ^ 1u);
                const uint32_t fb = mapa_shared(full_bar(stage), 0);     // the leader's barrier
                if (elect_one()) {
                    if (leader) mbar_arrive_expect_tx(full_bar(stage), 2 * STAGE_BYTES);
                    tma_load_2d_cg2(smem_a + stage * A_BYTES, &p.tmA, fb, kb * BK, m0);
                    tma_load_2d_cg2(smem_b + stage * B_BYTES, &p.tmB, fb, kb * BK, n0);
                }
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader) {
            constexpr uint32_t idesc = umma_idesc_f16(FmtOf2<T>::v, 2 * BM, BN);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint64_t adesc = umma_desc_sw128(smem_a + stage * A_BYTES);
                    const uint64_t bdesc = umma_desc_sw128(smem_b + stage * B_BYTES);
                    if (elect_one()) {                          // the same (lowest) lane every time: commits track its MMAs
#pragma unroll
                        for (int k = 0; k < BK / UK; ++k)
                            umma_f16_ss_cg2(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
                        umma_commit_cg2(empty_bar(stage), 3);   // frees this stage in both CTAs
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
                if (elect_one()) umma_commit_cg2(tfull_bar(acc), 3);     // accumulator complete -> both epilogues
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
        __syncwarp();
    } else if (warp < 2 + EPI_WARPS) {
        // ===================== epilogue (warps 2..9, both CTAs) =====================
        const int e = warp - 2;
        const int quarter = warp & 3;                     // TMEM lane quarter this warp may access
        const int half = e >> 2;                          // which interleaved set of column blocks
        const uint32_t buf0 = epi_base + static_cast<uint32_t>(e * EPI_BUFS_W) * EPI_BUF;
        const uint32_t buf16 = buf0 + 2 * EPI_BUF;       // MODE 5 only: 32 rows x 64 columns of the 16-bit copy
        const uint32_t my_row = static_cast<uint32_t>(lane) * 128u;
        const uint32_t swz = static_cast<uint32_t>(lane & 7);
        const uint32_t tempty_leader0 = mapa_shared(tempty_bar(0), 0), tempty_leader1 = mapa_shared(tempty_bar(1), 0);
        constexpr bool OUT16 = MODE == 0 || MODE == 6 || MODE == 8 || MODE == 9;
        constexpr bool RES = MODE == 2 || MODE == 3 || MODE == 5 || MODE == 7 || MODE == 9;      // a tile comes in through tmRes
        constexpr bool TWO_OUT = MODE == 8;               // pre-activation -> tmOut (slot 0), activation -> tmOut16 (slot 1)
        constexpr bool DACT = MODE == 9;                  // the tmRes tile is the 16-bit pre-activation u: out = (acc + bias) * gelu'(u)
        constexpr int ACT_F = TWO_OUT ? ACT_NONE : ACT;
        constexpr bool DUAL = MODE == 5;
        constexpr bool PLANES = MODE == 7;
        constexpr bool FOLD = MODE == 6;
        const uint32_t sw64 = static_cast<uint32_t>((lane >> 1) & 3);      // 64-byte swizzle of the 32 x 32 16-bit plane tiles
        constexpr int COLS_PER_BLOCK = OUT16 ? 64 : 32;          // 128 B of output per row
        constexpr int BLOCKS = BN / COLS_PER_BLOCK;
        // column block handled in this warp's i-th step: every other block; MODE 5 takes adjacent PAIRS of 32-column blocks
        // (0,1),(4,5) / (2,3),(6,7) so that each pair is one 64-column group of the 16-bit copy
        auto block_of = [&](int i) { return DUAL ? ((i >> 1) * 4 + 2 * half + (i & 1)) : (half + 2 * i); };
        int acc = 0; uint32_t acc_phase = 0;
        uint32_t nbuf = 0;                                // running buffer counter (selects buffer and its load parity)
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
            const int tm = tile / p.tiles_n, tn = tile % p.tiles_n;
            const int row0 = tm * (2 * BM) + static_cast<int>(rank) * BM + quarter * 32;
            const int n0 = tn * BN;
            const bool rows_live = row0 < p.M;            // warp-uniform: nothing to store for a fully out-of-range slice
            const bool row_ok = row0 + lane < p.M;
            // number of column blocks of this tile that exist (N % 64 == 0, so blocks are all-or-nothing)
            int my_blocks = 0;
            for (int i = 0; i < BLOCKS / 2; ++i) my_blocks += (n0 + block_of(i) * COLS_PER_BLOCK < p.N) ? 1 : 0;
            if (!rows_live) my_blocks = 0;

            // LayerNorm of the residual row on read: r -> (r * rstd - mean * rstd) * gamma + beta   (MODE 3, MODE 5 with rstats)
            float ln_rstd = 1.0f, ln_nmr = 0.0f;
            const bool ln_res = MODE == 3 || ((DUAL || PLANES) && p.rstats != nullptr);
            if (ln_res && row_ok) {
                const float2 st = p.rstats[row0 + lane];
                ln_rstd = st.y; ln_nmr = -st.x * st.y;
            }
            // folded LayerNorm of the A rows (MODE 6): out = act(fa * acc + fb * cs[n] + bw[n])
            float fa = 1.0f, fb = 0.0f;
            if (FOLD && row_ok) {
                const float2 st = p.rstats[row0 + lane];
                fa = st.y; fb = -st.x * st.y;
            }
            float st_sum = 0.0f, st_sq = 0.0f;            // MODE 5: sum / sum of squares of this thread's z columns in this tile
            float2 st_sum2 = make_float2(0.0f, 0.0f), st_sq2 = make_float2(0.0f, 0.0f);      // MODE 7: the same as (even, odd) column pairs

            if (RES && my_blocks > 0) {             // prefetch the residual of the first block
                const uint32_t b = nbuf & 1u;
                if (lane == 0) {
                    tma_store_wait_read<0>();             // that buffer's previous store has drained
                    mbar_arrive_expect_tx(epild_bar(e, b), EPI_BUF);
                    tma_load_2d(buf0 + b * EPI_BUF, &p.tmRes, epild_bar(e, b), n0 + block_of(0) * COLS_PER_BLOCK, row0);
                    if (PLANES) tma_load_2d(buf0 + b * EPI_BUF + EPI_BUF / 2, &p.tmResLo, epild_bar(e, b), n0 + block_of(0) * COLS_PER_BLOCK, row0);
                }
                __syncwarp();
            }
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * BN);

            for (int i = 0; i < my_blocks; ++i) {
                const int cb = block_of(i);
                const int col0 = n0 + cb * COLS_PER_BLOCK;
                const uint32_t b = TWO_OUT ? 0u : (nbuf & 1u);
                const uint32_t buf = buf0 + b * EPI_BUF;
                if (RES) {
                    if (i + 1 < my_blocks) {              // prefetch the next block's residual into the other buffer
                        if (lane == 0) {
                            tma_store_wait_read<0>();
                            mbar_arrive_expect_tx(epild_bar(e, b ^ 1u), EPI_BUF);
                            tma_load_2d(buf0 + (b ^ 1u) * EPI_BUF, &p.tmRes, epild_bar(e, b ^ 1u), n0 + block_of(i + 1) * COLS_PER_BLOCK, row0);
                            if (PLANES)
                                tma_load_2d(buf0 + (b ^ 1u) * EPI_BUF + EPI_BUF / 2, &p.tmResLo, epild_bar(e, b ^ 1u), n0 + block_of(i + 1) * COLS_PER_BLOCK, row0);
                        }
                        __syncwarp();
                    }
                    mbar_wait(epild_bar(e, b), (nbuf >> 1) & 1u);
                } else {
                    if (lane == 0) {                      // the store(s) that last used this buffer have read it
                        if (TWO_OUT) tma_store_wait_read<0>(); else tma_store_wait_read<EPI_BUFS - 1>();
                    }
                    __syncwarp();
                }
#pragma unroll
                for (int part = 0; part < COLS_PER_BLOCK / 32; ++part) {
                    uint32_t v[32];
                    tmem_ld_32x32(t_addr + static_cast<uint32_t>(cb * COLS_PER_BLOCK + part * 32), v);
                    tmem_ld_wait();
                    float f[32];
                    const int n = col0 + part * 32;
                    const float2 fa2 = make_float2(fa, fa), fb2 = make_float2(fb, fb);
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n + j));
                        float2 lo = make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1]));
                        float2 hi = make_float2(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                        if (FOLD) {
                            const float4 c4 = __ldg(reinterpret_cast<const float4*>(p.cs + n + j));
                            lo = __ffma2_rn(fa2, lo, __ffma2_rn(fb2, make_float2(c4.x, c4.y), make_float2(b4.x, b4.y)));
                            hi = __ffma2_rn(fa2, hi, __ffma2_rn(fb2, make_float2(c4.z, c4.w), make_float2(b4.z, b4.w)));
                        } else {
                            lo = __fadd2_rn(lo, make_float2(b4.x, b4.y));
                            hi = __fadd2_rn(hi, make_float2(b4.z, b4.w));
                        }
                        lo = apply_act2<ACT_F>(lo); hi = apply_act2<ACT_F>(hi);
                        f[j] = lo.x; f[j + 1] = lo.y; f[j + 2] = hi.x; f[j + 3] = hi.y;
                    }
                    if (OUT16) {
                        // 32 columns -> 64 B = logical 16-byte chunks part*4 .. part*4+3 of this row
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const uint32_t chunk = static_cast<uint32_t>(part * 4 + c);
                            const uint32_t addr = buf + my_row + ((chunk ^ swz) << 4);
                            if (DACT) {                   // the same 16 bytes of the pre-activation tile that the result overwrites
                                const uint4 uq = lds_u128(addr);
                                const uint32_t uw[4] = {uq.x, uq.y, uq.z, uq.w};
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    const float2 d2 = __fmul2_rn(make_float2(f[8 * c + 2 * k], f[8 * c + 2 * k + 1]), gelu_grad_fast2(unpack2<T>(uw[k])));
                                    f[8 * c + 2 * k] = d2.x; f[8 * c + 2 * k + 1] = d2.y;
                                }
                            }
                            uint4 q;
                            q.x = pack2<T>(f[8 * c], f[8 * c + 1]); q.y = pack2<T>(f[8 * c + 2], f[8 * c + 3]);
                            q.z = pack2<T>(f[8 * c + 4], f[8 * c + 5]); q.w = pack2<T>(f[8 * c + 6], f[8 * c + 7]);
                            sts_u128(addr, q);
                            if (TWO_OUT) {                // act() of the ROUNDED pre-activation: what a separate pass over `out` would compute
                                const uint32_t qw[4] = {q.x, q.y, q.z, q.w};
                                uint32_t hw[4];
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    const float2 h2 = apply_act2<ACT>(unpack2<T>(qw[k]));
                                    hw[k] = pack2<T>(h2.x, h2.y);
                                }
                                sts_u128(addr + EPI_BUF, make_uint4(hw[0], hw[1], hw[2], hw[3]));
                            }
                        }
                    } else if (PLANES) {
                        // residual = hi + lo (two 32 x 32 16-bit tiles, 64-byte rows, 64-byte swizzle); the output planes overwrite them.
                        // The eight epilogue warps are the serial resource of out_proj (K = 1024: 6 us of MMA per tile against 10 us of
                        // epilogue, issue slots 46 % busy with two warps per scheduler), so the arithmetic runs on packed fp32 pairs
                        // (FADD2 / FFMA2) and gamma / beta come as 16-byte loads off one hoisted pointer.
                        const float2 rs2 = make_float2(ln_rstd, ln_rstd), nm2 = make_float2(ln_nmr, ln_nmr);
                        const float2 neg1 = make_float2(-1.0f, -1.0f);
                        const float4* g4p = reinterpret_cast<const float4*>(p.rgamma + n);
                        const float4* e4p = reinterpret_cast<const float4*>(p.rbeta + n);
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const uint32_t a_hi = buf + static_cast<uint32_t>(lane) * 64u + ((static_cast<uint32_t>(c) ^ sw64) << 4);
                            const uint32_t a_lo = a_hi + EPI_BUF / 2;
                            const uint4 qh = lds_u128(a_hi), ql = lds_u128(a_lo);
                            const uint32_t hw[4] = {qh.x, qh.y, qh.z, qh.w}, lw[4] = {ql.x, ql.y, ql.z, ql.w};
                            float2 g2[4], e2[4];
                            if (ln_res) {
                                const float4 ga = __ldg(g4p + 2 * c), gb = __ldg(g4p + 2 * c + 1);
                                const float4 ea = __ldg(e4p + 2 * c), eb = __ldg(e4p + 2 * c + 1);
                                g2[0] = make_float2(ga.x, ga.y); g2[1] = make_float2(ga.z, ga.w); g2[2] = make_float2(gb.x, gb.y); g2[3] = make_float2(gb.z, gb.w);
                                e2[0] = make_float2(ea.x, ea.y); e2[1] = make_float2(ea.z, ea.w); e2[2] = make_float2(eb.x, eb.y); e2[3] = make_float2(eb.z, eb.w);
                            }
                            uint32_t oh[4], ol[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                float2 r = __fadd2_rn(unpack2<T>(hw[j]), unpack2<T>(lw[j]));
                                if (ln_res) r = __ffma2_rn(__ffma2_rn(r, rs2, nm2), g2[j], e2[j]);
                                const float2 o = __fadd2_rn(make_float2(f[8 * c + 2 * j], f[8 * c + 2 * j + 1]), r);
                                oh[j] = pack2<T>(o.x, o.y);
                                const float2 d = __ffma2_rn(unpack2<T>(oh[j]), neg1, o);      // o - hi, exact product
                                ol[j] = pack2<T>(d.x, d.y);
                                st_sum2 = __fadd2_rn(st_sum2, o);
                                st_sq2 = __ffma2_rn(o, o, st_sq2);
                            }
                            sts_u128(a_hi, make_uint4(oh[0], oh[1], oh[2], oh[3]));
                            sts_u128(a_lo, make_uint4(ol[0], ol[1], ol[2], ol[3]));
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const uint32_t addr = buf + my_row + ((static_cast<uint32_t>(c) ^ swz) << 4);
                            float4 o = make_float4(f[4 * c], f[4 * c + 1], f[4 * c + 2], f[4 * c + 3]);
                            if (RES) {
                                float4 r;
                                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                             : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr) : "memory");
                                float2 r01 = make_float2(r.x, r.y), r23 = make_float2(r.z, r.w);
                                if (ln_res) {
                                    const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.rgamma + n) + c);
                                    const float4 e4 = __ldg(reinterpret_cast<const float4*>(p.rbeta + n) + c);
                                    const float2 rs2 = make_float2(ln_rstd, ln_rstd), nm2 = make_float2(ln_nmr, ln_nmr);
                                    r01 = __ffma2_rn(__ffma2_rn(r01, rs2, nm2), make_float2(g4.x, g4.y), make_float2(e4.x, e4.y));
                                    r23 = __ffma2_rn(__ffma2_rn(r23, rs2, nm2), make_float2(g4.z, g4.w), make_float2(e4.z, e4.w));
                                }
                                const float2 o01 = __fadd2_rn(make_float2(o.x, o.y), r01), o23 = __fadd2_rn(make_float2(o.z, o.w), r23);
                                o = make_float4(o01.x, o01.y, o23.x, o23.y);
                            }
                            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};"
                                         ::"r"(addr), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w) : "memory");
                            if (DUAL) { f[4 * c] = o.x; f[4 * c + 1] = o.y; f[4 * c + 2] = o.z; f[4 * c + 3] = o.w; }
                        }
                        if (DUAL) {
                            // 16-bit copy of the 32 z values: half of a 64-column (128-byte) row of buf16
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                uint4 q;
                                q.x = pack2<T>(f[8 * c], f[8 * c + 1]); q.y = pack2<T>(f[8 * c + 2], f[8 * c + 3]);
                                q.z = pack2<T>(f[8 * c + 4], f[8 * c + 5]); q.w = pack2<T>(f[8 * c + 6], f[8 * c + 7]);
                                const uint32_t chunk = static_cast<uint32_t>((i & 1) * 4 + c);
                                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                                             ::"r"(buf16 + my_row + ((chunk ^ swz) << 4)), "r"(q.x), "r"(q.y), "r"(q.z), "r"(q.w) : "memory");
                            }
#pragma unroll
                            for (int j = 0; j < 32; ++j) { st_sum += f[j]; st_sq = fmaf(f[j], f[j], st_sq); }
                        }
                    }
                }
                fence_proxy_async_smem();                 // generic-proxy smem writes -> visible to the TMA store
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&p.tmOut, buf, col0, row0);
                    if (TWO_OUT) tma_store_2d(&p.tmOut16, buf + EPI_BUF, col0, row0);
                    if (DUAL && (i & 1)) tma_store_2d(&p.tmOut16, buf16, col0 - 32, row0);
                    if (PLANES) tma_store_2d(&p.tmOutLo, buf + EPI_BUF / 2, col0, row0);
                    tma_store_commit();
                }
                ++nbuf;
            }
            if (PLANES) { st_sum = st_sum2.x + st_sum2.y; st_sq = st_sq2.x + st_sq2.y; }
            if ((DUAL || PLANES) && row_ok)              // also when this half owns no column of a ragged last tile: (0, 0)
                p.opart[static_cast<size_t>(tn * 2 + half) * p.M + row0 + lane] = make_float2(st_sum, st_sq);
            // all of this warp's TMEM reads of the accumulator are done
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(acc ? tempty_leader1 : tempty_leader0);
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
        if (lane == 0) tma_store_wait<0>();               // stores complete before the CTA (and its smem) goes away
        __syncwarp();
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_cg2(tmem_base, TMEM_COLS);
    }
}

template <typename T, int MODE, int ACT>
cudaError_t launch_mode(const Umma2Params& p, int num_sms, cudaStream_t s) {
    auto kern = linear_umma2_kernel<T, MODE, ACT>;
    static SmemAttrCache cache;
    if (cudaError_t e = ensure_dynamic_smem(kern, SMEM_BYTES, cache); e != cudaSuccess) return e;
    const int tiles = p.tiles_m * p.tiles_n;
    if (tiles <= 0) return cudaSuccess;
    int clusters = num_sms / 2;
    if (clusters > tiles) clusters = tiles;
    return launch_maybe_pdl(kern, 2u * clusters, THREADS, SMEM_BYTES, s, pdl_enabled(), p);
}

}  // namespace

bool pdl_enabled() {
    static const bool on = [] { const char* e = std::getenv("TIM_B200_PDL"); return e ? std::atoi(e) != 0 : true; }();
    return on;
}

bool umma2_supported(int M, int N, int K) { return M > 0 && N >= 128 && (N % 64) == 0 && K >= 64 && (K % 8) == 0; }

template <typename T>
cudaError_t launch_linear_umma2(Umma2Params p, int mode, int act, int num_sms, cudaStream_t s) {
    p.tiles_m = (p.M + 2 * BM - 1) / (2 * BM);
    p.tiles_n = (p.N + BN - 1) / BN;
#define TIM_U2(MODE_, ACT_) if (mode == MODE_ && act == ACT_) return launch_mode<T, MODE_, ACT_>(p, num_sms, s);
    TIM_U2(0, ACT_NONE) TIM_U2(0, ACT_RELU) TIM_U2(0, ACT_GELU)
    TIM_U2(1, ACT_NONE) TIM_U2(1, ACT_RELU) TIM_U2(1, ACT_GELU)
    TIM_U2(2, ACT_NONE) TIM_U2(2, ACT_RELU) TIM_U2(2, ACT_GELU)
    TIM_U2(3, ACT_NONE)
    TIM_U2(5, ACT_NONE)
    TIM_U2(6, ACT_NONE) TIM_U2(6, ACT_GELU)
    TIM_U2(7, ACT_NONE)
    TIM_U2(8, ACT_GELU)
    TIM_U2(9, ACT_NONE)
#undef TIM_U2
    return cudaErrorInvalidValue;
}
template cudaError_t launch_linear_umma2<__half>(Umma2Params, int, int, int, cudaStream_t);
template cudaError_t launch_linear_umma2<__nv_bfloat16>(Umma2Params, int, int, int, cudaStream_t);

}  // namespace tim
