// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma /
// commit / ld / fences), ldmatrix + mma.sync, cp.async. No CUTLASS dependency; encodings were checked
// against cute/arch/mma_sm100_desc.hpp (SmemDescriptor / InstrDescriptor bit layouts).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tim {

#ifndef TIM_SPIN_LIMIT
#define TIM_SPIN_LIMIT (1u << 24)   // bounded mbarrier spin: a protocol bug traps instead of hanging the box
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(pred));
    return pred != 0;
}

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0, spins = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred P;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, P;\n\t}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) break;
        if (++spins > TIM_SPIN_LIMIT) __trap();
    }
}
// generic-proxy smem writes -> visible to the async proxy (TMA / tcgen05 operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ------------------------------------------------------------------ programmatic dependent launch
// A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may become resident while its predecessor in the stream is
// still draining: everything in front of pdl_wait() (barrier init, TMEM allocation, descriptor prefetch) overlaps the predecessor's tail;
// pdl_wait() returns when the predecessor grid has COMPLETED and its memory is visible. pdl_launch_dependents() is the predecessor's side:
// once every CTA has issued it (or exited) the dependent grid may be scheduled onto whatever resources come free. Both are no-ops in a
// kernel launched the ordinary way / without a dependent.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ------------------------------------------------------------------ TMA
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// L2 prefetch of a tensor box (no shared-memory destination, no completion tracking)
__device__ __forceinline__ void tma_prefetch_l2_3d(const void* tmap, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];"
                 ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// ------------------------------------------------------------------ tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {   // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {                                 // whole warp
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {     // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers fp16 and bf16 operands, fp32 accumulate.
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrives when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// K-major operand tile, 128-byte swizzle, rows of 64 16-bit elements (128 B), 8-row groups 1024 B apart.
// SmemDescriptor: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout [61,64) (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// MN-major operand tile (the contraction index k runs over ROWS of 128 bytes; 64 16-bit elements of the M/N index are
// contiguous inside a row), 128-byte swizzle: canonical layout ((64, n), (8, k)) : ((1, LBO), (128 B, SBO)) - 8-row
// groups along k are SBO = 1024 B apart, successive 64-wide blocks of the M/N index are `lbo_bytes` apart
// (cute/atom/mma_traits_sm100.hpp, make_umma_desc<Major::MN>). Used for V in the attention P.V product.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
    return static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu) | (static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// InstrDescriptor kind::f16: c_format F32 (1<<4) | a_format [7,10) | b_format [10,13) (0 = F16, 1 = BF16) |
// a_major/b_major = 0 (K-major) | N>>3 at [17,23) | M>>4 at [24,29)
__host__ __device__ constexpr uint32_t umma_idesc_f16(uint32_t fmt, uint32_t M, uint32_t N) {
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives lane (32*(warp%4) + i), columns [c, c+32)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
// 32 lanes x 8 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------ clusters / CTA pairs (cta_group::2)
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {       // every thread of every CTA in the cluster
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address -> shared::cluster address of the same offset in CTA `rank`
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
// Arrive on a (possibly remote) barrier. Default semantics (release at CTA scope), as CUTLASS' ClusterBarrier::arrive does:
// the only thing published through these barriers is "my tcgen05.ld reads of the accumulator are done", which the preceding
// tcgen05.fence::before_thread_sync orders; the cluster-scope release used before cost a MEMBAR + ERRBAR per tile and warp
// (15-20 % of the epilogue warps' stall samples in profiles/r01e_ncu_full_linear_umma2).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// 2-CTA TMA load: data lands in this CTA's smem, the transaction bytes are signalled on `bar` (a shared::cluster
// address, normally the leader CTA's full barrier)
__device__ __forceinline__ void tma_load_2d_cg2(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_cg2(uint32_t dst_smem, uint32_t ncols) {   // one warp in EACH CTA of the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_cg2() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A (128 rows from each CTA) * B (N/2 rows from each CTA); issued by the leader CTA only
__device__ __forceinline__ void umma_f16_ss_cg2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives (once the previously issued MMAs retire) on the barrier at this smem offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_cg2(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}

// ------------------------------------------------------------------ TMA stores (smem -> global, bulk async group)
__device__ __forceinline__ void tma_store_2d(const void* tmap, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const void* tmap, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_store_wait_read() {   // <= N groups still reading their smem source
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N> __device__ __forceinline__ void tma_store_wait() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ------------------------------------------------------------------ shared-memory accesses by 32-bit shared address
// 16-byte read-only global load that stays where it is written (asm volatile): the compiler neither hoists it into a region where the
// register file is already full nor sinks it next to its first use, so a batch of them is in flight together
__device__ __forceinline__ uint4 ldg_nc_u128_pinned(const void* ptr) {
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ptr) : "memory");
    return v;
}
__device__ __forceinline__ uint4 lds_u128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u128(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

__device__ __forceinline__ unsigned long long lds_u64(uint32_t addr) {
    unsigned long long v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u64(uint32_t addr, unsigned long long v) {
    asm volatile("st.shared.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}

// ------------------------------------------------------------------ legacy warp MMA (attention core)
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
template <typename T> struct MmaSync;
template <> struct MmaSync<__half> {
    __device__ static __forceinline__ void run(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    }
};
template <> struct MmaSync<__nv_bfloat16> {
    __device__ static __forceinline__ void run(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    }
};

__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ------------------------------------------------------------------ 16-bit helpers
template <typename T> struct Half2Of;
template <> struct Half2Of<__half> { using type = __half2; };
template <> struct Half2Of<__nv_bfloat16> { using type = __nv_bfloat162; };

template <typename T> __device__ __forceinline__ T from_float(float v);
template <> __device__ __forceinline__ float from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_float<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_float<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <typename T> __device__ __forceinline__ float to_float(T v);
template <> __device__ __forceinline__ float to_float<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_float<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_float<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ uint32_t pack2(float lo, float hi);
template <> __device__ __forceinline__ uint32_t pack2<__half>(float lo, float hi) {
    __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
template <typename T> __device__ __forceinline__ float2 unpack2(uint32_t v);
template <> __device__ __forceinline__ float2 unpack2<__half>(uint32_t v) {
    return __half22float2(*reinterpret_cast<__half2*>(&v));
}
template <> __device__ __forceinline__ float2 unpack2<__nv_bfloat16>(uint32_t v) {
    return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&v));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ float ex2_approx(float x) {      // 2^x, one MUFU.EX2 (ftz; -inf -> 0)
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// dropout keep decisions of the element pair (2 * pair, 2 * pair + 1): see DropSite in kernels.h
__device__ __forceinline__ uint32_t drop_hash(uint32_t pair, uint32_t key) {
    uint32_t x = pair ^ key;
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ void drop_pair(uint32_t pair, uint32_t key, uint32_t thr, float scale, float& m0, float& m1) {
    const uint32_t h = drop_hash(pair, key);
    m0 = (h & 0xffffu) >= thr ? scale : 0.0f;
    m1 = (h >> 16) >= thr ? scale : 0.0f;
}
__device__ __forceinline__ float drop_one(uint32_t e, uint32_t key, uint32_t thr, float scale) {
    const uint32_t h = drop_hash(e >> 1, key);
    return ((e & 1u) ? (h >> 16) : (h & 0xffffu)) >= thr ? scale : 0.0f;
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// erf-GELU for the 16-bit epilogues: gelu(x) = relu(x) - 0.5 |x| erfc(|x| / sqrt2), with erfc(a / sqrt2) = 2^q(a) on [0, 4 sqrt2]
// (degree-6 fit of log2 erfc in a = |x| directly; one MUFU.EX2, no branch, 12 instructions). Evaluated in fp32 against float64
// over [-9, 9]: max abs error 4.6e-6, max relative error 3.1e-5 where |gelu| > 1e-3 - an order of magnitude below the 16-bit
// rounding of the value it feeds (the fp32 parity path uses erff()).
__device__ __forceinline__ float gelu_erf_fast(float x) {
    const float ax = fabsf(x);
    const float u = fminf(ax, 5.656854249f);
    float p = 2.513894565e-05f;
    p = fmaf(p, u, -6.454259847e-04f);
    p = fmaf(p, u, 7.399560496e-03f);
    p = fmaf(p, u, -5.173896880e-02f);
    p = fmaf(p, u, -4.605998700e-01f);
    p = fmaf(p, u, -1.150469307e+00f);
    p = fmaf(p, u, -4.401278411e-05f);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(p));
    return fmaf(-0.5f * ax, e, fmaxf(x, 0.0f));
}

// d/dx [x Phi(x)] = Phi(x) + x phi(x) for the 16-bit training paths: Phi(x) from the same degree-6 fit of log2 erfc(|x| / sqrt2) as
// gelu_erf_fast (relative error 3e-5 of erfc), phi(x) = 2^(-x^2 log2(e) / 2) / sqrt(2 pi): two MUFU.EX2 and ~16 FMAs, no erff() / expf()
// call sequences - the first version of act_bwd_kernel was instruction-bound on those (781 us for 2.5 GB,
// profiles/r02b_launches_train_cfg2.csv). Used by the row kernels (train_rows.cu) and by the dgrad epilogue of gemm_umma2.cu (mode 9).
__device__ __forceinline__ float gelu_grad_fast(float x) {
    const float ax = fabsf(x);
    const float u = fminf(ax, 5.656854249f);
    float p = 2.513894565e-05f;
    p = fmaf(p, u, -6.454259847e-04f);
    p = fmaf(p, u, 7.399560496e-03f);
    p = fmaf(p, u, -5.173896880e-02f);
    p = fmaf(p, u, -4.605998700e-01f);
    p = fmaf(p, u, -1.150469307e+00f);
    p = fmaf(p, u, -4.401278411e-05f);
    float e, g;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(p));
    const float h = 0.5f * e;                                   // 0.5 erfc(|x| / sqrt2) = Phi(-|x|)
    const float cdf = x >= 0.0f ? 1.0f - h : h;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(g) : "f"(-0.72134752044448170368f * x * x));
    return fmaf(x, 0.39894228040143267794f * g, cdf);
}

// The same two functions on a pair of values with the packed fp32 instructions of sm_100 (FFMA2 / FMUL2 / FADD2: two lanes per issue
// slot) - for the GEMM epilogues, where eight warps per SM are the serial resource and instruction count is what they are bound by.
// Same polynomial, same MUFU.EX2 calls: bit-identical to the scalar versions.
__device__ __forceinline__ float2 gelu_poly2(float2 u) {
    float2 p = make_float2(2.513894565e-05f, 2.513894565e-05f);
    p = __ffma2_rn(p, u, make_float2(-6.454259847e-04f, -6.454259847e-04f));
    p = __ffma2_rn(p, u, make_float2(7.399560496e-03f, 7.399560496e-03f));
    p = __ffma2_rn(p, u, make_float2(-5.173896880e-02f, -5.173896880e-02f));
    p = __ffma2_rn(p, u, make_float2(-4.605998700e-01f, -4.605998700e-01f));
    p = __ffma2_rn(p, u, make_float2(-1.150469307e+00f, -1.150469307e+00f));
    p = __ffma2_rn(p, u, make_float2(-4.401278411e-05f, -4.401278411e-05f));
    return p;
}
__device__ __forceinline__ float2 ex2_approx2(float2 x) {
    float2 r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(x.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(x.y));
    return r;
}
__device__ __forceinline__ float2 gelu_erf_fast2(float2 x) {
    const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
    const float2 e = ex2_approx2(gelu_poly2(make_float2(fminf(ax.x, 5.656854249f), fminf(ax.y, 5.656854249f))));
    return __ffma2_rn(__fmul2_rn(ax, make_float2(-0.5f, -0.5f)), e, make_float2(fmaxf(x.x, 0.0f), fmaxf(x.y, 0.0f)));
}
__device__ __forceinline__ float2 gelu_grad_fast2(float2 x) {
    const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
    const float2 e = ex2_approx2(gelu_poly2(make_float2(fminf(ax.x, 5.656854249f), fminf(ax.y, 5.656854249f))));
    const float2 h = __fmul2_rn(e, make_float2(0.5f, 0.5f));
    const float2 hc = __ffma2_rn(h, make_float2(-1.0f, -1.0f), make_float2(1.0f, 1.0f));      // 1 - h
    const float2 cdf = make_float2(x.x >= 0.0f ? hc.x : h.x, x.y >= 0.0f ? hc.y : h.y);
    const float2 g = ex2_approx2(__fmul2_rn(__fmul2_rn(x, make_float2(-0.72134752044448170368f, -0.72134752044448170368f)), x));
    return __ffma2_rn(x, __fmul2_rn(g, make_float2(0.39894228040143267794f, 0.39894228040143267794f)), cdf);
}

}  // namespace tim
