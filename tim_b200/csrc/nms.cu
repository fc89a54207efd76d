// 1-D soft-NMS / NMS of detection proposals on the device. The reference's only native code is a scalar
// CPU extension (detection/eval_detection/csrc/nms_cpu.cpp: softnms_1d_cpu :67-160, nms_1d_cpu :19-58) called once per class of
// every video from a joblib pool (nms.py:123-155, format_predictions_epic.py:146-157). Here one call handles every
// (video, class) group: one CTA per group, the group's proposals in shared memory (global scratch for groups above NMS_SMEM_CAP).
//
// The reference algorithm is a selection sort whose result depends on the ORDER the live entries are stored in (ties take the
// first maximum in current array order, and a deleted entry is overwritten by the last live one), so that order is reproduced
// exactly: each round does (1) a block-wide arg-max of the live tail, first maximum in position order, whose reduction carries
// the pick's segment along so that every thread ends up holding it; (2) the swap to the front, folded into (3) the decay of every
// other live score, all in the reference's fp32 operation order with explicit round-to-nearest intrinsics; (4) the reference's "move the last live entry into the hole" deletions, done for all holes of the round at once:
// with n' live entries left, the k-th hole below n' (ascending) receives the k-th surviving entry at or above n' counted from the
// END — which is what the sequential scan produces (tests/test_oracle_golden.py checks that identity against a literal transcription).
// Picks and indices are identical to the reference; the gaussian weight is exp() taken in double and rounded once to fp32
// (glibc's expf, which the reference calls, is within one ulp of that).
#include <cstdio>
#include <string>

#include "../../include/tim_b200.h"
#include "kernels.h"

namespace tim {
namespace {

// Two launch shapes cover all groups; a CTA whose group belongs to the other shape exits at once (no host round trip to sort
// groups by size): groups of up to NMS_SMALL_MAX proposals run on 128 threads with 24 KB of shared memory (many CTAs per SM - the
// median (video, class) group of an EPIC evaluation holds about ten proposals), larger ones on 1024 threads with 192 KB (one CTA
// per SM; the largest groups are the critical path of the launch, and a round costs two block barriers plus n / 1024 updates per
// thread). Groups above NMS_BIG_CAP work from global scratch.
constexpr int NMS_SMALL_MAX = 1024;
constexpr int NMS_BIG_CAP = 8192;

struct NmsParams {
    const float* segs;              // [N, 2]
    const float* scores;            // [N]
    const long long* offs;          // [G + 1] group boundaries (ascending, offs[0] = 0, offs[G] = N)
    int G;
    float thr, sigma, min_score;
    int method;                     // 0 vanilla, 1 linear, 2 gaussian (softnms_1d_cpu); ignored when hard != 0
    int hard;                       // 1: nms_1d_cpu semantics (suppressed entries are deleted, scores unchanged, ties by input index)
    long long max_num;              // hard mode: cap on the picks per group (<= 0: none)
    float* dets;                    // [N, 3] rows (start, end, score) of group g at offs[g] .. offs[g] + kept[g]
    long long* inds;                // [N] index INSIDE the group of every pick, same placement
    int* kept;                      // [G]
    float* ws;                      // [6, N] scratch for groups larger than NMS_BIG_CAP
    long long N;
};

// arg-max candidate with its payload, so that after the reduction every thread holds the pick in registers
struct Cand {
    float s; int tie, pos;          // pos < 0: empty
    float x1, x2, ar; int id;
};
__device__ __forceinline__ bool better(float s, int tie, float bs, int btie) { return s > bs || (s == bs && tie < btie); }
__device__ __forceinline__ void take(Cand& a, const Cand& b) {
    if (b.pos >= 0 && (a.pos < 0 || better(b.s, b.tie, a.s, a.tie))) a = b;
}
__device__ __forceinline__ Cand warp_best(Cand c) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        Cand o;
        o.s = __shfl_xor_sync(0xffffffffu, c.s, d); o.tie = __shfl_xor_sync(0xffffffffu, c.tie, d);
        o.pos = __shfl_xor_sync(0xffffffffu, c.pos, d); o.x1 = __shfl_xor_sync(0xffffffffu, c.x1, d);
        o.x2 = __shfl_xor_sync(0xffffffffu, c.x2, d); o.ar = __shfl_xor_sync(0xffffffffu, c.ar, d);
        o.id = __shfl_xor_sync(0xffffffffu, c.id, d);
        take(c, o);
    }
    return c;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

template <int THREADS, int CAP, bool BIG>
__global__ void __launch_bounds__(THREADS) softnms_kernel(const NmsParams p) {
    constexpr int NW = THREADS / 32;
    extern __shared__ float nms_smem[];
    __shared__ Cand red[2][NW], red_scan[NW];
    __shared__ int red_cnt[2][NW], scan_h[NW], scan_m[NW];

    const int g = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long o = p.offs[g];
    int n = static_cast<int>(p.offs[g + 1] - o);
    if (n <= 0) { if (!BIG && tid == 0) p.kept[g] = 0; return; }
    if ((n > NMS_SMALL_MAX) != BIG) return;                  // the other launch shape owns this group

    float *x1, *x2, *sc, *ar;
    int *id, *lst;
    if (n <= CAP) {
        x1 = nms_smem; x2 = x1 + CAP; sc = x2 + CAP; ar = sc + CAP;
        id = reinterpret_cast<int*>(ar + CAP); lst = id + CAP;
    } else {
        x1 = p.ws + o; x2 = x1 + p.N; sc = x2 + p.N; ar = sc + p.N;
        id = reinterpret_cast<int*>(ar + p.N); lst = id + p.N;
    }
    const float ninf = __int_as_float(0xff800000);
    for (int q = tid; q < n; q += THREADS) {
        const float2 s = reinterpret_cast<const float2*>(p.segs)[o + q];
        float v = p.scores[o + q];
        if (p.hard && p.min_score > 0.0f && !(v > p.min_score)) v = ninf;      // nms.py:15-19: scores <= min_score never enter
        x1[q] = s.x; x2[q] = s.y; sc[q] = v;
        ar[q] = __fadd_rn(__fsub_rn(s.y, s.x), 1e-6f);                        // nms_cpu.cpp:77
        id[q] = q;
    }
    __syncthreads();

    const Cand none = {ninf, 0x7fffffff, -1, 0.0f, 0.0f, 0.0f, 0};
    // arg-max of the live tail [i, n) from scratch (round 0, and after a round that deleted entries): per-thread scan, warp
    // reduction, then every warp reduces the warp results, so the pick ends up in every thread's registers
    auto scan_pick = [&](int from) -> Cand {
        Cand c = none;
        for (int q = from + tid; q < n; q += THREADS) {
            const float s = sc[q];
            const int tie = p.hard ? id[q] : q;
            if (c.pos < 0 || better(s, tie, c.s, c.tie)) { c.s = s; c.tie = tie; c.pos = q; }
        }
        if (c.pos >= 0) { c.x1 = x1[c.pos]; c.x2 = x2[c.pos]; c.ar = ar[c.pos]; c.id = id[c.pos]; }
        c = warp_best(c);
        if (lane == 0) red_scan[warp] = c;
        __syncthreads();
        c = lane < NW ? red_scan[lane] : none;
        return warp_best(c);
    };

    int i = 0, par = 0;
    Cand c = scan_pick(0);                                  // (1) first maximum of the live tail, first in position order
    while (i < n) {
        const int m = c.pos;
        if (p.hard && (c.s == ninf || (p.max_num > 0 && i >= p.max_num))) break;   // nothing above min_score left / max_num reached
        if (tid == 0) {
            float* d = p.dets + (o + i) * 3;
            d[0] = c.x1; d[1] = c.x2; d[2] = c.s;
            p.inds[o + i] = c.id;
        }
        const float ix1 = c.x1, ix2 = c.x2, iarea = c.ar;

        // (2) + (3) swap the pick to position i (nms_cpu.cpp:105-122) and decay the rest (:126-144) in one pass: the thread that
        // owns position m moves the entry of position i there (nobody else touches i or m in this pass). The same pass tracks the
        // maximum of the decayed scores: when the round deletes nothing, positions do not move and that maximum IS the next
        // round's pick - one block barrier per round.
        int my_removed = 0;
        Cand nc = none;
        for (int q = i + 1 + tid; q < n; q += THREADS) {
            float qx1, qx2, qar, s;
            int qid = 0;
            if (q == m) {
                qx1 = x1[i]; qx2 = x2[i]; qar = ar[i]; s = sc[i]; qid = id[i];
                x1[i] = c.x1; x2[i] = c.x2; ar[i] = c.ar; sc[i] = c.s; id[i] = c.id;
                x1[q] = qx1; x2[q] = qx2; ar[q] = qar; id[q] = qid;
            } else {
                qx1 = x1[q]; qx2 = x2[q]; qar = ar[q]; s = sc[q];
                if (p.hard) qid = id[q];
            }
            const float xx1 = fmaxf(ix1, qx1), xx2 = fminf(ix2, qx2);
            const float inter = fmaxf(0.0f, __fsub_rn(xx2, xx1));
            const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(iarea, qar), inter));
            if (p.hard) {
                if (ovr >= p.thr) s = ninf;
                my_removed += (s == ninf);
            } else {
                float w = 1.0f;
                if (p.method == 0) { if (ovr >= p.thr) w = 0.0f; }
                else if (p.method == 1) { if (ovr >= p.thr) w = __fsub_rn(1.0f, ovr); }
                else { w = static_cast<float>(exp(static_cast<double>(__fdiv_rn(-__fmul_rn(ovr, ovr), p.sigma)))); }
                s = __fmul_rn(s, w);
                my_removed += (s < p.min_score);
            }
            sc[q] = s;
            const int tie = p.hard ? qid : q;
            if (nc.pos < 0 || better(s, tie, nc.s, nc.tie)) { nc.s = s; nc.tie = tie; nc.pos = q; nc.x1 = qx1; nc.x2 = qx2; nc.ar = qar; }
        }
        if (nc.pos >= 0) nc.id = id[nc.pos];                 // written, if at all, by this very thread (q == m)
        nc = warp_best(nc);
        my_removed = warp_sum(my_removed);
        if (lane == 0) { red[par][warp] = nc; red_cnt[par][warp] = my_removed; }
        __syncthreads();
        const int R = warp_sum(lane < NW ? red_cnt[par][lane] : 0);
        if (R == 0) {
            c = lane < NW ? red[par][lane] : none;
            c = warp_best(c);
            ++i; par ^= 1;                                  // the buffers of this parity are rewritten two rounds from now
            continue;
        }

        // (4) deletions of the round (nms_cpu.cpp:147-156): holes below n2 take the survivors at or above n2, last first
        const int len = n - (i + 1);
        const int n2 = n - R;
        const int chunk = (len + THREADS - 1) / THREADS;
        const int q0 = i + 1 + tid * chunk, q1 = min(q0 + chunk, n);
        int nh = 0, nm = 0;
        for (int q = q0; q < q1; ++q) {
            const bool rem = p.hard ? (sc[q] == ninf) : (sc[q] < p.min_score);
            nh += (rem && q < n2);
            nm += (!rem && q >= n2);
        }
        int ph = nh, pm = nm;                              // inclusive warp scans
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int a = __shfl_up_sync(0xffffffffu, ph, d), b = __shfl_up_sync(0xffffffffu, pm, d);
            if (lane >= d) { ph += a; pm += b; }
        }
        if (lane == 31) { scan_h[warp] = ph; scan_m[warp] = pm; }
        __syncthreads();
        int wh = lane < NW ? scan_h[lane] : 0, wm = lane < NW ? scan_m[lane] : 0;   // inclusive scan over the warp totals
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int a = __shfl_up_sync(0xffffffffu, wh, d), b = __shfl_up_sync(0xffffffffu, wm, d);
            if (lane >= d) { wh += a; wm += b; }
        }
        const int K = __shfl_sync(0xffffffffu, wh, NW - 1);
        const int bh = warp ? __shfl_sync(0xffffffffu, wh, warp - 1) : 0, bm = warp ? __shfl_sync(0xffffffffu, wm, warp - 1) : 0;
        int rh = bh + ph - nh, rm = bm + pm - nm;          // exclusive ranks of this thread's first hole / mover
        int* holes = lst;
        int* movers = lst + (n + 1) / 2;                   // K <= len / 2
        for (int q = q0; q < q1; ++q) {
            const bool rem = p.hard ? (sc[q] == ninf) : (sc[q] < p.min_score);
            if (rem && q < n2) holes[rh++] = q;
            if (!rem && q >= n2) movers[rm++] = q;
        }
        __syncthreads();
        for (int k = tid; k < K; k += THREADS) {
            const int dst = holes[k], src = movers[K - 1 - k];
            x1[dst] = x1[src]; x2[dst] = x2[src]; sc[dst] = sc[src]; ar[dst] = ar[src]; id[dst] = id[src];
        }
        n = n2;
        ++i; par ^= 1;
        __syncthreads();
        if (i < n) c = scan_pick(i);
    }
    if (tid == 0) p.kept[g] = i;
}

int fail(const char* msg) { set_global_error(msg); return TIM_ERR_INVALID; }

int launch(const NmsParams& p, cudaStream_t s, const char* who) {
    auto small = softnms_kernel<128, NMS_SMALL_MAX, false>;
    auto big = softnms_kernel<1024, NMS_BIG_CAP, true>;
    constexpr size_t small_bytes = NMS_SMALL_MAX * 6 * 4, big_bytes = NMS_BIG_CAP * 6 * 4;
    static SmemAttrCache cache_small, cache_big;
    cudaError_t e = ensure_dynamic_smem(small, small_bytes, cache_small);
    if (e == cudaSuccess) e = ensure_dynamic_smem(big, big_bytes, cache_big);
    if (e == cudaSuccess) {
        // two launches on the caller's stream; every CTA of the shape that does not own its group returns immediately
        big<<<static_cast<unsigned>(p.G), 1024, big_bytes, s>>>(p);
        small<<<static_cast<unsigned>(p.G), 128, small_bytes, s>>>(p);
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) { set_global_error((std::string(who) + ": " + cudaGetErrorString(e)).c_str()); return TIM_ERR_CUDA; }
    return TIM_OK;
}

}  // namespace
}  // namespace tim

extern "C" {

size_t tim_nms_workspace_bytes(int64_t N) { return N > 0 ? static_cast<size_t>(N) * 6 * 4 : 0; }

int tim_softnms_1d(const float* segs, const float* scores, const int64_t* group_offsets, int G, int64_t N, float iou_threshold,
                   float sigma, float min_score, int method, float* dets, int64_t* inds, int* kept, void* workspace,
                   size_t workspace_bytes, void* stream) {
    using namespace tim;
    if (G == 0) return TIM_OK;
    if (!segs || !scores || !group_offsets || !dets || !inds || !kept) return fail("tim_softnms_1d: NULL argument");
    if (G < 0 || N < 0) return fail("tim_softnms_1d: negative size");
    if (N > 0x7fffffffLL) return fail("tim_softnms_1d: more than 2^31 - 1 proposals (positions inside a group are 32-bit)");
    if (method < 0 || method > 2) return fail("tim_softnms_1d: method must be 0 (vanilla), 1 (linear) or 2 (gaussian)");
    if (method == 2 && !(sigma > 0.0f)) return fail("tim_softnms_1d: sigma must be positive for the gaussian method");
    if (!workspace || workspace_bytes < tim_nms_workspace_bytes(N)) return fail("tim_softnms_1d: workspace smaller than tim_nms_workspace_bytes(N)");
    NmsParams p{segs, scores, reinterpret_cast<const long long*>(group_offsets), G, iou_threshold, sigma, min_score, method, 0, 0,
                dets, reinterpret_cast<long long*>(inds), kept, static_cast<float*>(workspace), N};
    return launch(p, static_cast<cudaStream_t>(stream), "tim_softnms_1d");
}

int tim_nms_1d(const float* segs, const float* scores, const int64_t* group_offsets, int G, int64_t N, float iou_threshold,
               float min_score, int64_t max_num, float* dets, int64_t* inds, int* kept, void* workspace, size_t workspace_bytes,
               void* stream) {
    using namespace tim;
    if (G == 0) return TIM_OK;
    if (!segs || !scores || !group_offsets || !dets || !inds || !kept) return fail("tim_nms_1d: NULL argument");
    if (G < 0 || N < 0) return fail("tim_nms_1d: negative size");
    if (N > 0x7fffffffLL) return fail("tim_nms_1d: more than 2^31 - 1 proposals (positions inside a group are 32-bit)");
    if (!workspace || workspace_bytes < tim_nms_workspace_bytes(N)) return fail("tim_nms_1d: workspace smaller than tim_nms_workspace_bytes(N)");
    NmsParams p{segs, scores, reinterpret_cast<const long long*>(group_offsets), G, iou_threshold, 1.0f, min_score, 0, 1, max_num,
                dets, reinterpret_cast<long long*>(inds), kept, static_cast<float*>(workspace), N};
    return launch(p, static_cast<cudaStream_t>(stream), "tim_nms_1d");
}

}  // extern "C"
