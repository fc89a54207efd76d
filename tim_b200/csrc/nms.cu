// 1-D soft-NMS / NMS of detection proposals on the device. The reference's only native code is a scalar
// CPU extension (detection/eval_detection/csrc/nms_cpu.cpp: softnms_1d_cpu :67-160, nms_1d_cpu :19-58) called once per class of
// every video from a joblib pool (nms.py:123-155, format_predictions_epic.py:146-157). Here ONE launch handles every
// (video, class) group: one CTA per group, the group's proposals in shared memory (global scratch for groups above NMS_SMEM_CAP).
//
// The reference algorithm is a selection sort whose result depends on the ORDER the live entries are stored in (ties take the
// first maximum in current array order, and a deleted entry is overwritten by the last live one), so that order is reproduced
// exactly: each round does (1) a block-wide arg-max of the live tail, first maximum in position order; (2) the swap to the
// front; (3) the decay of every other live score, all in the reference's fp32 operation order with explicit round-to-nearest
// intrinsics; (4) the reference's "move the last live entry into the hole" deletions, done for all holes of the round at once:
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

constexpr int NMS_THREADS = 256;
constexpr int NMS_WARPS = NMS_THREADS / 32;
constexpr int NMS_SMEM_CAP = 4096;                       // proposals of one group held in shared memory
constexpr int NMS_SMEM_BYTES = NMS_SMEM_CAP * 6 * 4;     // x1, x2, score, area, index, hole / mover lists

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
    float* ws;                      // [6, N] scratch for groups larger than NMS_SMEM_CAP
    long long N;
};

__device__ __forceinline__ bool better(float s, int tie, float bs, int btie) { return s > bs || (s == bs && tie < btie); }

__global__ void __launch_bounds__(NMS_THREADS) softnms_kernel(const NmsParams p) {
    extern __shared__ float nms_smem[];
    __shared__ float red_s[NMS_WARPS];
    __shared__ int red_tie[NMS_WARPS], red_pos[NMS_WARPS];
    __shared__ int scan_h[NMS_WARPS], scan_m[NMS_WARPS];
    __shared__ float pick[3];
    __shared__ int pick_pos, removed_cnt;

    const int g = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long o = p.offs[g];
    int n = static_cast<int>(p.offs[g + 1] - o);
    if (n <= 0) { if (tid == 0) p.kept[g] = 0; return; }

    float *x1, *x2, *sc, *ar;
    int *id, *lst;
    if (n <= NMS_SMEM_CAP) {
        x1 = nms_smem; x2 = x1 + NMS_SMEM_CAP; sc = x2 + NMS_SMEM_CAP; ar = sc + NMS_SMEM_CAP;
        id = reinterpret_cast<int*>(ar + NMS_SMEM_CAP); lst = id + NMS_SMEM_CAP;
    } else {
        x1 = p.ws + o; x2 = x1 + p.N; sc = x2 + p.N; ar = sc + p.N;
        id = reinterpret_cast<int*>(ar + p.N); lst = id + p.N;
    }
    const float ninf = __int_as_float(0xff800000);
    for (int q = tid; q < n; q += NMS_THREADS) {
        const float2 s = reinterpret_cast<const float2*>(p.segs)[o + q];
        float v = p.scores[o + q];
        if (p.hard && p.min_score > 0.0f && !(v > p.min_score)) v = ninf;      // nms.py:15-19: scores <= min_score never enter
        x1[q] = s.x; x2[q] = s.y; sc[q] = v;
        ar[q] = __fadd_rn(__fsub_rn(s.y, s.x), 1e-6f);                        // nms_cpu.cpp:77
        id[q] = q;
    }
    __syncthreads();

    int i = 0;
    while (i < n) {
        // (1) first maximum of the live tail [i, n)
        float bs = ninf; int btie = 0x7fffffff, bpos = -1;
        for (int q = i + tid; q < n; q += NMS_THREADS) {
            const float s = sc[q];
            const int tie = p.hard ? id[q] : q;
            if (bpos < 0 || better(s, tie, bs, btie)) { bs = s; btie = tie; bpos = q; }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const float os = __shfl_xor_sync(0xffffffffu, bs, d);
            const int ot = __shfl_xor_sync(0xffffffffu, btie, d), op = __shfl_xor_sync(0xffffffffu, bpos, d);
            if (op >= 0 && (bpos < 0 || better(os, ot, bs, btie))) { bs = os; btie = ot; bpos = op; }
        }
        if (lane == 0) { red_s[warp] = bs; red_tie[warp] = btie; red_pos[warp] = bpos; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < NMS_WARPS; ++w)
                if (red_pos[w] >= 0 && (bpos < 0 || better(red_s[w], red_tie[w], bs, btie))) { bs = red_s[w]; btie = red_tie[w]; bpos = red_pos[w]; }
            // (2) swap the pick to position i and emit it (nms_cpu.cpp:105-122)
            const int m = bpos;
            const float mx1 = x1[m], mx2 = x2[m], msc = sc[m], mar = ar[m];
            const int mid = id[m];
            x1[m] = x1[i]; x2[m] = x2[i]; sc[m] = sc[i]; ar[m] = ar[i]; id[m] = id[i];
            x1[i] = mx1; x2[i] = mx2; sc[i] = msc; ar[i] = mar; id[i] = mid;
            pick[0] = mx1; pick[1] = mx2; pick[2] = mar;
            const bool stop = p.hard && (msc == ninf || (p.max_num > 0 && i >= p.max_num));
            pick_pos = stop ? -1 : m;
            removed_cnt = 0;
            if (!stop) {
                float* d = p.dets + (o + i) * 3;
                d[0] = mx1; d[1] = mx2; d[2] = msc;
                p.inds[o + i] = mid;
            }
        }
        __syncthreads();
        if (pick_pos < 0) break;                          // hard mode: nothing above min_score left / max_num reached
        const float ix1 = pick[0], ix2 = pick[1], iarea = pick[2];

        // (3) decay the rest (nms_cpu.cpp:126-144)
        int my_removed = 0;
        for (int q = i + 1 + tid; q < n; q += NMS_THREADS) {
            const float xx1 = fmaxf(ix1, x1[q]), xx2 = fminf(ix2, x2[q]);
            const float inter = fmaxf(0.0f, __fsub_rn(xx2, xx1));
            const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(iarea, ar[q]), inter));
            float s = sc[q];
            if (p.hard) {
                if (ovr >= p.thr) s = ninf;
                my_removed += (s == ninf);
            } else {
                float w = 1.0f;
                if (p.method == 0) { if (ovr >= p.thr) w = 0.0f; }
                else if (p.method == 1) { if (ovr >= p.thr) w = __fsub_rn(1.0f, ovr); }
                else if (p.method == 2) { w = static_cast<float>(exp(static_cast<double>(__fdiv_rn(-__fmul_rn(ovr, ovr), p.sigma)))); }
                s = __fmul_rn(s, w);
                my_removed += (s < p.min_score);
            }
            sc[q] = s;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) my_removed += __shfl_xor_sync(0xffffffffu, my_removed, d);
        if (lane == 0 && my_removed) atomicAdd(&removed_cnt, my_removed);
        __syncthreads();
        const int R = removed_cnt;
        if (R == 0) { ++i; continue; }

        // (4) deletions of the round (nms_cpu.cpp:147-156): holes below n2 take the survivors at or above n2, last first
        const int len = n - (i + 1);
        const int n2 = n - R;
        const int chunk = (len + NMS_THREADS - 1) / NMS_THREADS;
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
        int bh = 0, bm = 0, K = 0;
        for (int w = 0; w < NMS_WARPS; ++w) {
            if (w < warp) { bh += scan_h[w]; bm += scan_m[w]; }
            K += scan_h[w];
        }
        int rh = bh + ph - nh, rm = bm + pm - nm;          // exclusive ranks of this thread's first hole / mover
        int* holes = lst;
        int* movers = lst + (n + 1) / 2;                   // K <= len / 2
        for (int q = q0; q < q1; ++q) {
            const bool rem = p.hard ? (sc[q] == ninf) : (sc[q] < p.min_score);
            if (rem && q < n2) holes[rh++] = q;
            if (!rem && q >= n2) movers[rm++] = q;
        }
        __syncthreads();
        for (int k = tid; k < K; k += NMS_THREADS) {
            const int dst = holes[k], src = movers[K - 1 - k];
            x1[dst] = x1[src]; x2[dst] = x2[src]; sc[dst] = sc[src]; ar[dst] = ar[src]; id[dst] = id[src];
        }
        n = n2;
        ++i;
        __syncthreads();
    }
    if (tid == 0) p.kept[g] = i;
}

int fail(const char* msg) { set_global_error(msg); return TIM_ERR_INVALID; }

int launch(const NmsParams& p, cudaStream_t s, const char* who) {
    static SmemAttrCache cache;
    cudaError_t e = ensure_dynamic_smem(softnms_kernel, NMS_SMEM_BYTES, cache);
    if (e == cudaSuccess) {
        softnms_kernel<<<static_cast<unsigned>(p.G), NMS_THREADS, NMS_SMEM_BYTES, s>>>(p);
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
    if (!workspace || workspace_bytes < tim_nms_workspace_bytes(N)) return fail("tim_nms_1d: workspace smaller than tim_nms_workspace_bytes(N)");
    NmsParams p{segs, scores, reinterpret_cast<const long long*>(group_offsets), G, iou_threshold, 1.0f, min_score, 0, 1, max_num,
                dets, reinterpret_cast<long long*>(inds), kept, static_cast<float*>(workspace), N};
    return launch(p, static_cast<cudaStream_t>(stream), "tim_nms_1d");
}

}  // extern "C"
