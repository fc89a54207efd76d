// Detection query labelling on the device: the IoU / arg-max / label gathering the reference does by
// materialising [B, Nq, Na] tensors with repeat_interleave (detection/time_interval_machine/models/tim.py:186-270), and the
// smoothed one-hot targets of assign_positive_labels (:158-185), as two small HBM-bound kernels. Results are bit-identical to
// the reference: every fp32 operation is done in the reference's order with explicit round-to-nearest intrinsics (no FMA
// contraction), arg-max takes the first maximum and treats NaN as maximal like torch.argmax.
#include <cstdio>
#include <string>

#include "../../include/tim_b200.h"
#include "kernels.h"

namespace tim {
namespace {

// one thread per (clip, query)
__global__ void __launch_bounds__(256) label_queries_kernel(const float* __restrict__ queries, const float* __restrict__ gt,
                                                            const long long* __restrict__ gt_labels, int B, int Nq, int Na, int Nl,
                                                            float thr, float* __restrict__ targets, long long* __restrict__ ids,
                                                            float* __restrict__ ious) {
    const long long i = blockIdx.x * 256LL + threadIdx.x;
    if (i >= 1LL * B * Nq) return;
    const int b = static_cast<int>(i / Nq);
    const float2* g = reinterpret_cast<const float2*>(gt) + static_cast<size_t>(b) * Na;
    // negative_offsets = |min(min_a gt_start, 0)|  (tim.py:200)
    float mn = g[0].x;
    for (int a = 1; a < Na; ++a) mn = fminf(mn, g[a].x);
    const float off = fabsf(fminf(mn, 0.0f));
    const float2 q = reinterpret_cast<const float2*>(queries)[i];
    const float qs = __fadd_rn(q.x, off), qe = __fadd_rn(q.y, off);
    const float qlen = __fsub_rn(qe, qs);
    float best = 0.0f, best_s = 0.0f, best_e = 0.0f;
    int best_a = -1;
    for (int a = 0; a < Na; ++a) {
        const float gs = __fadd_rn(g[a].x, off), ge = __fadd_rn(g[a].y, off);
        const float inter = fmaxf(__fsub_rn(fminf(qe, ge), fmaxf(qs, gs)), 0.0f);
        const float uni = __fsub_rn(__fadd_rn(__fsub_rn(ge, gs), qlen), inter);
        const float iou = __fdiv_rn(inter, uni);
        // first maximum; NaN beats every number (torch.argmax), a later NaN does not beat an earlier one
        const bool better = best_a < 0 || iou > best || (iou != iou && best == best);
        if (better) { best = iou; best_a = a; best_s = gs; best_e = ge; }
    }
    const bool neg = best < thr;                       // NaN: not negative, as `ious < iou_threshold` is False
    const float inf = __int_as_float(0x7f800000);
    reinterpret_cast<float2*>(targets)[i] = neg ? make_float2(inf, inf) : make_float2(best_s, best_e);
    ious[i] = best;
    const long long* lab = gt_labels + (static_cast<size_t>(b) * Na + best_a) * Nl;
    for (int k = 0; k < Nl; ++k) ids[i * Nl + k] = neg ? -1LL : lab[k];
}

// out[row, c] = (c == id ? s : 0) + (1 - s) / (C + 1); id == -1 selects the dropped column C   (tim.py:172-182)
// A CTA owns SL_ROWS consecutive rows = one contiguous span of the output, written with coalesced stores; 32-bit index math.
constexpr int SL_ROWS = 64;
__global__ void __launch_bounds__(256) smooth_labels_kernel(const long long* __restrict__ ids, int stride, int col, long long rows, int C,
                                                            float s, float base, float* __restrict__ out) {
    __shared__ int sid[SL_ROWS];
    const long long r0 = static_cast<long long>(blockIdx.x) * SL_ROWS;
    const int nr = static_cast<int>(min(static_cast<long long>(SL_ROWS), rows - r0));
    for (int r = threadIdx.x; r < nr; r += 256) sid[r] = static_cast<int>(ids[(r0 + r) * stride + col]);
    __syncthreads();
    float* o = out + r0 * C;
    const unsigned n = static_cast<unsigned>(nr) * static_cast<unsigned>(C);
    for (unsigned j = threadIdx.x; j < n; j += 256) {
        const unsigned r = j / static_cast<unsigned>(C);
        const int c = static_cast<int>(j - r * static_cast<unsigned>(C));
        o[j] = __fadd_rn(sid[r] == c ? s : 0.0f, base);
    }
}

int fail(const char* msg) { set_global_error(msg); return TIM_ERR_INVALID; }

}  // namespace
}  // namespace tim

extern "C" {

int tim_label_queries(const float* queries, const float* gt_segs, const int64_t* gt_labels, int B, int Nq, int Na, int Nl,
                      float iou_threshold, float* targets, int64_t* label_ids, float* ious, void* stream) {
    using namespace tim;
    if (!queries || !gt_segs || !gt_labels || !targets || !label_ids || !ious) return fail("tim_label_queries: NULL argument");
    if (B <= 0 || Nq <= 0 || Na <= 0 || Nl <= 0) return fail("tim_label_queries: B, Nq, Na, Nl must be positive");
    const long long n = 1LL * B * Nq;
    label_queries_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        queries, gt_segs, reinterpret_cast<const long long*>(gt_labels), B, Nq, Na, Nl, iou_threshold, targets,
        reinterpret_cast<long long*>(label_ids), ious);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_global_error((std::string("tim_label_queries: ") + cudaGetErrorString(e)).c_str()); return TIM_ERR_CUDA; }
    return TIM_OK;
}

int tim_smooth_labels(const int64_t* label_ids, int stride, int col, int64_t rows, int num_classes, double smoothing, float* out,
                      void* stream) {
    using namespace tim;
    if (!label_ids || !out) return fail("tim_smooth_labels: NULL argument");
    if (rows <= 0 || num_classes <= 0 || stride <= 0 || col < 0 || col >= stride) return fail("tim_smooth_labels: bad shape");
    // (1 - s) / (C + 1) in double, rounded once: what Python float arithmetic followed by the fp32 tensor add does
    // `smoothing` arrives as the double the reference holds (self.label_smoothing): the product uses its fp32 rounding, the
    // additive term is formed in double and rounded once
    const float base = static_cast<float>((1.0 - smoothing) / (num_classes + 1));
    if (num_classes > (1 << 20)) return fail("tim_smooth_labels: num_classes too large");
    const long long blocks = (rows + SL_ROWS - 1) / SL_ROWS;
    if (blocks > 0x7fffffffLL) return fail("tim_smooth_labels: too many rows");
    smooth_labels_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const long long*>(label_ids), stride, col, rows, num_classes, static_cast<float>(smoothing), base, out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_global_error((std::string("tim_smooth_labels: ") + cudaGetErrorString(e)).c_str()); return TIM_ERR_CUDA; }
    return TIM_OK;
}

}  // extern "C"
