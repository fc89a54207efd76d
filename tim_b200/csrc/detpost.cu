// Detection post-processing in front of the NMS, on the device:
//   tim_det_decode            <- FeatureMeter.update (detection/time_interval_machine/utils/meters.py:652-724): sigmoid of the class
//                                logits, regression outputs clamped to [0, max_time] and mapped back to seconds of the video
//   tim_det_count / _emit     <- the thresholding loop of detection/eval_detection/format_predictions.py:103-125 (and
//                                format_predictions_epic.py:120-143): proposals rounded to 3 decimals, empty ones dropped, one
//                                detection per (proposal, class) whose score exceeds the threshold, in (proposal, class) order
// The reference does the first on the host after a .cpu() of every batch and the second in a Python loop over all proposals of
// the dataset. Arithmetic follows the reference's dtypes: the de-normalisation multiplies in fp32 and adds the window start in
// double (the loader's metadata collates to float64, so `v_proposals * win_size + win_starts[:, None]` promotes to float64), the
// rounding to 3 decimals is numpy's multiply / rint / divide in double, the segment handed to the NMS is that value cast to fp32.
// HBM-bound elementwise / compaction kernels, one warp per proposal row, coalesced over classes.
#include <string>

#include "../../include/tim_b200.h"
#include "kernels.h"

namespace tim {
namespace {

constexpr int DP_THREADS = 256;
constexpr int DP_ROWS = DP_THREADS / 32;       // proposal rows per CTA (one warp each)

__device__ __forceinline__ float sigmoid_f32(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }

__global__ void __launch_bounds__(DP_THREADS) det_decode_kernel(const float* __restrict__ logits, const float* __restrict__ reg,
                                                                const double* __restrict__ win_start, int rows_per_window,
                                                                long long R, int C, float win_size, float max_time,
                                                                float* __restrict__ preds, double* __restrict__ props) {
    const long long row = static_cast<long long>(blockIdx.x) * DP_ROWS + (threadIdx.x >> 5);
    if (row >= R) return;
    const int lane = threadIdx.x & 31;
    if (preds) {
        const float* in = logits + row * C;
        float* out = preds + row * C;
        for (int c = lane; c < C; c += 32) out[c] = sigmoid_f32(in[c]);
    }
    if (props && lane < 2) {
        // torch.clamp(min=0, max=max_time) is min(max(x, 0), max_time) (NaN propagates); then fp32 multiply, float64 add
        float v = reg[row * 2 + lane];
        v = v != v ? v : fminf(fmaxf(v, 0.0f), max_time);
        props[row * 2 + lane] = static_cast<double>(__fmul_rn(v, win_size)) + win_start[row / rows_per_window];
    }
}

// np.round(x, 3) on float64: multiply, rint (half to even), divide
__device__ __forceinline__ double round3(double x) { return __ddiv_rn(rint(__dmul_rn(x, 1000.0)), 1000.0); }

template <bool EMIT>
__global__ void __launch_bounds__(DP_THREADS) det_threshold_kernel(const float* __restrict__ preds, const double* __restrict__ props,
                                                                   long long R, int C, float thr, int* __restrict__ counts,
                                                                   const long long* __restrict__ offsets, long long* __restrict__ out_row,
                                                                   long long* __restrict__ out_cls, float* __restrict__ out_score,
                                                                   float* __restrict__ out_seg) {
    const long long row = static_cast<long long>(blockIdx.x) * DP_ROWS + (threadIdx.x >> 5);
    if (row >= R) return;
    const int lane = threadIdx.x & 31;
    const double p0 = round3(props[row * 2]), p1 = round3(props[row * 2 + 1]);
    const bool live = (p1 - p0) > 0.0;                       // format_predictions.py:108 (False for NaN)
    int n = 0;
    if (live) {
        const float* in = preds + row * C;
        const long long base = EMIT ? offsets[row] : 0;
        const float s0 = static_cast<float>(p0), s1 = static_cast<float>(p1);      // torch.FloatTensor(segs)
        for (int c0 = 0; c0 < C; c0 += 32) {
            const int c = c0 + lane;
            const float s = c < C ? in[c] : 0.0f;
            const bool hit = c < C && s > thr;
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (EMIT && hit) {
                const long long at = base + n + __popc(m & ((1u << lane) - 1u));
                out_row[at] = row; out_cls[at] = c; out_score[at] = s;
                out_seg[at * 2] = s0; out_seg[at * 2 + 1] = s1;
            }
            n += __popc(m);
        }
    }
    if (!EMIT && lane == 0) counts[row] = n;
}

int fail(const char* msg) { set_global_error(msg); return TIM_ERR_INVALID; }
int cuda_status(const char* who) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_global_error((std::string(who) + ": " + cudaGetErrorString(e)).c_str()); return TIM_ERR_CUDA; }
    return TIM_OK;
}
bool grid_ok(int64_t R) { return (R + DP_ROWS - 1) / DP_ROWS <= 0x7fffffffLL; }

}  // namespace
}  // namespace tim

extern "C" {

int tim_det_decode(const float* logits, const float* reg, const double* win_start, int rows_per_window, int64_t R, int C, float win_size,
                   float max_time, float* preds, double* proposals, void* stream) {
    using namespace tim;
    if (R == 0) return TIM_OK;
    if (R < 0 || C < 0 || rows_per_window <= 0 || !grid_ok(R)) return fail("tim_det_decode: bad shape");
    if ((preds && (!logits || C <= 0)) || (proposals && (!reg || !win_start)) || (!preds && !proposals))
        return fail("tim_det_decode: NULL argument");
    det_decode_kernel<<<static_cast<unsigned>((R + DP_ROWS - 1) / DP_ROWS), DP_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
        logits, reg, win_start, rows_per_window, R, C, win_size, max_time, preds, proposals);
    return cuda_status("tim_det_decode");
}

int tim_det_count(const float* preds, const double* proposals, int64_t R, int C, float score_threshold, int* counts, void* stream) {
    using namespace tim;
    if (R == 0) return TIM_OK;
    if (!preds || !proposals || !counts) return fail("tim_det_count: NULL argument");
    if (R < 0 || C <= 0 || !grid_ok(R)) return fail("tim_det_count: bad shape");
    det_threshold_kernel<false><<<static_cast<unsigned>((R + DP_ROWS - 1) / DP_ROWS), DP_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
        preds, proposals, R, C, score_threshold, counts, nullptr, nullptr, nullptr, nullptr, nullptr);
    return cuda_status("tim_det_count");
}

int tim_det_emit(const float* preds, const double* proposals, int64_t R, int C, float score_threshold, const int64_t* offsets,
                 int64_t* out_row, int64_t* out_cls, float* out_score, float* out_seg, void* stream) {
    using namespace tim;
    if (R == 0) return TIM_OK;
    if (!preds || !proposals || !offsets || !out_row || !out_cls || !out_score || !out_seg) return fail("tim_det_emit: NULL argument");
    if (R < 0 || C <= 0 || !grid_ok(R)) return fail("tim_det_emit: bad shape");
    det_threshold_kernel<true><<<static_cast<unsigned>((R + DP_ROWS - 1) / DP_ROWS), DP_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
        preds, proposals, R, C, score_threshold, nullptr, reinterpret_cast<const long long*>(offsets), reinterpret_cast<long long*>(out_row),
        reinterpret_cast<long long*>(out_cls), out_score, out_seg);
    return cuda_status("tim_det_emit");
}

}  // extern "C"
