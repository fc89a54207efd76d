// fp32 CUDA-core GEMM with the same row mapping and epilogue as the tcgen05 kernel. Used only in
// compute_dtype = fp32 (the <= 1e-5 parity mode of BASELINE.json config 1); not a performance path.
//   C[rows, N] = A[rows, K] * W[N, K]^T   (+bias, ReLU / erf-GELU, +residual)
#include "kernels.h"
#include "ptx.cuh"

namespace tim {
namespace {

constexpr int TM = 64, TN = 64, TK = 16;

__global__ void __launch_bounds__(256) linear_simt_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W,
                                                          int N, int K, RowMap rm, Epilogue ep, int tiles_r) {
    __shared__ float As[TK][TM + 4];
    __shared__ float Ws[TK][TN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    // rows of this tile: logical rows are enumerated group-major, box_r x box_g per tile (box <= 128 -> two 64-row halves)
    const int tiles_per_box = (rm.box_r * rm.box_g + TM - 1) / TM;
    const int tm = blockIdx.y / tiles_per_box, half = blockIdx.y % tiles_per_box;
    const int g0 = (tm / tiles_r) * rm.box_g;
    const int r0 = (tm % tiles_r) * rm.box_r;
    const int n0 = blockIdx.x * TN;

    auto row_of = [&](int i, bool& ok, size_t& arow, size_t& orow) {
        const int li = half * TM + i;
        const int gi = li / rm.box_r, ri = li - gi * rm.box_r;
        ok = (gi < rm.box_g) && (g0 + gi < rm.G) && (r0 + ri < rm.R);
        arow = static_cast<size_t>(g0 + gi) * rm.a_group_rows + rm.a_row_off + r0 + ri;
        orow = static_cast<size_t>(g0 + gi) * rm.out_group_rows + rm.out_row_off + r0 + ri;
    };

    // loader mapping: thread loads 4 consecutive k of one row
    const int lrow = tid >> 2, lk = (tid & 3) * 4;
    bool a_ok; size_t a_row, dummy;
    row_of(lrow, a_ok, a_row, dummy);
    const bool w_ok = (n0 + lrow) < N;
    const float* wp = W + static_cast<size_t>(n0 + lrow) * K;
    const float* ap = A + a_row * lda;

    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += TK) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + lk + j;
            As[lk + j][lrow] = (a_ok && k < K) ? ap[k] : 0.0f;
            Ws[lk + j][lrow] = (w_ok && k < K) ? wp[k] : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            float a[4], w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = As[k][ty * 4 + i]; w[i] = Ws[k][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        bool ok; size_t arow, orow;
        row_of(ty * 4 + i, ok, arow, orow);
        if (!ok) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j];
            if (ep.bias) v += ep.bias[n];
            if (ep.act == ACT_RELU) v = fmaxf(v, 0.0f);
            else if (ep.act == ACT_GELU) v = gelu_erf(v);
            if (ep.resid) v += ep.resid[orow * ep.ldr + n];
            reinterpret_cast<float*>(ep.out)[orow * ep.ldo + n] = v;
        }
    }
}

}  // namespace

cudaError_t launch_linear_simt(const float* A, int lda, const float* W, int N, int K, RowMap rm, Epilogue ep, cudaStream_t s) {
    if (!ep.out_fp32) return cudaErrorInvalidValue;
    const int tiles_r = (rm.R + rm.box_r - 1) / rm.box_r;
    const int tiles_g = (rm.G + rm.box_g - 1) / rm.box_g;
    const int tiles_per_box = (rm.box_r * rm.box_g + TM - 1) / TM;
    const long long tiles_m = 1LL * tiles_r * tiles_g * tiles_per_box;
    if (tiles_m <= 0 || N <= 0) return cudaSuccess;
    if (tiles_m > 65535) return cudaErrorInvalidValue;     // fp32 parity mode is for small problems
    dim3 grid((N + TN - 1) / TN, static_cast<unsigned>(tiles_m));
    linear_simt_kernel<<<grid, 256, 0, s>>>(A, lda, W, N, K, rm, ep, tiles_r);
    return cudaGetLastError();
}

}  // namespace tim
