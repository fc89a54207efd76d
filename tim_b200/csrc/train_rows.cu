// HBM-bound row / elementwise kernels of the TIM training leg (backward of LayerNorm, GELU / ReLU, token assembly, the tiny
// first time-MLP layer and last regression layer, bias / CLS / modality-encoding gradients as column sums, operand packing).
// One warp per row with 128-bit accesses where rows are wide; column sums accumulate per CTA in shared memory and reach the
// gradient buffer with one atomicAdd per column and CTA. Nothing here is GEMM-shaped.
//
// What they replace: the autograd nodes torch records for  nn.LayerNorm / nn.GELU / nn.ReLU / torch.cat / expand / bias adds  of
//   recognition/.../models/tim.py:66-74, helpers/encodings.py:140-251, helpers/transformers.py:92-111, helpers/head.py
// when recognition/scripts/train.py:354-366 (detection/scripts/train.py:372-384) runs backward().
#include <type_traits>

#include "kernels.h"
#include "ptx.cuh"

namespace tim {
namespace {

constexpr int RW = 8;                      // warps (rows in flight) per CTA
constexpr float kInvSqrt2 = 0.70710678118654752440f;
constexpr float kInvSqrt2Pi = 0.39894228040143267794f;

__device__ __forceinline__ float gelu_grad(float x) {          // d/dx [x Phi(x)] = Phi(x) + x phi(x)   (fp32 parity path)
    const float cdf = 0.5f * (1.0f + erff(x * kInvSqrt2));
    return fmaf(x * kInvSqrt2Pi, __expf(-0.5f * x * x), cdf);
}
template <typename T> __device__ __forceinline__ void load4(const T* p, float (&v)[4]);
template <> __device__ __forceinline__ void load4<float>(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <> __device__ __forceinline__ void load4<__half>(const __half* p, float (&v)[4]) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const float2 a = unpack2<__half>(u.x), b = unpack2<__half>(u.y);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
template <> __device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[4]) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const float2 a = unpack2<__nv_bfloat16>(u.x), b = unpack2<__nv_bfloat16>(u.y);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
template <typename T> __device__ __forceinline__ void store4t(T* p, const float (&v)[4]);
template <> __device__ __forceinline__ void store4t<float>(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void store4t<__half>(__half* p, const float (&v)[4]) {
    *reinterpret_cast<uint2*>(p) = make_uint2(pack2<__half>(v[0], v[1]), pack2<__half>(v[2], v[3]));
}
template <> __device__ __forceinline__ void store4t<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[4]) {
    *reinterpret_cast<uint2*>(p) = make_uint2(pack2<__nv_bfloat16>(v[0], v[1]), pack2<__nv_bfloat16>(v[2], v[3]));
}

// ---------------------------------------------------------------------------------------------------------------------
// LayerNorm backward over rows: y = (z - mean) * rstd * gamma + beta.   dy (fp32, [M, n]) is overwritten by dz;
//   dz = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma,  xhat = (z - mean) * rstd
// also: dz16 (operand copy for the GEMMs that follow), dgamma += sum dy * xhat, dbeta += sum dy, dbias += sum dz (the bias of the
// linear layer whose output was added into z). Statistics are recomputed from z (two-pass; z and dy are read from memory once and
// stay in registers). Rows of at most 128 * V floats; grid-stride over rows, the three column sums accumulate in registers per
// lane, are combined across the CTA's warps in shared memory and reach the gradient buffer with one atomicAdd per column and CTA.
// ---------------------------------------------------------------------------------------------------------------------
template <typename T, int V>
__global__ void __launch_bounds__(RW * 32) ln_bwd_kernel(float* __restrict__ dy, int ldd, const float* __restrict__ z, int ldz,
                                                         const float* __restrict__ gamma, T* __restrict__ dz16, int ld16,
                                                         float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                         float* __restrict__ dbias, int M, int n, DropSite drop) {
    extern __shared__ float sred[];                      // [RW][n]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float inv_n = 1.0f / static_cast<float>(n);
    float4 ag[V], ab[V], az[V];
#pragma unroll
    for (int i = 0; i < V; ++i) ag[i] = ab[i] = az[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int row = blockIdx.x * RW + warp; row < M; row += gridDim.x * RW) {
        const float* zr = z + static_cast<size_t>(row) * ldz;
        float* dr = dy + static_cast<size_t>(row) * ldd;
        float4 zz[V], dd[V];
        float s = 0.0f;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const int c = lane * 4 + i * 128;
            const bool ok = c < n;
            zz[i] = ok ? *reinterpret_cast<const float4*>(zr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            dd[i] = ok ? *reinterpret_cast<const float4*>(dr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            s += (zz[i].x + zz[i].y) + (zz[i].z + zz[i].w);
        }
        const float mean = warp_sum(s) * inv_n;
        float q = 0.0f;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            if (lane * 4 + i * 128 < n) {
                const float a = zz[i].x - mean, b = zz[i].y - mean, c2 = zz[i].z - mean, d2 = zz[i].w - mean;
                q += (a * a + b * b) + (c2 * c2 + d2 * d2);
            }
        }
        const float rstd = rsqrtf(warp_sum(q) * inv_n + 1e-5f);
        float sg = 0.0f, sgx = 0.0f;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            if (lane * 4 + i * 128 < n) {
                // zz becomes xhat
                zz[i].x = (zz[i].x - mean) * rstd; zz[i].y = (zz[i].y - mean) * rstd;
                zz[i].z = (zz[i].z - mean) * rstd; zz[i].w = (zz[i].w - mean) * rstd;
                ag[i].x = fmaf(dd[i].x, zz[i].x, ag[i].x); ag[i].y = fmaf(dd[i].y, zz[i].y, ag[i].y);
                ag[i].z = fmaf(dd[i].z, zz[i].z, ag[i].z); ag[i].w = fmaf(dd[i].w, zz[i].w, ag[i].w);
                ab[i].x += dd[i].x; ab[i].y += dd[i].y; ab[i].z += dd[i].z; ab[i].w += dd[i].w;
                // dd becomes g = dy * gamma
                const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + lane * 4 + i * 128));
                dd[i].x *= gm.x; dd[i].y *= gm.y; dd[i].z *= gm.z; dd[i].w *= gm.w;
                sg += (dd[i].x + dd[i].y) + (dd[i].z + dd[i].w);
                sgx = fmaf(dd[i].x, zz[i].x, fmaf(dd[i].y, zz[i].y, fmaf(dd[i].z, zz[i].z, fmaf(dd[i].w, zz[i].w, sgx))));
            }
        }
        const float mg = warp_sum(sg) * inv_n, mgx = warp_sum(sgx) * inv_n;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const int c = lane * 4 + i * 128;
            if (c < n) {
                float o[4];
                o[0] = rstd * (dd[i].x - mg - zz[i].x * mgx); o[1] = rstd * (dd[i].y - mg - zz[i].y * mgx);
                o[2] = rstd * (dd[i].z - mg - zz[i].z * mgx); o[3] = rstd * (dd[i].w - mg - zz[i].w * mgx);
                store4t<float>(dr + c, o);               // the residual branch keeps the un-masked gradient
                if (drop.thr) {                          // gradient w.r.t. the dropped sub-layer output: dz o mask
                    const uint32_t e = static_cast<uint32_t>(row) * static_cast<uint32_t>(n) + static_cast<uint32_t>(c);
                    float m0, m1, m2, m3;
                    drop_pair(e >> 1, drop.key, drop.thr, drop.scale, m0, m1);
                    drop_pair((e >> 1) + 1, drop.key, drop.thr, drop.scale, m2, m3);
                    o[0] *= m0; o[1] *= m1; o[2] *= m2; o[3] *= m3;
                }
                az[i].x += o[0]; az[i].y += o[1]; az[i].z += o[2]; az[i].w += o[3];
                if (dz16) store4t<T>(dz16 + static_cast<size_t>(row) * ld16 + c, o);
            }
        }
    }
    // CTA-wide combination of the three column-sum vectors, one after the other through the same [RW][n] buffer
#pragma unroll 1
    for (int which = 0; which < 3; ++which) {
        float* dst = which == 0 ? dgamma : (which == 1 ? dbeta : dbias);
        if (!dst) continue;                                  // uniform over the grid
        __syncthreads();
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const int c = lane * 4 + i * 128;
            if (c < n) *reinterpret_cast<float4*>(sred + warp * n + c) = which == 0 ? ag[i] : (which == 1 ? ab[i] : az[i]);
        }
        __syncthreads();
        for (int c = threadIdx.x; c < n; c += blockDim.x) {
            float t = 0.0f;
#pragma unroll
            for (int w = 0; w < RW; ++w) t += sred[w * n + c];
            atomicAdd(dst + c, t);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// activation backward, elementwise over [rows, cols] with the bias gradient (column sums of the result) on the side:
//   out = d * f'(a)      MODE 0: f = erf-GELU, a = pre-activation;  MODE 1: f = ReLU, a = the POST-activation (a > 0 <=> pre > 0)
// TD: type of the incoming gradient, TA: type of the saved activation, TO: type of the result (out may alias d when TO == TD).
// Each thread owns 4 consecutive columns and walks rows with stride gridDim.y; per-thread column sums -> one atomicAdd each.
// ---------------------------------------------------------------------------------------------------------------------
template <typename TD, typename TA, typename TO, int MODE>
__global__ void __launch_bounds__(256) act_bwd_kernel(const TD* __restrict__ d, const TA* __restrict__ a, TO* __restrict__ out,
                                                      int rows, int cols, float* __restrict__ dbias, DropSite drop) {
    const int c = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (c >= cols) return;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int r = blockIdx.y; r < rows; r += gridDim.y) {
        const size_t off = static_cast<size_t>(r) * cols + c;
        float dv[4], av[4], o[4];
        load4<TD>(d + off, dv); load4<TA>(a + off, av);
        if (drop.thr) {                                  // the activation's output was dropped in the forward: d <- d o mask
            float m0, m1, m2, m3;
            drop_pair(static_cast<uint32_t>(off >> 1), drop.key, drop.thr, drop.scale, m0, m1);
            drop_pair(static_cast<uint32_t>(off >> 1) + 1, drop.key, drop.thr, drop.scale, m2, m3);
            dv[0] *= m0; dv[1] *= m1; dv[2] *= m2; dv[3] *= m3;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            o[j] = MODE == 0 ? dv[j] * (sizeof(TA) == 2 ? gelu_grad_fast(av[j]) : gelu_grad(av[j])) : (av[j] > 0.0f ? dv[j] : 0.0f);
            if (sizeof(TO) == 2) o[j] = to_float<TO>(from_float<TO>(o[j]));     // the bias gradient sums what the GEMMs will see
            acc[j] += o[j];
        }
        store4t<TO>(out + off, o);
    }
    if (dbias) {
#pragma unroll
        for (int j = 0; j < 4; ++j) atomicAdd(dbias + c + j, acc[j]);
    }
}

// all-16-bit version (the FFN's GELU backward, [M, FF]): 8 columns = 16 bytes per thread and row, two rows in flight
template <typename T, int MODE>
__global__ void __launch_bounds__(256) act_bwd16_kernel(const T* __restrict__ d, const T* __restrict__ a, T* __restrict__ out, int rows, int cols,
                                                        float* __restrict__ dbias, DropSite drop) {
    const int c = (blockIdx.x * 256 + threadIdx.x) * 8;
    if (c >= cols) return;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int r = blockIdx.y; r < rows; r += gridDim.y) {
        const size_t off = static_cast<size_t>(r) * cols + c;
        const uint4 dq = *reinterpret_cast<const uint4*>(d + off);
        const uint4 aq = *reinterpret_cast<const uint4*>(a + off);
        const uint32_t dw[4] = {dq.x, dq.y, dq.z, dq.w}, aw[4] = {aq.x, aq.y, aq.z, aq.w};
        uint32_t ow[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float2 dv = unpack2<T>(dw[j]);
            const float2 av = unpack2<T>(aw[j]);
            if (drop.thr) {
                float m0, m1;
                drop_pair(static_cast<uint32_t>(off >> 1) + j, drop.key, drop.thr, drop.scale, m0, m1);
                dv.x *= m0; dv.y *= m1;
            }
            const float o0 = MODE == 0 ? dv.x * gelu_grad_fast(av.x) : (av.x > 0.0f ? dv.x : 0.0f);
            const float o1 = MODE == 0 ? dv.y * gelu_grad_fast(av.y) : (av.y > 0.0f ? dv.y : 0.0f);
            ow[j] = pack2<T>(o0, o1);
            const float2 q = unpack2<T>(ow[j]);                 // the bias gradient sums what the GEMMs will see
            acc[2 * j] += q.x; acc[2 * j + 1] += q.y;
        }
        *reinterpret_cast<uint4*>(out + off) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
    if (dbias) {
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(dbias + c + j, acc[j]);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) gelu_fwd_kernel(const T* __restrict__ u, T* __restrict__ h, size_t n4, DropSite drop) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        float v[4], o[4];
        load4<T>(u + 4 * i, v);
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = sizeof(T) == 4 ? gelu_erf(v[j]) : gelu_erf_fast(v[j]);
        if (drop.thr) {                                  // the FFN's inner dropout (transformers.py:107)
            float m0, m1, m2, m3;
            drop_pair(static_cast<uint32_t>(2 * i), drop.key, drop.thr, drop.scale, m0, m1);
            drop_pair(static_cast<uint32_t>(2 * i) + 1, drop.key, drop.thr, drop.scale, m2, m3);
            o[0] *= m0; o[1] *= m1; o[2] *= m2; o[3] *= m3;
        }
        store4t<T>(h + 4 * i, o);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) drop_cast_kernel(const float* __restrict__ in, T* __restrict__ out, size_t n2, DropSite drop) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n2; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const float2 v = *reinterpret_cast<const float2*>(in + 2 * i);
        float m0, m1;
        drop_pair(static_cast<uint32_t>(i), drop.key, drop.thr, drop.scale, m0, m1);
        out[2 * i] = from_float<T>(v.x * m0);
        out[2 * i + 1] = from_float<T>(v.y * m1);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) drop_apply_kernel(float* __restrict__ x32, T* __restrict__ x16, size_t n2, DropSite drop) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n2; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        float2 v = *reinterpret_cast<float2*>(x32 + 2 * i);
        float m0, m1;
        drop_pair(static_cast<uint32_t>(i), drop.key, drop.thr, drop.scale, m0, m1);
        v.x *= m0; v.y *= m1;
        *reinterpret_cast<float2*>(x32 + 2 * i) = v;
        if (x16) { x16[2 * i] = from_float<T>(v.x); x16[2 * i + 1] = from_float<T>(v.y); }
    }
}

// z = R(resid) + a o mask (one warp per row, 16-byte accesses)
__global__ void __launch_bounds__(RW * 32) residual_drop_kernel(const float* a, const float* __restrict__ resid, const float2* __restrict__ rstats,
                                                                const float* __restrict__ rgamma, const float* __restrict__ rbeta, float* z, int M,
                                                                int n, DropSite drop) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * RW + (threadIdx.x >> 5);
    if (row >= M) return;
    float rs = 1.0f, nm = 0.0f;
    if (rstats) { const float2 st = rstats[row]; rs = st.y; nm = -st.x * st.y; }
    const size_t base = static_cast<size_t>(row) * n;
    for (int c = lane * 4; c < n; c += 128) {
        float av[4], rv[4], o[4];
        load4<float>(a + base + c, av); load4<float>(resid + base + c, rv);
        if (rstats) {
            float g[4], b[4];
            load4<float>(rgamma + c, g); load4<float>(rbeta + c, b);
#pragma unroll
            for (int j = 0; j < 4; ++j) rv[j] = fmaf(fmaf(rv[j], rs, nm), g[j], b[j]);
        }
        float m0, m1, m2, m3;
        const uint32_t e = static_cast<uint32_t>(base + c);
        drop_pair(e >> 1, drop.key, drop.thr, drop.scale, m0, m1);
        drop_pair((e >> 1) + 1, drop.key, drop.thr, drop.scale, m2, m3);
        o[0] = fmaf(av[0], m0, rv[0]); o[1] = fmaf(av[1], m1, rv[1]); o[2] = fmaf(av[2], m2, rv[2]); o[3] = fmaf(av[3], m3, rv[3]);
        store4t<float>(z + base + c, o);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// column sums over a strided set of rows: out[c] += sum_{g < G, r < R} x[(g * group_rows + row_off + r) * ld + col_off + c]
// (bias gradients: G = 1; CLS / modality-encoding gradients: the rows of one token group in every clip)
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ x, int ld, int G, int group_rows, int row_off, int R,
                                                     int col_off, int ncols, float* __restrict__ out) {
    const int c = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (c >= ncols) return;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const long long total = static_cast<long long>(G) * R;
    for (long long i = blockIdx.y; i < total; i += gridDim.y) {
        const long long g = i / R, r = i - g * R;
        float v[4];
        load4<T>(x + static_cast<size_t>(g * group_rows + row_off + r) * ld + col_off + c, v);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] += v[j];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) atomicAdd(out + c + j, acc[j]);
}

// plain [rows, ncols] 16-bit matrix (the in_proj bias gradient over dqkv): 8 columns = 16 bytes per thread and row, no index
// arithmetic in the loop (the generic kernel above spends a 64-bit division per row: 62 % issue-active at 2.2 TB/s)
template <typename T>
__global__ void __launch_bounds__(256) colsum16_kernel(const T* __restrict__ x, int ld, int rows, int ncols, float* __restrict__ out) {
    const int c = (blockIdx.x * 256 + threadIdx.x) * 8;
    if (c >= ncols) return;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const T* p = x + static_cast<size_t>(blockIdx.y) * ld + c;
    const size_t step = static_cast<size_t>(gridDim.y) * ld;
    for (int r = blockIdx.y; r < rows; r += gridDim.y, p += step) {
        const uint4 q = *reinterpret_cast<const uint4*>(p);
        const float2 a = unpack2<T>(q.x), b = unpack2<T>(q.y), c2 = unpack2<T>(q.z), d2 = unpack2<T>(q.w);
        acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y; acc[4] += c2.x; acc[5] += c2.y; acc[6] += d2.x; acc[7] += d2.y;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(out + c + j, acc[j]);
}

// scalar version for rows whose width / pitch is not a multiple of 4 (class counts such as 97 or 3806): one thread per column
template <typename T>
__global__ void __launch_bounds__(256) colsum_scalar_kernel(const T* __restrict__ x, int ld, long long rows, int ncols, float* __restrict__ out) {
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= ncols) return;
    float acc = 0.0f;
    for (long long r = blockIdx.y; r < rows; r += gridDim.y) acc += to_float<T>(x[static_cast<size_t>(r) * ld + c]);
    atomicAdd(out + c, acc);
}

// Wt[k, n] = T(W[n, k]) for n < N, 0 for N <= n < Np  (the K-major operand of the dgrad GEMMs: dX = dY * W = dY * (W^T)^T)
template <typename T>
__global__ void __launch_bounds__(256) transpose_pack_kernel(const float* __restrict__ W, T* __restrict__ Wt, int N, int K, int Np) {
    __shared__ float tile[32][33];
    const int n0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 32; i += 8) {
        const int n = n0 + i, k = k0 + tx;
        tile[i][tx] = (n < N && k < K) ? W[static_cast<size_t>(n) * K + k] : 0.0f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int k = k0 + i, n = n0 + tx;
        if (k < K && n < Np) Wt[static_cast<size_t>(k) * Np + n] = from_float<T>(tile[tx][i]);
    }
}

// out[r, c] = T(in[r, c]) for c < C, 0 for C <= c < Cp
template <typename T>
__global__ void __launch_bounds__(256) cast_pad_kernel(const float* __restrict__ in, T* __restrict__ out, size_t rows, int C, int Cp) {
    const size_t total = rows * static_cast<size_t>(Cp);
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const size_t r = i / Cp;
        const int c = static_cast<int>(i - r * Cp);
        out[i] = from_float<T>(c < C ? in[r * C + c] : 0.0f);
    }
}

// dense copy of one token group's rows: out[b * Q + r, :] = x[b * Qt + off + r, :]   (one warp per row)
template <typename T>
__global__ void __launch_bounds__(RW * 32) gather_group_kernel(const T* __restrict__ x, T* __restrict__ out, int B, int Qt, int off, int Q, int E) {
    const int lane = threadIdx.x & 31;
    const long long row = static_cast<long long>(blockIdx.x) * RW + (threadIdx.x >> 5);
    if (row >= static_cast<long long>(B) * Q) return;
    const long long b = row / Q, r = row - b * Q;
    const T* src = x + static_cast<size_t>(b * Qt + off + r) * E;
    T* dst = out + static_cast<size_t>(row) * E;
    for (int c = lane * 4; c < E; c += 128) {
        float v[4];
        load4<T>(src + c, v);
        store4t<T>(dst + c, v);
    }
}
// dx[b * Qt + off + r, :] += in[b * Q + r, :]   (fp32; every destination row belongs to exactly one source row of this launch)
__global__ void __launch_bounds__(RW * 32) scatter_add_group_kernel(const float* __restrict__ in, float* __restrict__ dx, int B, int Qt, int off,
                                                                    int Q, int E) {
    const int lane = threadIdx.x & 31;
    const long long row = static_cast<long long>(blockIdx.x) * RW + (threadIdx.x >> 5);
    if (row >= static_cast<long long>(B) * Q) return;
    const long long b = row / Q, r = row - b * Q;
    const float* src = in + static_cast<size_t>(row) * E;
    float* dst = dx + static_cast<size_t>(b * Qt + off + r) * E;
    for (int c = lane * 4; c < E; c += 128) {
        float4 a = *reinterpret_cast<const float4*>(src + c);
        const float4 d = *reinterpret_cast<const float4*>(dst + c);
        a.x += d.x; a.y += d.y; a.z += d.z; a.w += d.w;
        *reinterpret_cast<float4*>(dst + c) = a;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// token assembly backward, gather form (deterministic): one warp per time row (b, t)
//   dte[b, t, :]  = sum over the token rows that consumed time row t of dtok[row, d:2d]      (a query's encoding feeds up to
//                   three CLS tokens: verb / noun / action share it, encodings.py:207-236)
//   demb_v / demb_a[b * F + f, :] = dtok[feature row, 0:d]       (gradient of the embedder's LayerNorm output)
// CLS-parameter and modality-encoding gradients are column sums over row groups (colsum_kernel).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RW * 32) assemble_bwd_kernel(const AssembleBwdParams p) {
    const int lane = threadIdx.x & 31;
    const long long tr = static_cast<long long>(blockIdx.x) * RW + (threadIdx.x >> 5);
    if (tr >= static_cast<long long>(p.B) * p.T) return;
    const int b = static_cast<int>(tr / p.T), t = static_cast<int>(tr - static_cast<long long>(b) * p.T);
    const int d = p.d, E = 2 * p.d, Ft = p.Fv + p.Fa;
    const size_t n_feat = static_cast<size_t>(p.B) * Ft;
    float* out = p.dte + static_cast<size_t>(tr) * d;
    if (t < Ft) {
        const float* src = p.dtok + (static_cast<size_t>(b) * Ft + t) * E;
        float* de = t < p.Fv ? p.demb_v + (static_cast<size_t>(b) * p.Fv + t) * d : p.demb_a + (static_cast<size_t>(b) * p.Fa + (t - p.Fv)) * d;
        for (int c = lane * 4; c < d; c += 128) {
            *reinterpret_cast<float4*>(de + c) = *reinterpret_cast<const float4*>(src + c);
            *reinterpret_cast<float4*>(out + c) = *reinterpret_cast<const float4*>(src + d + c);
        }
        return;
    }
    for (int c = lane * 4; c < d; c += 128) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int start = 0;
        for (int g = 0; g < p.n_groups; ++g) {
            const int q = t - p.groups[g].te_off;
            if (q >= 0 && q < p.groups[g].count) {
                const float4 v = *reinterpret_cast<const float4*>(p.dtok + (n_feat + static_cast<size_t>(b) * p.Qt + start + q) * E + d + c);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
            start += p.groups[g].count;
        }
        *reinterpret_cast<float4*>(out + c) = acc;
    }
}

// first time-MLP layer backward: dW0[c, j] += sum_m d1[m, c] * times[m, j]   (Linear(2, d), tim.py:67)
template <typename T>
__global__ void __launch_bounds__(256) time_l0_bwd_kernel(const T* __restrict__ d1, const float* __restrict__ times, float* __restrict__ dW0,
                                                          int M, int d) {
    const int c = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (c >= d) return;
    float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
    for (int m = blockIdx.y; m < M; m += gridDim.y) {
        float v[4];
        load4<T>(d1 + static_cast<size_t>(m) * d + c, v);
        const float2 t = *reinterpret_cast<const float2*>(times + 2 * static_cast<size_t>(m));
#pragma unroll
        for (int j = 0; j < 4; ++j) { a0[j] = fmaf(v[j], t.x, a0[j]); a1[j] = fmaf(v[j], t.y, a1[j]); }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { atomicAdd(dW0 + 2 * (c + j), a0[j]); atomicAdd(dW0 + 2 * (c + j) + 1, a1[j]); }
}

// last regression layer backward (detection/.../helpers/head.py:101-103: Linear(E/2, 2) + Sigmoid), one warp per row:
//   dp[j] = dout[r, j] * y[r, j] * (1 - y[r, j]);  dW4[j, k] += dp[j] * h[r, k];  db4[j] += dp[j];
//   dh[r, k] = (h[r, k] > 0) * (dp[0] W4[0, k] + dp[1] W4[1, k])    (the ReLU in front of the layer is applied here)
// dW4 / db4 / db2 (column sums of dh) accumulate per CTA in shared memory.
template <typename T>
__global__ void __launch_bounds__(RW * 32) reg_final_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ y, const T* __restrict__ h,
                                                                const float* __restrict__ W4, float* __restrict__ dW4, float* __restrict__ db4,
                                                                T* __restrict__ dh, float* __restrict__ db2, int rows, int K) {
    extern __shared__ float sacc[];                      // [3][K]: dW4 row 0, dW4 row 1, db2;  + 2 floats db4
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 3 * K + 2; i += blockDim.x) sacc[i] = 0.0f;
    __syncthreads();
    for (int r = blockIdx.x * RW + warp; r < rows; r += gridDim.x * RW) {
        const float y0 = y[2 * static_cast<size_t>(r)], y1 = y[2 * static_cast<size_t>(r) + 1];
        const float p0 = dout[2 * static_cast<size_t>(r)] * y0 * (1.0f - y0), p1 = dout[2 * static_cast<size_t>(r) + 1] * y1 * (1.0f - y1);
        if (lane == 0) { atomicAdd(&sacc[3 * K], p0); atomicAdd(&sacc[3 * K + 1], p1); }
        for (int k = lane; k < K; k += 32) {
            const float hv = to_float<T>(h[static_cast<size_t>(r) * K + k]);
            atomicAdd(&sacc[k], p0 * hv);
            atomicAdd(&sacc[K + k], p1 * hv);
            float g = hv > 0.0f ? fmaf(p0, W4[k], p1 * W4[K + k]) : 0.0f;
            const T gq = from_float<T>(g);
            dh[static_cast<size_t>(r) * K + k] = gq;
            atomicAdd(&sacc[2 * K + k], to_float<T>(gq));
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K; i += blockDim.x) {
        atomicAdd(dW4 + i, sacc[i]);
        atomicAdd(dW4 + K + i, sacc[K + i]);
        atomicAdd(db2 + i, sacc[2 * K + i]);
    }
    if (threadIdx.x < 2) atomicAdd(db4 + threadIdx.x, sacc[3 * K + threadIdx.x]);
}

__global__ void __launch_bounds__(256) axpy_kernel(float* __restrict__ y, const float* __restrict__ x, size_t n4) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        float4 a = *reinterpret_cast<float4*>(y + 4 * i);
        const float4 b = *reinterpret_cast<const float4*>(x + 4 * i);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        *reinterpret_cast<float4*>(y + 4 * i) = a;
    }
}

inline int row_grid(long long rows, int cap = 148 * 8) {
    long long g = (rows + RW - 1) / RW;
    if (g > cap) g = cap;
    return g > 0 ? static_cast<int>(g) : 1;
}
inline int flat_grid(size_t n, int cap = 148 * 16) {
    size_t g = (n + 255) / 256;
    if (g > static_cast<size_t>(cap)) g = cap;
    return g ? static_cast<int>(g) : 1;
}
// rows walked per column slab: enough CTAs to fill the machine, few enough that the final atomics stay cheap
inline dim3 col_grid(int cols, long long rows) {
    const int gx = (cols + 1023) / 1024;
    long long gy = (148 * 8 + gx - 1) / gx;
    if (gy > rows) gy = rows;
    if (gy < 1) gy = 1;
    return dim3(gx, static_cast<unsigned>(gy));
}

}  // namespace

template <typename T>
cudaError_t launch_ln_bwd(float* dy, int ldd, const float* z, int ldz, const float* gamma, T* dz16, int ld16, float* dgamma, float* dbeta,
                          float* dbias, int M, int n, cudaStream_t s, DropSite drop) {
    if (M <= 0) return cudaSuccess;
    if (drop.thr && (static_cast<unsigned long long>(M) * n > 0xffffffffull || ld16 != n)) return cudaErrorInvalidValue;
    if ((n & 3) || (ldd & 3) || (ldz & 3) || (dz16 && (ld16 & 3)) || n > 2048) return cudaErrorInvalidValue;
    const size_t smem = static_cast<size_t>(RW) * n * sizeof(float);
    const int grid = row_grid(M, 148 * 2);
#define TIM_LNB(V_)                                                                                                     \
    do {                                                                                                                \
        auto kern = ln_bwd_kernel<T, V_>;                                                                               \
        static SmemAttrCache cache;                                                                                     \
        if (cudaError_t e = ensure_dynamic_smem(kern, smem, cache); e != cudaSuccess) return e;                         \
        kern<<<grid, RW * 32, smem, s>>>(dy, ldd, z, ldz, gamma, dz16, ld16, dgamma, dbeta, dbias, M, n, drop);         \
    } while (0)
    if (n <= 512) TIM_LNB(4);
    else if (n <= 1024) TIM_LNB(8);
    else if (n <= 1536) TIM_LNB(12);
    else TIM_LNB(16);
#undef TIM_LNB
    return cudaGetLastError();
}
template cudaError_t launch_ln_bwd<float>(float*, int, const float*, int, const float*, float*, int, float*, float*, float*, int, int, cudaStream_t, DropSite);
template cudaError_t launch_ln_bwd<__half>(float*, int, const float*, int, const float*, __half*, int, float*, float*, float*, int, int, cudaStream_t, DropSite);
template cudaError_t launch_ln_bwd<__nv_bfloat16>(float*, int, const float*, int, const float*, __nv_bfloat16*, int, float*, float*, float*, int, int, cudaStream_t, DropSite);

template <typename TD, typename TA, typename TO>
cudaError_t launch_act_bwd(int mode, const TD* d, const TA* a, TO* out, int rows, int cols, float* dbias, cudaStream_t s, DropSite drop) {
    if (rows <= 0 || cols <= 0) return cudaSuccess;
    if (cols & 3) return cudaErrorInvalidValue;
    if (drop.thr && static_cast<unsigned long long>(rows) * cols > 0xffffffffull) return cudaErrorInvalidValue;
    if constexpr (std::is_same<TD, TA>::value && std::is_same<TD, TO>::value && sizeof(TD) == 2) {
        if ((cols & 7) == 0 && ((reinterpret_cast<uintptr_t>(d) | reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
            const int gx = (cols + 2047) / 2048;
            long long gy = (148 * 8 + gx - 1) / gx;
            if (gy > rows) gy = rows;
            const dim3 grid(gx, static_cast<unsigned>(gy));
            if (mode == 0) act_bwd16_kernel<TD, 0><<<grid, 256, 0, s>>>(d, a, out, rows, cols, dbias, drop);
            else act_bwd16_kernel<TD, 1><<<grid, 256, 0, s>>>(d, a, out, rows, cols, dbias, drop);
            return cudaGetLastError();
        }
    }
    const dim3 grid = col_grid(cols, rows);
    if (mode == 0) act_bwd_kernel<TD, TA, TO, 0><<<grid, 256, 0, s>>>(d, a, out, rows, cols, dbias, drop);
    else act_bwd_kernel<TD, TA, TO, 1><<<grid, 256, 0, s>>>(d, a, out, rows, cols, dbias, drop);
    return cudaGetLastError();
}
#define TIM_ACT_BWD(TD, TA, TO) template cudaError_t launch_act_bwd<TD, TA, TO>(int, const TD*, const TA*, TO*, int, int, float*, cudaStream_t, DropSite);
TIM_ACT_BWD(float, float, float)
TIM_ACT_BWD(__half, __half, __half)
TIM_ACT_BWD(__nv_bfloat16, __nv_bfloat16, __nv_bfloat16)
TIM_ACT_BWD(float, float, __half)
TIM_ACT_BWD(float, float, __nv_bfloat16)
#undef TIM_ACT_BWD

template <typename T>
cudaError_t launch_gelu_fwd(const T* u, T* h, size_t n, cudaStream_t s, DropSite drop) {
    if (!n) return cudaSuccess;
    if ((n & 3) || (drop.thr && n > 0xffffffffull)) return cudaErrorInvalidValue;
    gelu_fwd_kernel<T><<<flat_grid(n / 4), 256, 0, s>>>(u, h, n / 4, drop);
    return cudaGetLastError();
}
template cudaError_t launch_gelu_fwd<float>(const float*, float*, size_t, cudaStream_t, DropSite);
template cudaError_t launch_gelu_fwd<__half>(const __half*, __half*, size_t, cudaStream_t, DropSite);
template cudaError_t launch_gelu_fwd<__nv_bfloat16>(const __nv_bfloat16*, __nv_bfloat16*, size_t, cudaStream_t, DropSite);

template <typename T>
cudaError_t launch_drop_cast(const float* in, T* out, size_t n, DropSite drop, cudaStream_t s) {
    if (!n) return cudaSuccess;
    if ((n & 1) || n > 0xffffffffull) return cudaErrorInvalidValue;
    drop_cast_kernel<T><<<flat_grid(n / 2), 256, 0, s>>>(in, out, n / 2, drop);
    return cudaGetLastError();
}
template cudaError_t launch_drop_cast<float>(const float*, float*, size_t, DropSite, cudaStream_t);
template cudaError_t launch_drop_cast<__half>(const float*, __half*, size_t, DropSite, cudaStream_t);
template cudaError_t launch_drop_cast<__nv_bfloat16>(const float*, __nv_bfloat16*, size_t, DropSite, cudaStream_t);

template <typename T>
cudaError_t launch_drop_apply(float* x32, T* x16, size_t n, DropSite drop, cudaStream_t s) {
    if (!n || !drop.thr) return cudaSuccess;
    if ((n & 1) || n > 0xffffffffull) return cudaErrorInvalidValue;
    drop_apply_kernel<T><<<flat_grid(n / 2), 256, 0, s>>>(x32, x16, n / 2, drop);
    return cudaGetLastError();
}
template cudaError_t launch_drop_apply<float>(float*, float*, size_t, DropSite, cudaStream_t);
template cudaError_t launch_drop_apply<__half>(float*, __half*, size_t, DropSite, cudaStream_t);
template cudaError_t launch_drop_apply<__nv_bfloat16>(float*, __nv_bfloat16*, size_t, DropSite, cudaStream_t);

cudaError_t launch_residual_drop(const float* a, const float* resid, const float2* rstats, const float* rgamma, const float* rbeta, float* z,
                                 int M, int n, DropSite drop, cudaStream_t s) {
    if (M <= 0) return cudaSuccess;
    if ((n & 3) || static_cast<unsigned long long>(M) * n > 0xffffffffull) return cudaErrorInvalidValue;
    residual_drop_kernel<<<(M + RW - 1) / RW, RW * 32, 0, s>>>(a, resid, rstats, rgamma, rbeta, z, M, n, drop);
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_colsum(const T* x, int ld, int G, int group_rows, int row_off, int R, int col_off, int ncols, float* out, cudaStream_t s) {
    if (G <= 0 || R <= 0 || ncols <= 0) return cudaSuccess;
    if ((ncols & 3) || (ld & 3) || (col_off & 3) || (reinterpret_cast<uintptr_t>(x) & 15)) {
        if (G != 1) return cudaErrorInvalidValue;                 // the scalar fallback covers plain matrices only
        const int gx = (ncols + 255) / 256;
        long long gy = (148 * 8 + gx - 1) / gx;
        if (gy > R) gy = R;
        colsum_scalar_kernel<T><<<dim3(gx, static_cast<unsigned>(gy)), 256, 0, s>>>(x + static_cast<size_t>(row_off) * ld + col_off, ld, R, ncols, out);
        return cudaGetLastError();
    }
    if constexpr (sizeof(T) == 2) {
        if (G == 1 && (ncols & 7) == 0 && (ld & 7) == 0 && (col_off & 7) == 0) {
            const int gx = (ncols + 2047) / 2048;
            long long gy = (148 * 8 + gx - 1) / gx;
            if (gy > R) gy = R;
            colsum16_kernel<T><<<dim3(gx, static_cast<unsigned>(gy)), 256, 0, s>>>(x + static_cast<size_t>(row_off) * ld + col_off, ld, R, ncols, out);
            return cudaGetLastError();
        }
    }
    colsum_kernel<T><<<col_grid(ncols, static_cast<long long>(G) * R), 256, 0, s>>>(x, ld, G, group_rows, row_off, R, col_off, ncols, out);
    return cudaGetLastError();
}
template cudaError_t launch_colsum<float>(const float*, int, int, int, int, int, int, int, float*, cudaStream_t);
template cudaError_t launch_colsum<__half>(const __half*, int, int, int, int, int, int, int, float*, cudaStream_t);
template cudaError_t launch_colsum<__nv_bfloat16>(const __nv_bfloat16*, int, int, int, int, int, int, int, float*, cudaStream_t);

template <typename T>
cudaError_t launch_transpose_pack(const float* W, T* Wt, int N, int K, int Np, cudaStream_t s) {
    if (N <= 0 || K <= 0) return cudaSuccess;
    dim3 grid((Np + 31) / 32, (K + 31) / 32);
    transpose_pack_kernel<T><<<grid, 256, 0, s>>>(W, Wt, N, K, Np);
    return cudaGetLastError();
}
template cudaError_t launch_transpose_pack<float>(const float*, float*, int, int, int, cudaStream_t);
template cudaError_t launch_transpose_pack<__half>(const float*, __half*, int, int, int, cudaStream_t);
template cudaError_t launch_transpose_pack<__nv_bfloat16>(const float*, __nv_bfloat16*, int, int, int, cudaStream_t);

template <typename T>
cudaError_t launch_cast_pad(const float* in, T* out, size_t rows, int C, int Cp, cudaStream_t s) {
    if (!rows || Cp <= 0) return cudaSuccess;
    cast_pad_kernel<T><<<flat_grid(rows * Cp), 256, 0, s>>>(in, out, rows, C, Cp);
    return cudaGetLastError();
}
template cudaError_t launch_cast_pad<__half>(const float*, __half*, size_t, int, int, cudaStream_t);
template cudaError_t launch_cast_pad<__nv_bfloat16>(const float*, __nv_bfloat16*, size_t, int, int, cudaStream_t);

template <typename T>
cudaError_t launch_gather_group(const T* x, T* out, int B, int Qt, int off, int Q, int E, cudaStream_t s) {
    const long long rows = static_cast<long long>(B) * Q;
    if (rows <= 0) return cudaSuccess;
    if (E & 3) return cudaErrorInvalidValue;
    gather_group_kernel<T><<<static_cast<unsigned>((rows + RW - 1) / RW), RW * 32, 0, s>>>(x, out, B, Qt, off, Q, E);
    return cudaGetLastError();
}
template cudaError_t launch_gather_group<float>(const float*, float*, int, int, int, int, int, cudaStream_t);
template cudaError_t launch_gather_group<__half>(const __half*, __half*, int, int, int, int, int, cudaStream_t);
template cudaError_t launch_gather_group<__nv_bfloat16>(const __nv_bfloat16*, __nv_bfloat16*, int, int, int, int, int, cudaStream_t);

cudaError_t launch_scatter_add_group(const float* in, float* dx, int B, int Qt, int off, int Q, int E, cudaStream_t s) {
    const long long rows = static_cast<long long>(B) * Q;
    if (rows <= 0) return cudaSuccess;
    if (E & 3) return cudaErrorInvalidValue;
    scatter_add_group_kernel<<<static_cast<unsigned>((rows + RW - 1) / RW), RW * 32, 0, s>>>(in, dx, B, Qt, off, Q, E);
    return cudaGetLastError();
}

cudaError_t launch_assemble_bwd(const AssembleBwdParams& p, cudaStream_t s) {
    const long long rows = static_cast<long long>(p.B) * p.T;
    if (rows <= 0) return cudaSuccess;
    if (p.d & 3) return cudaErrorInvalidValue;
    assemble_bwd_kernel<<<static_cast<unsigned>((rows + RW - 1) / RW), RW * 32, 0, s>>>(p);
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_time_l0_bwd(const T* d1, const float* times, float* dW0, int M, int d, cudaStream_t s) {
    if (M <= 0) return cudaSuccess;
    if (d & 3) return cudaErrorInvalidValue;
    time_l0_bwd_kernel<T><<<col_grid(d, M), 256, 0, s>>>(d1, times, dW0, M, d);
    return cudaGetLastError();
}
template cudaError_t launch_time_l0_bwd<float>(const float*, const float*, float*, int, int, cudaStream_t);
template cudaError_t launch_time_l0_bwd<__half>(const __half*, const float*, float*, int, int, cudaStream_t);
template cudaError_t launch_time_l0_bwd<__nv_bfloat16>(const __nv_bfloat16*, const float*, float*, int, int, cudaStream_t);

template <typename T>
cudaError_t launch_reg_final_bwd(const float* dout, const float* y, const T* h, const float* W4, float* dW4, float* db4, T* dh, float* db2,
                                 int rows, int K, cudaStream_t s) {
    if (rows <= 0) return cudaSuccess;
    const size_t smem = (static_cast<size_t>(3) * K + 2) * sizeof(float);
    if (smem > 48 * 1024) return cudaErrorInvalidValue;
    reg_final_bwd_kernel<T><<<row_grid(rows, 148 * 2), RW * 32, smem, s>>>(dout, y, h, W4, dW4, db4, dh, db2, rows, K);
    return cudaGetLastError();
}
template cudaError_t launch_reg_final_bwd<float>(const float*, const float*, const float*, const float*, float*, float*, float*, float*, int, int, cudaStream_t);
template cudaError_t launch_reg_final_bwd<__half>(const float*, const float*, const __half*, const float*, float*, float*, __half*, float*, int, int, cudaStream_t);
template cudaError_t launch_reg_final_bwd<__nv_bfloat16>(const float*, const float*, const __nv_bfloat16*, const float*, float*, float*, __nv_bfloat16*, float*, int, int, cudaStream_t);

cudaError_t launch_axpy(float* y, const float* x, size_t n, cudaStream_t s) {
    if (!n) return cudaSuccess;
    if (n & 3) return cudaErrorInvalidValue;
    axpy_kernel<<<flat_grid(n / 4), 256, 0, s>>>(y, x, n / 4);
    return cudaGetLastError();
}

}  // namespace tim
