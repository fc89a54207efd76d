// HBM-bound row kernels of the TIM forward: first time-MLP layer, LayerNorm, token assembly, casts, last regression
// layer. One warp per row, 128-bit loads where the row is wide, warp-shuffle reductions; nothing here is GEMM-shaped.
//
// Reference ops replaced:
//   time_mlp layer 0 (Linear(2, d) + ReLU)          recognition/.../models/tim.py:66-68
//   LayerNorm (norm1 / norm2 / time_mlp.6 / embedder.3)   transformers.py:105,109 ; tim.py:73 ; encodings.py:144,152
//   token concat / CLS expand / modality add        encodings.py:181-251 (and the uni-modal variants :41-75,102-121)
//   reg head last layer + Sigmoid                   detection/.../helpers/head.py:101-103
#include "kernels.h"
#include "ptx.cuh"

namespace tim {
namespace {

constexpr int ROWS_PER_CTA = 8;      // 8 warps, one row each

template <typename T> __device__ __forceinline__ void store4(T* p, float a, float b, float c, float d);

// One warp per time row: lane l writes columns 4l, 4l + 128, ... (8- or 16-byte stores), the two inputs of the row are loaded once.
template <typename TO>
__global__ void __launch_bounds__(ROWS_PER_CTA * 32) time_l1_kernel(const float* __restrict__ times, const float* __restrict__ W,
                                                                    const float* __restrict__ b, TO* __restrict__ out, int M, int d) {
    const int lane = threadIdx.x & 31;
    const int m = blockIdx.x * ROWS_PER_CTA + (threadIdx.x >> 5);
    if (m >= M) return;
    const float2 t = *reinterpret_cast<const float2*>(times + 2 * static_cast<size_t>(m));
    TO* o = out + static_cast<size_t>(m) * d;
    for (int c = lane * 4; c < d; c += 128) {
        const float4 w01 = __ldg(reinterpret_cast<const float4*>(W + 2 * c));          // (w[c][0], w[c][1], w[c+1][0], w[c+1][1])
        const float4 w23 = __ldg(reinterpret_cast<const float4*>(W + 2 * c + 4));
        const float4 bb = __ldg(reinterpret_cast<const float4*>(b + c));
        // same association as x @ W.T + b: (t0*w0 + t1*w1) + b
        const float v0 = fmaf(t.y, w01.y, t.x * w01.x) + bb.x, v1 = fmaf(t.y, w01.w, t.x * w01.z) + bb.y;
        const float v2 = fmaf(t.y, w23.y, t.x * w23.x) + bb.z, v3 = fmaf(t.y, w23.w, t.x * w23.z) + bb.w;
        store4<TO>(o + c, fmaxf(v0, 0.0f), fmaxf(v1, 0.0f), fmaxf(v2, 0.0f), fmaxf(v3, 0.0f));
    }
}

// two-pass (mean, then centred variance) LayerNorm statistics of one row held by a warp
__device__ __forceinline__ void row_stats(const float* __restrict__ row, int n, int lane, float& mean, float& rstd) {
    float s = 0.0f;
    for (int c = lane * 4; c < n; c += 128) {
        const float4 v = *reinterpret_cast<const float4*>(row + c);
        s += (v.x + v.y) + (v.z + v.w);
    }
    mean = warp_sum(s) / static_cast<float>(n);
    float q = 0.0f;
    for (int c = lane * 4; c < n; c += 128) {
        const float4 v = *reinterpret_cast<const float4*>(row + c);
        const float a = v.x - mean, b2 = v.y - mean, c2 = v.z - mean, d2 = v.w - mean;
        q += (a * a + b2 * b2) + (c2 * c2 + d2 * d2);
    }
    rstd = rsqrtf(warp_sum(q) / static_cast<float>(n) + 1e-5f);
}

template <> __device__ __forceinline__ void store4<float>(float* p, float a, float b, float c, float d) {
    *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
template <> __device__ __forceinline__ void store4<__half>(__half* p, float a, float b, float c, float d) {
    *reinterpret_cast<uint2*>(p) = make_uint2(pack2<__half>(a, b), pack2<__half>(c, d));
}
template <> __device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, float a, float b, float c, float d) {
    *reinterpret_cast<uint2*>(p) = make_uint2(pack2<__nv_bfloat16>(a, b), pack2<__nv_bfloat16>(c, d));
}

template <typename T>
__global__ void __launch_bounds__(ROWS_PER_CTA * 32) layernorm_kernel(const float* __restrict__ in, int ldi,
                                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                      float* __restrict__ out32, int ld32, T* __restrict__ out16, int ld16,
                                                                      int M, int n, float2* __restrict__ stats) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * ROWS_PER_CTA + (threadIdx.x >> 5);
    if (row >= M) return;
    const float* x = in + static_cast<size_t>(row) * ldi;
    float mean, rstd;
    row_stats(x, n, lane, mean, rstd);
    if (stats && lane == 0) stats[row] = make_float2(mean, rstd);
    for (int c = lane * 4; c < n; c += 128) {
        const float4 v = *reinterpret_cast<const float4*>(x + c);
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
        const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
        const float y0 = (v.x - mean) * rstd * g.x + b.x, y1 = (v.y - mean) * rstd * g.y + b.y;
        const float y2 = (v.z - mean) * rstd * g.z + b.z, y3 = (v.w - mean) * rstd * g.w + b.w;
        if (out32) store4<float>(out32 + static_cast<size_t>(row) * ld32 + c, y0, y1, y2, y3);
        if (out16) store4<T>(out16 + static_cast<size_t>(row) * ld16 + c, y0, y1, y2, y3);
    }
}

// Same, for rows of at most 128 * V floats: the row is read from memory ONCE and stays in registers for the two-pass
// statistics and the normalisation (the generic kernel above re-reads it twice through L2, which is what bounds it).
template <typename T, int V>
__global__ void __launch_bounds__(ROWS_PER_CTA * 32) layernorm_reg_kernel(const float* __restrict__ in, int ldi,
                                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                          float* __restrict__ out32, int ld32, T* __restrict__ out16, int ld16,
                                                                          int M, int n, float2* __restrict__ stats) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * ROWS_PER_CTA + (threadIdx.x >> 5);
    if (row >= M) return;
    const float* x = in + static_cast<size_t>(row) * ldi;
    float4 v[V];
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const int c = lane * 4 + i * 128;
        v[i] = c < n ? *reinterpret_cast<const float4*>(x + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(s) / static_cast<float>(n);
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        if (lane * 4 + i * 128 < n) {
            const float a = v[i].x - mean, b2 = v[i].y - mean, c2 = v[i].z - mean, d2 = v[i].w - mean;
            q += (a * a + b2 * b2) + (c2 * c2 + d2 * d2);
        }
    }
    const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(n) + 1e-5f);
    if (stats && lane == 0) stats[row] = make_float2(mean, rstd);
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const int c = lane * 4 + i * 128;
        if (c < n) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
            const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
            const float y0 = (v[i].x - mean) * rstd * g.x + b.x, y1 = (v[i].y - mean) * rstd * g.y + b.y;
            const float y2 = (v[i].z - mean) * rstd * g.z + b.z, y3 = (v[i].w - mean) * rstd * g.w + b.w;
            if (out32) store4<float>(out32 + static_cast<size_t>(row) * ld32 + c, y0, y1, y2, y3);
            if (out16) store4<T>(out16 + static_cast<size_t>(row) * ld16 + c, y0, y1, y2, y3);
        }
    }
}

// One warp per token row of the two-stream buffer.
//   feature row (b, f):  [ LN(emb[b, f]) | te[b, f_t] ] + mod
//   query row  (b, q):   [ cls          | te[b, t_q] ] + mod
template <typename T>
__global__ void __launch_bounds__(ROWS_PER_CTA * 32) assemble_kernel(const AssembleParams p) {
    const int lane = threadIdx.x & 31;
    const size_t row = static_cast<size_t>(blockIdx.x) * ROWS_PER_CTA + (threadIdx.x >> 5);
    const int Ft = p.Fv + p.Fa;
    const size_t n_feat = static_cast<size_t>(p.B) * Ft;
    const size_t n_rows = n_feat + static_cast<size_t>(p.B) * p.Qt;
    if (row >= n_rows) return;
    const int d = p.d, E = 2 * p.d;
    const float* left = nullptr;      // d values for the left half (pre-LN embedder row, or the CLS parameter)
    const float* g = nullptr; const float* bt = nullptr;
    const float* mod = nullptr;
    int b, te_row;
    if (row < n_feat) {
        b = static_cast<int>(row / Ft);
        const int f = static_cast<int>(row - static_cast<size_t>(b) * Ft);
        te_row = f;                   // time rows are ordered [vis F | aud F | queries], like the token rows
        if (f < p.Fv) {
            left = p.emb_v + (static_cast<size_t>(b) * p.Fv + f) * d; g = p.ln_v_g; bt = p.ln_v_b; mod = p.mod_v;
        } else {
            left = p.emb_a + (static_cast<size_t>(b) * p.Fa + (f - p.Fv)) * d; g = p.ln_a_g; bt = p.ln_a_b; mod = p.mod_a;
        }
    } else {
        const size_t qr = row - n_feat;
        b = static_cast<int>(qr / p.Qt);
        int q = static_cast<int>(qr - static_cast<size_t>(b) * p.Qt);
        int gi = 0;
        while (gi < p.n_groups - 1 && q >= p.groups[gi].count) { q -= p.groups[gi].count; ++gi; }
        left = p.groups[gi].cls; mod = p.groups[gi].mod;
        te_row = p.groups[gi].te_off + q;
    }
    float mean = 0.0f, rstd = 1.0f;
    if (g) row_stats(left, d, lane, mean, rstd);
    const float* te = p.te + (static_cast<size_t>(b) * p.T + te_row) * d;
    float* o32 = p.x32 ? p.x32 + row * E : nullptr;      // not written on the two-plane residual stream: (x16, xlo) carry the tokens
    T* o16 = p.x16 ? reinterpret_cast<T*>(p.x16) + row * E : nullptr;
    T* olo = p.xlo ? reinterpret_cast<T*>(p.xlo) + row * E : nullptr;
    for (int c = lane * 4; c < E; c += 128) {
        float4 v;
        if (c < d) {
            v = *reinterpret_cast<const float4*>(left + c);
            if (g) {
                const float4 gg = __ldg(reinterpret_cast<const float4*>(g + c));
                const float4 bb = __ldg(reinterpret_cast<const float4*>(bt + c));
                v.x = (v.x - mean) * rstd * gg.x + bb.x; v.y = (v.y - mean) * rstd * gg.y + bb.y;
                v.z = (v.z - mean) * rstd * gg.z + bb.z; v.w = (v.w - mean) * rstd * gg.w + bb.w;
            }
        } else {
            v = *reinterpret_cast<const float4*>(te + (c - d));
        }
        if (mod) {
            const float4 m = __ldg(reinterpret_cast<const float4*>(mod + c));
            v.x += m.x; v.y += m.y; v.z += m.z; v.w += m.w;
        }
        if (o32) store4<float>(o32 + c, v.x, v.y, v.z, v.w);
        if (o16) store4<T>(o16 + c, v.x, v.y, v.z, v.w);
        if (olo) {
            const float h0 = to_float<T>(from_float<T>(v.x)), h1 = to_float<T>(from_float<T>(v.y));
            const float h2 = to_float<T>(from_float<T>(v.z)), h3 = to_float<T>(from_float<T>(v.w));
            store4<T>(olo + c, v.x - h0, v.y - h1, v.z - h2, v.w - h3);
        }
    }
}

// LayerNorm of rows stored as two 16-bit planes (z = hi + lo): the row is read once into registers (so out16 may alias hi)
template <typename T, int V>
__global__ void __launch_bounds__(ROWS_PER_CTA * 32) layernorm_planes_kernel(const T* __restrict__ hi, const T* __restrict__ lo, int ldi,
                                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                             float* __restrict__ out32, int ld32, T* out16, int ld16, int M, int n) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * ROWS_PER_CTA + (threadIdx.x >> 5);
    if (row >= M) return;
    const T* ph = hi + static_cast<size_t>(row) * ldi;
    const T* pl = lo + static_cast<size_t>(row) * ldi;
    float4 v[V];
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const int c = lane * 4 + i * 128;
        if (c < n) {
            const uint2 a = *reinterpret_cast<const uint2*>(ph + c), b = *reinterpret_cast<const uint2*>(pl + c);
            const float2 a0 = unpack2<T>(a.x), a1 = unpack2<T>(a.y), b0 = unpack2<T>(b.x), b1 = unpack2<T>(b.y);
            v[i] = make_float4(a0.x + b0.x, a0.y + b0.y, a1.x + b1.x, a1.y + b1.y);
        } else {
            v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(s) / static_cast<float>(n);
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        if (lane * 4 + i * 128 < n) {
            const float a = v[i].x - mean, b2 = v[i].y - mean, c2 = v[i].z - mean, d2 = v[i].w - mean;
            q += (a * a + b2 * b2) + (c2 * c2 + d2 * d2);
        }
    }
    const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(n) + 1e-5f);
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const int c = lane * 4 + i * 128;
        if (c < n) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
            const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
            const float y0 = (v[i].x - mean) * rstd * g.x + b.x, y1 = (v[i].y - mean) * rstd * g.y + b.y;
            const float y2 = (v[i].z - mean) * rstd * g.z + b.z, y3 = (v[i].w - mean) * rstd * g.w + b.w;
            if (out32) store4<float>(out32 + static_cast<size_t>(row) * ld32 + c, y0, y1, y2, y3);
            if (out16) store4<T>(out16 + static_cast<size_t>(row) * ld16 + c, y0, y1, y2, y3);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) cast_kernel(const float* __restrict__ in, T* __restrict__ out, size_t n4, size_t scale_n4, float scale) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        float4 v = *reinterpret_cast<const float4*>(in + 4 * i);
        if (i < scale_n4) { v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale; }
        store4<T>(out + 4 * i, v.x, v.y, v.z, v.w);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) cast_tail_kernel(const float* __restrict__ in, T* __restrict__ out, size_t start, size_t n,
                                                        size_t scale_n, float scale) {
    const size_t i = start + blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i < n) out[i] = from_float<T>(i < scale_n ? in[i] * scale : in[i]);
}

// LayerNorm statistics of the rows a folded-LayerNorm producer GEMM wrote (gemm_umma2.cu, mode 5): the GEMM leaves per-tile
// partial (sum, sum of squares); var = E[z^2] - mean^2 in fp32 (relative error ~ 6e-8 * (1 + mean^2 / var), fine for
// residual-stream rows whose mean is far below their spread; clamped at 0).
__global__ void __launch_bounds__(256) row_stats_finalize_kernel(const float2* __restrict__ part, int P, float inv_width,
                                                                 float2* __restrict__ stats, int M, float alarm_ratio,
                                                                 int* __restrict__ alarm) {
    pdl_launch_dependents();          // the consumer GEMM behind this kernel may set itself up while these few CTAs run (ptx.cuh)
    const int row = blockIdx.x * 256 + threadIdx.x;
    if (row >= M) return;
    float s = 0.0f, q = 0.0f;
    for (int p = 0; p < P; ++p) {
        const float2 v = part[static_cast<size_t>(p) * M + row];
        s += v.x; q += v.y;
    }
    const float mean = s * inv_width;
    const float var = fmaxf(q * inv_width - mean * mean, 0.0f);
    stats[row] = make_float2(mean, rsqrtf(var + 1e-5f));
    // the folded path rounds z, not LN(z), to 16 bits: its relative error grows with sqrt(1 + mean^2 / var). Rows whose mean
    // dwarfs their spread raise a flag; the host side then takes the un-folded flow from the next call on (api.cu).
    if (alarm && mean * mean > alarm_ratio * (var + 1e-5f)) *alarm = 1;
}

// Folding a LayerNorm's affine into the weight of the linear layer that follows it (one warp per output row n):
//   Wf[n,k] = T(s_n * gamma[k] * W[n,k]),  cs[n] = sum_k float(Wf[n,k]),  bw[n] = s_n * (sum_k beta[k] * W[n,k] + bias[n])
// so that  LN(z) W^T + bias = rstd * (z Wf^T - mean * cs) + bw   (s_n: optional row scale, the folded q scaling of in_proj)
template <typename T>
__global__ void __launch_bounds__(ROWS_PER_CTA * 32) fold_ln_kernel(const float* __restrict__ W, const float* __restrict__ bias,
                                                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                    T* __restrict__ Wf, float* __restrict__ cs, float* __restrict__ bw,
                                                                    int N, int K, int scale_rows, float scale) {
    const int lane = threadIdx.x & 31;
    const int n = blockIdx.x * ROWS_PER_CTA + (threadIdx.x >> 5);
    if (n >= N) return;
    const float sn = n < scale_rows ? scale : 1.0f;
    const float* w = W + static_cast<size_t>(n) * K;
    float c = 0.0f, b = 0.0f;
    for (int k = lane; k < K; k += 32) {
        const T q = from_float<T>(sn * gamma[k] * w[k]);
        Wf[static_cast<size_t>(n) * K + k] = q;
        c += to_float<T>(q);
        b = fmaf(beta[k], w[k], b);
    }
    c = warp_sum(c); b = warp_sum(b);
    if (lane == 0) { cs[n] = c; bw[n] = sn * (b + bias[n]); }
}

// Window gather on the device (the reference gathers `feats[video][feat_indices, aug_indices]` on the host,
// recognition/.../datasets/sliding_window.py:356-375, and ships the result over PCIe): out[m, :] = TO(bank[rows[m], :]) for a
// feature bank resident in HBM (fp32 or 16-bit). One warp per row, 16-byte loads; a row index outside the bank yields zeros AND is counted
// in *bad (the reference's host-side indexing raises: the context turns the count into an error, tim_index_check).
template <typename TI, typename TO>
__global__ void __launch_bounds__(ROWS_PER_CTA * 32) gather_rows_kernel(const TI* __restrict__ bank, long long bank_rows,
                                                                        const long long* __restrict__ rows, TO* __restrict__ out,
                                                                        long long M, int D, int* __restrict__ bad) {
    const int lane = threadIdx.x & 31;
    const long long m = static_cast<long long>(blockIdx.x) * ROWS_PER_CTA + (threadIdx.x >> 5);
    if (m >= M) return;
    const long long r = rows[m];
    const bool ok = r >= 0 && r < bank_rows;
    if (!ok && lane == 0 && bad) atomicAdd(bad, 1);
    const TI* src = bank + (ok ? r : 0) * D;
    TO* dst = out + m * D;
    for (int c = lane * 4; c < D; c += 128) {
        float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
        if (ok) {
            if constexpr (sizeof(TI) == 4) {
                const float4 v = *reinterpret_cast<const float4*>(src + c);
                v0 = v.x; v1 = v.y; v2 = v.z; v3 = v.w;
            } else {
                const uint2 u = *reinterpret_cast<const uint2*>(src + c);
                const float2 a = unpack2<TI>(u.x), b = unpack2<TI>(u.y);
                v0 = a.x; v1 = a.y; v2 = b.x; v3 = b.y;
            }
        }
        store4<TO>(dst + c, v0, v1, v2, v3);
    }
}

template <typename T>
__global__ void __launch_bounds__(ROWS_PER_CTA * 32) reg_final_kernel(const T* __restrict__ h, int ldh, const float* __restrict__ W,
                                                                      const float* __restrict__ b, float* __restrict__ out, int rows, int K) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * ROWS_PER_CTA + (threadIdx.x >> 5);
    if (row >= rows) return;
    const T* x = h + static_cast<size_t>(row) * ldh;
    float a0 = 0.0f, a1 = 0.0f;
    for (int c = lane; c < K; c += 32) {
        const float v = to_float<T>(x[c]);
        a0 = fmaf(v, W[c], a0);
        a1 = fmaf(v, W[K + c], a1);
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1);
    if (lane == 0) {
        out[2 * static_cast<size_t>(row)] = 1.0f / (1.0f + expf(-(a0 + b[0])));
        out[2 * static_cast<size_t>(row) + 1] = 1.0f / (1.0f + expf(-(a1 + b[1])));
    }
}

inline int grid_for(size_t n, int block, int cap = 148 * 16) {
    size_t g = (n + block - 1) / block;
    if (g > static_cast<size_t>(cap)) g = cap;
    return g ? static_cast<int>(g) : 1;
}

}  // namespace

template <typename TO>
cudaError_t launch_time_l1(const float* times, const float* W, const float* b, TO* out, int M, int d, cudaStream_t s) {
    if (M <= 0) return cudaSuccess;
    if (d % 4) return cudaErrorInvalidValue;
    time_l1_kernel<TO><<<(M + ROWS_PER_CTA - 1) / ROWS_PER_CTA, ROWS_PER_CTA * 32, 0, s>>>(times, W, b, out, M, d);
    return cudaGetLastError();
}
template cudaError_t launch_time_l1<float>(const float*, const float*, const float*, float*, int, int, cudaStream_t);
template cudaError_t launch_time_l1<__half>(const float*, const float*, const float*, __half*, int, int, cudaStream_t);
template cudaError_t launch_time_l1<__nv_bfloat16>(const float*, const float*, const float*, __nv_bfloat16*, int, int, cudaStream_t);

template <typename T>
cudaError_t launch_layernorm(const float* in, int ldi, const float* gamma, const float* beta, float* out32, int ld32, T* out16,
                             int ld16, int M, int n, cudaStream_t s, float2* stats) {
    if (M <= 0) return cudaSuccess;
    if ((n & 3) || (ldi & 3) || (out32 && (ld32 & 3)) || (out16 && (ld16 & 3))) return cudaErrorInvalidValue;
    const int grid = (M + ROWS_PER_CTA - 1) / ROWS_PER_CTA;
    if (n <= 512) layernorm_reg_kernel<T, 4><<<grid, ROWS_PER_CTA * 32, 0, s>>>(in, ldi, gamma, beta, out32, ld32, out16, ld16, M, n, stats);
    else if (n <= 1024) layernorm_reg_kernel<T, 8><<<grid, ROWS_PER_CTA * 32, 0, s>>>(in, ldi, gamma, beta, out32, ld32, out16, ld16, M, n, stats);
    else if (n <= 2048) layernorm_reg_kernel<T, 16><<<grid, ROWS_PER_CTA * 32, 0, s>>>(in, ldi, gamma, beta, out32, ld32, out16, ld16, M, n, stats);
    else layernorm_kernel<T><<<grid, ROWS_PER_CTA * 32, 0, s>>>(in, ldi, gamma, beta, out32, ld32, out16, ld16, M, n, stats);
    return cudaGetLastError();
}
template cudaError_t launch_layernorm<float>(const float*, int, const float*, const float*, float*, int, float*, int, int, int, cudaStream_t, float2*);
template cudaError_t launch_layernorm<__half>(const float*, int, const float*, const float*, float*, int, __half*, int, int, int, cudaStream_t, float2*);
template cudaError_t launch_layernorm<__nv_bfloat16>(const float*, int, const float*, const float*, float*, int, __nv_bfloat16*, int, int, int, cudaStream_t, float2*);

template <typename T>
cudaError_t launch_layernorm_planes(const T* hi, const T* lo, int ldi, const float* gamma, const float* beta, float* out32, int ld32, T* out16,
                                    int ld16, int M, int n, cudaStream_t s) {
    if (M <= 0) return cudaSuccess;
    if ((n & 3) || (ldi & 3) || (out32 && (ld32 & 3)) || (out16 && (ld16 & 3)) || n > 2048) return cudaErrorInvalidValue;
    const int grid = (M + ROWS_PER_CTA - 1) / ROWS_PER_CTA;
    if (n <= 512) layernorm_planes_kernel<T, 4><<<grid, ROWS_PER_CTA * 32, 0, s>>>(hi, lo, ldi, gamma, beta, out32, ld32, out16, ld16, M, n);
    else if (n <= 1024) layernorm_planes_kernel<T, 8><<<grid, ROWS_PER_CTA * 32, 0, s>>>(hi, lo, ldi, gamma, beta, out32, ld32, out16, ld16, M, n);
    else layernorm_planes_kernel<T, 16><<<grid, ROWS_PER_CTA * 32, 0, s>>>(hi, lo, ldi, gamma, beta, out32, ld32, out16, ld16, M, n);
    return cudaGetLastError();
}
template cudaError_t launch_layernorm_planes<__half>(const __half*, const __half*, int, const float*, const float*, float*, int, __half*, int, int, int, cudaStream_t);
template cudaError_t launch_layernorm_planes<__nv_bfloat16>(const __nv_bfloat16*, const __nv_bfloat16*, int, const float*, const float*, float*, int, __nv_bfloat16*, int, int, int, cudaStream_t);

template <typename T>
cudaError_t launch_assemble(const AssembleParams& p, cudaStream_t s) {
    if (p.d & 3) return cudaErrorInvalidValue;
    const size_t rows = static_cast<size_t>(p.B) * (p.Fv + p.Fa + p.Qt);
    if (!rows) return cudaSuccess;
    assemble_kernel<T><<<static_cast<unsigned>((rows + ROWS_PER_CTA - 1) / ROWS_PER_CTA), ROWS_PER_CTA * 32, 0, s>>>(p);
    return cudaGetLastError();
}
template cudaError_t launch_assemble<float>(const AssembleParams&, cudaStream_t);
template cudaError_t launch_assemble<__half>(const AssembleParams&, cudaStream_t);
template cudaError_t launch_assemble<__nv_bfloat16>(const AssembleParams&, cudaStream_t);

template <typename T>
cudaError_t launch_cast(const float* in, T* out, size_t rows, size_t cols, size_t scale_rows, float scale, cudaStream_t s) {
    const size_t n = rows * cols, scale_n = scale_rows * cols;
    if (!n) return cudaSuccess;
    const size_t n4 = n / 4;
    // the scaled prefix must end on a float4 boundary for the vector body (true for every packed weight: cols % 4 == 0)
    if (scale_n % 4) return cudaErrorInvalidValue;
    if (n4) {
        cast_kernel<T><<<grid_for(n4, 256), 256, 0, s>>>(in, out, n4, scale_n / 4, scale);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    if (n % 4) cast_tail_kernel<T><<<1, 256, 0, s>>>(in, out, n4 * 4, n, scale_n, scale);
    return cudaGetLastError();
}
template cudaError_t launch_cast<__half>(const float*, __half*, size_t, size_t, size_t, float, cudaStream_t);
template cudaError_t launch_cast<__nv_bfloat16>(const float*, __nv_bfloat16*, size_t, size_t, size_t, float, cudaStream_t);

cudaError_t launch_scale_copy(const float* in, float* out, size_t rows, size_t cols, size_t scale_rows, float scale, cudaStream_t s) {
    const size_t n = rows * cols, scale_n = scale_rows * cols;
    if (!n) return cudaSuccess;
    if (scale_n % 4) return cudaErrorInvalidValue;
    const size_t n4 = n / 4;
    if (n4) {
        cast_kernel<float><<<grid_for(n4, 256), 256, 0, s>>>(in, out, n4, scale_n / 4, scale);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    if (n % 4) cast_tail_kernel<float><<<1, 256, 0, s>>>(in, out, n4 * 4, n, scale_n, scale);
    return cudaGetLastError();
}

template <typename TO>
cudaError_t launch_gather_rows(const void* bank, int bank_dtype, long long bank_rows, const long long* rows, TO* out, long long M, int D,
                               cudaStream_t s, int* bad) {
    if (M <= 0) return cudaSuccess;
    if (D % 4 || bank_rows <= 0) return cudaErrorInvalidValue;
    const unsigned grid = static_cast<unsigned>((M + ROWS_PER_CTA - 1) / ROWS_PER_CTA);
    switch (bank_dtype) {      // TIM_FP32 = 0, TIM_BF16 = 1, TIM_FP16 = 2
        case 0: gather_rows_kernel<float, TO><<<grid, ROWS_PER_CTA * 32, 0, s>>>(static_cast<const float*>(bank), bank_rows, rows, out, M, D, bad); break;
        case 1: gather_rows_kernel<__nv_bfloat16, TO><<<grid, ROWS_PER_CTA * 32, 0, s>>>(static_cast<const __nv_bfloat16*>(bank), bank_rows, rows, out, M, D, bad); break;
        case 2: gather_rows_kernel<__half, TO><<<grid, ROWS_PER_CTA * 32, 0, s>>>(static_cast<const __half*>(bank), bank_rows, rows, out, M, D, bad); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}
template cudaError_t launch_gather_rows<float>(const void*, int, long long, const long long*, float*, long long, int, cudaStream_t, int*);
template cudaError_t launch_gather_rows<__half>(const void*, int, long long, const long long*, __half*, long long, int, cudaStream_t, int*);
template cudaError_t launch_gather_rows<__nv_bfloat16>(const void*, int, long long, const long long*, __nv_bfloat16*, long long, int, cudaStream_t, int*);

cudaError_t launch_row_stats_finalize(const float2* part, int P, int width, float2* stats, int M, float alarm_ratio, int* alarm,
                                      cudaStream_t s) {
    if (M <= 0) return cudaSuccess;
    row_stats_finalize_kernel<<<(M + 255) / 256, 256, 0, s>>>(part, P, 1.0f / static_cast<float>(width), stats, M, alarm_ratio, alarm);
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_fold_ln(const float* W, const float* bias, const float* gamma, const float* beta, T* Wf, float* cs, float* bw,
                           int N, int K, int scale_rows, float scale, cudaStream_t s) {
    if (N <= 0) return cudaSuccess;
    fold_ln_kernel<T><<<(N + ROWS_PER_CTA - 1) / ROWS_PER_CTA, ROWS_PER_CTA * 32, 0, s>>>(W, bias, gamma, beta, Wf, cs, bw, N, K, scale_rows, scale);
    return cudaGetLastError();
}
template cudaError_t launch_fold_ln<__half>(const float*, const float*, const float*, const float*, __half*, float*, float*, int, int, int, float, cudaStream_t);
template cudaError_t launch_fold_ln<__nv_bfloat16>(const float*, const float*, const float*, const float*, __nv_bfloat16*, float*, float*, int, int, int, float, cudaStream_t);

template <typename T>
cudaError_t launch_reg_final(const T* h, int ldh, const float* W, const float* b, float* out, int rows, int K, cudaStream_t s) {
    if (rows <= 0) return cudaSuccess;
    reg_final_kernel<T><<<(rows + ROWS_PER_CTA - 1) / ROWS_PER_CTA, ROWS_PER_CTA * 32, 0, s>>>(h, ldh, W, b, out, rows, K);
    return cudaGetLastError();
}
template cudaError_t launch_reg_final<float>(const float*, int, const float*, const float*, float*, int, int, cudaStream_t);
template cudaError_t launch_reg_final<__half>(const __half*, int, const float*, const float*, float*, int, int, cudaStream_t);
template cudaError_t launch_reg_final<__nv_bfloat16>(const __nv_bfloat16*, int, const float*, const float*, float*, int, int, cudaStream_t);

}  // namespace tim
