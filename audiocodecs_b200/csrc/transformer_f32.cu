// Mimi transformer pieces that are not GEMMs, fp32 (exact-parity path): LayerNorm, RoPE + causal sliding-window
// attention, depthwise transposed-conv upsample.  The GEMMs (q/k/v/o, fc1+GELU, fc2, with the layer-scale folded
// into the weights and the residual add in the epilogue) run on the conv kernels.
// Replaces MimiTransformerLayer / MimiAttention / MimiRotaryEmbedding (HF/mimi/modeling_mimi.py:926-993,645-736,
// 515-577) and the `upsample` MimiConvTranspose1d (HF/mimi:1433-1441).
#include "common.cuh"

namespace {

// ---------------------------------------------------------------- LayerNorm over the channel axis, one warp per row
__global__ void layernorm_f32_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                     float* __restrict__ y, long long rows, int C, float eps) {
    const long long row = (long long)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const float* xr = x + row * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += xr[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / C;
    float v = 0.f;
    for (int c = lane; c < C; c += 32) { const float d = xr[c] - mean; v += d * d; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const float rstd = rsqrtf(v / C + eps);
    for (int c = lane; c < C; c += 32) y[row * C + c] = (xr[c] - mean) * rstd * w[c] + b[c];
}

// ---------------------------------------------------------------- attention
// qkv [B][T][3*H*D] (q | k | v, heads contiguous inside each), out [B][T][H*D].  One CTA = (64 queries, head, clip): the
// keys max(0, q0-window+1) .. q0+63 are staged ONCE with RoPE applied (rotate-half form), then every warp takes four
// queries at a time: scores with lane = key (one 16-byte K read feeds four queries), a two-pass softmax through a
// per-warp probability buffer, and P.V with lane = output dim pair.  ~3.3x fewer shared-memory instructions per
// query than one-query-at-a-time; the math (fp32, same masks, max-subtracted softmax) is the reference's.
constexpr int D = 64;
constexpr int QT = 64;         // queries per CTA
constexpr int QW = 4;          // queries per warp iteration
constexpr int KS = D + 4;      // K / V row stride in floats: 16-byte aligned rows, conflict-free quarter-warp LDS.128

__global__ void __launch_bounds__(256)
attention_f32_kernel(const float* __restrict__ qkv, const float* __restrict__ inv_freq, float* __restrict__ out, int T, int H,
                     int window, float scaling) {
    extern __shared__ __align__(16) float sm[];
    const int q0 = blockIdx.x * QT, h = blockIdx.y, b = blockIdx.z;
    const int k_lo = max(0, q0 - window + 1);
    const int k_hi = min(T, q0 + QT);  // exclusive
    const int nk = k_hi - k_lo;
    const int kcap = window + QT;      // rows reserved per K / V buffer and per probability row
    float* Ks = sm;                          // [kcap][KS]
    float* Vs = Ks + (size_t)kcap * KS;      // [kcap][KS]
    float* Qs = Vs + (size_t)kcap * KS;      // [QT][D]
    float* Ps = Qs + QT * D;                 // [8 warps][QW][kcap]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t row_stride = (size_t)3 * H * D;
    const float* base = qkv + (size_t)b * T * row_stride;
    for (int e = tid; e < nk * (D / 2); e += 256) {
        const int j = e / (D / 2), i = e % (D / 2);
        const int pos = k_lo + j;
        const float* kr = base + (size_t)pos * row_stride + (size_t)H * D + h * D;
        const float* vr = base + (size_t)pos * row_stride + (size_t)2 * H * D + h * D;
        const float ang = (float)pos * inv_freq[i];
        float sn, cs;
        sincosf(ang, &sn, &cs);
        const float x1 = kr[i], x2 = kr[i + D / 2];
        Ks[j * KS + i] = x1 * cs - x2 * sn;            // q*cos + rotate_half(q)*sin, first half: -x2
        Ks[j * KS + i + D / 2] = x2 * cs + x1 * sn;    // second half: +x1
        Vs[j * KS + i] = vr[i];
        Vs[j * KS + i + D / 2] = vr[i + D / 2];
    }
    for (int e = tid; e < QT * (D / 2); e += 256) {
        const int qi = e / (D / 2), i = e % (D / 2);
        const int pos = q0 + qi;
        float a = 0.f, c2 = 0.f;
        if (pos < T) {
            const float* qr = base + (size_t)pos * row_stride + h * D;
            const float ang = (float)pos * inv_freq[i];
            float sn, cs;
            sincosf(ang, &sn, &cs);
            const float x1 = qr[i], x2 = qr[i + D / 2];
            a = (x1 * cs - x2 * sn) * scaling;           // the 1/sqrt(D) scaling is folded into q
            c2 = (x2 * cs + x1 * sn) * scaling;
        }
        Qs[qi * D + i] = a;
        Qs[qi * D + i + D / 2] = c2;
    }
    __syncthreads();
    float* P = Ps + (size_t)warp * QW * kcap;
    for (int qg = warp * QW; qg < QT; qg += 8 * QW) {
        if (q0 + qg >= T) break;
        // local key ranges of the four queries: [lo_i, hi_i]; the union [lo_0, hi_3] is scanned once
        int lo[QW], hi[QW];
#pragma unroll
        for (int i = 0; i < QW; ++i) {
            const int pos = min(q0 + qg + i, T - 1);
            lo[i] = max(0, pos - window + 1) - k_lo;
            hi[i] = pos - k_lo;
        }
        const bool live[QW] = {true, q0 + qg + 1 < T, q0 + qg + 2 < T, q0 + qg + 3 < T};
        float mx[QW] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        for (int j = lo[0] + lane; j <= hi[QW - 1]; j += 32) {
            float s[QW] = {0.f, 0.f, 0.f, 0.f};
            const float4* kr = reinterpret_cast<const float4*>(Ks + (size_t)j * KS);
#pragma unroll 4
            for (int d4 = 0; d4 < D / 4; ++d4) {
                const float4 k = kr[d4];
#pragma unroll
                for (int i = 0; i < QW; ++i) {
                    const float4 q = *reinterpret_cast<const float4*>(Qs + (qg + i) * D + d4 * 4);
                    s[i] = fmaf(q.x, k.x, s[i]); s[i] = fmaf(q.y, k.y, s[i]); s[i] = fmaf(q.z, k.z, s[i]); s[i] = fmaf(q.w, k.w, s[i]);
                }
            }
#pragma unroll
            for (int i = 0; i < QW; ++i) {
                const bool ok = j >= lo[i] && j <= hi[i];
                P[i * kcap + j] = ok ? s[i] : -INFINITY;
                if (ok) mx[i] = fmaxf(mx[i], s[i]);
            }
        }
        float inv[QW];
#pragma unroll
        for (int i = 0; i < QW; ++i) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], o));
        }
        __syncwarp();
        float sum[QW] = {0.f, 0.f, 0.f, 0.f};
        for (int j = lo[0] + lane; j <= hi[QW - 1]; j += 32) {
#pragma unroll
            for (int i = 0; i < QW; ++i) {
                const float e = expf(P[i * kcap + j] - mx[i]);  // masked entries: exp(-inf) = 0
                P[i * kcap + j] = e;
                sum[i] += e;
            }
        }
#pragma unroll
        for (int i = 0; i < QW; ++i) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum[i] += __shfl_xor_sync(0xffffffffu, sum[i], o);
            inv[i] = 1.0f / sum[i];
        }
        __syncwarp();
        float o0[QW] = {0.f, 0.f, 0.f, 0.f}, o1[QW] = {0.f, 0.f, 0.f, 0.f};  // dims lane, lane+32
        for (int j = lo[0]; j <= hi[QW - 1]; ++j) {
            const float v0 = Vs[(size_t)j * KS + lane], v1 = Vs[(size_t)j * KS + lane + 32];
#pragma unroll
            for (int i = 0; i < QW; ++i) {
                const float pj = P[i * kcap + j];
                o0[i] = fmaf(pj, v0, o0[i]);
                o1[i] = fmaf(pj, v1, o1[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < QW; ++i) {
            if (!live[i]) continue;
            float* orow = out + ((size_t)b * T + q0 + qg + i) * H * D + h * D;
            orow[lane] = o0[i] * inv[i];
            orow[lane + 32] = o1[i] * inv[i];
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------- depthwise ConvTranspose1d(k=4, s=2, groups=C), trim right 2
__global__ void upsample_dw_f32_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y,
                                       long long total, int L, int C) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // over [B][2L][C]
    if (i >= total) return;
    const int c = (int)(i % C);
    const long long t = (i / C) % (2 * L);
    const long long b = i / ((long long)C * 2 * L);
    const int q = (int)(t >> 1), r = (int)(t & 1);
    const float* xb = x + b * (long long)L * C;
    float v = xb[(long long)q * C + c] * w[c * 4 + r];
    if (q > 0) v = fmaf(xb[(long long)(q - 1) * C + c], w[c * 4 + r + 2], v);
    y[i] = v;
}

}  // namespace

extern "C" int ac_layernorm_f32(const float* x, const float* w, const float* b, float* y, int64_t rows, int32_t C, float eps,
                                void* stream) {
    AC_REQUIRE(x && w && b && y && rows > 0 && C > 0, "ac_layernorm_f32: bad arguments");
    const int wpb = 8;
    layernorm_f32_kernel<<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(x, w, b, y, rows, C, eps);
    return ac::finish_launch("ac_layernorm_f32");
}

extern "C" int ac_attention_f32(const float* qkv, const float* inv_freq, float* out, int32_t batch, int32_t T, int32_t heads,
                                int32_t head_dim, int32_t window, float scaling, void* stream) {
    AC_REQUIRE(qkv && inv_freq && out, "ac_attention_f32: null pointer");
    AC_REQUIRE(head_dim == D, "ac_attention_f32: head_dim %d (built for %d)", head_dim, D);
    AC_REQUIRE(batch > 0 && T > 0 && heads > 0 && window > 0 && window <= 256, "ac_attention_f32: bad sizes");
    const size_t smem = ((size_t)2 * (window + QT) * KS + QT * D + 8 * (size_t)QW * (window + QT)) * 4;
    static size_t set = 0;
    if (smem > set) {
        cudaError_t e = cudaFuncSetAttribute(attention_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { ac::set_error("ac_attention_f32: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        set = smem;
    }
    dim3 grid((T + QT - 1) / QT, heads, batch);
    attention_f32_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(qkv, inv_freq, out, T, heads, window, scaling);
    return ac::finish_launch("ac_attention_f32");
}

extern "C" int ac_upsample_dw_f32(const float* x, const float* w, float* y, int32_t batch, int32_t L, int32_t C, void* stream) {
    AC_REQUIRE(x && w && y && batch > 0 && L > 0 && C > 0, "ac_upsample_dw_f32: bad arguments");
    const long long total = (long long)batch * 2 * L * C;
    upsample_dw_f32_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, w, y, total, L, C);
    return ac::finish_launch("ac_upsample_dw_f32");
}
