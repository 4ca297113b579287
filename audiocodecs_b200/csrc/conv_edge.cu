// The two HBM-bound edge layers of every codec on the bf16 path (ac_conv_first_bf16 / ac_conv_last_bf16):
//   first: waveform fp32 [B][T] (Cin = 1) -> C channels, bf16, raw and/or activated copy (HF/encodec:289, HF/mimi:461,
//          HF/dac:449);  7 MACs per output value -- not tensor-core work, the cost is the 2 x C x 2 bytes written per sample.
//   last : C channels bf16 -> waveform fp32 (Cout = 1) (+ tanh for DAC) (HF/encodec:341, HF/mimi:1169, HF/dac:434-437).
// Both keep the reference's padding rule (reflect with the tiny-input zero extension / zero) by index arithmetic on
// a shared-memory tile, so no padded copy of the 10 s x 64-clip tensors is ever made.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace {

constexpr int MAXK = 8;

__device__ __forceinline__ int pad_index(int pos, int L, int mode, int reflect_len) {
    if (pos >= 0 && pos < L) return pos;
    if (mode == AC_PAD_ZERO) return -1;
    if (mode == AC_PAD_REPLICATE) return pos < 0 ? 0 : L - 1;
    if (pos < 0) pos = -pos;
    if (pos >= reflect_len) pos = 2 * (reflect_len - 1) - pos;
    return (pos < 0 || pos >= L) ? -1 : pos;
}

__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// ELU through one bare MUFU.EX2 (same form as tcc::elu_ex2 of the tensor-path epilogues)
__device__ __forceinline__ float elu_fast(float x) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 1.4426950408889634f));
    return x > 0.f ? x : e - 1.0f;
}

// Snake with the hardware sine, as in the tensor-path epilogues (the output is rounded to bf16)
__device__ __forceinline__ float snake_fast(float x, float a, float ra) {
    const float s = __sinf(a * x);
    return fmaf(ra, s * s, x);
}

struct FirstP {
    const float* x; const float* w; const float* bias; const float* alpha; const int* vlen;
    __nv_bfloat16* y; __nv_bfloat16* y_act;
    __nv_bfloat16* y_lo; __nv_bfloat16* y_act_lo;   // optional lo planes: bf16(v - float(bf16(v)))
    long long y_bs, ya_bs;
    int T, K, pad_left, pad_mode, reflect_len, act;
    int out_f16;   // hi planes are written as IEEE fp16 (saturating) instead of bf16; lo planes stay bf16
};

__device__ __forceinline__ uint32_t pack2_f16(float a, float b) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}

// 8 values -> the 16 bytes of their 16-bit roundings; with `lo` also the 16 bytes of the rounding residuals (bf16)
__device__ __forceinline__ void store8_split(const float (&v)[8], __nv_bfloat16* hi, __nv_bfloat16* lo, bool f16) {
    uint32_t q[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) q[i] = f16 ? pack2_f16(v[2 * i], v[2 * i + 1]) : pack2(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(hi) = make_uint4(q[0], q[1], q[2], q[3]);
    if (lo) {
        uint32_t r[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 f;
            if (f16) f = __half22float2(*reinterpret_cast<const __half2*>(&q[i]));
            else { const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&q[i]); f = make_float2(__low2float(h2), __high2float(h2)); }
            r[i] = pack2(v[2 * i] - f.x, v[2 * i + 1] - f.y);
        }
        *reinterpret_cast<uint4*>(lo) = make_uint4(r[0], r[1], r[2], r[3]);
    }
}

// block = (C/8) x 32 threads; one block pass = 32 time steps x C channels; a block owns TILE time steps.
// The kernel was instruction-issue bound (ncu: 68 % issue slots, 265 instructions per 8 outputs against 1.9 GB of stores):
// ACT and the tap count are compile-time (KT = 0: run-time taps) and ELU goes through one MUFU.EX2 (abs error ~1e-7, far
// below the bf16 rounding of the output), which halves the instruction count.
template <int C, int ACT, int KT>
__global__ void __launch_bounds__(C * 4) conv_first_kernel(const FirstP p) {
    constexpr int G = C / 8;
    constexpr int TILE = 2048;
    __shared__ float xs[TILE + MAXK];
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * TILE;
    const int tid = threadIdx.x;
    const int grp = tid % G;   // 8-channel group
    const int tl = tid / G;    // time lane 0..31
    const int vlen = p.vlen ? p.vlen[b] : p.T;
    const float* xb = p.x + (long long)b * p.T;
    const int K = KT ? KT : p.K;
    for (int i = tid; i < TILE + K - 1; i += blockDim.x) {
        const int src = pad_index(t0 + i - p.pad_left, p.T, p.pad_mode, p.reflect_len);
        xs[i] = (src >= 0 && src < vlen) ? __ldg(xb + src) : 0.f;
    }
    float w[MAXK][8], bias[8], al[8], ral[8];
#pragma unroll
    for (int j = 0; j < MAXK; ++j)
#pragma unroll
        for (int c = 0; c < 8; ++c) w[j][c] = j < K ? __ldg(p.w + j * C + grp * 8 + c) : 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        bias[c] = p.bias ? __ldg(p.bias + grp * 8 + c) : 0.f;
        al[c] = ACT == AC_ACT_SNAKE ? __ldg(p.alpha + grp * 8 + c) : 0.f;
        ral[c] = 1.0f / (al[c] + 1e-9f);  // HF/dac:85-99
    }
    __syncthreads();
    for (int tt = tl; tt < TILE; tt += 32) {
        const int t = t0 + tt;
        if (t >= p.T) break;
        float acc[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = bias[c];
#pragma unroll
        for (int j = 0; j < MAXK; ++j) {
            if (j < K) {
                const float xv = xs[tt + j];
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[c] = fmaf(xv, w[j][c], acc[c]);
            }
        }
        const long long off = (long long)t * C + grp * 8;
        if (p.y) store8_split(acc, p.y + (long long)b * p.y_bs + off, p.y_lo ? p.y_lo + (long long)b * p.y_bs + off : nullptr, p.out_f16 != 0);
        if (p.y_act) {
            float a[8];
#pragma unroll
            for (int c = 0; c < 8; ++c)
                a[c] = ACT == AC_ACT_ELU ? elu_fast(acc[c])
                                         : (ACT == AC_ACT_SNAKE ? snake_fast(acc[c], al[c], ral[c]) : acc[c]);
            store8_split(a, p.y_act + (long long)b * p.ya_bs + off, p.y_act_lo ? p.y_act_lo + (long long)b * p.ya_bs + off : nullptr, p.out_f16 != 0);
        }
    }
}

struct LastP {
    const __nv_bfloat16* x; const float* w; const float* bias; float* y;
    long long x_bs;
    int T, K, pad_left, pad_mode, reflect_len, epi;
};

// block = 256 threads = 256 outputs; input tile [(256 + K - 1) rows][C] staged as padded rows of bf16x2 words.
template <int C>
__global__ void __launch_bounds__(256) conv_last_kernel(const LastP p) {
    constexpr int WPR = C / 2;        // 32-bit words per row
    constexpr int RS = WPR + 1;       // padded row stride (words): conflict-free for row-per-thread reads
    extern __shared__ uint32_t smem_u[];
    uint32_t* xs = smem_u;                                    // [(256 + MAXK) rows][RS]
    float* ws = reinterpret_cast<float*>(smem_u + (256 + MAXK) * RS);  // [K][C]
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * 256;
    const int tid = threadIdx.x;
    const uint32_t* xb = reinterpret_cast<const uint32_t*>(p.x + (long long)b * p.x_bs);
    const int rows = 256 + p.K - 1;
    for (int i = tid; i < rows * WPR; i += 256) {
        const int r = i / WPR, cw = i % WPR;
        const int src = pad_index(t0 + r - p.pad_left, p.T, p.pad_mode, p.reflect_len);
        xs[r * RS + cw] = src >= 0 ? __ldg(xb + (long long)src * WPR + cw) : 0u;
    }
    for (int i = tid; i < p.K * C; i += 256) ws[i] = __ldg(p.w + i);
    __syncthreads();
    const int t = t0 + tid;
    if (t >= p.T) return;
    float acc = p.bias ? __ldg(p.bias) : 0.f;
    for (int j = 0; j < p.K; ++j) {
        const uint32_t* row = xs + (tid + j) * RS;
        const float* wj = ws + j * C;
#pragma unroll 8
        for (int cw = 0; cw < WPR; ++cw) {
            const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&row[cw]);
            const float2 wv = *reinterpret_cast<const float2*>(&wj[2 * cw]);
            acc = fmaf(__low2float(h), wv.x, acc);
            acc = fmaf(__high2float(h), wv.y, acc);
        }
    }
    if (p.epi == AC_EPI_TANH) acc = tanhf(acc);
    p.y[(long long)b * p.T + t] = acc;
}

}  // namespace

extern "C" int ac_conv_first_bf16(const float* x, const float* w, const float* bias, const float* alpha, const int32_t* vlen,
                                  void* y, void* y_act, void* y_lo, void* y_act_lo, int64_t y_bstride, int64_t y_act_bstride,
                                  int32_t batch, int32_t T, int32_t C, int32_t K, int32_t pad_left, int32_t pad_mode,
                                  int32_t reflect_len, int32_t act, int32_t out_f16, void* stream) {
    AC_REQUIRE(x && w && (y || y_act), "ac_conv_first_bf16: null pointer");
    AC_REQUIRE(batch > 0 && batch <= 65535 && T > 0 && K >= 1 && K <= MAXK, "ac_conv_first_bf16: bad sizes");
    AC_REQUIRE(act != AC_ACT_SNAKE || alpha, "ac_conv_first_bf16: snake needs alpha");
    AC_REQUIRE((!y_lo || y) && (!y_act_lo || y_act), "ac_conv_first_bf16: lo plane without its hi plane");
    FirstP p{x, w, bias, alpha, vlen, (__nv_bfloat16*)y, (__nv_bfloat16*)y_act, (__nv_bfloat16*)y_lo, (__nv_bfloat16*)y_act_lo,
             y_bstride, y_act_bstride,
             T, K, pad_left, pad_mode, reflect_len < T ? T : reflect_len, act, out_f16 ? 1 : 0};
    dim3 grid((T + 2047) / 2048, batch);
    cudaStream_t s = (cudaStream_t)stream;
#define AC_FIRST(CC)                                                                                     \
    do {                                                                                                 \
        if (act == AC_ACT_ELU && K == 7) conv_first_kernel<CC, AC_ACT_ELU, 7><<<grid, CC * 4, 0, s>>>(p); \
        else if (act == AC_ACT_ELU) conv_first_kernel<CC, AC_ACT_ELU, 0><<<grid, CC * 4, 0, s>>>(p);      \
        else if (act == AC_ACT_SNAKE && K == 7) conv_first_kernel<CC, AC_ACT_SNAKE, 7><<<grid, CC * 4, 0, s>>>(p); \
        else if (act == AC_ACT_SNAKE) conv_first_kernel<CC, AC_ACT_SNAKE, 0><<<grid, CC * 4, 0, s>>>(p);  \
        else conv_first_kernel<CC, AC_ACT_NONE, 0><<<grid, CC * 4, 0, s>>>(p);                            \
    } while (0)
    switch (C) {
        case 32: AC_FIRST(32); break;
        case 64: AC_FIRST(64); break;
        case 96: AC_FIRST(96); break;
        default: ac::set_error("ac_conv_first_bf16: unsupported channel count %d (32/64/96)", C); return -2;
    }
#undef AC_FIRST
    return ac::finish_launch("ac_conv_first_bf16");
}

extern "C" int ac_conv_last_bf16(const void* x, const float* w, const float* bias, float* y, int64_t x_bstride, int32_t batch,
                                 int32_t T, int32_t C, int32_t K, int32_t pad_left, int32_t pad_mode, int32_t reflect_len,
                                 int32_t epi, void* stream) {
    AC_REQUIRE(x && w && y, "ac_conv_last_bf16: null pointer");
    AC_REQUIRE(batch > 0 && batch <= 65535 && T > 0 && K >= 1 && K <= MAXK, "ac_conv_last_bf16: bad sizes");
    LastP p{(const __nv_bfloat16*)x, w, bias, y, x_bstride, T, K, pad_left, pad_mode, reflect_len < T ? T : reflect_len, epi};
    dim3 grid((T + 255) / 256, batch);
    cudaStream_t s = (cudaStream_t)stream;
    auto smem_for = [&](int c) { return (size_t)(256 + MAXK) * (c / 2 + 1) * 4 + (size_t)MAXK * c * 4; };
    cudaError_t e = cudaSuccess;
    switch (C) {
        case 32: conv_last_kernel<32><<<grid, 256, smem_for(32), s>>>(p); break;
        case 64: conv_last_kernel<64><<<grid, 256, smem_for(64), s>>>(p); break;
        case 96:
            e = cudaFuncSetAttribute(conv_last_kernel<96>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_for(96));
            if (e != cudaSuccess) { ac::set_error("ac_conv_last_bf16: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
            conv_last_kernel<96><<<grid, 256, smem_for(96), s>>>(p);
            break;
        default: ac::set_error("ac_conv_last_bf16: unsupported channel count %d (32/64/96)", C); return -2;
    }
    return ac::finish_launch("ac_conv_last_bf16");
}
