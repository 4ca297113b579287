// fp32 SIMT tap-GEMM 1-D convolution (see ac_conv1d_f32 in include/audiocodecs_b200.h).
//
// One CTA computes a BM x BN tile of the per-clip output matrix [m_rows][n_cols]; the contraction
// runs over (tap j, input channel c) in chunks of BK through shared memory.  Input rows are
// gathered with the layer's padding rule (zero / reflect / replicate) and the prologue activation
// (ELU / Snake) is applied while staging, so padded copies and activated copies of the activation
// tensors are never materialised (the reference materialises both: HF/encodec:160-162, :268).
#include <cuda_bf16.h>

#include "common.cuh"

namespace {

constexpr int BK = 16;
constexpr int THREADS = 256;

__device__ __forceinline__ int pad_index(int pos, int L, int mode, int reflect_len) {
    // returns the source row, or -1 for a zero
    if (pos >= 0 && pos < L) return pos;
    if (mode == AC_PAD_ZERO) return -1;
    if (mode == AC_PAD_REPLICATE) return pos < 0 ? 0 : L - 1;
    // reflect (no edge repeat) over the zero-extended length reflect_len >= L (HF/encodec:148-155)
    int Lr = reflect_len;
    if (pos < 0) pos = -pos;
    if (pos >= Lr) pos = 2 * (Lr - 1) - pos;
    if (pos < 0 || pos >= L) return -1;
    return pos;
}

template <bool BF>
__device__ __forceinline__ float load_x(const void* base, int64_t idx) {
    if (BF) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[idx]);
    return __ldg(reinterpret_cast<const float*>(base) + idx);
}

template <int BM, int BN, bool XBF>
__global__ void __launch_bounds__(THREADS) conv1d_f32_kernel(const ac_conv_f32 p) {
    constexpr int TM = 4, TN = 4;
    static_assert((BM / TM) * (BN / TN) == THREADS, "tile/thread mismatch");
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];

    const int b = blockIdx.z;
    const int m0 = blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int tid = threadIdx.x;
    const int tx = tid % (BN / TN);
    const int ty = tid / (BN / TN);

    const int64_t xoff = (int64_t)b * p.x_bstride;
    const int vlen = p.vlen ? p.vlen[b] : p.x_rows;
    const int Ktot = p.taps * p.cin;

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < Ktot; k0 += BK) {
        // ---- stage A: BM rows x BK contraction entries (gather + activation)
        for (int e = tid; e < BM * BK; e += THREADS) {
            const int kk = e % BK;
            const int r = e / BK;
            const int k = k0 + kk;
            float v = 0.f;
            const int m = m0 + r;
            if (k < Ktot && m < p.m_rows) {
                const int j = k / p.cin;
                const int c = k - j * p.cin;
                const int pos = m * p.stride + j * p.dilation - p.pad_left;
                const int src = pad_index(pos, p.x_rows, p.pad_mode, p.reflect_len);
                if (src >= 0 && src < vlen) {
                    v = load_x<XBF>(p.x, xoff + (int64_t)src * p.x_rstride + c);
                    if (p.act == AC_ACT_ELU) v = ac::elu1(v);
                    else if (p.act == AC_ACT_SNAKE) v = ac::snake(v, __ldg(p.alpha + c));
                } else if (p.act == AC_ACT_SNAKE || p.act == AC_ACT_ELU) {
                    v = 0.f;  // act(0) == 0 for both
                }
            }
            As[kk][r] = v;
        }
        // ---- stage B: BK x BN weights
        for (int e = tid; e < BK * BN; e += THREADS) {
            const int n = e % BN;
            const int kk = e / BN;
            const int k = k0 + kk;
            float v = 0.f;
            if (k < Ktot && n0 + n < p.n_cols) v = __ldg(p.w + (int64_t)k * p.n_cols + n0 + n);
            Bs[kk][n] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * TM]);
            const float4 w = *reinterpret_cast<const float4*>(&Bs[kk][tx * TN]);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
        }
        __syncthreads();
    }

    // ---- epilogue: bias, activation, flat-shifted store (+ residual)
    float* __restrict__ yb = p.y ? p.y + (int64_t)b * p.y_bstride : nullptr;
    __nv_bfloat16* ybf = p.y_bf16 ? reinterpret_cast<__nv_bfloat16*>(p.y_bf16) + (int64_t)b * p.y_bf16_bstride : nullptr;
    __nv_bfloat16* yact = p.y_act_bf16 ? reinterpret_cast<__nv_bfloat16*>(p.y_act_bf16) + (int64_t)b * p.y_act_bstride : nullptr;
    const float* rb = p.res ? p.res + (int64_t)b * p.res_bstride : nullptr;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + ty * TM + i;
        if (m >= p.m_rows) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + tx * TN + j;
            if (n >= p.n_cols) continue;
            float v = acc[i][j] + (p.bias ? __ldg(p.bias + n) : 0.f);
            if (p.epi == AC_EPI_TANH) v = tanhf(v);
            else if (p.epi == AC_EPI_GELU) v = ac::gelu_erf(v);
            const int64_t flat = (int64_t)m * p.n_cols + n - p.out_shift;
            if (flat < 0 || flat >= p.out_valid) continue;
            if (rb) v += rb[flat];
            if (yb) yb[flat] = v;
            if (ybf) ybf[flat] = __float2bfloat16(v);
            if (yact) yact[flat] = __float2bfloat16(p.act2 == AC_ACT_ELU ? ac::elu1(v) : v);
        }
    }
}

}  // namespace

extern "C" int ac_conv1d_f32(const ac_conv_f32* p, void* stream) {
    AC_REQUIRE(p && p->x && p->w && (p->y || p->y_bf16 || p->y_act_bf16), "ac_conv1d_f32: null pointer");
    AC_REQUIRE(p->batch > 0 && p->m_rows > 0 && p->n_cols > 0 && p->taps > 0 && p->cin > 0,
               "ac_conv1d_f32: empty problem (batch %d rows %d cols %d)", p->batch, p->m_rows, p->n_cols);
    AC_REQUIRE(p->act != AC_ACT_SNAKE || p->alpha, "ac_conv1d_f32: snake needs alpha");
    AC_REQUIRE(p->batch <= 65535, "ac_conv1d_f32: batch %d > 65535", p->batch);
    cudaStream_t s = (cudaStream_t)stream;
    const bool xbf = p->x_is_bf16 != 0;
    if (p->n_cols <= 16) {
        dim3 grid((p->m_rows + 255) / 256, (p->n_cols + 15) / 16, p->batch);
        if (xbf) conv1d_f32_kernel<256, 16, true><<<grid, THREADS, 0, s>>>(*p);
        else conv1d_f32_kernel<256, 16, false><<<grid, THREADS, 0, s>>>(*p);
    } else if (p->n_cols <= 32) {
        dim3 grid((p->m_rows + 127) / 128, (p->n_cols + 31) / 32, p->batch);
        if (xbf) conv1d_f32_kernel<128, 32, true><<<grid, THREADS, 0, s>>>(*p);
        else conv1d_f32_kernel<128, 32, false><<<grid, THREADS, 0, s>>>(*p);
    } else {
        dim3 grid((p->m_rows + 63) / 64, (p->n_cols + 63) / 64, p->batch);
        if (xbf) conv1d_f32_kernel<64, 64, true><<<grid, THREADS, 0, s>>>(*p);
        else conv1d_f32_kernel<64, 64, false><<<grid, THREADS, 0, s>>>(*p);
    }
    return ac::finish_launch("ac_conv1d_f32");
}
