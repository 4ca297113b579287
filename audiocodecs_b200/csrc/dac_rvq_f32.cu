// DAC residual vector quantizer, fp32 (ac_dac_rvq_encode_f32 / ac_dac_rvq_decode_f32).
//
// Reference per stage (dac/nn/quantize.py @1.0.0; twin HF/dac/modeling_dac.py:122-170,281-343): z_e = in_proj(res)
// (1x1 conv 1024->8 + bias), L2-normalise z_e and the codebook rows, idx = argmax -(|a|^2 - 2 a.b) + |b|^2,
// z_q = codebook[idx], STE value z_e + (z_q - z_e), out = out_proj(.) (8->1024 + bias), res -= out.
// The reference launches ~12 kernels per stage and re-reads the [B,1024,N] residual (3.5 MB/clip) each time; here a
// CTA keeps FR frames' 1024-wide residual in shared memory across ALL stages and only the 8-dim projections,
// the codebook (32 KB) and the out_proj matrix stream through per stage.
#include "common.cuh"

namespace {

constexpr int FR = 16;      // frames per CTA
constexpr int HD = 1024;    // residual width
constexpr int CD = 8;       // codebook dim
constexpr int THREADS = 256;

__global__ void __launch_bounds__(THREADS)
dac_rvq_encode_kernel(const float* __restrict__ z, const float* __restrict__ w_in, const float* __restrict__ b_in,
                      const float* __restrict__ cb, const float* __restrict__ w_out, const float* __restrict__ b_out,
                      int64_t* __restrict__ codes, float* __restrict__ zq_out, int64_t rows, int n_codes, int stages,
                      int code_stride) {
    extern __shared__ __align__(16) float sm[];
    float* R = sm;                       // [FR][HD] residual
    float* ZE = R + FR * HD;             // [FR][CD] projected latents
    float* A = ZE + FR * CD;             // [FR][CD] normalised
    float* ZQ = A + FR * CD;             // [FR][CD] STE value
    float* bestv = ZQ + FR * CD;         // [FR][THREADS/FR]
    int* besti = reinterpret_cast<int*>(bestv + FR * (THREADS / FR));
    int* sel = besti + FR * (THREADS / FR);  // [FR]
    const int tid = threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.x * FR;
    for (int e = tid; e < FR * HD; e += THREADS) {
        const int f = e / HD, c = e % HD;
        R[e] = (r0 + f < rows) ? z[(r0 + f) * HD + c] : 0.f;
    }
    __syncthreads();
    for (int k = 0; k < stages; ++k) {
        const float* Wi = w_in + (size_t)k * CD * HD;     // [CD][HD]
        const float* Cb = cb + (size_t)k * n_codes * CD;  // [n_codes][CD]
        const float* Wo = w_out + (size_t)k * HD * CD;    // [HD][CD]
        // ---- z_e[f][d] = b_in[d] + sum_c Wi[d][c] * R[f][c]   (128 dot products, 2 threads each... one warp per (f,d) pair set)
        for (int pr = tid >> 5; pr < FR * CD; pr += THREADS / 32) {
            const int f = pr / CD, d = pr % CD;
            const int lane = tid & 31;
            float s = 0.f;
            for (int c = lane; c < HD; c += 32) s = fmaf(__ldg(Wi + (size_t)d * HD + c), R[f * HD + c], s);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) ZE[f * CD + d] = s + __ldg(b_in + k * CD + d);
        }
        __syncthreads();
        if (tid < FR) {  // F.normalize(z_e): x / max(|x|, 1e-12)
            float n = 0.f;
            for (int d = 0; d < CD; ++d) n += ZE[tid * CD + d] * ZE[tid * CD + d];
            const float inv = 1.0f / fmaxf(sqrtf(n), 1e-12f);
            for (int d = 0; d < CD; ++d) A[tid * CD + d] = ZE[tid * CD + d] * inv;
        }
        __syncthreads();
        // ---- argmax over codes: thread (f, lane16) scans codes lane16, lane16+16, ...
        {
            const int f = tid / (THREADS / FR), l = tid % (THREADS / FR);
            float a[CD], an = 0.f;
#pragma unroll
            for (int d = 0; d < CD; ++d) { a[d] = A[f * CD + d]; an += a[d] * a[d]; }
            float bv = -INFINITY;
            int bi = 0x7fffffff;
            for (int c = l; c < n_codes; c += THREADS / FR) {
                float b[CD], bn = 0.f;
#pragma unroll
                for (int d = 0; d < CD; ++d) { b[d] = __ldg(Cb + (size_t)c * CD + d); bn += b[d] * b[d]; }
                const float inv = 1.0f / fmaxf(sqrtf(bn), 1e-12f);
                float dot = 0.f, bnn = 0.f;
#pragma unroll
                for (int d = 0; d < CD; ++d) { const float bd = b[d] * inv; dot = fmaf(a[d], bd, dot); bnn += bd * bd; }
                const float score = -(an - 2.f * dot) + bnn;  // HF/dac:165
                if (score > bv || (score == bv && c < bi)) { bv = score; bi = c; }
            }
            bestv[tid] = bv;
            besti[tid] = bi;
        }
        __syncthreads();
        if (tid < FR) {
            float bv = -INFINITY;
            int bi = 0x7fffffff;
            for (int l = 0; l < THREADS / FR; ++l) {
                const float v = bestv[tid * (THREADS / FR) + l];
                const int i = besti[tid * (THREADS / FR) + l];
                if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
            }
            sel[tid] = bi;
            if (r0 + tid < rows) codes[(r0 + tid) * code_stride + k] = (int64_t)bi;
            for (int d = 0; d < CD; ++d) {
                const float ze = ZE[tid * CD + d];
                const float zq = __ldg(Cb + (size_t)bi * CD + d);
                ZQ[tid * CD + d] = ze + (zq - ze);  // straight-through arithmetic kept in eval
            }
        }
        __syncthreads();
        // ---- out = out_proj(zq) ; res -= out
        for (int e = tid; e < FR * HD; e += THREADS) {
            const int f = e / HD, c = e % HD;
            float o = __ldg(b_out + (size_t)k * HD + c);
#pragma unroll
            for (int d = 0; d < CD; ++d) o = fmaf(__ldg(Wo + (size_t)c * CD + d), ZQ[f * CD + d], o);
            R[e] -= o;
            if (zq_out && r0 + f < rows) {
                float* q = zq_out + (r0 + f) * HD + c;
                *q = (k == 0 ? 0.f : *q) + o;
            }
        }
        __syncthreads();
    }
}

// from_codes: out[row][c] = sum_k (b_out[k][c] + sum_d Wo[k][c][d] * cb[k][code][d]), accumulated in stage order from 0
__global__ void dac_rvq_decode_kernel(const int64_t* __restrict__ codes, const float* __restrict__ cb, const float* __restrict__ w_out,
                                      const float* __restrict__ b_out, float* __restrict__ out, int64_t rows, int n_codes,
                                      int stages, int code_stride, int* err_flag) {
    __shared__ float zp[16][CD];
    const int64_t row = blockIdx.x;
    const int tid = threadIdx.x;
    if (tid < stages * CD) {
        const int k = tid / CD, d = tid % CD;
        int64_t c = codes[row * code_stride + k];
        if (c < 0 || c >= n_codes) { if (err_flag) atomicExch(err_flag, 1); c = 0; }
        zp[k][d] = cb[((size_t)k * n_codes + c) * CD + d];
    }
    __syncthreads();
    for (int c = tid; c < HD; c += blockDim.x) {
        float acc = 0.f;
        for (int k = 0; k < stages; ++k) {
            float o = __ldg(b_out + (size_t)k * HD + c);
            const float* Wo = w_out + ((size_t)k * HD + c) * CD;
#pragma unroll
            for (int d = 0; d < CD; ++d) o = fmaf(__ldg(Wo + d), zp[k][d], o);
            acc += o;
        }
        out[row * HD + c] = acc;
    }
}

}  // namespace

extern "C" int ac_dac_rvq_encode_f32(const float* z, const float* w_in, const float* b_in, const float* codebooks,
                                     const float* w_out, const float* b_out, int64_t* codes, float* zq_out, int64_t rows,
                                     int32_t hidden, int32_t cb_dim, int32_t n_codes, int32_t stages, int32_t code_stride,
                                     void* stream) {
    AC_REQUIRE(z && w_in && b_in && codebooks && w_out && b_out && codes, "ac_dac_rvq_encode_f32: null pointer");
    AC_REQUIRE(hidden == HD && cb_dim == CD, "ac_dac_rvq_encode_f32: built for hidden %d / codebook dim %d", HD, CD);
    AC_REQUIRE(rows > 0 && stages > 0 && stages <= 16 && n_codes > 0, "ac_dac_rvq_encode_f32: bad sizes");
    const size_t smem = (size_t)(FR * HD + 3 * FR * CD + FR * (THREADS / FR)) * 4 + (size_t)(FR * (THREADS / FR) + FR) * 4;
    static bool set = false;
    if (!set) {
        cudaError_t e = cudaFuncSetAttribute(dac_rvq_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { ac::set_error("ac_dac_rvq_encode_f32: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        set = true;
    }
    dac_rvq_encode_kernel<<<(unsigned)((rows + FR - 1) / FR), THREADS, smem, (cudaStream_t)stream>>>(
        z, w_in, b_in, codebooks, w_out, b_out, codes, zq_out, rows, n_codes, stages, code_stride);
    return ac::finish_launch("ac_dac_rvq_encode_f32");
}

extern "C" int ac_dac_rvq_decode_f32(const int64_t* codes, const float* codebooks, const float* w_out, const float* b_out,
                                     float* out, int64_t rows, int32_t hidden, int32_t cb_dim, int32_t n_codes, int32_t stages,
                                     int32_t code_stride, int32_t* err_flag, void* stream) {
    AC_REQUIRE(codes && codebooks && w_out && b_out && out, "ac_dac_rvq_decode_f32: null pointer");
    AC_REQUIRE(hidden == HD && cb_dim == CD, "ac_dac_rvq_decode_f32: built for hidden %d / codebook dim %d", HD, CD);
    AC_REQUIRE(rows > 0 && stages > 0 && stages <= 16, "ac_dac_rvq_decode_f32: bad sizes");
    dac_rvq_decode_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(codes, codebooks, w_out, b_out, out, rows, n_codes,
                                                                             stages, code_stride, err_flag);
    return ac::finish_launch("ac_dac_rvq_decode_f32");
}
