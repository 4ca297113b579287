// Residual vector quantizer, fp32 (ac_rvq_encode_f32 / ac_rvq_decode_f32).
//
// Encode: one CTA keeps a tile of FR frames' residual resident in shared memory across ALL stages
// (the reference re-reads and re-writes the [B,128,N] residual from memory per stage and launches a
// GEMM + pow/sum/neg/max + embedding + sub per stage: HF/encodec:364-369,424-438).  Per stage the
// codebook streams through shared memory in chunks of CC codes; each thread owns a 2-frame x 4-code
// micro-tile of dot products, keeps a running best per frame, then the CTA reduces to the argmin with
// the reference's tie rule (lowest index) and subtracts the selected codeword in place.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace {

constexpr int FR = 32;    // frames per CTA
constexpr int CC = 128;   // codes per chunk
constexpr int THREADS = 256;

template <int D>
__global__ void __launch_bounds__(THREADS)
rvq_encode_f32_kernel(const float* __restrict__ x, const float* __restrict__ cbs, const float* __restrict__ cbn,
                      int64_t* __restrict__ codes_out, float* __restrict__ res_out, int64_t rows, int n_codes,
                      int stages, int code_stride, int code_offset, int metric) {
    extern __shared__ __align__(16) float smem[];
    float* Rs = smem;                       // [FR][D+1]
    float* Es = Rs + FR * (D + 1);          // [CC][D+1]
    float* best_v = Es + CC * (D + 1);      // [FR][32]
    int* best_i = reinterpret_cast<int*>(best_v + FR * 32);  // [FR][32]
    float* xnorm = reinterpret_cast<float*>(best_i + FR * 32);  // [FR]
    int* sel = reinterpret_cast<int*>(xnorm + FR);              // [FR]

    const int tid = threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.x * FR;
    const int fg = tid / 32;        // frame group: frames fg*4 .. fg*4+3   (8 groups x 4 frames)
    const int cl = tid % 32;        // code lane: codes cl*4 .. cl*4+3 of the chunk

    for (int e = tid; e < FR * D; e += THREADS) {
        const int f = e / D, d = e % D;
        Rs[f * (D + 1) + d] = (r0 + f < rows) ? x[(r0 + f) * D + d] : 0.f;
    }
    __syncthreads();

    for (int k = 0; k < stages; ++k) {
        const float* E = cbs + (size_t)k * n_codes * D;
        const float* En = cbn + (size_t)k * n_codes;
        // |r|^2 per frame, summed in index order
        if (tid < FR) {
            float s = 0.f;
            for (int d = 0; d < D; ++d) { const float v = Rs[tid * (D + 1) + d]; s += v * v; }
            xnorm[tid] = s;
        }
        float bv[4];
        int bi[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { bv[i] = -INFINITY; bi[i] = 0x7fffffff; }

        for (int c0 = 0; c0 < n_codes; c0 += CC) {
            __syncthreads();  // previous chunk consumed (and xnorm written)
            for (int e = tid; e < CC * D; e += THREADS) {
                const int c = e / D, d = e % D;
                Es[c * (D + 1) + d] = (c0 + c < n_codes) ? __ldg(E + (size_t)(c0 + c) * D + d) : 0.f;
            }
            __syncthreads();
            float acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 4
            for (int d = 0; d < D; ++d) {
                float rv[4], ev[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) rv[i] = Rs[(fg * 4 + i) * (D + 1) + d];
#pragma unroll
                for (int j = 0; j < 4; ++j) ev[j] = Es[(cl + 32 * j) * (D + 1) + d];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(rv[i], ev[j], acc[i][j]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = c0 + cl + 32 * j;
                if (c >= n_codes) continue;
                const float en = __ldg(En + c);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float xn = xnorm[fg * 4 + i];
                    float score;  // larger is better
                    if (metric == 0) score = -((xn - 2.f * acc[i][j]) + en);       // HF/encodec:367
                    else score = -sqrtf(fmaxf((xn + en) - 2.f * acc[i][j], 0.f));  // cdist (HF/mimi:1200)
                    if (score > bv[i] || (score == bv[i] && c < bi[i])) { bv[i] = score; bi[i] = c; }
                }
            }
        }
        // ---- reduce across the 32 code lanes of each frame (warp = one frame group)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float v = bv[i];
            int ix = bi[i];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, v, off);
                const int oi = __shfl_xor_sync(0xffffffffu, ix, off);
                if (ov > v || (ov == v && oi < ix)) { v = ov; ix = oi; }
            }
            // a NaN residual never beats -inf: the index stays at its sentinel -> code 0 (the reference returns an arbitrary
            // index there); the gather below must stay in range
            if (cl == 0) sel[fg * 4 + i] = (unsigned)ix < (unsigned)n_codes ? ix : 0;
        }
        __syncthreads();
        // ---- emit code, subtract the selected codeword
        if (tid < FR && r0 + tid < rows)
            codes_out[(r0 + tid) * code_stride + code_offset + k] = (int64_t)sel[tid];
        for (int e = tid; e < FR * D; e += THREADS) {
            const int f = e / D, d = e % D;
            Rs[f * (D + 1) + d] -= __ldg(E + (size_t)sel[f] * D + d);
        }
        __syncthreads();
    }
    if (res_out)
        for (int e = tid; e < FR * D; e += THREADS) {
            const int f = e / D, d = e % D;
            if (r0 + f < rows) res_out[(r0 + f) * D + d] = Rs[f * (D + 1) + d];
        }
}

template <int VEC>
__global__ void rvq_decode_f32_kernel(const int64_t* __restrict__ codes, const float* __restrict__ cbs,
                                      float* __restrict__ out, int64_t rows, int dim, int n_codes, int stages,
                                      int code_stride, int code_offset, int* err_flag,
                                      __nv_bfloat16* __restrict__ out_bf = nullptr, int rows_per_clip = 1, long long bstride = 0,
                                      __nv_bfloat16* __restrict__ out_lo = nullptr, int out_f16 = 0) {
    // one thread per VEC consecutive output channels; consecutive threads cover one row contiguously
    const int per_row = dim / VEC;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t row = gid / per_row;
    if (row >= rows) return;
    const int d0 = (int)(gid % per_row) * VEC;
    float acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
    for (int k = 0; k < stages; ++k) {
        const int64_t c = __ldg(codes + row * code_stride + code_offset + k);
        if (c < 0 || c >= n_codes) {
            if (err_flag) atomicExch(err_flag, 1);
            continue;
        }
        const float* e = cbs + ((size_t)k * n_codes + c) * dim + d0;
        if (VEC == 4) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(e));
            acc[0] += q.x; acc[1] += q.y; acc[2] += q.z; acc[3] += q.w;
        } else {
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[v] += __ldg(e + v);
        }
    }
    if (out_bf) {
        const long long off = (row / rows_per_clip) * bstride + (row % rows_per_clip) * (long long)dim + d0;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            float hv;
            if (out_f16) {  // fp16 hi plane (saturating), bf16 lo plane
                const __half hh = __float2half_rn(fminf(fmaxf(acc[v], -65504.f), 65504.f));
                reinterpret_cast<__half*>(out_bf)[off + v] = hh;
                hv = __half2float(hh);
            } else {
                const __nv_bfloat16 hb = __float2bfloat16(acc[v]);
                out_bf[off + v] = hb;
                hv = __bfloat162float(hb);
            }
            if (out_lo) out_lo[off + v] = __float2bfloat16(acc[v] - hv);
        }
    }
    if (!out) return;
    float* o = out + row * dim + d0;
    if (VEC == 4) *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    else
#pragma unroll
        for (int v = 0; v < VEC; ++v) o[v] = acc[v];
}

template <int D>
int launch_encode(const float* x, const float* cbs, const float* cbn, int64_t* codes_out, float* res_out,
                  int64_t rows, int n_codes, int stages, int code_stride, int code_offset, int metric,
                  cudaStream_t s) {
    const size_t smem = (size_t)(FR * (D + 1) + CC * (D + 1) + FR * 32) * 4 + FR * 32 * 4 + FR * 8;
    cudaError_t e = cudaFuncSetAttribute(rvq_encode_f32_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { ac::set_error("ac_rvq_encode_f32: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
    const int64_t grid = (rows + FR - 1) / FR;
    rvq_encode_f32_kernel<D><<<(unsigned)grid, THREADS, smem, s>>>(x, cbs, cbn, codes_out, res_out, rows, n_codes,
                                                                   stages, code_stride, code_offset, metric);
    return ac::finish_launch("ac_rvq_encode_f32");
}

}  // namespace

extern "C" int ac_rvq_encode_f32(const float* x, const float* codebooks, const float* cb_norm, int64_t* codes_out,
                                 float* residual_out, int64_t rows, int32_t dim, int32_t codes, int32_t stages,
                                 int32_t code_stride, int32_t code_offset, int32_t metric, void* stream) {
    AC_REQUIRE(x && codebooks && cb_norm && codes_out, "ac_rvq_encode_f32: null pointer");
    AC_REQUIRE(rows > 0 && stages > 0 && codes > 0, "ac_rvq_encode_f32: empty problem");
    AC_REQUIRE(metric == 0 || metric == 1, "ac_rvq_encode_f32: metric %d", metric);
    cudaStream_t s = (cudaStream_t)stream;
    switch (dim) {
        case 128: return launch_encode<128>(x, codebooks, cb_norm, codes_out, residual_out, rows, codes, stages, code_stride, code_offset, metric, s);
        case 256: return launch_encode<256>(x, codebooks, cb_norm, codes_out, residual_out, rows, codes, stages, code_stride, code_offset, metric, s);
        case 8:   return launch_encode<8>(x, codebooks, cb_norm, codes_out, residual_out, rows, codes, stages, code_stride, code_offset, metric, s);
        default: ac::set_error("ac_rvq_encode_f32: unsupported dim %d (8/128/256)", dim); return -2;
    }
}

extern "C" int ac_rvq_decode_f32(const int64_t* codes, const float* codebooks, float* out, int64_t rows, int32_t dim,
                                 int32_t n_codes, int32_t stages, int32_t code_stride, int32_t code_offset,
                                 int32_t* err_flag, void* stream) {
    AC_REQUIRE(codes && codebooks && out, "ac_rvq_decode_f32: null pointer");
    AC_REQUIRE(rows > 0 && stages > 0 && dim > 0, "ac_rvq_decode_f32: empty problem");
    cudaStream_t s = (cudaStream_t)stream;
    const int threads = 256;
    if (dim % 4 == 0) {
        const int64_t total = rows * (dim / 4);
        rvq_decode_f32_kernel<4><<<(unsigned)((total + threads - 1) / threads), threads, 0, s>>>(
            codes, codebooks, out, rows, dim, n_codes, stages, code_stride, code_offset, err_flag);
    } else {
        const int64_t total = rows * dim;
        rvq_decode_f32_kernel<1><<<(unsigned)((total + threads - 1) / threads), threads, 0, s>>>(
            codes, codebooks, out, rows, dim, n_codes, stages, code_stride, code_offset, err_flag);
    }
    return ac::finish_launch("ac_rvq_decode_f32");
}

extern "C" int ac_rvq_decode_bf16(const int64_t* codes, const float* codebooks, void* out_bf16, void* out_lo, int64_t rows,
                                  int32_t rows_per_clip, int64_t bstride, int32_t dim, int32_t n_codes, int32_t stages,
                                  int32_t code_stride, int32_t code_offset, int32_t* err_flag, int32_t out_f16, void* stream) {
    AC_REQUIRE(codes && codebooks && out_bf16, "ac_rvq_decode_bf16: null pointer");
    AC_REQUIRE(rows > 0 && stages > 0 && dim > 0 && dim % 4 == 0 && rows_per_clip > 0, "ac_rvq_decode_bf16: bad sizes");
    const int threads = 256;
    const int64_t total = rows * (dim / 4);
    rvq_decode_f32_kernel<4><<<(unsigned)((total + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(
        codes, codebooks, nullptr, rows, dim, n_codes, stages, code_stride, code_offset, err_flag,
        (__nv_bfloat16*)out_bf16, rows_per_clip, bstride, (__nv_bfloat16*)out_lo, out_f16);
    return ac::finish_launch("ac_rvq_decode_bf16");
}
