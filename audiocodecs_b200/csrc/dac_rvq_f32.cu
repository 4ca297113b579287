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
            if ((unsigned)bi >= (unsigned)n_codes) bi = 0;  // all-NaN scores: code 0, gather in range
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

// from_codes: out[row][c] = sum_k (b_out[k][c] + sum_d Wo[k][c][d] * cb[k][code][d]), accumulated in stage order from 0.
// A [rows x 8K] x [8K x 1024] GEMM with gathered rows: a CTA takes DR rows x 256 channels; the rows' code vectors sit in
// shared memory (read as broadcasts), thread c keeps the DR running sums in registers and streams its 8 weights per stage
// once per CTA instead of once per row (the first version re-read the 288 KB out_proj tensor for every row: 16 GB of
// L2 traffic at 55 104 rows).  The per-element arithmetic (fma order d = 0..7 from the bias, stage sums added in order) is
// unchanged, so the result is bit-identical to it.
constexpr int DR = 32;
__global__ void __launch_bounds__(256)
dac_rvq_decode_kernel(const int64_t* __restrict__ codes, const float* __restrict__ cb, const float* __restrict__ w_out,
                      const float* __restrict__ b_out, float* __restrict__ out, int64_t rows, int n_codes, int stages,
                      int code_stride, int* err_flag) {
    extern __shared__ __align__(16) float zp[];  // [DR][stages][CD]
    const int64_t row0 = (int64_t)blockIdx.y * DR;
    const int tid = threadIdx.x;
    const int c = blockIdx.x * 256 + tid;
    for (int e = tid; e < DR * stages; e += 256) {
        const int r = e / stages, k = e % stages;
        float4 lo4 = make_float4(0.f, 0.f, 0.f, 0.f), hi4 = lo4;
        if (row0 + r < rows) {
            int64_t code = codes[(row0 + r) * code_stride + k];
            if (code < 0 || code >= n_codes) { if (err_flag) atomicExch(err_flag, 1); code = 0; }
            const float4* src = reinterpret_cast<const float4*>(cb + ((size_t)k * n_codes + code) * CD);
            lo4 = __ldg(src); hi4 = __ldg(src + 1);
        }
        reinterpret_cast<float4*>(zp)[e * 2] = lo4;
        reinterpret_cast<float4*>(zp)[e * 2 + 1] = hi4;
    }
    __syncthreads();
    float acc[DR];
#pragma unroll
    for (int r = 0; r < DR; ++r) acc[r] = 0.f;
    for (int k = 0; k < stages; ++k) {
        const float bias = __ldg(b_out + (size_t)k * HD + c);
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(w_out + ((size_t)k * HD + c) * CD));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(w_out + ((size_t)k * HD + c) * CD) + 1);
#pragma unroll
        for (int r = 0; r < DR; ++r) {
            const float4 a = reinterpret_cast<const float4*>(zp)[(r * stages + k) * 2];
            const float4 b = reinterpret_cast<const float4*>(zp)[(r * stages + k) * 2 + 1];
            float o = bias;
            o = fmaf(w0.x, a.x, o); o = fmaf(w0.y, a.y, o); o = fmaf(w0.z, a.z, o); o = fmaf(w0.w, a.w, o);
            o = fmaf(w1.x, b.x, o); o = fmaf(w1.y, b.y, o); o = fmaf(w1.z, b.z, o); o = fmaf(w1.w, b.w, o);
            acc[r] += o;
        }
    }
#pragma unroll
    for (int r = 0; r < DR; ++r)
        if (row0 + r < rows) out[(row0 + r) * HD + c] = acc[r];
}

// ---------------------------------------------------------------------------------------------------------------------
// Encode from PROJECTED latents (the bf16 tensor path).  in_proj is linear, so the 1024-wide residual never has to exist:
//   z_e[k] = in_proj_k(z - sum_{j<k} out_j) = P[k] + c[k] - sum_{j<k} M[k][j] zq[j]
// with P = all stages' in_proj of z at once (one [rows x 1024] x [1024 x 8S] tensor-core GEMM, bias included),
// M[k][j] = W_in_k W_out_j (8x8), c[k] = -sum_{j<k} W_in_k b_out_j, zq[j] the straight-through value of stage j.
// What remains per row is 8-dimensional: a warp takes 4 rows (lane = row r, dim d), scans the L2-normalised codebook
// with lane = code (each code's 32 bytes are read once for the 4 rows), keeps the reference's score formula
// -(|a|^2 - 2 a.b) + |b|^2 (HF/dac:165), first-index tie-break and STE arithmetic z_e + (z_q - z_e).
// The sums are associated differently from the 1024-wide chain, so this path is used where the latents are already
// approximate (precision="bf16"); the exact-order kernel above stays the fp32 path.
constexpr int PW = 8, RPW = 4;  // warps per CTA, rows per warp
__global__ void __launch_bounds__(PW * 32)
dac_rvq_encode_proj_kernel(const float* __restrict__ P, int ldp, const float* __restrict__ cconst, const float* __restrict__ M,
                           const float* __restrict__ cbn, const float* __restrict__ cbn2, const float* __restrict__ cb,
                           int64_t* __restrict__ codes, int64_t rows, int n_codes, int stages, int stages_total, int code_stride) {
    extern __shared__ __align__(16) float zq_s[];  // [PW][stages][32]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = lane >> 3, d = lane & 7;
    const int64_t row = ((int64_t)blockIdx.x * PW + warp) * RPW + r;
    const bool live = row < rows;
    float* zq_w = zq_s + (size_t)warp * stages * 32;
    for (int k = 0; k < stages; ++k) {
        float ze = (live ? __ldg(P + row * ldp + k * CD + d) : 0.f) + __ldg(cconst + k * CD + d);
        for (int j = 0; j < k; ++j) {
            const float* m = M + (((size_t)k * stages_total + j) * CD + d) * CD;
            const float* q = zq_w + j * 32 + (lane & ~7);
#pragma unroll
            for (int e = 0; e < CD; ++e) ze = fmaf(-__ldg(m + e), q[e], ze);
        }
        float n = ze * ze;
        n += __shfl_xor_sync(0xffffffffu, n, 1); n += __shfl_xor_sync(0xffffffffu, n, 2); n += __shfl_xor_sync(0xffffffffu, n, 4);
        const float a = ze * (1.0f / fmaxf(sqrtf(n), 1e-12f));  // F.normalize
        float an = a * a;
        an += __shfl_xor_sync(0xffffffffu, an, 1); an += __shfl_xor_sync(0xffffffffu, an, 2); an += __shfl_xor_sync(0xffffffffu, an, 4);
        float av[RPW][CD], anv[RPW], bv[RPW];
        int bi[RPW];
#pragma unroll
        for (int rr = 0; rr < RPW; ++rr) {
#pragma unroll
            for (int dd = 0; dd < CD; ++dd) av[rr][dd] = __shfl_sync(0xffffffffu, a, rr * 8 + dd);
            anv[rr] = __shfl_sync(0xffffffffu, an, rr * 8);
            bv[rr] = -INFINITY;
            bi[rr] = 0x7fffffff;
        }
        const float* Cn = cbn + (size_t)k * n_codes * CD;
        for (int c = lane; c < n_codes; c += 32) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(Cn + (size_t)c * CD));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(Cn + (size_t)c * CD) + 1);
            const float bn2 = __ldg(cbn2 + (size_t)k * n_codes + c);
#pragma unroll
            for (int rr = 0; rr < RPW; ++rr) {
                float dot = 0.f;
                dot = fmaf(av[rr][0], b0.x, dot); dot = fmaf(av[rr][1], b0.y, dot); dot = fmaf(av[rr][2], b0.z, dot);
                dot = fmaf(av[rr][3], b0.w, dot); dot = fmaf(av[rr][4], b1.x, dot); dot = fmaf(av[rr][5], b1.y, dot);
                dot = fmaf(av[rr][6], b1.z, dot); dot = fmaf(av[rr][7], b1.w, dot);
                const float score = -(anv[rr] - 2.f * dot) + bn2;
                if (score > bv[rr]) { bv[rr] = score; bi[rr] = c; }  // ascending c per lane: strict > keeps the first index
            }
        }
        int mine = 0;
#pragma unroll
        for (int rr = 0; rr < RPW; ++rr) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv[rr], o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi[rr], o);
                if (ov > bv[rr] || (ov == bv[rr] && oi < bi[rr])) { bv[rr] = ov; bi[rr] = oi; }
            }
            if (rr == r) mine = bi[rr];
        }
        if (mine >= n_codes) mine = 0;  // all-NaN scores: keep the gather in range
        const float zq = __ldg(cb + ((size_t)k * n_codes + mine) * CD + d);
        __syncwarp();
        zq_w[k * 32 + lane] = ze + (zq - ze);  // straight-through arithmetic kept in eval
        __syncwarp();
        if (live && d == 0) codes[row * code_stride + k] = (int64_t)mine;
    }
}

}  // namespace

extern "C" int ac_dac_rvq_encode_f32(const float* z, const float* w_in, const float* b_in, const float* codebooks,
                                     const float* w_out, const float* b_out, int64_t* codes, float* zq_out, int64_t rows,
                                     int32_t hidden, int32_t cb_dim, int32_t n_codes, int32_t stages, int32_t code_stride,
                                     void* stream) {
    AC_REQUIRE(z && w_in && b_in && codebooks && w_out && b_out && codes, "ac_dac_rvq_encode_f32: null pointer");
    AC_REQUIRE(hidden == HD && cb_dim == CD, "ac_dac_rvq_encode_f32: built for hidden %d / codebook dim %d", HD, CD);
    AC_REQUIRE(rows > 0 && stages > 0 && stages <= 32 && n_codes > 0, "ac_dac_rvq_encode_f32: bad sizes");
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
    AC_REQUIRE(rows > 0 && stages > 0 && stages <= 32, "ac_dac_rvq_decode_f32: bad sizes");
    const int64_t slice = (int64_t)65535 * DR;  // grid.y limit: launch in slices of at most 65535 row blocks
    for (int64_t r0 = 0; r0 < rows; r0 += slice) {
        const int64_t n = rows - r0 < slice ? rows - r0 : slice;
        dac_rvq_decode_kernel<<<dim3(HD / 256, (unsigned)((n + DR - 1) / DR)), 256, (size_t)DR * stages * CD * 4, (cudaStream_t)stream>>>(
            codes + r0 * code_stride, codebooks, w_out, b_out, out + r0 * HD, n, n_codes, stages, code_stride, err_flag);
    }
    return ac::finish_launch("ac_dac_rvq_decode_f32");
}

extern "C" int ac_dac_rvq_encode_proj_f32(const float* proj, int32_t ld_proj, const float* cconst, const float* cross,
                                          const float* cb_normed, const float* cb_norm2, const float* codebooks, int64_t* codes,
                                          int64_t rows, int32_t cb_dim, int32_t n_codes, int32_t stages, int32_t stages_total,
                                          int32_t code_stride, void* stream) {
    AC_REQUIRE(proj && cconst && cross && cb_normed && cb_norm2 && codebooks && codes, "ac_dac_rvq_encode_proj_f32: null pointer");
    AC_REQUIRE(cb_dim == CD, "ac_dac_rvq_encode_proj_f32: built for codebook dim %d", CD);
    AC_REQUIRE(rows > 0 && stages > 0 && stages <= stages_total && stages_total <= 32 && n_codes > 0 && ld_proj >= stages * CD,
               "ac_dac_rvq_encode_proj_f32: bad sizes");
    const int64_t per_cta = PW * RPW;
    dac_rvq_encode_proj_kernel<<<(unsigned)((rows + per_cta - 1) / per_cta), PW * 32, (size_t)PW * stages * 32 * 4, (cudaStream_t)stream>>>(
        proj, ld_proj, cconst, cross, cb_normed, cb_norm2, codebooks, codes, rows, n_codes, stages, stages_total, code_stride);
    return ac::finish_launch("ac_dac_rvq_encode_proj_f32");
}
