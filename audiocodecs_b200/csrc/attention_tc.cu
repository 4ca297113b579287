// Causal sliding-window attention with RoPE on tcgen05 tensor cores (ac_attention_tc in include/audiocodecs_b200.h).
// Replaces MimiAttention / MimiRotaryEmbedding on the bf16 tensor path (HF/mimi/modeling_mimi.py:645-736, 515-577).
//
// One CTA = (128 queries, head, clip), thread r = query row r = TMEM lane r.  Keys are walked in blocks of 64:
//   stage   K (RoPE applied) and V of the block from the fp32 qkv tensor into shared memory as split bf16 (hi + lo planes,
//           128-byte rows in the 128B-swizzle layout the MMA descriptors expect)
//   S       = Q K^T   three tcgen05.mma products per k-step (hi*hi + lo*hi + hi*lo: ~2^-17 relative, fp32 accumulate)
//   softmax online, in the exp2 domain (1/sqrt(D)*log2(e) folded into Q), one row per thread straight out of TMEM: no
//           shuffles; P is written back as split bf16 into the A-operand layout
//   PV      = P V     same three products; V stays in its natural [key][d] layout and is consumed as an MN-major B
//           operand, so nothing is transposed; the running output row lives in 64 fp32 registers (rescaled by
//           exp2(m_old - m_new) per block)
// The result is written as the split-bf16 activation the output projection GEMM reads (and/or fp32).
// Two CTAs fit per SM (97 KB of shared memory, 128 TMEM columns each), which hides the serial stage->MMA->softmax chain.
#include <cuda_bf16.h>

#include "common.cuh"
#include "sm100.cuh"
#include "tc_common.cuh"

namespace {

using namespace sm100;

constexpr int D = 64;      // head dim
constexpr int QT = 128;    // queries per CTA (UMMA M)
constexpr int KB = 64;     // keys per block (UMMA N of S, K of PV)
constexpr int THREADS = 128;
constexpr uint32_t Q_PLANE = QT * 128, KV_PLANE = KB * 128, P_PLANE = QT * 128;
constexpr uint32_t SMEM_TILES = 2 * Q_PLANE + 4 * KV_PLANE + 2 * P_PLANE;  // 96 KB
constexpr uint32_t TMEM_COLS = 128;  // S: [0,64), PV: [64,128)

struct AttnParams {
    const float* qkv;    // [B][T][3*H*D]
    const float* rope;   // [T][D]: cos (D/2) | sin (D/2)
    float* out32;        // [B][T][H*D] or null
    __nv_bfloat16* out_hi;  // [B] x out_bs + [T][H*D], or null
    __nv_bfloat16* out_lo;
    int out_f16;            // hi plane written as fp16 instead of bf16
    long long out_bs;
    int T, H, window;
    float qscale;        // 1/sqrt(D) * log2(e)
};

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// byte offset of bf16 element `el` of row `row` in a [rows][64] plane with 128-byte rows and the 128B swizzle
__device__ __forceinline__ uint32_t sw_off(int row, int el) {
    return (uint32_t)row * 128u + ((((uint32_t)el >> 3) ^ ((uint32_t)row & 7u)) << 4) + ((uint32_t)el & 7u) * 2u;
}

__device__ __forceinline__ void store4_split(uint8_t* hi, uint8_t* lo, uint32_t off, float a, float b, float c, float d) {
    const uint32_t h0 = tcc::pack_bf16(a, b), h1 = tcc::pack_bf16(c, d);
    const __nv_bfloat162 x0 = *reinterpret_cast<const __nv_bfloat162*>(&h0), x1 = *reinterpret_cast<const __nv_bfloat162*>(&h1);
    *reinterpret_cast<uint2*>(hi + off) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(lo + off) =
        make_uint2(tcc::pack_bf16(a - __low2float(x0), b - __high2float(x0)), tcc::pack_bf16(c - __low2float(x1), d - __high2float(x1)));
}

// rows [pos0, pos0+rows) of q or k (column offset `col` of the fused tensor) with RoPE -> split planes; rows outside
// [0, pos_end) become zeros
__device__ __forceinline__ void stage_rope(const AttnParams& p, const float* base, size_t row_stride, int col, int pos0, int rows,
                                           int pos_end, float scale, uint8_t* hi, uint8_t* lo) {
    for (int e = threadIdx.x; e < rows * 8; e += THREADS) {
        const int j = e >> 3, c = e & 7, pos = pos0 + j;
        float4 r1 = make_float4(0.f, 0.f, 0.f, 0.f), r2 = r1;
        if (pos >= 0 && pos < pos_end) {
            const float* xr = base + (size_t)pos * row_stride + col;
            const float4 x1 = __ldg(reinterpret_cast<const float4*>(xr) + c), x2 = __ldg(reinterpret_cast<const float4*>(xr + D / 2) + c);
            const float* rr = p.rope + (size_t)pos * D;
            const float4 cs = __ldg(reinterpret_cast<const float4*>(rr) + c), sn = __ldg(reinterpret_cast<const float4*>(rr + D / 2) + c);
            // x*cos + rotate_half(x)*sin: first half -x2*sin, second half +x1*sin (HF/mimi:560-577)
            r1 = make_float4((x1.x * cs.x - x2.x * sn.x) * scale, (x1.y * cs.y - x2.y * sn.y) * scale,
                             (x1.z * cs.z - x2.z * sn.z) * scale, (x1.w * cs.w - x2.w * sn.w) * scale);
            r2 = make_float4((x2.x * cs.x + x1.x * sn.x) * scale, (x2.y * cs.y + x1.y * sn.y) * scale,
                             (x2.z * cs.z + x1.z * sn.z) * scale, (x2.w * cs.w + x1.w * sn.w) * scale);
        }
        store4_split(hi, lo, sw_off(j, 4 * c), r1.x, r1.y, r1.z, r1.w);
        store4_split(hi, lo, sw_off(j, D / 2 + 4 * c), r2.x, r2.y, r2.z, r2.w);
    }
}

__device__ __forceinline__ void stage_plain(const float* base, size_t row_stride, int col, int pos0, int rows, int pos_end, uint8_t* hi,
                                            uint8_t* lo) {
    for (int e = threadIdx.x; e < rows * 16; e += THREADS) {
        const int j = e >> 4, c = e & 15, pos = pos0 + j;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (pos >= 0 && pos < pos_end) x = __ldg(reinterpret_cast<const float4*>(base + (size_t)pos * row_stride + col) + c);
        store4_split(hi, lo, sw_off(j, 4 * c), x.x, x.y, x.z, x.w);
    }
}

__global__ void __launch_bounds__(THREADS, 2)
attention_tc_kernel(const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align_smem_1024(smem_raw);
    uint8_t* q_hi = smem;
    uint8_t* q_lo = q_hi + Q_PLANE;
    uint8_t* k_hi = q_lo + Q_PLANE;
    uint8_t* k_lo = k_hi + KV_PLANE;
    uint8_t* v_hi = k_lo + KV_PLANE;
    uint8_t* v_lo = v_hi + KV_PLANE;
    uint8_t* p_hi = v_lo + KV_PLANE;
    uint8_t* p_lo = p_hi + P_PLANE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(p_lo + P_PLANE);  // [0] S complete, [1] PV complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

    const int r = threadIdx.x, warp = r >> 5;
    const int q0 = blockIdx.x * QT, h = blockIdx.y, b = blockIdx.z;
    const int T = p.T;
    const size_t row_stride = (size_t)3 * p.H * D;
    const float* base = p.qkv + (size_t)b * T * row_stride;
    const int key_lo = max(0, q0 - p.window + 1);
    const int key_hi = min(T, q0 + QT);  // exclusive
    const int nblk = (key_hi - key_lo + KB - 1) / KB;

    if (r == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, TMEM_COLS);
    stage_rope(p, base, row_stride, h * D, q0, QT, T, p.qscale, q_hi, q_lo);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);

    const uint32_t idesc_s = make_idesc_bf16(QT, KB);
    const uint32_t idesc_pv = make_idesc_bf16(QT, D) | (1u << 16);  // B (= V, [key][d]) is MN-major
    const uint64_t dq_hi = make_smem_desc(smem_u32(q_hi), 128), dq_lo = make_smem_desc(smem_u32(q_lo), 128);
    const uint64_t dk_hi = make_smem_desc(smem_u32(k_hi), 128), dk_lo = make_smem_desc(smem_u32(k_lo), 128);
    const uint64_t dp_hi = make_smem_desc(smem_u32(p_hi), 128), dp_lo = make_smem_desc(smem_u32(p_lo), 128);
    // MN-major 128B-swizzle operand: 64 d-values (128 B) contiguous per key, 8-key groups 1024 B apart (stride byte
    // offset); one 64-wide atom along N so the leading offset is never used
    const uint64_t dv_hi = make_smem_desc(smem_u32(v_hi), 128), dv_lo = make_smem_desc(smem_u32(v_lo), 128);

    float o[D];
#pragma unroll
    for (int i = 0; i < D; ++i) o[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    const int qp = q0 + r;

    for (int blk = 0; blk < nblk; ++blk) {
        const int kb = key_lo + blk * KB;
        stage_rope(p, base, row_stride, (p.H + h) * D, kb, KB, key_hi, 1.0f, k_hi, k_lo);
        stage_plain(base, row_stride, (2 * p.H + h) * D, kb, KB, key_hi, v_hi, v_lo);
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (warp == 0) {
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < D / 16; ++ks) {
                    umma_bf16(tmem_base, dq_hi + 2 * ks, dk_hi + 2 * ks, idesc_s, ks != 0);
                    umma_bf16(tmem_base, dq_lo + 2 * ks, dk_hi + 2 * ks, idesc_s, 1);
                    umma_bf16(tmem_base, dq_hi + 2 * ks, dk_lo + 2 * ks, idesc_s, 1);
                }
                umma_commit(&bars[0]);
            }
            __syncwarp();
        }
        mbar_wait(&bars[0], blk & 1);
        tc_fence_after();
        uint32_t sv[KB];
#pragma unroll
        for (int c = 0; c < KB / 16; ++c) tmem_ld16(lane_addr + c * 16, *reinterpret_cast<uint32_t(*)[16]>(&sv[c * 16]));
        tmem_ld_wait();
        // causal + sliding window + end of sequence (HF/mimi:1096-1102: key <= query and key > query - window)
        float mx = m_run;
#pragma unroll
        for (int c = 0; c < KB; ++c) {
            const int kp = kb + c;
            const bool ok = kp <= qp && kp > qp - p.window && kp < key_hi;
            const float s = ok ? __uint_as_float(sv[c]) : -INFINITY;
            sv[c] = __float_as_uint(s);
            mx = fmaxf(mx, s);
        }
        const float m_safe = mx == -INFINITY ? 0.f : mx;
        const float alpha = ex2(m_run - m_safe);
        m_run = mx;
        float sum = 0.f;
#pragma unroll
        for (int c8 = 0; c8 < KB / 8; ++c8) {
            float pv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                pv[i] = ex2(__uint_as_float(sv[c8 * 8 + i]) - m_safe);
                sum += pv[i];
            }
            const uint4 hi = tcc::pack8(pv);
            const uint32_t off = (uint32_t)r * 128u + (((uint32_t)c8 ^ ((uint32_t)r & 7u)) << 4);
            *reinterpret_cast<uint4*>(p_hi + off) = hi;
            *reinterpret_cast<uint4*>(p_lo + off) = tcc::pack_lo(pv, hi);
        }
        l_run = l_run * alpha + sum;
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (warp == 0) {
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < KB / 16; ++ks) {
                    const uint64_t vstep = (uint64_t)((16u * 128u) >> 4) * ks;  // 16 keys = two 8-key groups further
                    umma_bf16(tmem_base + KB, dp_hi + 2 * ks, dv_hi + vstep, idesc_pv, ks != 0);
                    umma_bf16(tmem_base + KB, dp_lo + 2 * ks, dv_hi + vstep, idesc_pv, 1);
                    umma_bf16(tmem_base + KB, dp_hi + 2 * ks, dv_lo + vstep, idesc_pv, 1);
                }
                umma_commit(&bars[1]);
            }
            __syncwarp();
        }
        mbar_wait(&bars[1], blk & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < D / 16; ++c) {
            uint32_t ov[16];
            tmem_ld16(lane_addr + KB + c * 16, ov);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[c * 16 + i] = o[c * 16 + i] * alpha + __uint_as_float(ov[i]);
        }
        tc_fence_before();
    }

    if (qp < T) {
        const float inv = l_run > 0.f ? 1.0f / l_run : 0.f;
        const size_t col = (size_t)h * D;
        const int C = p.H * D;
#pragma unroll
        for (int c = 0; c < D / 16; ++c) {
            float y[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) y[i] = o[c * 16 + i] * inv;
            if (p.out_hi) {
                const size_t off = (size_t)b * p.out_bs + (size_t)qp * C + col + c * 16;
                tcc::store_bf16x16(y, p.out_hi + off, p.out_lo ? p.out_lo + off : nullptr, p.out_f16 != 0);
            }
            if (p.out32) {
                float4* dst = reinterpret_cast<float4*>(p.out32 + ((size_t)b * T + qp) * C + col + c * 16);
#pragma unroll
                for (int i = 0; i < 4; ++i) dst[i] = make_float4(y[4 * i], y[4 * i + 1], y[4 * i + 2], y[4 * i + 3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// LayerNorm over the channel axis with the result written as the split-bf16 activation the next GEMM reads (one warp per
// row, two-pass mean / variance in fp32 as ac_layernorm_f32; C <= 1024, a multiple of 128).
__global__ void __launch_bounds__(256)
layernorm_split_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                       __nv_bfloat16* __restrict__ o_hi, __nv_bfloat16* __restrict__ o_lo, long long rows, int rows_per_clip, int C,
                       long long out_bs, float eps, int out_f16) {
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const int nv = C / 128;  // float4 per lane
    const float4* xr = reinterpret_cast<const float4*>(x + row * C);
    float4 v[8];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (k < nv) { v[k] = __ldg(xr + lane + 32 * k); s += (v[k].x + v[k].y) + (v[k].z + v[k].w); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / C;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (k < nv) {
            const float d0 = v[k].x - mean, d1 = v[k].y - mean, d2 = v[k].z - mean, d3 = v[k].w - mean;
            q += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / C + eps);
    const long long off = (row / rows_per_clip) * out_bs + (row % rows_per_clip) * C;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (k < nv) {
            const int c = 4 * (lane + 32 * k);
            const float4 ww = __ldg(reinterpret_cast<const float4*>(w + c)), bb = __ldg(reinterpret_cast<const float4*>(b + c));
            const float y0 = (v[k].x - mean) * rstd * ww.x + bb.x, y1 = (v[k].y - mean) * rstd * ww.y + bb.y;
            const float y2 = (v[k].z - mean) * rstd * ww.z + bb.z, y3 = (v[k].w - mean) * rstd * ww.w + bb.w;
            const uint32_t h0 = tcc::pack16(y0, y1, out_f16 != 0), h1 = tcc::pack16(y2, y3, out_f16 != 0);
            *reinterpret_cast<uint2*>(o_hi + off + c) = make_uint2(h0, h1);
            if (o_lo) {
                const float2 a0 = tcc::unpack16(h0, out_f16 != 0), a1 = tcc::unpack16(h1, out_f16 != 0);
                *reinterpret_cast<uint2*>(o_lo + off + c) = make_uint2(tcc::pack_bf16(y0 - a0.x, y1 - a0.y), tcc::pack_bf16(y2 - a1.x, y3 - a1.y));
            }
        }
}

__global__ void rope_table_kernel(const float* __restrict__ inv_freq, float* __restrict__ table, int T, int half) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= T * half) return;
    const int pos = e / half, i = e % half;
    float sn, cs;
    sincosf((float)pos * inv_freq[i], &sn, &cs);
    table[(size_t)pos * 2 * half + i] = cs;
    table[(size_t)pos * 2 * half + half + i] = sn;
}

}  // namespace

extern "C" int ac_layernorm_split_bf16(const float* x, const float* w, const float* b, void* out_hi, void* out_lo, int32_t batch,
                                       int32_t rows_per_clip, int32_t C, int64_t out_bstride, float eps, int32_t out_f16, void* stream) {
    AC_REQUIRE(x && w && b && out_hi && batch > 0 && rows_per_clip > 0, "ac_layernorm_split_bf16: bad arguments");
    AC_REQUIRE(C > 0 && C <= 1024 && C % 128 == 0 && out_bstride % 4 == 0, "ac_layernorm_split_bf16: C %d (multiple of 128, <= 1024)", C);
    const long long rows = (long long)batch * rows_per_clip;
    layernorm_split_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
        x, w, b, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo, rows, rows_per_clip, C, out_bstride, eps, out_f16);
    return ac::finish_launch("ac_layernorm_split_bf16");
}

extern "C" int ac_rope_table_f32(const float* inv_freq, float* table, int32_t T, int32_t half, void* stream) {
    AC_REQUIRE(inv_freq && table && T > 0 && half > 0, "ac_rope_table_f32: bad arguments");
    rope_table_kernel<<<(T * half + 255) / 256, 256, 0, (cudaStream_t)stream>>>(inv_freq, table, T, half);
    return ac::finish_launch("ac_rope_table_f32");
}

extern "C" int ac_attention_tc(const float* qkv, const float* rope, float* out32, void* out_hi, void* out_lo, int64_t out_bstride,
                               int32_t batch, int32_t T, int32_t heads, int32_t head_dim, int32_t window, float scaling, int32_t out_f16,
                               void* stream) {
    AC_REQUIRE(qkv && rope && (out32 || out_hi), "ac_attention_tc: null pointer");
    AC_REQUIRE(out_hi || !out_lo, "ac_attention_tc: lo plane without hi plane");
    AC_REQUIRE(head_dim == D, "ac_attention_tc: head_dim %d (built for %d)", head_dim, D);
    AC_REQUIRE(batch > 0 && batch <= 65535 && T > 0 && heads > 0 && heads <= 65535 && window > 0, "ac_attention_tc: bad sizes");
    AC_REQUIRE(((heads * D) % 16) == 0 && (out_bstride % 16) == 0, "ac_attention_tc: rows must be 32-byte aligned");
    const size_t smem = 1024 + SMEM_TILES + 64;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { ac::set_error("ac_attention_tc: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        configured = true;
    }
    AttnParams p{};
    p.qkv = qkv; p.rope = rope; p.out32 = out32;
    p.out_hi = (__nv_bfloat16*)out_hi; p.out_lo = (__nv_bfloat16*)out_lo; p.out_bs = out_bstride; p.out_f16 = out_f16 ? 1 : 0;
    p.T = T; p.H = heads; p.window = window;
    p.qscale = scaling * 1.4426950408889634f;
    dim3 grid((T + QT - 1) / QT, heads, batch);
    attention_tc_kernel<<<grid, THREADS, smem, (cudaStream_t)stream>>>(p);
    return ac::finish_launch("ac_attention_tc");
}
