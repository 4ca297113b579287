// Helpers shared by the tcgen05 tap-GEMM kernels (conv_tc.cu, resunit_tc.cu): bf16 packing, 256-bit epilogue stores,
// activation math, and the host-side tensor-map plumbing.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace tcc {

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// Operand planes come in two 16-bit formats.  The "hi" plane of an activation / weight is bf16 (8 significant bits, fp32's
// range) or IEEE fp16 (11 significant bits; values beyond +-65504 saturate instead of becoming inf, tiny ones go subnormal
// with 2^-25 absolute error); the optional "lo" plane is always bf16(v - float(hi)).  One fp16 product carries 2^-12
// operand rounding -- better than the error-compensated bf16 pair-of-products at half the tensor work -- and the
// (fp16 hi, bf16 lo) pair carries ~2^-21.  `f16` selects the hi-plane format.
__device__ __forceinline__ uint32_t pack_f16(float a, float b) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));  // first source operand -> upper half
    return r;
}
__device__ __forceinline__ uint32_t pack16(float a, float b, bool f16) { return f16 ? pack_f16(a, b) : pack_bf16(a, b); }
__device__ __forceinline__ float2 unpack16(uint32_t v, bool f16) {
    if (f16) return __half22float2(*reinterpret_cast<const __half2*>(&v));
    const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&v);
    return make_float2(__low2float(h2), __high2float(h2));
}

__device__ __forceinline__ void add_bf16x8(float (&o)[8], const __nv_bfloat16* ptr, bool f16 = false) {
    const uint4 r = *reinterpret_cast<const uint4*>(ptr);
    const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = unpack16(rr[i], f16);
        o[2 * i] += f.x;
        o[2 * i + 1] += f.y;
    }
}

// lo plane of 8 values: bf16(v - float(hi)) where hi is the already-packed 16-bit rounding of v
__device__ __forceinline__ uint4 pack_lo(const float (&v)[8], const uint4& hi, bool f16 = false) {
    const uint32_t h[4] = {hi.x, hi.y, hi.z, hi.w};
    uint32_t r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = unpack16(h[i], f16);
        r[i] = pack_bf16(v[2 * i] - f.x, v[2 * i + 1] - f.y);
    }
    return make_uint4(r[0], r[1], r[2], r[3]);
}

__device__ __forceinline__ void st_global_256(void* ptr, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e, uint32_t f,
                                              uint32_t g, uint32_t h) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(ptr), "r"(a), "r"(b), "r"(c), "r"(d), "r"(e), "r"(f),
                 "r"(g), "r"(h) : "memory");
}

__device__ __forceinline__ void add_bf16x16(float (&o)[16], const __nv_bfloat16* ptr, bool f16 = false) {
    const uint4 r0 = reinterpret_cast<const uint4*>(ptr)[0], r1 = reinterpret_cast<const uint4*>(ptr)[1];
    const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float2 f = unpack16(rr[i], f16);
        o[2 * i] += f.x;
        o[2 * i + 1] += f.y;
    }
}

__device__ __forceinline__ void add_bf16x16(float (&o)[16], const uint4& r0, const uint4& r1, bool f16 = false) {
    const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float2 f = unpack16(rr[i], f16);
        o[2 * i] += f.x;
        o[2 * i + 1] += f.y;
    }
}

// 16 values -> one 32-byte store of their 16-bit roundings (+ one of the rounding residuals, bf16, when `lo` is given)
__device__ __forceinline__ void store_bf16x16(const float (&o)[16], __nv_bfloat16* hi, __nv_bfloat16* lo, bool f16 = false) {
    uint32_t q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) q[i] = pack16(o[2 * i], o[2 * i + 1], f16);
    st_global_256(hi, q[0], q[1], q[2], q[3], q[4], q[5], q[6], q[7]);
    if (lo) {
        uint32_t r[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float2 f = unpack16(q[i], f16);
            r[i] = pack_bf16(o[2 * i] - f.x, o[2 * i + 1] - f.y);
        }
        st_global_256(lo, r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7]);
    }
}

__device__ __forceinline__ uint4 pack8(const float (&o)[8], bool f16 = false) {
    return make_uint4(pack16(o[0], o[1], f16), pack16(o[2], o[3], f16), pack16(o[4], o[5], f16), pack16(o[6], o[7], f16));
}

// ELU on the epilogue: exp through one bare MUFU.EX2 (ex2.approx.ftz: no range fix-up code -- results below 2^-126 flush
// to zero, i.e. ELU = -1 exactly as fp32 rounds it; exp2f() spent ~5 extra instructions per element on that range handling,
// profiles/r02 source page); branch-free, abs error ~1e-7 (cancellation in exp(x)-1 near 0), far below the 16-bit rounding
// that follows.
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float elu_ex2(float x) {
    const float e = ex2_approx(x * 1.4426950408889634f) - 1.0f;
    return x > 0.f ? x : e;
}


// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

inline CUtensorMapSwizzle swizzle_for(int bk) {
    return bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (bk == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

inline int sm_count() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    return sms;
}

inline uint32_t round_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }


}  // namespace tcc
