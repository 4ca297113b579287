// Halo fill for channels-last bf16 activation buffers (ac_pad_halo_bf16).  The reference materialises a
// whole padded copy of every activation (F.pad reflect, HF/encodec:139-162); here only the few halo rows
// next to the valid region are written, in place, and every consumer reads them through its TMA view.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace {
__global__ void pad_halo_bf16_kernel(__nv_bfloat16* data, __nv_bfloat16* data_lo, int rows, int ch, long long bstride, int halo_l,
                                     int halo_r, int mode, int reflect_len) {
    const int b = blockIdx.y;
    __nv_bfloat16* base = (blockIdx.z ? data_lo : data) + (long long)b * bstride;   // z = 1: the lo plane of the same tensor
    const int total = (halo_l + halo_r) * ch;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int hr = e / ch, c = e % ch;
        const int pos = hr < halo_l ? hr - halo_l : rows + (hr - halo_l);  // row index relative to valid row 0
        int src = -1;
        if (mode == AC_PAD_REPLICATE) src = pos < 0 ? 0 : rows - 1;
        else if (mode == AC_PAD_REFLECT) {
            int q = pos < 0 ? -pos : pos;
            if (q >= reflect_len) q = 2 * (reflect_len - 1) - q;
            src = (q >= 0 && q < rows) ? q : -1;
        }
        base[(long long)pos * ch + c] = src >= 0 ? base[(long long)src * ch + c] : __float2bfloat16(0.f);
    }
}

// out = act(a + b) on split-bf16 tensors (hi [+lo] planes), 8 elements per thread.
__global__ void add_act_bf16_kernel(const __nv_bfloat16* a_hi, const __nv_bfloat16* a_lo, const __nv_bfloat16* b_hi,
                                    const __nv_bfloat16* b_lo, __nv_bfloat16* o_hi, __nv_bfloat16* o_lo, long long per_clip,
                                    long long a_bs, long long b_bs, long long o_bs, int act, int fmt) {
    const bool a_f16 = fmt & 1, b_f16 = fmt & 2, o_f16 = fmt & 4;  // hi-plane formats (lo planes are bf16)
    const int clip = blockIdx.y;
    for (long long e = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8; e < per_clip; e += (long long)gridDim.x * blockDim.x * 8) {
        float v[8];
        auto acc = [&](const __nv_bfloat16* ptr, bool first, bool f16) {
            const uint4 r = *reinterpret_cast<const uint4*>(ptr);
            const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 f = tcc::unpack16(w[i], f16);
                v[2 * i] = first ? f.x : v[2 * i] + f.x;
                v[2 * i + 1] = first ? f.y : v[2 * i + 1] + f.y;
            }
        };
        acc(a_hi + clip * a_bs + e, true, a_f16);
        if (a_lo) acc(a_lo + clip * a_bs + e, false, false);
        acc(b_hi + clip * b_bs + e, false, b_f16);
        if (b_lo) acc(b_lo + clip * b_bs + e, false, false);
        if (act == AC_ACT_ELU) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = ac::elu_fast(v[i]);
        }
        const uint4 q = tcc::pack8(v, o_f16);
        *reinterpret_cast<uint4*>(o_hi + clip * o_bs + e) = q;
        if (o_lo) *reinterpret_cast<uint4*>(o_lo + clip * o_bs + e) = tcc::pack_lo(v, q, o_f16);
    }
}

// fp32 -> split bf16 planes (hi = bf16(x), lo = bf16(x - hi)), 8 elements per thread
__global__ void f32_to_split_kernel(const float* x, __nv_bfloat16* o_hi, __nv_bfloat16* o_lo, long long per_clip, long long x_bs,
                                    long long o_bs, int out_f16) {
    const int clip = blockIdx.y;
    for (long long e = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8; e < per_clip; e += (long long)gridDim.x * blockDim.x * 8) {
        const float4 a = *reinterpret_cast<const float4*>(x + clip * x_bs + e);
        const float4 b = *reinterpret_cast<const float4*>(x + clip * x_bs + e + 4);
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        const uint4 q = tcc::pack8(v, out_f16 != 0);
        *reinterpret_cast<uint4*>(o_hi + clip * o_bs + e) = q;
        if (o_lo) *reinterpret_cast<uint4*>(o_lo + clip * o_bs + e) = tcc::pack_lo(v, q, out_f16 != 0);
    }
}
}  // namespace

extern "C" int ac_f32_to_split_bf16(const float* x, void* out_hi, void* out_lo, int32_t batch, int64_t per_clip, int64_t x_bstride,
                                    int64_t out_bstride, int32_t out_f16, void* stream) {
    AC_REQUIRE(x && out_hi && batch > 0 && batch <= 65535 && per_clip > 0 && per_clip % 8 == 0 && x_bstride % 8 == 0 &&
                   out_bstride % 8 == 0, "ac_f32_to_split_bf16: bad arguments");
    long long blocks = (per_clip / 8 + 255) / 256;
    if (blocks > 1024) blocks = 1024;
    f32_to_split_kernel<<<dim3((unsigned)blocks, batch), 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo,
                                                                                        per_clip, x_bstride, out_bstride, out_f16);
    return ac::finish_launch("ac_f32_to_split_bf16");
}

extern "C" int ac_add_act_bf16(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, void* out_hi, void* out_lo,
                               int32_t batch, int64_t per_clip, int64_t a_bstride, int64_t b_bstride, int64_t out_bstride,
                               int32_t act, int32_t fmt, void* stream) {
    AC_REQUIRE(a_hi && b_hi && out_hi && batch > 0 && batch <= 65535 && per_clip > 0 && per_clip % 8 == 0 && a_bstride % 8 == 0 &&
                   b_bstride % 8 == 0 && out_bstride % 8 == 0, "ac_add_act_bf16: bad arguments");
    AC_REQUIRE(act == AC_ACT_NONE || act == AC_ACT_ELU, "ac_add_act_bf16: act %d", act);
    long long blocks = (per_clip / 8 + 255) / 256;
    if (blocks > 1024) blocks = 1024;
    add_act_bf16_kernel<<<dim3((unsigned)blocks, batch), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)a_hi, (const __nv_bfloat16*)a_lo, (const __nv_bfloat16*)b_hi, (const __nv_bfloat16*)b_lo,
        (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo, per_clip, a_bstride, b_bstride, out_bstride, act, fmt);
    return ac::finish_launch("ac_add_act_bf16");
}

extern "C" int ac_pad_halo2_bf16(void* data, void* data_lo, int32_t batch, int32_t rows, int32_t ch, int64_t batch_stride,
                                 int32_t halo_l, int32_t halo_r, int32_t mode, int32_t reflect_len, void* stream) {
    AC_REQUIRE(data && batch > 0 && batch <= 65535 && rows > 0 && ch > 0, "ac_pad_halo_bf16: bad arguments");
    if (halo_l + halo_r <= 0) return 0;
    const int total = (halo_l + halo_r) * ch;
    dim3 grid((total + 255) / 256 > 64 ? 64 : (total + 255) / 256, batch, data_lo ? 2 : 1);
    pad_halo_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)data, (__nv_bfloat16*)data_lo, rows, ch, batch_stride,
                                                                  halo_l, halo_r, mode, reflect_len < rows ? rows : reflect_len);
    return ac::finish_launch("ac_pad_halo_bf16");
}

extern "C" int ac_pad_halo_bf16(void* data, int32_t batch, int32_t rows, int32_t ch, int64_t batch_stride,
                                int32_t halo_l, int32_t halo_r, int32_t mode, int32_t reflect_len, void* stream) {
    return ac_pad_halo2_bf16(data, nullptr, batch, rows, ch, batch_stride, halo_l, halo_r, mode, reflect_len, stream);
}
