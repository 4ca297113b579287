// Halo fill for channels-last bf16 activation buffers (ac_pad_halo_bf16).  The reference materialises a
// whole padded copy of every activation (F.pad reflect, HF/encodec:139-162); here only the few halo rows
// next to the valid region are written, in place, and every consumer reads them through its TMA view.
#include <cuda_bf16.h>

#include "common.cuh"

namespace {
__global__ void pad_halo_bf16_kernel(__nv_bfloat16* data, int rows, int ch, long long bstride, int halo_l, int halo_r,
                                     int mode, int reflect_len) {
    const int b = blockIdx.y;
    __nv_bfloat16* base = data + (long long)b * bstride;
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
}  // namespace

extern "C" int ac_pad_halo_bf16(void* data, int32_t batch, int32_t rows, int32_t ch, int64_t batch_stride,
                                int32_t halo_l, int32_t halo_r, int32_t mode, int32_t reflect_len, void* stream) {
    AC_REQUIRE(data && batch > 0 && batch <= 65535 && rows > 0 && ch > 0, "ac_pad_halo_bf16: bad arguments");
    if (halo_l + halo_r <= 0) return 0;
    const int total = (halo_l + halo_r) * ch;
    dim3 grid((total + 255) / 256 > 64 ? 64 : (total + 255) / 256, batch);
    pad_halo_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)data, rows, ch, batch_stride, halo_l, halo_r,
                                                                  mode, reflect_len < rows ? rows : reflect_len);
    return ac::finish_launch("ac_pad_halo_bf16");
}
