// Polyphase windowed-sinc resampler (ac_resample_f32): the FIR the reference rebuilds and applies
// with F.conv1d on every call (TA/functional.py:1405-1432; R/audiocodecs/codec.py:59-63,95-99).
// HBM-bound: one thread per output sample, taps served from L1/L2 (<= 441 x 174 floats).
#include "common.cuh"

namespace {
__global__ void resample_f32_kernel(const float* __restrict__ x, const float* __restrict__ taps, float* __restrict__ y,
                                    int64_t in_len, int64_t out_len, int orig, int n_phase, int n_taps, int width) {
    const int b = blockIdx.y;
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= out_len) return;
    const int64_t f = o / n_phase;
    const int ph = (int)(o - f * n_phase);
    const float* xb = x + (int64_t)b * in_len;
    const float* tp = taps + (size_t)ph * n_taps;
    const int64_t start = f * orig - width;  // index into the unpadded signal of tap 0
    float acc = 0.f;
    for (int j = 0; j < n_taps; ++j) {
        const int64_t i = start + j;
        const float v = (i >= 0 && i < in_len) ? __ldg(xb + i) : 0.f;
        acc = fmaf(v, __ldg(tp + j), acc);
    }
    y[(int64_t)b * out_len + o] = acc;
}
}  // namespace

extern "C" int ac_resample_f32(const float* x, const float* taps, float* y, int32_t batch, int64_t in_len,
                               int64_t out_len, int32_t orig, int32_t n_phase, int32_t n_taps, int32_t width,
                               void* stream) {
    AC_REQUIRE(x && taps && y, "ac_resample_f32: null pointer");
    AC_REQUIRE(batch > 0 && batch <= 65535 && in_len > 0 && out_len > 0, "ac_resample_f32: empty problem");
    const int threads = 256;
    dim3 grid((unsigned)((out_len + threads - 1) / threads), batch);
    resample_f32_kernel<<<grid, threads, 0, (cudaStream_t)stream>>>(x, taps, y, in_len, out_len, orig, n_phase, n_taps, width);
    return ac::finish_launch("ac_resample_f32");
}
