// Shared helpers for the audiocodecs_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/audiocodecs_b200.h"

namespace ac {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

// Call after every kernel launch: records the launch and converts a launch error into a return code.
inline int finish_launch(const char* what) {
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

#define AC_REQUIRE(cond, ...)              \
    do {                                   \
        if (!(cond)) {                     \
            ac::set_error(__VA_ARGS__);    \
            return -1;                     \
        }                                  \
    } while (0)

__device__ __forceinline__ float elu1(float x) { return x > 0.f ? x : expm1f(x); }
// ELU for the tensor-path epilogues: exp via MUFU (+ a 5-term series near 0 where exp(x)-1 cancels); abs error < 3e-7,
// far below the bf16 (2^-9) / split-bf16 (2^-17) rounding that follows.  expm1f costs ~4x more issue slots.
__device__ __forceinline__ float elu_fast(float x) {
    const float series = x * (1.0f + x * (0.5f + x * (0.16666667f + x * (0.041666668f + x * 0.0083333338f))));
    const float big = __expf(x) - 1.0f;
    const float neg = x > -0.0625f ? series : big;
    return x > 0.f ? x : neg;
}
__device__ __forceinline__ float snake(float x, float a) {
    float s = sinf(a * x);
    return x + (1.0f / (a + 1e-9f)) * (s * s);
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

}  // namespace ac
