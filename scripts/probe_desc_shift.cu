// Probe (B200): can a tcgen05 shared-memory descriptor start at an arbitrary ROW of a TMA-written swizzled tile?
// A tile [256 rows x bk] bf16 is loaded once by TMA; the MMA reads rows j..j+127 through a descriptor whose start
// address is advanced by j rows, with the base-offset field either 0 or (addr >> 7) & 7.  Expected result is exact
// (small integers).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -I../audiocodecs_b200/csrc probe_desc_shift.cu -lcuda
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "sm100.cuh"

using namespace sm100;

constexpr int ROWS = 256, N = 16;

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap,
                                             int bk, int shift, int bo_mode, float* out) {
    extern __shared__ __align__(1024) uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a_s = smem;
    uint8_t* b_s = smem + 64 * 1024;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 80 * 1024);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(slot, 32);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar[0], (ROWS + N) * bk * 2);
        tma_load_2d(a_s, &amap, &bar[0], 0, 0);
        tma_load_2d(b_s, &bmap, &bar[0], 0, 0);
        mbar_wait(&bar[0], 0);
        tc_fence_after();
        const uint32_t sw = bk * 2;
        const uint32_t a_addr = smem_u32(a_s) + shift * sw;
        uint64_t adesc = make_smem_desc(a_addr, sw);
        if (bo_mode == 1) adesc |= (uint64_t)((a_addr >> 7) & 7) << 49;
        const uint64_t bdesc = make_smem_desc(smem_u32(b_s), sw);
        const uint32_t idesc = make_idesc_bf16(128, N);
        for (int k = 0; k < bk / 16; ++k) umma_bf16(tmem, adesc + 2 * k, bdesc + 2 * k, idesc, k != 0);
        umma_commit(&bar[1]);
    }
    mbar_wait(&bar[1], 0);
    tc_fence_after();
    uint32_t v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16), v);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) out[(warp * 32 + lane) * N + i] = __uint_as_float(v[i]);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 32);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
    EncodeTiledFn encode = reinterpret_cast<EncodeTiledFn>(ptr);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    float* d_out;
    cudaMalloc(&d_out, 128 * N * 4);
    for (int bk : {64, 32, 16}) {
        std::vector<__nv_bfloat16> A(ROWS * bk), B(N * bk);
        std::vector<float> Af(ROWS * bk), Bf(N * bk);
        srand(bk);
        for (int i = 0; i < ROWS * bk; ++i) { Af[i] = (float)(rand() % 7 - 3); A[i] = __float2bfloat16(Af[i]); }
        for (int i = 0; i < N * bk; ++i) { Bf[i] = (float)(rand() % 5 - 2); B[i] = __float2bfloat16(Bf[i]); }
        __nv_bfloat16 *dA, *dB;
        cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2);
        cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
        CUtensorMapSwizzle sw = bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (bk == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
        CUtensorMap am, bm;
        cuuint32_t est[2] = {1, 1};
        {
            cuuint64_t gd[2] = {(cuuint64_t)bk, ROWS}; cuuint64_t gs[1] = {(cuuint64_t)bk * 2}; cuuint32_t box[2] = {(cuuint32_t)bk, ROWS};
            CUresult r = encode(&am, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, gd, gs, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r) { printf("encode A failed %d\n", (int)r); return 1; }
        }
        {
            cuuint64_t gd[2] = {(cuuint64_t)bk, N}; cuuint64_t gs[1] = {(cuuint64_t)bk * 2}; cuuint32_t box[2] = {(cuuint32_t)bk, N};
            CUresult r = encode(&bm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, gd, gs, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r) { printf("encode B failed %d\n", (int)r); return 1; }
        }
        for (int bo = 0; bo < 2; ++bo) {
            printf("bk=%d base_offset_mode=%d :", bk, bo);
            for (int shift : {0, 1, 2, 3, 4, 5, 7, 8, 9, 18, 27, 54, 100}) {
                probe<<<1, 128, 100 * 1024>>>(am, bm, bk, shift, bo, d_out);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf(" shift %d: CUDA error %s\n", shift, cudaGetErrorString(e)); return 1; }
                std::vector<float> out(128 * N);
                cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost);
                int bad = 0;
                for (int m = 0; m < 128; ++m)
                    for (int n = 0; n < N; ++n) {
                        float ref = 0;
                        for (int k = 0; k < bk; ++k) ref += Af[(m + shift) * bk + k] * Bf[n * bk + k];
                        if (ref != out[m * N + n]) ++bad;
                    }
                printf(" %d:%s", shift, bad ? "BAD" : "ok");
            }
            printf("\n");
        }
        cudaFree(dA); cudaFree(dB);
    }
    return 0;
}
