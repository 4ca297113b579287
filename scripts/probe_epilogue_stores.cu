// Probe (B200), prepared for the next round: how fast can 16 epilogue warps per SM write a
// [rows][C] bf16 tensor when every lane owns one accumulator ROW (the tcgen05.ld 32x32b layout)?
//   mode 0  what the epilogues do today: each lane stores 32 bytes of its own row per item (st.global.v8.b32): a warp
//           instruction touches 32 different 128-byte lines -> ~16 L1 wavefronts per request (ncu on DAC's k1 convs: the LSU
//           wavefront pipe at 75 % is what bounds them at ~3.8 TB/s of DRAM traffic)
//   mode 1  the same data staged per warp in shared memory ([32 rows][64 cols] bf16 = 4 KB, 128B swizzle, two buffers) and
//           written by ONE cp.async.bulk.tensor (TMA) store per 4 items
//   mode 2  upper bound: the same bytes with lanes writing CONSECUTIVE 32-byte pieces (fully coalesced, wrong layout)
// Prints GB/s per mode; mode 1 also verifies the written tensor.  Rows x C sized like DAC's decoder block 1 (384 channels).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../audiocodecs_b200/csrc probe_epilogue_stores.cu -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "sm100.cuh"
#include "tc_common.cuh"

using namespace sm100;

constexpr int WARPS = 16, TILE_M = 128, COLS_PER_ITEM = 16, ITEMS_PER_BOX = 4;  // a TMA box = 32 rows x 64 columns (128 B rows)

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

// value written at (row, col): exactly representable in bf16
__device__ __host__ inline float val(long long row, int col) { return (float)((row * 7 + col * 3) % 251) - 125.f; }

__global__ void __launch_bounds__(WARPS * 32, 1)
store_probe(const __grid_constant__ CUtensorMap omap, __nv_bfloat16* out, long long rows, int C, int mode) {
    extern __shared__ __align__(1024) uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int quarter = warp & 3, slot = warp >> 2;          // as the epilogues: 4 warps per TMEM lane quarter
    uint8_t* stage = smem + (size_t)warp * 2 * 4096;           // two 4 KB buffers per warp
    const int chunks = C / COLS_PER_ITEM;
    const long long tiles = rows / TILE_M;
    int buf = 0;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long row = tile * TILE_M + quarter * 32 + lane;
        if (mode == 1) {
            // slot s takes the column boxes s, s+4, ...: four ADJACENT 16-column items fill one 64-column box
            for (int box = slot; box < chunks / ITEMS_PER_BOX; box += WARPS / 4) {
                bulk_wait_read<1>();  // the store that last read this buffer (two boxes ago) has drained it
                __syncwarp();
                uint8_t* dst = stage + buf * 4096;
#pragma unroll
                for (int it = 0; it < ITEMS_PER_BOX; ++it) {
                    const int col = (box * ITEMS_PER_BOX + it) * COLS_PER_ITEM;
                    uint32_t q[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) q[i] = tcc::pack_bf16(val(row, col + 2 * i), val(row, col + 2 * i + 1));
                    // row `lane` of the box, 16-byte units 2*it and 2*it+1, 128B swizzle (unit ^= row & 7)
                    uint8_t* rp = dst + lane * 128;
                    *reinterpret_cast<uint4*>(rp + (((2 * it) ^ (lane & 7)) << 4)) = make_uint4(q[0], q[1], q[2], q[3]);
                    *reinterpret_cast<uint4*>(rp + (((2 * it + 1) ^ (lane & 7)) << 4)) = make_uint4(q[4], q[5], q[6], q[7]);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&omap, dst, box * 64, (int)(tile * TILE_M + quarter * 32));
                    bulk_commit();
                }
                buf ^= 1;
            }
        } else {
            for (int item = slot; item < chunks; item += WARPS / 4) {
                const int col = item * COLS_PER_ITEM;
                uint32_t q[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) q[i] = tcc::pack_bf16(val(row, col + 2 * i), val(row, col + 2 * i + 1));
                __nv_bfloat16* p = mode == 0 ? out + row * C + col
                                             : out + ((tile * chunks + item) * 4 + quarter) * 512 + lane * 16;  // 1 KB contiguous per warp
                tcc::st_global_256(p, q[0], q[1], q[2], q[3], q[4], q[5], q[6], q[7]);
            }
        }
    }
    if (mode == 1) bulk_wait_read<0>();
}

int main() {
    const int C = 384;
    const long long rows = 64LL * 55104 / TILE_M * TILE_M;  // 3.5 M rows: 2.7 GB
    __nv_bfloat16* out;
    cudaMalloc(&out, (size_t)rows * C * 2);
    CUtensorMap omap;
    {
        cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)rows};
        cuuint64_t gstr[1] = {(cuuint64_t)C * 2};
        cuuint32_t box[2] = {64, 32};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = tcc::get_encode()(&omap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("tensor map failed: %d\n", (int)r); return 1; }
    }
    const size_t smem = 1024 + (size_t)WARPS * 2 * 4096;
    cudaFuncSetAttribute(store_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const char* names[3] = {"lane = row, 32 B per lane (today)", "smem-staged TMA store (32 x 64 boxes)", "coalesced upper bound"};
    for (int mode = 0; mode < 3; ++mode) {
        for (int rep = 0; rep < 2; ++rep) store_probe<<<148, WARPS * 32, smem>>>(omap, out, rows, C, mode);
        cudaEventRecord(e0);
        for (int rep = 0; rep < 5; ++rep) store_probe<<<148, WARPS * 32, smem>>>(omap, out, rows, C, mode);
        cudaEventRecord(e1);
        cudaError_t err = cudaDeviceSynchronize();
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("mode %d  %-42s %8.3f ms  %8.1f GB/s  (%s)\n", mode, names[mode], ms / 5, (double)rows * C * 2 / (ms / 5 * 1e-3) / 1e9,
               cudaGetErrorString(err));
        if (mode == 1) {
            std::vector<__nv_bfloat16> h((size_t)1024 * C);
            const long long r0 = rows - 1024;
            cudaMemcpy(h.data(), out + r0 * C, h.size() * 2, cudaMemcpyDeviceToHost);
            long long bad = 0;
            for (long long r = 0; r < 1024; ++r)
                for (int c = 0; c < C; ++c) bad += __bfloat162float(h[r * C + c]) != val(r0 + r, c);
            printf("        verification of the TMA-stored tensor (last 1024 rows): %lld mismatches\n", bad);
        }
    }
    return 0;
}
