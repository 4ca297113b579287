// Probe (B200) for the LSTM recurrence redesign:
//  (1) tcgen05.mma with the A operand in TENSOR MEMORY (written by tcgen05.st, lane = row, two bf16 per 32-bit column)
//      and the B operand in shared memory in the un-swizzled K-major core-matrix layout; which of the descriptor's
//      two byte offsets is the K-direction stride is determined empirically (variant 0: LBO = K stride; 1: SBO = K stride).
//  (2) cp.async.bulk shared::cta -> shared::cluster with complete_tx on the REMOTE CTA's mbarrier (cluster of 2).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -I../audiocodecs_b200/csrc probe_tmem_a.cu -lcuda
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "sm100.cuh"

using namespace sm100;

constexpr int M = 128, N = 16, K = 64;

__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__global__ void __launch_bounds__(128) probe_ts(const __nv_bfloat16* A, const __nv_bfloat16* B, int variant, float* out) {
    __shared__ __align__(1024) uint8_t b_s[N * K * 2];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&slot, 64);
    // B -> smem, un-swizzled K-major: core matrix (8 rows x 16 B) at k_grp*KSTR + n_grp*NSTR
    const uint32_t KSTR = 256, NSTR = 128;
    for (int e = threadIdx.x; e < N * K; e += 128) {
        const int n = e / K, k = e % K;
        *reinterpret_cast<__nv_bfloat16*>(b_s + (k / 8) * KSTR + (n / 8) * NSTR + (n % 8) * 16 + (k % 8) * 2) = B[e];
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    // A -> TMEM columns [32, 64): thread = row, 32 words = 64 bf16
    {
        const int row = warp * 32 + lane;
        uint32_t w[32];
        for (int j = 0; j < 32; ++j) {
            const __nv_bfloat162 h = __halves2bfloat162(A[row * K + 2 * j], A[row * K + 2 * j + 1]);
            w[j] = *reinterpret_cast<const uint32_t*>(&h);
        }
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + 32;
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,"
            "%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
            ::"r"(taddr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "r"(w[8]), "r"(w[9]),
              "r"(w[10]), "r"(w[11]), "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15]), "r"(w[16]), "r"(w[17]), "r"(w[18]), "r"(w[19]),
              "r"(w[20]), "r"(w[21]), "r"(w[22]), "r"(w[23]), "r"(w[24]), "r"(w[25]), "r"(w[26]), "r"(w[27]), "r"(w[28]), "r"(w[29]),
              "r"(w[30]), "r"(w[31])
            : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 0) {
        uint64_t d = 0;
        d |= (uint64_t)((smem_u32(b_s) & 0x3FFFFu) >> 4);
        const uint32_t lbo = variant == 0 ? KSTR : NSTR, sbo = variant == 0 ? NSTR : KSTR;
        d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
        d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
        d |= 1ull << 46;
        const uint32_t idesc = make_idesc_bf16(M, N);
        for (int k = 0; k < K / 16; ++k)
            umma_bf16_ts(tmem, tmem + 32 + k * 8, d + ((2 * KSTR) >> 4) * k, idesc, k != 0);
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    uint32_t v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16), v);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) out[(warp * 32 + lane) * N + i] = __uint_as_float(v[i]);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 64);
}

// ---------------------------------------------------------------- (2) bulk DSMEM copy with remote complete_tx
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) probe_bulk(int* result) {
    __shared__ __align__(128) uint32_t buf[256];  // 1 KB
    __shared__ uint64_t bar;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    for (int i = threadIdx.x; i < 256; i += 128) buf[i] = rank == 0 ? 0xABC00000u + i : 0u;
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
    if (rank == 1 && threadIdx.x == 0) mbar_arrive_expect_tx(&bar, 1024);
    asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
    if (rank == 0 && threadIdx.x == 0) {
        fence_proxy_async();
        uint32_t dst, rbar;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(dst) : "r"(smem_u32(buf)), "r"(1));
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(smem_u32(&bar)), "r"(1));
        asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "r"(smem_u32(buf)), "r"(1024), "r"(rbar) : "memory");
    }
    if (rank == 1) {
        mbar_wait(&bar, 0);
        int bad = 0;
        for (int i = threadIdx.x; i < 256; i += 128) bad += buf[i] != 0xABC00000u + i;
        if (bad) atomicAdd(result, bad);
        if (threadIdx.x == 0) atomicAdd(result + 1, 1);
    }
    asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
}

int main() {
    std::vector<__nv_bfloat16> A(M * K), B(N * K);
    std::vector<float> Af(M * K), Bf(N * K);
    srand(7);
    for (int i = 0; i < M * K; ++i) { Af[i] = (float)(rand() % 7 - 3); A[i] = __float2bfloat16(Af[i]); }
    for (int i = 0; i < N * K; ++i) { Bf[i] = (float)(rand() % 5 - 2); B[i] = __float2bfloat16(Bf[i]); }
    __nv_bfloat16 *dA, *dB;
    float* d_out;
    cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&d_out, M * N * 4);
    cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
    for (int variant = 0; variant < 2; ++variant) {
        cudaMemset(d_out, 0, M * N * 4);
        probe_ts<<<1, 128>>>(dA, dB, variant, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("ts variant %d: CUDA error %s\n", variant, cudaGetErrorString(e)); return 1; }
        std::vector<float> out(M * N);
        cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < N; ++n) {
                float ref = 0;
                for (int k = 0; k < K; ++k) ref += Af[m * K + k] * Bf[n * K + k];
                if (ref != out[m * N + n]) ++bad;
            }
        printf("A-in-TMEM + unswizzled B, variant %d (%s = K stride): %s (%d wrong)\n", variant, variant == 0 ? "LBO" : "SBO",
               bad ? "BAD" : "ok", bad);
    }
    int* d_res;
    cudaMalloc(&d_res, 8);
    cudaMemset(d_res, 0, 8);
    probe_bulk<<<2, 128>>>(d_res);
    cudaError_t e = cudaDeviceSynchronize();
    int res[2] = {-1, -1};
    cudaMemcpy(res, d_res, 8, cudaMemcpyDeviceToHost);
    printf("bulk DSMEM copy with remote complete_tx: %s (err %s, mismatches %d, receiver done %d)\n",
           (e == cudaSuccess && res[0] == 0 && res[1] == 1) ? "ok" : "BAD", cudaGetErrorString(e), res[0], res[1]);
    return 0;
}
