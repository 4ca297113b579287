// LSTM layer recurrence on tcgen05 tensor cores, one thread-block CLUSTER of 16 CTAs per 16 clips
// (ac_lstm_tc in include/audiocodecs_b200.h; replaces EncodecLSTM's nn.LSTM, HF/encodec:236-249).
//
// The 750-step chain is latency-bound, so everything a step needs stays on chip:
//   * CTA r of the cluster owns hidden units [32r, 32r+32): its 128 gate rows (i,f,g,o x 32 units) of W_hh
//     (bf16, 128 KB) are TMA-loaded once into shared memory as the A operand [128 x 512], K-major SW128.
//   * h[t-1] of the cluster's 16 clips is the B operand [16 x 512] (bf16, 16 KB, double-buffered by step
//     parity).  One step = 32 tcgen05.mma (M=128, N=16, K=16) into a 16-column TMEM accumulator:
//     gates^T[128 gate rows][16 clips].
//   * epilogue warp g (TMEM lane quarter g) holds gate g of 32 units: adds the hoisted input projection
//     (pre, prefetched one step ahead), applies sigmoid/tanh, the four gates meet through shared memory,
//     c stays in fp32 registers, h = o*tanh(c).
//   * the CTA's new h slice (16 clips x 32 units) is written straight into the B buffers of all 16 CTAs
//     (st.shared::cluster, distributed shared memory) and each CTA signals every peer's mbarrier with a
//     cluster-scope release-arrive; the MMA thread of each CTA acquires it.  No global-memory round trip and
//     no grid-wide barrier on the critical path.
// Outputs (bf16 hi [+lo] planes for the next layer's GEMM, or act(h + skip) for the consumer conv) are
// fire-and-forget global stores.
#include <cuda_bf16.h>

#include "common.cuh"
#include "sm100.cuh"

namespace {

using namespace sm100;

constexpr int CL = 16;        // CTAs per cluster
constexpr int HID = 512;      // hidden size (EnCodec)
constexpr int UPC = HID / CL; // 32 units per CTA
constexpr int NB = 16;        // clips per cluster (UMMA N)
constexpr int THREADS = 192;
constexpr uint32_t A_BYTES = 128 * HID * 2;         // 131072
constexpr uint32_t B_BYTES = NB * HID * 2;          // 16384 per parity
constexpr uint32_t GS_FLOATS = 4 * UPC * 17;        // gate exchange, padded

struct LstmTcParams {
    const float* pre;            // [B][T][4*HID]
    __nv_bfloat16* out_hi;       // [B][T][HID] or null
    __nv_bfloat16* out_lo;
    const __nv_bfloat16* skip_hi;
    const __nv_bfloat16* skip_lo;
    __nv_bfloat16* fin_hi;
    __nv_bfloat16* fin_lo;
    long long skip_bs, fin_bs;
    int fin_act, batch, steps;
    long long* dbg;  // optional [steps][8] clock64 samples from cluster 0 / CTA 0 (profiling aid)
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    // non-.aligned forms: the role branches above may leave a warp's lanes un-converged
    asm volatile("barrier.cluster.arrive.release;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint4 v) {
    asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote_release(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0, ok = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) break;
        if (++spins > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }  // the 4 epilogue warps

__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) {
    const float e = __expf(-2.0f * fabsf(x));
    const float t = __fdividef(1.0f - e, 1.0f + e);
    return copysignf(t, x);
}
__device__ __forceinline__ float elu_f(float x) { return x > 0.f ? x : expm1f(x); }

template <int NBV>  // clips handled per cluster (8 or 16); the UMMA N stays 16, unused B rows are zero
__global__ void __launch_bounds__(THREADS, 1)
lstm_tc_kernel(const __grid_constant__ CUtensorMap wmap, const LstmTcParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a_s = smem;                         // W_hh slice: 8 chunks of [128 rows][64] SW128 (16 KB each)
    uint8_t* b_s = smem + A_BYTES;               // h operand: 2 parities x 8 chunks of [16 rows][64] SW128 (2 KB each)
    float* gs = reinterpret_cast<float*>(b_s + 2 * B_BYTES);          // [4][32][17] activated gates
    __nv_bfloat16* hs = reinterpret_cast<__nv_bfloat16*>(gs + GS_FLOATS);  // [16 clips][32 units] new h slice
    uint64_t* bars = reinterpret_cast<uint64_t*>(hs + NB * UPC);
    uint64_t* w_full = bars;          // W_hh landed
    uint64_t* h_ready = bars + 1;     // [2] all 16 slices of h for this parity have arrived
    uint64_t* d_full = bars + 3;      // accumulator of the current step complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform role index
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int clip0 = cluster_id_x() * NBV;

    if (threadIdx.x == 0) {
        prefetch_tensormap(&wmap);
        mbar_init(w_full, 1);
        mbar_init(&h_ready[0], CL);
        mbar_init(&h_ready[1], CL);
        mbar_init(d_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 32);
    // h[-1] = 0: zero both parities of the B operand
    for (int i = threadIdx.x; i < (int)(2 * B_BYTES / 16); i += THREADS) reinterpret_cast<uint4*>(b_s)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_all();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // every CTA's barriers are initialised before any peer signals them
    tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // W_hh rows of gate g, units [32 rank, +32): global row g*HID + 32*rank ; smem rows g*32.. of every k-chunk
            mbar_arrive_expect_tx(w_full, A_BYTES);
            for (int kc = 0; kc < 8; ++kc)
                for (int g = 0; g < 4; ++g)
                    tma_load_2d(a_s + kc * 16384 + g * 32 * 128, &wmap, w_full, kc * 64, g * HID + (int)rank * UPC);
        }
    } else if (warp == 1) {
        // whole warp walks the loop (uniform control flow keeps the descriptors in uniform registers); one elected lane
        // issues the 32 tcgen05.mma of the step back to back
        const bool leader = elect_one();
        mbar_wait(w_full, 0);
        const uint32_t idesc = make_idesc_bf16(128, NB);
        const uint32_t a0 = smem_u32(a_s);
        const uint64_t desc_base = make_smem_desc(0, 128);
        for (int t = 0; t < p.steps; ++t) {
            const int par = t & 1;
            if (t > 0) mbar_wait_cluster(&h_ready[par], ((t - 1) >> 1) & 1);  // h[t-1] complete in buffer `par`
            tc_fence_after();
            if (p.dbg && blockIdx.x == 0 && leader) { p.dbg[t * 8 + 5] = clock64(); }
            const uint32_t b0 = smem_u32(b_s + par * B_BYTES);
            if (leader) {
#pragma unroll
                for (int kc = 0; kc < 8; ++kc) {
                    const uint64_t adesc = desc_base | (((a0 + kc * 16384) & 0x3FFFFu) >> 4);
                    const uint64_t bdesc = desc_base | (((b0 + kc * 2048) & 0x3FFFFu) >> 4);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kc | k) != 0);
                }
                umma_commit(d_full);
            }
            __syncwarp();
            if (p.dbg && blockIdx.x == 0 && leader) { p.dbg[t * 8 + 6] = clock64(); }
        }
    } else {
        // ======================================================== epilogue: 128 threads
        constexpr int CPT = NBV / 4;       // clips per thread in the cell update
        constexpr int GROUPS = 128 / (NBV * 4);  // thread groups sharing the 16 destination CTAs of the broadcast
        constexpr int DPT = CL / GROUPS;   // destinations per thread
        const int e = threadIdx.x - 64;
        const int g = warp & 3;            // TMEM lane quarter == gate index (rows g*32 + u)
        const int u = lane;
        const int bq = e >> 5;             // clips bq*CPT .. for the cell update
        const int gu = (int)rank * UPC + u;
        float c_state[CPT], h_prev[CPT];
#pragma unroll
        for (int i = 0; i < CPT; ++i) { c_state[i] = 0.f; h_prev[i] = 0.f; }
        float pre_next[NBV];
        auto load_pre = [&](int t) {
#pragma unroll
            for (int b = 0; b < NBV; ++b) {
                const int clip = clip0 + b;
                pre_next[b] = clip < p.batch ? __ldg(p.pre + ((size_t)clip * p.steps + t) * (4 * HID) + g * HID + gu) : 0.f;
            }
        };
        // global outputs of step t are written one step late, while this thread would otherwise idle on the MMA:
        // keeping them off the path between the DSMEM stores and the cluster-scope release (which waits for
        // every earlier store of the thread) is worth ~2000 cycles per step.
        auto emit = [&](int t) {
#pragma unroll
            for (int i = 0; i < CPT; ++i) {
                const int clip = clip0 + bq * CPT + i;
                if (clip >= p.batch) continue;
                const float h = h_prev[i];
                const __nv_bfloat16 hb = __float2bfloat16(h);
                const size_t o = ((size_t)clip * p.steps + t) * HID + gu;
                if (p.out_hi) {
                    p.out_hi[o] = hb;
                    if (p.out_lo) p.out_lo[o] = __float2bfloat16(h - __bfloat162float(hb));
                }
                if (p.fin_hi) {
                    float y = h;
                    const size_t so = (size_t)clip * p.skip_bs + (size_t)t * HID + gu;
                    if (p.skip_hi) y += __bfloat162float(p.skip_hi[so]);
                    if (p.skip_lo) y += __bfloat162float(p.skip_lo[so]);
                    if (p.fin_act == AC_ACT_ELU) y = elu_f(y);
                    const __nv_bfloat16 yb = __float2bfloat16(y);
                    const size_t fo = (size_t)clip * p.fin_bs + (size_t)t * HID + gu;
                    p.fin_hi[fo] = yb;
                    if (p.fin_lo) p.fin_lo[fo] = __float2bfloat16(y - __bfloat162float(yb));
                }
            }
        };
        load_pre(0);
        for (int t = 0; t < p.steps; ++t) {
            float pre_cur[NBV];
#pragma unroll
            for (int b = 0; b < NBV; ++b) pre_cur[b] = pre_next[b];
            if (t + 1 < p.steps) load_pre(t + 1);  // in flight during this step's MMA
            if (t > 0) emit(t - 1);
            const bool dbg = p.dbg && blockIdx.x == 0 && e == 0;
            if (dbg) p.dbg[t * 8 + 0] = clock64();
            mbar_wait(d_full, t & 1);
            tc_fence_after();
            if (dbg) p.dbg[t * 8 + 1] = clock64();
            uint32_t v[NBV];
            if constexpr (NBV == 16) tmem_ld16(tmem_d + ((uint32_t)(g * 32) << 16), v);
            else tmem_ld8(tmem_d + ((uint32_t)(g * 32) << 16), v);
            tmem_ld_wait();
            tc_fence_before();
            if (g == 2) {  // warp-uniform: the g gate is tanh, i/f/o are sigmoids
#pragma unroll
                for (int b = 0; b < NBV; ++b) gs[(g * UPC + u) * 17 + b] = fast_tanh(__uint_as_float(v[b]) + pre_cur[b]);
            } else {
#pragma unroll
                for (int b = 0; b < NBV; ++b) gs[(g * UPC + u) * 17 + b] = fast_sigmoid(__uint_as_float(v[b]) + pre_cur[b]);
            }
            epi_bar_sync();
            if (dbg) p.dbg[t * 8 + 2] = clock64();
#pragma unroll
            for (int i = 0; i < CPT; ++i) {
                const int b = bq * CPT + i;
                const float ig = gs[(0 * UPC + u) * 17 + b], fg = gs[(1 * UPC + u) * 17 + b];
                const float gg = gs[(2 * UPC + u) * 17 + b], og = gs[(3 * UPC + u) * 17 + b];
                c_state[i] = fg * c_state[i] + ig * gg;
                h_prev[i] = og * fast_tanh(c_state[i]);
                hs[b * UPC + u] = __float2bfloat16(h_prev[i]);
            }
            if (dbg) p.dbg[t * 8 + 3] = clock64();
            if (t + 1 < p.steps) {
                epi_bar_sync();  // hs complete (and gs reads finished)
                // broadcast the slice into parity (t+1)&1 of every CTA's B operand, in its swizzled K-major position
                const int npar = (t + 1) & 1;
                const int b = e / (4 * GROUPS);      // clip row
                const int jj = (e / GROUPS) & 3;     // 16-byte unit of the 64-byte slice row
                const int grp = e % GROUPS;
                const uint4 val = *reinterpret_cast<const uint4*>(hs + b * UPC + jj * 8);
                const int kc = (int)rank >> 1;
                const int unit = 4 * ((int)rank & 1) + jj;
                const uint32_t off = (uint32_t)(npar * B_BYTES + kc * 2048 + b * 128 + ((unit ^ (b & 7)) << 4));
                const uint32_t local = smem_u32(b_s) + off;
#pragma unroll
                for (int d = 0; d < DPT; ++d) st_cluster_v4(map_to_cta(local, (uint32_t)(grp * DPT + d)), val);
                fence_proxy_async_all();  // generic-proxy stores -> visible to the peers' async-proxy (tcgen05) reads
                epi_bar_sync();
                if (dbg) p.dbg[t * 8 + 4] = clock64();
                if (warp == 2 && lane < CL) mbar_arrive_remote_release(map_to_cta(smem_u32(&h_ready[npar]), (uint32_t)lane));
            }
        }
        emit(p.steps - 1);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // no CTA exits while a peer may still write into its shared memory
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem_d, 32);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

extern "C" int ac_lstm_tc(const ac_lstm_tc_desc* d, void* stream) {
    AC_REQUIRE(d && d->pre && d->w_hh_bf16, "ac_lstm_tc: null pointer");
    AC_REQUIRE(d->hidden == HID, "ac_lstm_tc: hidden %d (this kernel is built for %d)", d->hidden, HID);
    AC_REQUIRE(d->batch > 0 && d->steps > 0, "ac_lstm_tc: empty problem");
    AC_REQUIRE(d->out_hi || d->final_hi, "ac_lstm_tc: no output");
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            encode = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    AC_REQUIRE(encode, "ac_lstm_tc: cuTensorMapEncodeTiled not available");
    CUtensorMap wmap;
    {
        cuuint64_t gdim[2] = {HID, 4 * HID};
        cuuint64_t gstr[1] = {HID * 2};
        cuuint32_t box[2] = {64, UPC};
        cuuint32_t est[2] = {1, 1};
        CUresult r = encode(&wmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(d->w_hh_bf16), gdim, gstr, box, est,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        AC_REQUIRE(r == CUDA_SUCCESS, "ac_lstm_tc: cuTensorMapEncodeTiled failed: %d", (int)r);
    }
    const size_t smem = 1024 + A_BYTES + 2 * B_BYTES + GS_FLOATS * 4 + NB * UPC * 2 + 64;
    // Only 4 clusters of 16 CTAs are co-resident on a B200 (measured: 8 clusters ran as two waves), so a cluster
    // takes 16 clips unless the whole batch fits in 4 clusters of 8 (half the per-step epilogue work).
    const int nbv = d->batch <= 32 ? 8 : 16;
    static bool configured = false;
    if (!configured) {
        for (const void* fn : {(const void*)lstm_tc_kernel<8>, (const void*)lstm_tc_kernel<16>}) {
            cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            if (e != cudaSuccess) { ac::set_error("ac_lstm_tc: func attributes: %s", cudaGetErrorString(e)); return (int)e; }
        }
        configured = true;
    }
    LstmTcParams p{};
    p.pre = d->pre;
    p.out_hi = (__nv_bfloat16*)d->out_hi; p.out_lo = (__nv_bfloat16*)d->out_lo;
    p.skip_hi = (const __nv_bfloat16*)d->skip_hi; p.skip_lo = (const __nv_bfloat16*)d->skip_lo;
    p.fin_hi = (__nv_bfloat16*)d->final_hi; p.fin_lo = (__nv_bfloat16*)d->final_lo;
    p.skip_bs = d->skip_bstride; p.fin_bs = d->final_bstride;
    p.fin_act = d->final_act; p.batch = d->batch; p.steps = d->steps;
    p.dbg = (long long*)d->dbg;

    const int clusters = (d->batch + nbv - 1) / nbv;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(clusters * CL);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = nbv == 8 ? cudaLaunchKernelEx(&cfg, lstm_tc_kernel<8>, wmap, p) : cudaLaunchKernelEx(&cfg, lstm_tc_kernel<16>, wmap, p);
    ac::count_launch();
    if (e != cudaSuccess) { ac::set_error("ac_lstm_tc: launch: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}
