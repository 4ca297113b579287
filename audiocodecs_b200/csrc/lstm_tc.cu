// LSTM layer recurrence on tcgen05 tensor cores, one thread-block CLUSTER of 16 CTAs per <= 16 clips
// (ac_lstm_tc in include/audiocodecs_b200.h; replaces EncodecLSTM's nn.LSTM, HF/encodec:236-249).
//
// The 750-step chain is latency-bound, so everything a step needs stays on chip:
//   * CTA r of the cluster owns hidden units [32r, 32r+32): its 128 gate rows (i,f,g,o x 32 units) of W_hh (16-bit,
//     128 KB) live in TENSOR MEMORY for the whole kernel as the A operand of tcgen05.mma (lane = gate row, two values
//     per 32-bit column: 256 of the 512 columns).  Streaming them from shared memory cost ~1000 cycles per step
//     (128 KB at 128 B/clk); from TMEM the 32 MMAs of a step retire in a few hundred.
//   * the cluster's clips form TWO GROUPS of up to 8 that run as independent, interleaved pipelines: a step of a group is
//     MMA (tensor pipe) -> gates / cell update (that group's eight epilogue warps) -> all-gather of the new h (DSMEM), a
//     strictly serial chain of fixed latencies, so while one group's h is in flight the other group's MMAs and gate math
//     run.  clock64 profile of a group's step (7 clusters, 4-5 clips per group): MMA issue 371 + completion 55 -> tcgen05.ld
//     + gates + barrier 476 -> cell update 196 -> proxy fence + barrier + copy issue 602 -> DSMEM flight and skew 196 = 1 896
//     cycles, against 2 319 for one group of 16 clips per cluster (same 16 epilogue warps): 0.993 -> 0.852 ms per layer.
//     (With four epilogue warps per group the same pipeline was no faster than one group -- scripts/experiments/README.md.)
//   * h[t-1] of both groups is the B operand [16 x 512] (16 KB, double-buffered by step parity) in the un-swizzled
//     K-major core-matrix layout with the group as the OUTER index (8-row group stride 8 KB, K stride 128 B): the 32
//     units a CTA produces for one group are ONE contiguous 512-byte run of every peer's operand.  Each group's MMAs
//     compute all 16 columns (N = 16 is the minimum at M = 128) into their own accumulator; only the group's 8 are read.
//   * epilogue warp (group, gate g = TMEM lane quarter, half) holds gate g of 32 units for four of the group's clips: adds
//     the hoisted input projection (pre, prefetched one step ahead), applies sigmoid/tanh, the four gates meet through
//     shared memory, thread (g, half, unit) then owns the cell of clip g + 4*half: c in an fp32 register, h = o*tanh(c).
//   * the CTA's new h slice of a group is pushed to all 16 CTAs with bulk async copies (cp.async.bulk shared::cta ->
//     shared::cluster) that complete_tx on the DESTINATION's mbarrier: no per-thread remote stores, no cluster-scope
//     release fence and no separate arrive on the critical path.  Only the rows of clips that exist travel (`trim`).
//     (scripts/probe_tmem_a.cu pins the three hardware behaviours this relies on.)
//   * cudaOccupancyMaxActiveClusters reports 7 co-resident clusters of 16 on a B200 (profiles/r02_lstm_occupancy.txt), so a
//     batch is spread over up to 7 clusters per wave: 64 clips = 14 groups of 4-5 instead of 4 clusters x 16.
// Outputs (16-bit hi [+lo] planes for the next layer's GEMM, or act(h + skip) for the consumer conv) are
// fire-and-forget global stores issued one step late, off the critical path.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "sm100.cuh"

namespace {

using namespace sm100;

constexpr int CL = 16;        // CTAs per cluster
constexpr int HID = 512;      // hidden size (EnCodec)
constexpr int UPC = HID / CL; // 32 units per CTA
constexpr int NB = 16;        // UMMA N: two groups of 8 operand rows
constexpr int GRP = 8;        // clips per group (operand rows of one 8-row core-matrix group)
constexpr int EPI_WARPS = 16; // eight per group: two per TMEM lane quarter (= gate), each taking 4 of the group's 8 operand rows
constexpr int THREADS = 64 + 32 * EPI_WARPS;
constexpr uint32_t B_BYTES = NB * HID * 2;          // 16384 per parity
constexpr uint32_t G_BYTES = GRP * HID * 2;         // 8192: one group's rows
constexpr uint32_t SLICE_BYTES = GRP * UPC * 2;     // 512: one CTA's 32 units of one group's 8 clips (4 core matrices)
constexpr uint32_t GS_PAD = 9;
constexpr uint32_t GS_FLOATS = 4 * UPC * GS_PAD;    // gate exchange of one group, padded
constexpr uint32_t TMEM_COLS = 512;                 // D of group g: columns [16g, 16g+16); W_hh slice: columns [256,512)
constexpr uint32_t A_COL0 = 256;
// un-swizzled K-major operand: 8x8 core matrices of 128 B; K-direction stride 128 B, N-direction (8-clip group) stride 8 KB
constexpr uint32_t B_KSTR = 128, B_NSTR = G_BYTES;

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
    int n_groups;  // clip groups over the whole launch (two per cluster); group G takes batch/n_groups clips (+1 for the first batch%n_groups)
    int poll;      // every epilogue warp polls the accumulator barrier itself (else one per group does and releases the rest)
    int trim;      // push only the operand rows of clips that exist (4 copies of ng*16 bytes per peer instead of one of 512)
    int f16;    // operands (W_hh in tensor memory, h in shared memory) are fp16 instead of bf16
    int out_f16, skip_f16;  // hi-plane formats of out / final and of skip (lo planes are bf16)
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
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void grp_bar_sync(int grp) { asm volatile("bar.sync %0, 256;" ::"r"(1 + grp) : "memory"); }  // the 8 epilogue warps of a group
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&v)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
}

__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) {
    const float e = __expf(-2.0f * fabsf(x));
    const float t = __fdividef(1.0f - e, 1.0f + e);
    return copysignf(t, x);
}
__device__ __forceinline__ float elu_f(float x) { return x > 0.f ? x : expm1f(x); }
// one hi-plane element (bf16, or fp16 saturating at +-65504); returns the stored value
__device__ __forceinline__ float store_hi(__nv_bfloat16* dst, float v, bool f16) {
    if (f16) {
        const __half h = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
        *reinterpret_cast<__half*>(dst) = h;
        return __half2float(h);
    }
    const __nv_bfloat16 b = __float2bfloat16(v);
    *dst = b;
    return __bfloat162float(b);
}

__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&w)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,"
        "%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "r"(w[8]), "r"(w[9]),
          "r"(w[10]), "r"(w[11]), "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15]), "r"(w[16]), "r"(w[17]), "r"(w[18]), "r"(w[19]),
          "r"(w[20]), "r"(w[21]), "r"(w[22]), "r"(w[23]), "r"(w[24]), "r"(w[25]), "r"(w[26]), "r"(w[27]), "r"(w[28]), "r"(w[29]),
          "r"(w[30]), "r"(w[31])
        : "memory");
}
__device__ __forceinline__ void bulk_copy_to_peer(uint32_t dst_cluster, uint32_t src_local, uint32_t bytes, uint32_t bar_cluster) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_cluster), "r"(src_local), "r"(bytes), "r"(bar_cluster) : "memory");
}

__global__ void __launch_bounds__(THREADS, 1)
lstm_tc_kernel(const __nv_bfloat16* __restrict__ w_hh, const LstmTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align_smem_1024(smem_raw);
    uint8_t* b_s = smem;                         // h operand: 2 parities x (2 groups x 8 KB), un-swizzled K-major core matrices
    uint8_t* hs = b_s + 2 * B_BYTES;             // [parity][group] 512 B: this CTA's new h slice, already in operand layout
    float* gs = reinterpret_cast<float*>(hs + 4 * SLICE_BYTES);       // [group][4][32][9] activated gates
    uint64_t* bars = reinterpret_cast<uint64_t*>(gs + 2 * GS_FLOATS);
    uint64_t* h_ready = bars;         // [parity][group]: that group's h for this parity has landed (transaction bytes)
    uint64_t* d_full = bars + 4;      // [group]: accumulator of the group's current step complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform role index
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    // clips of the two groups of this cluster: group G of the launch takes base (+1 for the first `rem`) consecutive clips
    int ng[2], clip_base[2];
    {
        const int base = p.batch / p.n_groups, rem = p.batch % p.n_groups;
        for (int g = 0; g < 2; ++g) {
            const int G = (int)cluster_id_x() * 2 + g;
            ng[g] = G < p.n_groups ? base + (G < rem ? 1 : 0) : 0;
            clip_base[g] = G * base + (G < rem ? G : rem);
        }
    }
    // bytes one source CTA pushes per group and step: the whole 512-byte slice (trim 0), rows [0, n) of each of its four core
    // matrices (trim 1: four copies of n*16 bytes), or two copies that each span two core matrices up to row n of the second
    // (trim 3: 128 + n*16 bytes; the unused rows of the first travel along)
    auto slice_tx = [&](int n) { return p.trim == 1 ? 4u * n * 16u : (p.trim == 3 ? 2u * (128u + n * 16u) : SLICE_BYTES); };
    const uint32_t tx_bytes[2] = {CL * slice_tx(ng[0]), CL * slice_tx(ng[1])};

    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(&h_ready[i], 1);
        mbar_init(&d_full[0], 1);
        mbar_init(&d_full[1], 1);
        fence_barrier_init();
        // h[0] (written by step 0) lands in parity 1, h[1] in parity 0: arm both before any peer can send
        for (int i = 0; i < 4; ++i)
            if (ng[i & 1]) mbar_arrive_expect_tx(&h_ready[i], tx_bytes[i & 1]);
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    // h[-1] = 0: zero both parities of the B operand (rows of absent clips stay zero for the whole kernel)
    for (int i = threadIdx.x; i < (int)(2 * B_BYTES / 16); i += THREADS) reinterpret_cast<uint4*>(b_s)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_all();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp >= 2 && warp < 6) {
        // W_hh slice -> TMEM: thread (gate g = lane quarter, unit u) owns global row g*HID + 32*rank + u (1 KB = 256 words)
        const int g = warp & 3;
        const uint4* src = reinterpret_cast<const uint4*>(w_hh + ((size_t)g * HID + rank * UPC + lane) * HID);
        const uint32_t taddr = tmem_base + ((uint32_t)(g * 32) << 16) + A_COL0;
#pragma unroll 1
        for (int c = 0; c < 8; ++c) {
            uint32_t w[32];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const uint4 t = __ldg(src + c * 8 + q);
                w[4 * q] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w;
            }
            tmem_st32(taddr + c * 32, w);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // every CTA's barriers are armed and operands zeroed before any peer sends
    tc_fence_after();

    if (warp == 1) {
        // whole warp walks the loop (uniform control flow keeps the descriptors in uniform registers); one elected lane
        // issues the 32 tcgen05.mma of a group's step back to back.  The two groups alternate: while one group's gates and
        // all-gather run, the other's MMAs do.
        const bool leader = elect_one();
        // kind::f16 instruction descriptor: operand format bits 7-9 (A) / 10-12 (B) = 1 for bf16, 0 for fp16
        const uint32_t idesc = p.f16 ? make_idesc_f16(128, NB) : make_idesc_bf16(128, NB);
        uint64_t desc_base = 0;
        desc_base |= (uint64_t)((B_KSTR >> 4) & 0x3FFF) << 16;  // leading byte offset = K-direction core-matrix stride
        desc_base |= (uint64_t)((B_NSTR >> 4) & 0x3FFF) << 32;  // stride byte offset = 8-row group stride
        desc_base |= 1ull << 46;
        const uint32_t a_tmem = tmem_base + A_COL0;
        for (int t = 0; t < p.steps; ++t) {
            const int par = t & 1;
            const uint64_t bdesc = desc_base | (((smem_u32(b_s) + par * B_BYTES) & 0x3FFFFu) >> 4);
            for (int g = 0; g < 2; ++g) {
                if (!ng[g]) continue;
                if (t > 0) {
                    mbar_wait(&h_ready[par * 2 + g], ((t - 1) >> 1) & 1);  // the group's h[t-1] has landed in buffer `par`
                    if (leader && t + 1 < p.steps) mbar_arrive_expect_tx(&h_ready[par * 2 + g], tx_bytes[g]);  // re-arm for h[t+1]
                }
                tc_fence_after();
                if (p.dbg && blockIdx.x == 0 && leader && g == 0) { p.dbg[t * 8 + 5] = clock64(); }
                if (leader) {
#pragma unroll
                    for (int ks = 0; ks < HID / 16; ++ks)  // K = 16 per MMA: 8 TMEM columns of A, two core matrices of B
                        umma_bf16_ts(tmem_base + NB * g, a_tmem + ks * 8, bdesc + (uint64_t)((2 * B_KSTR) >> 4) * ks, idesc, ks != 0);
                    umma_commit(&d_full[g]);
                }
                __syncwarp();
                if (p.dbg && blockIdx.x == 0 && leader && g == 0) { p.dbg[t * 8 + 6] = clock64(); }
            }
        }
    } else if (warp >= 2) {
        // ======================================================== epilogue: 2 groups x 256 threads
        const int ew = warp - 2;           // 0..15
        const int q = warp & 3;            // TMEM lane quarter == gate index (rows q*32 + u)
        const int grp = ew >> 3;           // clip group of this warp
        const int sub = (ew >> 2) & 1;     // which 4 of the group's 8 operand rows this warp activates
        const int w8 = ew & 7;             // warp index inside the group
        const int n = ng[grp], cb = clip_base[grp];
        const int u = lane;
        const int gu = (int)rank * UPC + u;
        const bool dbg = p.dbg && blockIdx.x == 0 && threadIdx.x == 64;
        if (n > 0) {
            constexpr int HC = GRP / 2;    // operand rows (clips) per thread in the activation phase
            float* gsg = gs + grp * GS_FLOATS;
            // cell update: thread (q, sub, u) owns unit u of the group's clip q + 4*sub
            const int bc = q + 4 * sub;
            const bool has_cell = bc < n;
            float c_state = 0.f, h_prev = 0.f;
            float pre_next[HC];
            auto load_pre = [&](int t) {
#pragma unroll
                for (int b = 0; b < HC; ++b) {
                    const int clip = sub * HC + b;
                    pre_next[b] = clip < n ? __ldg(p.pre + ((size_t)(cb + clip) * p.steps + t) * (4 * HID) + q * HID + gu) : 0.f;
                }
            };
            // global outputs of step t are written one step late, while this thread would otherwise idle on the MMA
            auto emit = [&](int t) {
                if (!has_cell) return;
                const int clip = cb + bc;
                const float h = h_prev;
                const size_t o = ((size_t)clip * p.steps + t) * HID + gu;
                if (p.out_hi) {
                    const float hv = store_hi(p.out_hi + o, h, p.out_f16 != 0);
                    if (p.out_lo) p.out_lo[o] = __float2bfloat16(h - hv);
                }
                if (p.fin_hi) {
                    float y = h;
                    const size_t so = (size_t)clip * p.skip_bs + (size_t)t * HID + gu;
                    if (p.skip_hi) y += p.skip_f16 ? __half2float(reinterpret_cast<const __half*>(p.skip_hi)[so]) : __bfloat162float(p.skip_hi[so]);
                    if (p.skip_lo) y += __bfloat162float(p.skip_lo[so]);
                    if (p.fin_act == AC_ACT_ELU) y = elu_f(y);
                    const size_t fo = (size_t)clip * p.fin_bs + (size_t)t * HID + gu;
                    const float yv = store_hi(p.fin_hi + fo, y, p.out_f16 != 0);
                    if (p.fin_lo) p.fin_lo[fo] = __float2bfloat16(y - yv);
                }
            };
            // destinations of this CTA's slice of the group: the eight warps cover the 16 peers, lane l < 2 of warp w pushes the
            // 512-byte slice to peer 2*(w % 8) + l.  (Trimmed: lane l < 8 pushes rows [0, n) of core matrix l & 3 to peer 2*(w%8) + (l>>2).)
            const uint32_t dst_rank = (uint32_t)(w8 * 2 + (p.trim == 1 ? (lane >> 2) & 1 : (p.trim == 3 ? (lane >> 1) & 1 : lane & 1)));
            const uint32_t piece = p.trim == 1 ? (uint32_t)(lane & 3) * 128u : (p.trim == 3 ? (uint32_t)(lane & 1) * 256u : 0u);
            const uint32_t copy_bytes = p.trim == 1 ? (uint32_t)n * 16u : (p.trim == 3 ? 128u + (uint32_t)n * 16u : SLICE_BYTES);
            const bool sender = lane < (p.trim == 1 ? 8 : (p.trim == 3 ? 4 : 2));
            const uint32_t peer_b = map_to_cta(smem_u32(b_s) + grp * G_BYTES + rank * SLICE_BYTES + piece, dst_rank);
            const uint32_t peer_bar = map_to_cta(smem_u32(&h_ready[grp]), dst_rank);
            load_pre(0);
            for (int t = 0; t < p.steps; ++t) {
                float pre_cur[HC];
#pragma unroll
                for (int b = 0; b < HC; ++b) pre_cur[b] = pre_next[b];
                if (t + 1 < p.steps) load_pre(t + 1);  // in flight during this step's MMA
                if (t > 0) emit(t - 1);
                if (dbg) p.dbg[t * 8 + 0] = clock64();
                // p.poll = 0 (AC_LSTM_POLL=0): one warp of the group polls the accumulator barrier and the other seven sleep in the
                // hardware barrier -- tried against the power cap: +70 cycles per step, no measurable effect on the step
                if (p.poll || w8 == 0) mbar_wait(&d_full[grp], t & 1);
                if (!p.poll) grp_bar_sync(grp);
                tc_fence_after();
                if (dbg) p.dbg[t * 8 + 1] = clock64();
                uint32_t v[HC];
                tmem_ld4(tmem_base + ((uint32_t)(q * 32) << 16) + NB * grp + GRP * grp + sub * HC, v);  // 4 of the group's own 8 columns
                tmem_ld_wait();
                tc_fence_before();
                if (q == 2) {  // warp-uniform: the g gate is tanh, i/f/o are sigmoids
#pragma unroll
                    for (int b = 0; b < HC; ++b)
                        if (sub * HC + b < n) gsg[(q * UPC + u) * GS_PAD + sub * HC + b] = fast_tanh(__uint_as_float(v[b]) + pre_cur[b]);
                } else {
#pragma unroll
                    for (int b = 0; b < HC; ++b)
                        if (sub * HC + b < n) gsg[(q * UPC + u) * GS_PAD + sub * HC + b] = fast_sigmoid(__uint_as_float(v[b]) + pre_cur[b]);
                }
                grp_bar_sync(grp);
                if (dbg) p.dbg[t * 8 + 2] = clock64();
                const int npar = (t + 1) & 1;
                uint8_t* hsl = hs + (npar * 2 + grp) * SLICE_BYTES;
                if (has_cell) {
                    const float ig = gsg[(0 * UPC + u) * GS_PAD + bc], fg = gsg[(1 * UPC + u) * GS_PAD + bc];
                    const float gg = gsg[(2 * UPC + u) * GS_PAD + bc], og = gsg[(3 * UPC + u) * GS_PAD + bc];
                    c_state = fg * c_state + ig * gg;
                    h_prev = og * fast_tanh(c_state);
                    // operand layout of the slice: core matrix u/8 (128 B), row bc, element u%8
                    uint8_t* hp = hsl + (u >> 3) * B_KSTR + bc * 16 + (u & 7) * 2;
                    if (p.f16) *reinterpret_cast<__half*>(hp) = __float2half_rn(h_prev);
                    else *reinterpret_cast<__nv_bfloat16*>(hp) = __float2bfloat16(h_prev);
                }
                if (dbg) p.dbg[t * 8 + 3] = clock64();
                if (t + 1 < p.steps) {
                    fence_proxy_async();      // this thread's generic-proxy stores to hs (shared::cta only: a full proxy fence would also
                                              // wait for the deferred global stores) -> visible to the bulk-copy (async) proxy
                    grp_bar_sync(grp);        // slice complete (and gs reads finished)
                    if (sender) bulk_copy_to_peer(peer_b + npar * B_BYTES, smem_u32(hsl) + piece, copy_bytes, peer_bar + npar * 16);
                    if (dbg) p.dbg[t * 8 + 4] = clock64();
                }
            }
            emit(p.steps - 1);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // no CTA exits while a peer may still write into its shared memory
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

}  // namespace

// Diagnostic: how many clusters of `cluster_size` CTAs of this kernel the device can hold at once (cudaOccupancyMaxActiveClusters).
extern "C" int ac_lstm_tc_max_clusters(int32_t cluster_size, int32_t smem_bytes) {
    cudaError_t e = cudaFuncSetAttribute((const void*)lstm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute((const void*)lstm_tc_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) { ac::set_error("ac_lstm_tc_max_clusters: %s", cudaGetErrorString(e)); return -1; }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(cluster_size * 16);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = smem_bytes;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster_size; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = 0;
    e = cudaOccupancyMaxActiveClusters(&n, (const void*)lstm_tc_kernel, &cfg);
    if (e != cudaSuccess) { ac::set_error("ac_lstm_tc_max_clusters: %s", cudaGetErrorString(e)); cudaGetLastError(); return -1; }
    return n;
}

extern "C" int ac_lstm_tc(const ac_lstm_tc_desc* d, void* stream) {
    AC_REQUIRE(d && d->pre && d->w_hh_bf16, "ac_lstm_tc: null pointer");
    AC_REQUIRE(d->hidden == HID, "ac_lstm_tc: hidden %d (this kernel is built for %d)", d->hidden, HID);
    AC_REQUIRE(d->batch > 0 && d->steps > 0, "ac_lstm_tc: empty problem");
    AC_REQUIRE(d->out_hi || d->final_hi, "ac_lstm_tc: no output");
    // The kernel needs ~45 KB of shared memory, but every CTA allocates all 512 TMEM columns: two CTAs of different
    // clusters on one SM would block each other's tcgen05.alloc (cross-cluster deadlock), so the launch asks for more
    // than half an SM's shared memory and exactly one CTA fits per SM.
    const size_t smem_used = 1024 + 2 * B_BYTES + 4 * SLICE_BYTES + 2 * GS_FLOATS * 4 + 128;
    const size_t smem = smem_used > 120 * 1024 ? smem_used : 120 * 1024;
    static bool configured = false;
    if (!configured) {
        for (const void* fn : {(const void*)lstm_tc_kernel}) {
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
    p.f16 = d->operand_fp16 ? 1 : 0;
    p.out_f16 = d->out_fp16 ? 1 : 0; p.skip_f16 = d->skip_fp16 ? 1 : 0;

    // Clusters per wave.  7 are co-resident on a B200 and a lone layer is fastest spread over all of them (0.852 ms at 64 clips:
    // 4-5 clips per group) -- but inside the power-capped step the same layer takes ~0.89 ms either way (the chain scales with
    // the SM clock) and 112 busy SMs instead of 64 cost the FOLLOWING kernels 1-2 % of clock: A/B of the whole step on one
    // box, twice on two boxes: 16.50 / 16.55 ms with 4 clusters against 16.76 / 16.77 ms with 7.  So 4 is the default
    // (AC_LSTM_CLUSTERS=0: ask the occupancy API, =n: n).
    static int max_clusters = 0;
    if (!max_clusters) {
        const char* env = getenv("AC_LSTM_CLUSTERS");
        int n = env ? atoi(env) : 4;
        if (n <= 0) {
            cudaLaunchConfig_t q{};
            q.gridDim = dim3(CL * 16); q.blockDim = dim3(THREADS); q.dynamicSmemBytes = smem;
            cudaLaunchAttribute qa[1];
            qa[0].id = cudaLaunchAttributeClusterDimension;
            qa[0].val.clusterDim.x = CL; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
            q.attrs = qa; q.numAttrs = 1;
            if (cudaOccupancyMaxActiveClusters(&n, (const void*)lstm_tc_kernel, &q) != cudaSuccess || n <= 0) { cudaGetLastError(); n = 4; }
        }
        max_clusters = n;
    }
    const int waves = (d->batch + NB * max_clusters - 1) / (NB * max_clusters);
    int clusters = waves * max_clusters;
    if (clusters > (d->batch + 1) / 2) clusters = (d->batch + 1) / 2;   // at least one clip per group (a lone clip: one group)
    p.n_groups = d->batch < 2 * clusters ? d->batch : 2 * clusters;
    // rows of existing clips only (4 small copies per peer) pay off while a group holds <= 6 clips; with full groups the whole
    // 512-byte slice in one copy is faster (measured: 0.852 vs 0.882 ms at 4-5 clips per group, 0.942 vs 0.888 ms at 8)
    static int trim = -1;
    if (trim < 0) { const char* env = getenv("AC_LSTM_TRIM"); trim = env ? atoi(env) : 2; }
    static int poll = -1;
    if (poll < 0) { const char* env = getenv("AC_LSTM_POLL"); poll = env ? atoi(env) : 1; }
    p.poll = poll;
    p.trim = trim == 2 ? ((d->batch + p.n_groups - 1) / p.n_groups <= 6 ? 1 : 0) : trim;

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
    cudaError_t e = cudaLaunchKernelEx(&cfg, lstm_tc_kernel, (const __nv_bfloat16*)d->w_hh_bf16, p);
    ac::count_launch();
    if (e != cudaSuccess) { ac::set_error("ac_lstm_tc: launch: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}
