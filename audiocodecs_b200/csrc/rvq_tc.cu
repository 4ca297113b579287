// Residual vector quantizer encode on tcgen05 tensor cores (ac_rvq_encode_tc in include/audiocodecs_b200.h).
//
// Replaces EncodecResidualVectorQuantizer.encode + EncodecEuclideanCodebook.quantize (HF/encodec:364-369,424-438; metric 0)
// and MimiResidualVectorQuantizer.encode + MimiEuclideanCodebook.quantize (HF/mimi:1197-1202,1262-1280; metric 1):
//   per stage k:  idx = argmin_c dist(r, E_k[c])   (lowest index wins ties);   r -= E_k[idx]
//     metric 0: score = -(|r|^2 - 2 r.E + |E|^2), argmax      metric 1: cdist(r, E) = sqrt(max(|r|^2 + |E|^2 - 2 r.E, 0)), argmin
//
// One persistent CTA owns a tile of 128 frames for ALL stages and the fp32 residual never leaves the SM:
//   * the residual r[128 x D] lives in TENSOR MEMORY (lane = frame, one fp32 per column: D of the 512 columns), read and
//     updated by its owner threads with tcgen05.ld / tcgen05.st -- shared memory is left to the operand planes and to a
//     deep codebook ring (the first version kept r in shared memory and was bound by the latency of a 5-stage ring);
//   * distance GEMM  r . E_k^T  on tcgen05 with error-compensated bf16 operands (r_hi.E_hi + r_hi.E_lo + r_lo.E_hi, fp32
//     accumulate: ~2^-16 relative; a plain bf16 GEMM matches the fp32 argmin for only 38 % of EnCodec stage-0 frames,
//     SURVEY section 7).  The codebook (hi/lo bf16 planes, L2-resident) streams through the TMA ring in blocks of
//     64 codes x 64 dims, continuously across stages and tiles;
//   * two TMEM buffers of 128 distance columns: the eight epilogue warps (lane = frame, the two warps of a lane quarter
//     split the columns) scan buffer g -- running top-2 per frame with packed (truncated distance | index) integer keys,
//     3 ops per code -- while the MMA warp fills buffer g+1;
//   * the two candidates of a frame are re-scored in exact fp32 (SIMT, the reference's formula and tie rule), so the
//     emitted code is the fp32 argmin wherever the tensor-core ranking has the true winner in its top 2; the winner is
//     subtracted from the residual in place and the operand planes are re-split for the next stage.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "sm100.cuh"

namespace {

using namespace sm100;

constexpr int TM = 128;           // frames per tile
constexpr int CHUNK = 64;         // codes per ring block / per MMA N
constexpr int GROUP_MAX = 128;    // codes per TMEM distance buffer: 128 (one CTA per SM, 512 TMEM columns), or 64 with D = 128 --
                                  // residual 128 + 2 x 64 distance columns = 256: TWO CTAs share an SM and one CTA's serial tail
                                  // (candidate merge, exact re-score, subtract, re-split) runs under the other's distance GEMM
constexpr int MAX_RING = 10;
constexpr int EPI_WARPS = 8;
constexpr int THREADS = 64 + 32 * EPI_WARPS;
constexpr uint32_t PLANE_KB_BYTES = TM * 128;             // one [128 rows x 64] SW128 block of an operand plane: 16 KB
constexpr uint32_t E_BLOCK_BYTES = CHUNK * 128;           // [64 codes x 64 dims] SW128: 8 KB
constexpr uint32_t RING_STAGE_BYTES = 2 * E_BLOCK_BYTES;  // one k-block of a chunk: E_hi, E_lo = 16 KB
constexpr int XCH = 8;                                     // floats exchanged per (frame, half)
constexpr int MAX_CODES = 2048;

struct RvqParams {
    const float* x;          // [rows][D]
    const float* cb;         // [S][n_codes][D] fp32 (exact re-score + subtract)
    const float* cbn;        // [S][n_codes] |E|^2
    int64_t* codes;
    float* res_out;          // optional [rows][D]
    long long rows;
    int n_codes, stages, stage0, code_stride, code_offset, tiles, metric, ring;
    int lo_row0;             // first row of the lo plane in the codebook tensor map
};

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
          "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// (score, index) ordering of the reference: larger score first, lower index on ties
__device__ __forceinline__ bool better(float s, int i, float t, int j) { return s > t || (s == t && i < j); }

__device__ __forceinline__ void top2_insert(float s, int c, float& v1, int& i1, float& v2, int& i2) {
    if (better(s, c, v1, i1)) { v2 = v1; i2 = i1; v1 = s; i1 = c; }
    else if (better(s, c, v2, i2)) { v2 = s; i2 = c; }
}

// the reference's score of one candidate from the exact fp32 dot product (larger is better)
__device__ __forceinline__ float ref_score(float xn, float dot, float en, int metric) {
    if (metric == 0) return -((xn - 2.f * dot) + en);          // HF/encodec:367
    return -sqrtf(fmaxf((xn + en) - 2.f * dot, 0.f));          // cdist, HF/mimi:1200
}

template <int D, int GROUP>
__global__ void __launch_bounds__(THREADS, GROUP == 64 ? 2 : 1)
rvq_encode_tc_kernel(const __grid_constant__ CUtensorMap emap, const RvqParams p) {
    constexpr uint32_t R_COL = 2 * GROUP;            // residual columns start after the two distance buffers
    constexpr uint32_t TMEM_COLS = (2 * GROUP + D) <= 256 ? 256 : 512;
    constexpr int HG = GROUP / 2;                    // codes of a distance buffer scanned by one thread of a frame
    constexpr int KB = D / 64;                       // k-blocks of 64 dims
    constexpr int HD = D / 2;                        // dims owned by one thread of a frame
    constexpr uint32_t A_PLANE_BYTES = KB * PLANE_KB_BYTES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align_smem_1024(smem_raw);
    uint8_t* a_hi = smem;
    uint8_t* a_lo = a_hi + A_PLANE_BYTES;
    uint8_t* ring = a_lo + A_PLANE_BYTES;
    float* en_s = reinterpret_cast<float*>(ring + (size_t)p.ring * RING_STAGE_BYTES);  // [n_codes <= 2048]
    float* xch = en_s + p.n_codes;                          // [TM][2][XCH] exchange between the two halves of a frame
    uint64_t* bars = reinterpret_cast<uint64_t*>(xch + TM * 2 * XCH);
    uint64_t* full = bars;                      // [MAX_RING]
    uint64_t* empty = bars + MAX_RING;          // [MAX_RING]
    uint64_t* tfull = bars + 2 * MAX_RING;      // [2]
    uint64_t* tempty = tfull + 2;               // [2]
    uint64_t* a_ready = tempty + 2;             // operand planes of the next stage are written
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_ready + 1);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const int groups = p.n_codes / GROUP;

    if (threadIdx.x == 0) {
        prefetch_tensormap(&emap);
        for (int i = 0; i < p.ring; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], EPI_WARPS); }
        mbar_init(a_ready, EPI_WARPS);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================================================= codebook producer (TMA)
        if (lane == 0) {
            int rs = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x)
                for (int k = 0; k < p.stages; ++k)
                    for (int c0 = 0; c0 < p.n_codes; c0 += CHUNK)
                        for (int kb = 0; kb < KB; ++kb) {
                            mbar_wait(&empty[rs], ph ^ 1);
                            mbar_arrive_expect_tx(&full[rs], RING_STAGE_BYTES);
                            uint8_t* dst = ring + (size_t)rs * RING_STAGE_BYTES;
                            const int row = (p.stage0 + k) * p.n_codes + c0;
                            tma_load_2d(dst, &emap, &full[rs], kb * 64, row);
                            tma_load_2d(dst + E_BLOCK_BYTES, &emap, &full[rs], kb * 64, p.lo_row0 + row);
                            if (++rs == p.ring) { rs = 0; ph ^= 1; }
                        }
        }
    } else if (warp == 1) {
        // ================================================================= MMA issuer (warp-uniform loop, one elected lane)
        const bool leader = elect_one();
        const uint32_t idesc = make_idesc_bf16(TM, CHUNK);
        const uint64_t desc_base = make_smem_desc(0, 128);
        const uint32_t a_hi_u = smem_u32(a_hi), a_lo_u = smem_u32(a_lo), ring_u = smem_u32(ring);
        int rs = 0;
        uint32_t ph = 0, gcount = 0, acount = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x)
            for (int k = 0; k < p.stages; ++k) {
                mbar_wait(a_ready, acount & 1);
                ++acount;
                tc_fence_after();
                for (int g = 0; g < groups; ++g, ++gcount) {
                    const uint32_t buf = gcount & 1;
                    mbar_wait(&tempty[buf], ((gcount >> 1) & 1) ^ 1);
                    tc_fence_after();
                    for (int c = 0; c < GROUP / CHUNK; ++c) {
                        const uint32_t d_tmem = tmem_base + buf * GROUP + c * CHUNK;
#pragma unroll
                        for (int kb = 0; kb < KB; ++kb) {
                            mbar_wait(&full[rs], ph);
                            tc_fence_after();
                            const uint32_t e_u = ring_u + rs * RING_STAGE_BYTES;
                            if (leader) {
                                const uint64_t ah = desc_base | (((a_hi_u + kb * PLANE_KB_BYTES) & 0x3FFFFu) >> 4);
                                const uint64_t al = desc_base | (((a_lo_u + kb * PLANE_KB_BYTES) & 0x3FFFFu) >> 4);
                                const uint64_t eh = desc_base | ((e_u & 0x3FFFFu) >> 4);
                                const uint64_t el = desc_base | (((e_u + E_BLOCK_BYTES) & 0x3FFFFu) >> 4);
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks) {
                                    umma_bf16(d_tmem, ah + 2 * ks, eh + 2 * ks, idesc, (kb | ks) != 0);
                                    umma_bf16(d_tmem, ah + 2 * ks, el + 2 * ks, idesc, 1u);
                                    umma_bf16(d_tmem, al + 2 * ks, eh + 2 * ks, idesc, 1u);
                                }
                                umma_commit(&empty[rs]);
                            }
                            __syncwarp();
                            if (++rs == p.ring) { rs = 0; ph ^= 1; }
                        }
                    }
                    if (leader) umma_commit(&tfull[buf]);
                    __syncwarp();
                }
            }
    } else {
        // ================================================================= epilogue: 8 warps, thread = (frame, half)
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;            // which 64 of a buffer's 128 columns / which D/2 dims this thread owns
        const int r = quarter * 32 + lane;           // frame within the tile == TMEM lane
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const uint32_t r_addr = lane_addr + R_COL + half * HD;   // this thread's slice of the residual row
        float* my_x = xch + (r * 2 + half) * XCH;
        const float* peer_x = xch + (r * 2 + (half ^ 1)) * XCH;
        const int d0 = half * HD;
        uint32_t gcount = 0;

        // 16 residual values (dims d0 + c*16 ..) -> hi / lo operand planes in the SW128 K-major layout; returns their sum of squares
        auto split16 = [&](const uint32_t (&v)[16], int c) {
            float ss = 0.f;
            const int d = d0 + c * 16;
            uint8_t* row_hi = a_hi + (d >> 6) * PLANE_KB_BYTES + r * 128;
            uint8_t* row_lo = a_lo + (d >> 6) * PLANE_KB_BYTES + r * 128;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float f0 = __uint_as_float(v[u * 8 + 2 * j]), f1 = __uint_as_float(v[u * 8 + 2 * j + 1]);
                    ss = fmaf(f0, f0, ss);
                    ss = fmaf(f1, f1, ss);
                    hi[j] = pack2(f0, f1);
                    const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&hi[j]);
                    lo[j] = pack2(f0 - __low2float(h2), f1 - __high2float(h2));
                }
                const uint32_t off = (uint32_t)(((((d & 63) >> 3) + u) ^ (r & 7)) << 4);  // 16-byte unit index XOR (row & 7)
                *reinterpret_cast<uint4*>(row_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4*>(row_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
            return ss;
        };

        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
            const long long row = (long long)tile * TM + r;
            const bool valid = row < p.rows;
            // ---- load the frame's embedding (this thread's D/2 dims): residual -> TMEM, operand planes -> shared memory
            float xn_part = 0.f;
#pragma unroll 1
            for (int c = 0; c < HD / 16; ++c) {
                uint32_t v[16];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 t = valid ? __ldg(reinterpret_cast<const float4*>(p.x + row * D + d0 + c * 16) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                    v[4 * q] = __float_as_uint(t.x); v[4 * q + 1] = __float_as_uint(t.y);
                    v[4 * q + 2] = __float_as_uint(t.z); v[4 * q + 3] = __float_as_uint(t.w);
                }
                tmem_st16(r_addr + c * 16, v);
                xn_part += split16(v, c);
            }
            tmem_st_wait();
            for (int k = 0; k < p.stages; ++k) {
                float* en = en_s;
                for (int i = threadIdx.x - 64; i < p.n_codes; i += 32 * EPI_WARPS) en[i] = __ldg(p.cbn + (size_t)(p.stage0 + k) * p.n_codes + i);
                my_x[0] = xn_part;
                fence_proxy_async();  // operand planes (generic-proxy stores) -> visible to tcgen05 (async proxy)
                epi_bar();            // en / xn parts complete; every thread's plane stores are fenced
                if (lane == 0) mbar_arrive(a_ready);
                const float xn = half == 0 ? xn_part + peer_x[0] : peer_x[0] + xn_part;  // lower dims first, on both halves

                // Candidate selection: per frame the two smallest keys of this thread's n_codes/2 codes, where
                //   key = (bits(max(dist, 0)) & ~1023) | local_index,   dist = (|r|^2 + |E|^2) - 2 r.E  >= 0
                // (monotone in both metrics).  Non-negative floats order like their bit patterns, so ONE integer min/max
                // chain tracks value and index together (3 ops per code).  Dropping 10 mantissa bits (2^-13 relative) keeps
                // the true winner's key <= every other key; it can only fall out of the top 2 through a three-way tie of
                // truncated keys, i.e. inside the near-tie zone of the parity definition -- and the two candidates are
                // re-scored exactly below.
                uint32_t k1 = 0xFFFFFFFFu, k2 = 0xFFFFFFFFu;
                for (int g = 0; g < groups; ++g, ++gcount) {
                    const uint32_t buf = gcount & 1;
                    mbar_wait(&tfull[buf], (gcount >> 1) & 1);
                    tc_fence_after();
#pragma unroll 2
                    for (int cc = 0; cc < HG / 16; ++cc) {
                        const int col = half * HG + cc * 16;
                        uint32_t v[16];
                        tmem_ld16(lane_addr + buf * GROUP + col, v);
                        const float4* enp = reinterpret_cast<const float4*>(en + g * GROUP + col);
                        float e[16];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 t = enp[q];
                            e[4 * q] = xn + t.x; e[4 * q + 1] = xn + t.y; e[4 * q + 2] = xn + t.z; e[4 * q + 3] = xn + t.w;
                        }
                        const uint32_t lbase = (uint32_t)(g * HG + cc * 16);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float dist = fmaxf(fmaf(-2.f, __uint_as_float(v[j]), e[j]), 0.f);
                            const uint32_t key = (__float_as_uint(dist) & 0xFFFFFC00u) | (lbase + j);
                            const uint32_t t = max(k1, key);
                            k1 = min(k1, key);
                            k2 = min(k2, t);
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty[buf]);
                }
                // ---- merge the two halves' candidates -> the frame's top 2 (truncated distance, then GLOBAL index)
                auto glob = [&](uint32_t kk, int h) { const int l = (int)(kk & 1023u); return (l / HG) * GROUP + h * HG + (l % HG); };
                float v1 = -__uint_as_float(k1 & 0xFFFFFC00u), v2 = -__uint_as_float(k2 & 0xFFFFFC00u);  // larger is better
                int i1 = glob(k1, half), i2 = glob(k2, half);
                my_x[1] = v1; my_x[2] = __int_as_float(i1); my_x[3] = v2; my_x[4] = __int_as_float(i2);
                epi_bar();
                top2_insert(peer_x[1], __float_as_int(peer_x[2]), v1, i1, v2, i2);
                top2_insert(peer_x[3], __float_as_int(peer_x[4]), v1, i1, v2, i2);
                // ---- exact fp32 re-score of both candidates: partial dots over this thread's D/2 dims
                const float* E = p.cb + (size_t)(p.stage0 + k) * p.n_codes * D;
                const float* e1 = E + (size_t)i1 * D + d0;
                const float* e2 = E + (size_t)i2 * D + d0;
                float p1 = 0.f, p2 = 0.f;
#pragma unroll 1
                for (int c = 0; c < HD / 16; ++c) {
                    uint32_t v[16];
                    tmem_ld16(r_addr + c * 16, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 a = __ldg(reinterpret_cast<const float4*>(e1 + c * 16) + q);
                        const float4 bq = __ldg(reinterpret_cast<const float4*>(e2 + c * 16) + q);
                        const float r0 = __uint_as_float(v[4 * q]), r1 = __uint_as_float(v[4 * q + 1]);
                        const float r2 = __uint_as_float(v[4 * q + 2]), r3 = __uint_as_float(v[4 * q + 3]);
                        p1 = fmaf(r0, a.x, p1); p1 = fmaf(r1, a.y, p1); p1 = fmaf(r2, a.z, p1); p1 = fmaf(r3, a.w, p1);
                        p2 = fmaf(r0, bq.x, p2); p2 = fmaf(r1, bq.y, p2); p2 = fmaf(r2, bq.z, p2); p2 = fmaf(r3, bq.w, p2);
                    }
                }
                my_x[5] = p1; my_x[6] = p2;
                const float en1 = en[i1], en2 = en[i2];  // read before the barrier: en is rewritten for stage k+1 after it
                epi_bar();
                const float dot1 = half == 0 ? p1 + peer_x[5] : peer_x[5] + p1;
                const float dot2 = half == 0 ? p2 + peer_x[6] : peer_x[6] + p2;
                const float s1 = ref_score(xn, dot1, en1, p.metric), s2 = ref_score(xn, dot2, en2, p.metric);
                const int sel = better(s2, i2, s1, i1) ? i2 : i1;
                if (half == 0 && valid) p.codes[row * p.code_stride + p.code_offset + k] = (int64_t)sel;
                // ---- subtract the winner in place (this thread's D/2 dims) and re-split the operand planes
                const float* es = E + (size_t)sel * D + d0;
                const bool more = k + 1 < p.stages;
                xn_part = 0.f;
#pragma unroll 1
                for (int c = 0; c < HD / 16; ++c) {
                    uint32_t v[16];
                    tmem_ld16(r_addr + c * 16, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 a = __ldg(reinterpret_cast<const float4*>(es + c * 16) + q);
                        v[4 * q] = __float_as_uint(__uint_as_float(v[4 * q]) - a.x);
                        v[4 * q + 1] = __float_as_uint(__uint_as_float(v[4 * q + 1]) - a.y);
                        v[4 * q + 2] = __float_as_uint(__uint_as_float(v[4 * q + 2]) - a.z);
                        v[4 * q + 3] = __float_as_uint(__uint_as_float(v[4 * q + 3]) - a.w);
                    }
                    tmem_st16(r_addr + c * 16, v);
                    if (more) xn_part += split16(v, c);  // all MMAs of this stage have retired (last tfull seen)
                }
                tmem_st_wait();
            }
            if (p.res_out) {  // tcgen05.ld is warp-collective: every lane reads, only valid frames store
#pragma unroll 1
                for (int c = 0; c < HD / 16; ++c) {
                    uint32_t v[16];
                    tmem_ld16(r_addr + c * 16, v);
                    tmem_ld_wait();
                    if (valid) {
                        float4* dst = reinterpret_cast<float4*>(p.res_out + row * D + d0 + c * 16);
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            dst[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                                                 __uint_as_float(v[4 * q + 3]));
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int D, int GROUP>
int launch(const CUtensorMap& emap, RvqParams p, cudaStream_t stream) {
    constexpr int per_sm = GROUP == 64 ? 2 : 1;
    const size_t a_planes = 2 * (size_t)(D / 64) * PLANE_KB_BYTES;
    const size_t fixed = 1024 + a_planes + (size_t)p.n_codes * 4 + TM * 2 * XCH * 4 + (2 * MAX_RING + 5) * 8 + 16;
    const size_t budget = per_sm == 2 ? 113 * 1024 : 227 * 1024;
    int ring = (int)((budget - fixed) / RING_STAGE_BYTES);
    if (ring > MAX_RING) ring = MAX_RING;
    if (ring < 2) { ac::set_error("ac_rvq_encode_tc: shared memory"); return -1; }
    p.ring = ring;
    const size_t smem = fixed + (size_t)ring * RING_STAGE_BYTES;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(rvq_encode_tc_kernel<D, GROUP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget);
        if (e != cudaSuccess) { ac::set_error("ac_rvq_encode_tc: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        configured = true;
    }
    int sms = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = p.tiles < per_sm * sms ? p.tiles : per_sm * sms;
    rvq_encode_tc_kernel<D, GROUP><<<grid, THREADS, smem, stream>>>(emap, p);
    return ac::finish_launch("ac_rvq_encode_tc");
}

}  // namespace

extern "C" int ac_rvq_encode_tc(const float* x, const void* cb_split_bf16, const float* codebooks, const float* cb_norm,
                                int64_t* codes_out, float* residual_out, int64_t rows, int32_t dim, int32_t n_codes,
                                int32_t stages, int32_t stage0, int32_t stages_total, int32_t code_stride, int32_t code_offset,
                                int32_t metric, void* stream) {
    AC_REQUIRE(x && cb_split_bf16 && codebooks && cb_norm && codes_out, "ac_rvq_encode_tc: null pointer");
    AC_REQUIRE(rows > 0 && stages > 0 && stage0 >= 0 && stage0 + stages <= stages_total, "ac_rvq_encode_tc: bad stage range");
    AC_REQUIRE(dim == 128 || dim == 256, "ac_rvq_encode_tc: dim %d (128 or 256)", dim);
    AC_REQUIRE(n_codes % GROUP_MAX == 0 && n_codes >= 2 * GROUP_MAX && n_codes <= MAX_CODES,
               "ac_rvq_encode_tc: n_codes %d must be a multiple of %d in [%d, %d]", n_codes, GROUP_MAX, 2 * GROUP_MAX, MAX_CODES);
    AC_REQUIRE(metric == 0 || metric == 1, "ac_rvq_encode_tc: metric %d", metric);
    AC_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)codebooks & 15) == 0 && ((uintptr_t)cb_split_bf16 & 15) == 0,
               "ac_rvq_encode_tc: pointers must be 16-byte aligned");
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            encode = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    AC_REQUIRE(encode, "ac_rvq_encode_tc: cuTensorMapEncodeTiled not available");
    CUtensorMap emap;
    {
        cuuint64_t gdim[2] = {(cuuint64_t)dim, (cuuint64_t)2 * stages_total * n_codes};  // hi plane rows, then lo plane rows
        cuuint64_t gstr[1] = {(cuuint64_t)dim * 2};
        cuuint32_t box[2] = {64, CHUNK};
        cuuint32_t est[2] = {1, 1};
        CUresult r = encode(&emap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(cb_split_bf16), gdim, gstr, box, est,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        AC_REQUIRE(r == CUDA_SUCCESS, "ac_rvq_encode_tc: cuTensorMapEncodeTiled failed: %d", (int)r);
    }
    RvqParams p{};
    p.x = x; p.cb = codebooks; p.cbn = cb_norm; p.codes = codes_out; p.res_out = residual_out;
    p.rows = rows; p.n_codes = n_codes; p.stages = stages; p.stage0 = stage0; p.code_stride = code_stride; p.code_offset = code_offset;
    p.tiles = (int)((rows + TM - 1) / TM);
    p.metric = metric;
    p.lo_row0 = stages_total * n_codes;
    // D = 128: two CTAs per SM (64-code distance buffers) once there is more than one tile per SM to overlap; AC_RVQ_PAIR=0/1 forces
    static int pair = -1;
    if (pair < 0) { const char* env = getenv("AC_RVQ_PAIR"); pair = env ? atoi(env) : 2; }
    int sms = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const bool use_pair = dim == 128 && (pair == 1 || (pair == 2 && p.tiles > sms));
    if (use_pair) return launch<128, 64>(emap, p, (cudaStream_t)stream);
    return dim == 128 ? launch<128, 128>(emap, p, (cudaStream_t)stream) : launch<256, 128>(emap, p, (cudaStream_t)stream);
}
