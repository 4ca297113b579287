// bf16 tensor-core tap-GEMM convolution for sm_100a (ac_conv_tc in include/audiocodecs_b200.h).
//
//   acc[b][m][n] = sum over sources s, taps j, channels k :  A_s[b][m + j*dil_s + shift_s][k] * W[n][kcol(s,j,k)]
//
// A_s are channels-last bf16 activation views [batch][rows][phases*c0] described by 4-D TMA tensor maps
// (c0, phase, row, batch): a stride-s conv with kernel 2s is the 2-tap GEMM over the view whose row is s
// consecutive time steps (no im2col, no padded copy; zero padding is TMA out-of-bounds fill, reflect padding
// lives in the producer-written halo rows of the activation buffer).  W is the packed K-major weight matrix
// [n_total][k_total] (optionally the error-compensating pair W_hi, W_lo).
//
// Data movement (v2):
//  * an A block = [G*128 + (taps-1)*dil rows] x [bk channels] is loaded ONCE per tile and contraction chunk;
//    every tap and every one of the G 128-row sub-tiles reads it through a tcgen05 shared-memory descriptor
//    whose start address is advanced by whole rows (measured: the swizzle is a function of the absolute
//    shared-memory address, so a row-shifted start needs no base offset -- scripts/probe_desc_shift.cu).
//    A K-tap conv therefore moves its activations once instead of K times, and G sub-tiles share one barrier
//    round trip and one weight block.
//  * weights: small matrices (EnCodec/Mimi <=64-channel layers) are RESIDENT in shared memory for the whole
//    kernel; large ones stream through their own ring, one [n_tile x bk] block (hi + lo) per tap and chunk,
//    shared by the G sub-tiles.
// Roles of the persistent CTA (one per SM): warp 0 = A producer, warp 1 = tcgen05.mma issuer (one lane),
// warp 2 = W producer, warps 3-18 = epilogue (tcgen05.ld -> bias, residual, raw / activated / lo-plane bf16
// or fp32 stores; the activation written is the CONSUMER's, so no layer re-reads a tensor just to activate it).
// TMEM holds up to two accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Chunked accumulation (flush_adds > 0, the "exact" encoder): the tcgen05 fp32 accumulator TRUNCATES every add (measured,
// scripts/accum_probe.py: the error of a K-deep contraction is a shrink of ~2e-8 per tcgen05.mma into the accumulator -- 2e-5
// at K = 4096 with three products, whatever the operand precision).  So a tile's contraction is cut into partial sums of a
// few dozen MMAs each: the two accumulator stages alternate as PARTIAL accumulators, and the epilogue warps add every
// finished partial into a running sum kept in a third TMEM region with round-to-nearest fp32 adds (tcgen05.ld / add /
// tcgen05.st by the thread that owns those lanes and columns).  The error per layer then no longer grows with K.
#include <cuda_bf16.h>

#include "common.cuh"
#define AC_MBAR_SUSPEND_NS 20000u   // throughput kernel: sleeping waits (see sm100.cuh)
#include "sm100.cuh"
#include "tc_common.cuh"

namespace {

using namespace sm100;
using namespace tcc;

constexpr int TILE_M = 128;
constexpr int MAX_A_STAGES = 8;
constexpr int MAX_W_STAGES = 6;
constexpr int EPI_WARPS = 16;  // four warps per TMEM lane quarter
constexpr int FIRST_EPI_WARP = 3;
constexpr int THREADS = 32 * (FIRST_EPI_WARP + EPI_WARPS);
constexpr int MAX_SRC = 4;
constexpr int SMEM_LIMIT = 227 * 1024;

struct TcSrc {
    int c0;        // innermost tensor-map dim (channels per phase)
    int taps, dil, shift;
    int chunks;    // k-blocks per tap
    int kb0;       // first weight k-block of this source
    int has_lo;    // lo plane present: adds the A_lo * W_hi product
    int pieces, box_rows;  // the A block is loaded as `pieces` TMA boxes of box_rows rows
};

struct TcMaps {
    CUtensorMap a[MAX_SRC];
    CUtensorMap a_lo[MAX_SRC];
    CUtensorMap w;
};

struct TcParams {
    TcSrc src[MAX_SRC];
    int n_src, bk, num_kb;
    int n_total, n_tile, n_tiles, m_rows, m_groups, batch, G;
    int a_stages, w_stages, acc_stages;
    int w_resident, w_split;
    int f16;       // hi planes of the A sources and the W_hi / W_lo planes are fp16 (else bf16); lo planes of A are always bf16
    int w_hib;     // f16 only: an extra bf16(W) plane (after W_hi [W_lo]) that the A_lo products multiply
    int y_f16, ya_f16, res_f16;   // hi-plane formats of the outputs / the residual
    uint32_t a_stage_bytes, a_plane_bytes;  // lo block sits a_plane_bytes after the hi block
    uint32_t w_stage_bytes, w_plane_bytes;  // streaming ring: plane q (W_hi, [W_lo], [bf16 W_hi]) sits q * w_plane_bytes after W_hi
    uint32_t w_kb_bytes, w_res_plane;       // resident: block kb at kb*w_kb_bytes, plane q at q * w_res_plane
    uint32_t tmem_cols;
    const float* bias;
    const float* alpha;
    const __nv_bfloat16* res;
    const __nv_bfloat16* res_lo;  // optional lo plane of the residual
    const float* res32;           // optional fp32 residual
    __nv_bfloat16* y;
    __nv_bfloat16* y_act;
    __nv_bfloat16* y_lo;          // optional lo planes: lo = bf16(v - float(bf16(v)))
    __nv_bfloat16* y_act_lo;
    float* y32;
    int flush_blocks, n_partials; // chunked accumulation: a partial sum per `flush_blocks` (chunk, tap) blocks; n_partials per tile (1 = off)
    uint32_t run_col;             // TMEM column of the running sums (after the two partial stages)
    int w_rows;                   // rows of one weight plane in the W tensor map (n_total); W_lo starts at row w_rows
    int act, epi, act_mod;
    int wide_ok;                  // every output/residual row run of 16 columns is 32-byte aligned: 256-bit stores
    long long y_bs, ya_bs, y32_bs, res_bs, out_shift, out_valid;
};

// Epilogue warps: each item is one tcgen05.ld of 32 rows x 16 accumulator columns (lane = row).  The sixteen warps
// (four per TMEM lane quarter) interleave over the (sub-tile, column chunk) items of the tile.
template <int ACT>
__device__ __forceinline__ void epilogue(const TcParams& p, uint32_t tmem_base, uint64_t* tfull, uint64_t* tempty, int warp, int lane,
                                         int total_tiles, const float* bias_s, const float* alpha_s, const float* ralpha_s) {
    const int quarter = warp & 3;                    // TMEM lane quarter this warp may read
    const int slot = (warp - FIRST_EPI_WARP) >> 2;   // 0..3
    const int chunks_n = p.n_tile / 16;
    const int items = p.G * chunks_n;
    const bool has_res = p.res != nullptr, has_res_lo = p.res_lo != nullptr, has_res32 = p.res32 != nullptr;
    const bool periodic = p.act_mod < p.n_total;
    const bool wide = p.wide_ok != 0;  // bias / alpha are indexed by n % act_mod (transposed conv: n = phase*C + c)
    const bool any_res = has_res || has_res32;
    // The residual of an item (64 bytes per lane: bf16 hi + lo runs, or one fp32 run) is fetched ONE ITEM AHEAD -- the first
    // item's before the wait on the accumulator -- so its DRAM latency runs under the MMAs / the previous item's math instead
    // of stalling every item (ncu: long_scoreboard was the top stall of the residual-adding launches).
    uint4 rnext[4] = {};
    auto fetch_res = [&](int b, int mg, int nt, int g, int c) {
        const int m = (mg * p.G + g) * TILE_M + quarter * 32 + lane;
        const int n0 = nt * p.n_tile + c * 16;
        const long long flat = (long long)m * p.n_total + n0 - p.out_shift;
        if (!(m < p.m_rows && wide && flat >= 0 && flat + 16 <= p.out_valid && n0 + 16 <= p.n_total)) return;
        if (has_res) {
            const uint4* r = reinterpret_cast<const uint4*>(p.res + (long long)b * p.res_bs + flat);
            rnext[0] = r[0]; rnext[1] = r[1];
            if (has_res_lo) {
                const uint4* l = reinterpret_cast<const uint4*>(p.res_lo + (long long)b * p.res_bs + flat);
                rnext[2] = l[0]; rnext[3] = l[1];
            }
        } else {
            const uint4* r = reinterpret_cast<const uint4*>(p.res32 + (long long)b * p.res_bs + flat);
            rnext[0] = r[0]; rnext[1] = r[1]; rnext[2] = r[2]; rnext[3] = r[3];
        }
    };
    int it = 0;  // accumulator hand-overs so far (= tiles, or partial sums with chunked accumulation)
    const uint32_t run_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + p.run_col;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int nt = tile % p.n_tiles;
        const int rest = tile / p.n_tiles;
        const int mg = rest % p.m_groups;
        const int b = rest / p.m_groups;
        // chunked accumulation: fold every partial but the last into the running sums (this thread's own lanes / columns:
        // the items of a warp are the same for every partial, so no other warp ever touches these TMEM cells)
        for (int q = 0; q + 1 < p.n_partials; ++q, ++it) {
            const int pas = it & 1;
            mbar_wait(&tfull[pas], (it >> 1) & 1);
            tc_fence_after();
            const uint32_t paddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + pas * p.G * p.n_tile;
            for (int item = slot; item < items; item += EPI_WARPS / 4) {
                uint32_t v[16];
                tmem_ld16(paddr + item * 16, v);
                if (q > 0) {
                    uint32_t r[16];
                    tmem_ld16(run_addr + item * 16, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(r[i]));
                } else {
                    tmem_ld_wait();
                }
                tmem_st16(run_addr + item * 16, v);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[pas]);
        }
        const int as = p.acc_stages == 2 ? (it & 1) : 0;
        const uint32_t tphase = p.acc_stages == 2 ? ((it >> 1) & 1) : (it & 1);
        int g = 0, c = slot;
        while (c >= chunks_n) { c -= chunks_n; ++g; }
        if (any_res && slot < items) fetch_res(b, mg, nt, g, c);
        mbar_wait(&tfull[as], tphase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * p.G * p.n_tile;
        for (int item = slot; item < items; item += EPI_WARPS / 4) {
            uint32_t v[16];
            tmem_ld16(taddr + g * p.n_tile + c * 16, v);
            if (p.n_partials > 1) {  // last partial + running sums (item * 16 == g * n_tile + c * 16)
                uint32_t r[16];
                tmem_ld16(run_addr + g * p.n_tile + c * 16, r);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(r[i]));
            }
            const int m = (mg * p.G + g) * TILE_M + quarter * 32 + lane;
            const int n0 = nt * p.n_tile + c * 16;
            const long long flat = (long long)m * p.n_total + n0 - p.out_shift;
            const bool ok = m < p.m_rows && n0 < p.n_total;
            int ch0 = n0;  // channel of column n0
            if (periodic) ch0 = n0 % p.act_mod;
            c += EPI_WARPS / 4;
            while (c >= chunks_n) { c -= chunks_n; ++g; }
            const uint4 rcur[4] = {rnext[0], rnext[1], rnext[2], rnext[3]};
            if (any_res && item + EPI_WARPS / 4 < items) fetch_res(b, mg, nt, g, c);
            tmem_ld_wait();
            if (!ok) continue;
            if (wide && flat >= 0 && flat + 16 <= p.out_valid && n0 + 16 <= p.n_total) {
                // fast path: the lane's 16 columns are one aligned 32-byte (bf16) / 64-byte (fp32) run per output plane
                float o[16];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    int ch = ch0 + h * 8;
                    if (ch >= p.act_mod) ch -= p.act_mod;
                    const float4 b0 = *reinterpret_cast<const float4*>(bias_s + ch);
                    const float4 b1 = *reinterpret_cast<const float4*>(bias_s + ch + 4);
                    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) o[h * 8 + i] = __uint_as_float(v[h * 8 + i]) + bb[i];
                }
                if (p.epi == AC_EPI_GELU) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) o[i] = ac::gelu_erf(o[i]);
                } else if (p.epi == AC_EPI_TANH) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) o[i] = tanhf(o[i]);
                }
                if (has_res) {
                    add_bf16x16(o, rcur[0], rcur[1], p.res_f16 != 0);
                    if (has_res_lo) add_bf16x16(o, rcur[2], rcur[3]);
                }
                if (has_res32) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        o[4 * q] += __uint_as_float(rcur[q].x); o[4 * q + 1] += __uint_as_float(rcur[q].y);
                        o[4 * q + 2] += __uint_as_float(rcur[q].z); o[4 * q + 3] += __uint_as_float(rcur[q].w);
                    }
                }
                if (p.y32) {
                    float* d = p.y32 + (long long)b * p.y32_bs + flat;
                    st_global_256(d, __float_as_uint(o[0]), __float_as_uint(o[1]), __float_as_uint(o[2]), __float_as_uint(o[3]),
                                  __float_as_uint(o[4]), __float_as_uint(o[5]), __float_as_uint(o[6]), __float_as_uint(o[7]));
                    st_global_256(d + 8, __float_as_uint(o[8]), __float_as_uint(o[9]), __float_as_uint(o[10]), __float_as_uint(o[11]),
                                  __float_as_uint(o[12]), __float_as_uint(o[13]), __float_as_uint(o[14]), __float_as_uint(o[15]));
                }
                if (p.y) store_bf16x16(o, p.y + (long long)b * p.y_bs + flat, p.y_lo ? p.y_lo + (long long)b * p.y_bs + flat : nullptr, p.y_f16 != 0);
                if (p.y_act) {
                    if (ACT == AC_ACT_ELU) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) o[i] = elu_ex2(o[i]);
                    } else if (ACT == AC_ACT_SNAKE) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            int ch = ch0 + h * 8;
                            if (ch >= p.act_mod) ch -= p.act_mod;
                            const float4 a0 = *reinterpret_cast<const float4*>(alpha_s + ch);
                            const float4 a1 = *reinterpret_cast<const float4*>(alpha_s + ch + 4);
                            const float4 r0 = *reinterpret_cast<const float4*>(ralpha_s + ch);
                            const float4 r1 = *reinterpret_cast<const float4*>(ralpha_s + ch + 4);
                            const float al[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                            const float ra[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float sn = __sinf(al[i] * o[h * 8 + i]);
                                o[h * 8 + i] = fmaf(ra[i], sn * sn, o[h * 8 + i]);
                            }
                        }
                    }
                    store_bf16x16(o, p.y_act + (long long)b * p.ya_bs + flat, p.y_act_lo ? p.y_act_lo + (long long)b * p.ya_bs + flat : nullptr, p.ya_f16 != 0);
                }
                continue;
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {  // two 8-element vectors (16 B of bf16 each)
                const long long f = flat + h * 8;
                const int n = n0 + h * 8;
                if (f < 0 || f >= p.out_valid || n >= p.n_total) continue;
                int ch = ch0 + h * 8;
                if (ch >= p.act_mod) ch -= p.act_mod;
                float o[8];
                {
                    const float4 b0 = *reinterpret_cast<const float4*>(bias_s + ch);
                    const float4 b1 = *reinterpret_cast<const float4*>(bias_s + ch + 4);
                    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) o[i] = __uint_as_float(v[h * 8 + i]) + bb[i];
                }
                if (p.epi == AC_EPI_GELU) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) o[i] = ac::gelu_erf(o[i]);
                } else if (p.epi == AC_EPI_TANH) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) o[i] = tanhf(o[i]);
                }
                if (has_res) {
                    add_bf16x8(o, p.res + (long long)b * p.res_bs + f, p.res_f16 != 0);
                    if (has_res_lo) add_bf16x8(o, p.res_lo + (long long)b * p.res_bs + f);
                }
                if (has_res32) {
                    const float4* r = reinterpret_cast<const float4*>(p.res32 + (long long)b * p.res_bs + f);
                    const float4 r0 = r[0], r1 = r[1];
                    o[0] += r0.x; o[1] += r0.y; o[2] += r0.z; o[3] += r0.w;
                    o[4] += r1.x; o[5] += r1.y; o[6] += r1.z; o[7] += r1.w;
                }
                if (p.y32) {
                    float4* d = reinterpret_cast<float4*>(p.y32 + (long long)b * p.y32_bs + f);
                    d[0] = make_float4(o[0], o[1], o[2], o[3]);
                    d[1] = make_float4(o[4], o[5], o[6], o[7]);
                }
                if (p.y) {
                    const uint4 q = pack8(o, p.y_f16 != 0);
                    *reinterpret_cast<uint4*>(p.y + (long long)b * p.y_bs + f) = q;
                    if (p.y_lo) *reinterpret_cast<uint4*>(p.y_lo + (long long)b * p.y_bs + f) = pack_lo(o, q, p.y_f16 != 0);
                }
                if (p.y_act) {
                    float a[8];
                    if (ACT == AC_ACT_ELU) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) a[i] = elu_ex2(o[i]);
                    } else if (ACT == AC_ACT_SNAKE) {
                        // act_mod % 8 == 0, so a vector of 8 columns never wraps around the channel axis
                        const float4 a0 = *reinterpret_cast<const float4*>(alpha_s + ch);
                        const float4 a1 = *reinterpret_cast<const float4*>(alpha_s + ch + 4);
                        const float4 r0 = *reinterpret_cast<const float4*>(ralpha_s + ch);
                        const float4 r1 = *reinterpret_cast<const float4*>(ralpha_s + ch + 4);
                        const float al[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                        const float ra[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float sn = __sinf(al[i] * o[i]);
                            a[i] = fmaf(ra[i], sn * sn, o[i]);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) a[i] = o[i];
                    }
                    const uint4 q = pack8(a, p.ya_f16 != 0);
                    *reinterpret_cast<uint4*>(p.y_act + (long long)b * p.ya_bs + f) = q;
                    if (p.y_act_lo) *reinterpret_cast<uint4*>(p.y_act_lo + (long long)b * p.ya_bs + f) = pack_lo(a, q, p.ya_f16 != 0);
                }
            }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[as]);
    }
}

// MMA issuer warp.  The whole warp walks the (uniform) loop nest and waits on the barriers; one elected lane issues
// tcgen05.mma / tcgen05.commit.  Keeping control flow and operands warp-uniform lets ptxas hold the descriptors in
// uniform registers -- a divergent single-lane loop costs ~25 SASS instructions per MMA (measured: the issuer thread,
// not the tensor pipe or HBM, bounded the <=64-channel layers).
template <int KSTEPS>
__device__ __forceinline__ void mma_role(const TcParams& p, uint8_t* a_ring, uint8_t* w_area, uint64_t* a_full, uint64_t* a_empty,
                                         uint64_t* w_full, uint64_t* w_empty, uint64_t* tfull, uint64_t* tempty, uint64_t* wres_bar,
                                         uint32_t tmem_base, int total_tiles) {
    const bool leader = elect_one();
    int astage = 0, wstage = 0;
    uint32_t aphase = 0, wphase = 0;
    const uint32_t idesc = p.f16 ? make_idesc_f16(TILE_M, p.n_tile) : make_idesc_bf16(TILE_M, p.n_tile);  // hi-plane products
    const uint32_t idesc_lo = make_idesc_bf16(TILE_M, p.n_tile);  // A_lo (always bf16) x bf16(W_hi)
    const uint32_t hib_plane = p.w_hib ? (uint32_t)(1 + p.w_split) : 0u;  // which W plane the A_lo product reads
    const uint32_t row_bytes = KSTEPS * 32;
    const uint64_t desc_base = make_smem_desc(0, row_bytes);  // everything but the start address
    const uint32_t a_ring_u = smem_u32(a_ring), w_area_u = smem_u32(w_area);
    if (p.w_resident) { mbar_wait(wres_bar, 0); tc_fence_after(); }
    int it = 0;  // accumulator hand-overs so far (= tiles, or partial sums with chunked accumulation)
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        int as = p.acc_stages == 2 ? (it & 1) : 0;
        uint32_t tphase = p.acc_stages == 2 ? ((it >> 1) & 1) : (it & 1);
        mbar_wait(&tempty[as], tphase ^ 1);
        tc_fence_after();
        uint32_t d_base = tmem_base + as * p.G * p.n_tile;
        uint32_t acc = 0;  // first MMA into each accumulator overwrites
        int blk = 0;       // (chunk, tap) blocks issued into the current partial sum
        for (int s = 0; s < p.n_src; ++s) {
            const TcSrc& S = p.src[s];
            const int taps = S.taps, chunks = S.chunks, has_lo = S.has_lo;
            const uint32_t dil_bytes = S.dil * row_bytes;
            for (int cc = 0; cc < chunks; ++cc) {
                mbar_wait(&a_full[astage], aphase);
                tc_fence_after();
                const uint32_t a_hi = a_ring_u + astage * p.a_stage_bytes;
                for (int j = 0; j < taps; ++j) {
                    if (p.flush_blocks && blk == p.flush_blocks) {
                        // chunked accumulation: hand this partial sum to the epilogue warps, continue in the other stage
                        if (leader) umma_commit(&tfull[as]);
                        __syncwarp();
                        ++it;
                        as = it & 1;
                        tphase = (it >> 1) & 1;
                        mbar_wait(&tempty[as], tphase ^ 1);
                        tc_fence_after();
                        d_base = tmem_base + as * p.G * p.n_tile;
                        acc = 0;
                        blk = 0;
                    }
                    ++blk;
                    uint32_t w_hi, w_lo, w_hib;
                    if (p.w_resident) {
                        w_hi = w_area_u + (S.kb0 + j * chunks + cc) * p.w_kb_bytes;
                        w_lo = w_hi + p.w_res_plane;
                        w_hib = w_hi + hib_plane * p.w_res_plane;
                    } else {
                        mbar_wait(&w_full[wstage], wphase);
                        tc_fence_after();
                        w_hi = w_area_u + wstage * p.w_stage_bytes;
                        w_lo = w_hi + p.w_plane_bytes;
                        w_hib = w_hi + hib_plane * p.w_plane_bytes;
                    }
                    const uint64_t bdesc_hi = desc_base | ((w_hi & 0x3FFFFu) >> 4);
                    const uint64_t bdesc_lo = desc_base | ((w_lo & 0x3FFFFu) >> 4);
                    const uint64_t bdesc_hib = desc_base | ((w_hib & 0x3FFFFu) >> 4);
                    uint32_t a_addr = a_hi + j * dil_bytes;
                    uint32_t d_tmem = d_base;
                    for (int g = 0; g < p.G; ++g, a_addr += TILE_M * row_bytes, d_tmem += p.n_tile) {
                        const uint64_t adesc = desc_base | ((a_addr & 0x3FFFFu) >> 4);
                        if (leader) {
#pragma unroll
                            for (int k = 0; k < KSTEPS; ++k) umma_bf16(d_tmem, adesc + 2 * k, bdesc_hi + 2 * k, idesc, k == 0 ? acc : 1u);  // +32 B per K=16
                            if (p.w_split) {  // error-compensated weights: A * W_lo into the same accumulator
#pragma unroll
                                for (int k = 0; k < KSTEPS; ++k) umma_bf16(d_tmem, adesc + 2 * k, bdesc_lo + 2 * k, idesc, 1u);
                            }
                            if (has_lo) {  // split activations: A_lo * W_hi (A_lo * W_lo is below fp32 noise); bf16 x bf16
                                const uint64_t adesc_lo = desc_base | (((a_addr + p.a_plane_bytes) & 0x3FFFFu) >> 4);
#pragma unroll
                                for (int k = 0; k < KSTEPS; ++k) umma_bf16(d_tmem, adesc_lo + 2 * k, bdesc_hib + 2 * k, idesc_lo, 1u);
                            }
                        }
                    }
                    acc = 1u;
                    if (!p.w_resident) {
                        if (leader) umma_commit(&w_empty[wstage]);
                        if (++wstage == p.w_stages) { wstage = 0; wphase ^= 1; }
                    }
                }
                if (leader) umma_commit(&a_empty[astage]);  // frees the A block when the MMAs retire
                if (++astage == p.a_stages) { astage = 0; aphase ^= 1; }
            }
        }
        if (leader) umma_commit(&tfull[as]);  // accumulators complete
        __syncwarp();
    }
}

__global__ void __launch_bounds__(THREADS, 1)
conv_tc_kernel(const __grid_constant__ TcMaps maps, const __grid_constant__ TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    // dynamic smem base is only guaranteed 16-B aligned: round up to 1024 for the swizzled tiles
    uint8_t* smem = align_smem_1024(smem_raw);
    uint8_t* a_ring = smem;
    uint8_t* w_area = smem + (size_t)p.a_stages * p.a_stage_bytes;
    const size_t w_area_bytes = p.w_resident ? (size_t)p.w_res_plane * (1 + p.w_split + p.w_hib) : (size_t)p.w_stages * p.w_stage_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(w_area + w_area_bytes);
    uint64_t* a_full = bars;
    uint64_t* a_empty = a_full + MAX_A_STAGES;
    uint64_t* w_full = a_empty + MAX_A_STAGES;
    uint64_t* w_empty = w_full + MAX_W_STAGES;
    uint64_t* tfull = w_empty + MAX_W_STAGES;
    uint64_t* tempty = tfull + 2;
    uint64_t* wres_bar = tempty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wres_bar + 1);
    float* bias_s = reinterpret_cast<float*>(bars + 36);       // [act_mod]; 288 B into the (1024-aligned) barrier block
    float* alpha_s = bias_s + p.act_mod;                       // [act_mod] snake only
    float* ralpha_s = alpha_s + p.act_mod;                     // 1 / (alpha + 1e-9)

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform role index
    const int lane = threadIdx.x & 31;
    const int total_tiles = p.n_tiles * p.m_groups * p.batch;
    const uint32_t row_bytes = p.bk * 2;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < p.n_src; ++s) {
            prefetch_tensormap(&maps.a[s]);
            if (p.src[s].has_lo) prefetch_tensormap(&maps.a_lo[s]);
        }
        prefetch_tensormap(&maps.w);
        for (int i = 0; i < p.a_stages; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < p.w_stages; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], EPI_WARPS); }
        mbar_init(wres_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
    for (int i = threadIdx.x; i < p.act_mod; i += THREADS) {
        bias_s[i] = p.bias ? p.bias[i] : 0.f;
        if (p.act == AC_ACT_SNAKE) {
            const float a = p.alpha[i];
            alpha_s[i] = a;
            ralpha_s[i] = 1.0f / (a + 1e-9f);
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================================================= A producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int rest = tile / p.n_tiles;
                const int mg = rest % p.m_groups;
                const int b = rest / p.m_groups;
                for (int s = 0; s < p.n_src; ++s) {
                    const TcSrc& S = p.src[s];
                    const int row0 = mg * p.G * TILE_M + S.shift;
                    const uint32_t piece_bytes = S.box_rows * row_bytes;
                    for (int cc = 0; cc < S.chunks; ++cc) {
                        mbar_wait(&a_empty[stage], phase ^ 1);
                        mbar_arrive_expect_tx(&a_full[stage], S.pieces * piece_bytes * (1 + S.has_lo));
                        const int flat = cc * p.bk;
                        uint8_t* dst = a_ring + (size_t)stage * p.a_stage_bytes;
                        for (int pc = 0; pc < S.pieces; ++pc) {
                            tma_load_4d(dst + pc * piece_bytes, &maps.a[s], &a_full[stage], flat % S.c0, flat / S.c0,
                                        row0 + pc * S.box_rows, b);
                            if (S.has_lo)
                                tma_load_4d(dst + p.a_plane_bytes + pc * piece_bytes, &maps.a_lo[s], &a_full[stage], flat % S.c0,
                                            flat / S.c0, row0 + pc * S.box_rows, b);
                        }
                        if (++stage == p.a_stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 2) {
        // ================================================================= W producer
        if (lane == 0) {
            const uint32_t blk_bytes = p.n_tile * row_bytes;
            if (p.w_resident) {
                // whole matrix once: block kb = [n_total x bk] at kb * w_kb_bytes, lo plane after all hi blocks
                const int planes = 1 + p.w_split + p.w_hib;
                mbar_arrive_expect_tx(wres_bar, (uint32_t)p.num_kb * blk_bytes * planes);
                for (int kb = 0; kb < p.num_kb; ++kb)
                    for (int q = 0; q < planes; ++q)
                        tma_load_2d(w_area + (size_t)q * p.w_res_plane + (size_t)kb * p.w_kb_bytes, &maps.w, wres_bar, kb * p.bk, q * p.w_rows);
            } else {
                int stage = 0;
                uint32_t phase = 0;
                for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                    const int nt = tile % p.n_tiles;
                    for (int s = 0; s < p.n_src; ++s) {
                        const TcSrc& S = p.src[s];
                        for (int cc = 0; cc < S.chunks; ++cc) {
                            for (int j = 0; j < S.taps; ++j) {
                                const int kb = S.kb0 + j * S.chunks + cc;
                                mbar_wait(&w_empty[stage], phase ^ 1);
                                const int planes = 1 + p.w_split + p.w_hib;
                                mbar_arrive_expect_tx(&w_full[stage], blk_bytes * planes);
                                uint8_t* dst = w_area + (size_t)stage * p.w_stage_bytes;
                                for (int q = 0; q < planes; ++q)
                                    tma_load_2d(dst + (size_t)q * p.w_plane_bytes, &maps.w, &w_full[stage], kb * p.bk, q * p.w_rows + nt * p.n_tile);
                                if (++stage == p.w_stages) { stage = 0; phase ^= 1; }
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================================================================= MMA issuer
        if (p.bk == 64) mma_role<4>(p, a_ring, w_area, a_full, a_empty, w_full, w_empty, tfull, tempty, wres_bar, tmem_base, total_tiles);
        else if (p.bk == 32) mma_role<2>(p, a_ring, w_area, a_full, a_empty, w_full, w_empty, tfull, tempty, wres_bar, tmem_base, total_tiles);
        else mma_role<1>(p, a_ring, w_area, a_full, a_empty, w_full, w_empty, tfull, tempty, wres_bar, tmem_base, total_tiles);
    } else {
        // ================================================================= epilogue (warps 3..18)
        if (p.act == AC_ACT_ELU) epilogue<AC_ACT_ELU>(p, tmem_base, tfull, tempty, warp, lane, total_tiles, bias_s, alpha_s, ralpha_s);
        else if (p.act == AC_ACT_SNAKE) epilogue<AC_ACT_SNAKE>(p, tmem_base, tfull, tempty, warp, lane, total_tiles, bias_s, alpha_s, ralpha_s);
        else epilogue<AC_ACT_NONE>(p, tmem_base, tfull, tempty, warp, lane, total_tiles, bias_s, alpha_s, ralpha_s);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

// ------------------------------------------------------------------------------------------------ host
struct MergedSrc {
    const ac_tc_src* hi;
    const ac_tc_src* lo;
};

}  // namespace

extern "C" int ac_conv_tc(const ac_conv_tc_desc* d, void* stream) {
    AC_REQUIRE(d && d->w && d->n_src >= 1 && d->n_src <= MAX_SRC, "ac_conv_tc: bad descriptor");
    AC_REQUIRE(d->bk == 16 || d->bk == 32 || d->bk == 64, "ac_conv_tc: bk %d", d->bk);
    AC_REQUIRE(d->batch > 0 && d->m_rows > 0 && d->n_total >= 16 && d->n_total % 8 == 0 && d->n_total <= 8192,
               "ac_conv_tc: bad sizes (batch %d rows %d n %d)", d->batch, d->m_rows, d->n_total);
    AC_REQUIRE(d->y || d->y_act || d->y32, "ac_conv_tc: no output");
    AC_REQUIRE(d->out_shift % 8 == 0 && d->out_valid % 8 == 0, "ac_conv_tc: out_shift/out_valid must be multiples of 8");
    AC_REQUIRE(d->act != AC_ACT_SNAKE || (d->alpha && d->act_mod > 0 && d->act_mod % 8 == 0 && ((uintptr_t)d->alpha & 15) == 0),
               "ac_conv_tc: snake needs a 16-byte aligned alpha and act_mod % 8 == 0");
    AC_REQUIRE(!d->bias || ((uintptr_t)d->bias & 15) == 0, "ac_conv_tc: bias not 16-byte aligned");
    EncodeTiledFn encode = get_encode();
    AC_REQUIRE(encode, "ac_conv_tc: cuTensorMapEncodeTiled not available");

    // lo-plane twins (lo_of >= 0) are folded into their hi source: one A block carries both planes
    MergedSrc ms[MAX_SRC];
    int map_of[MAX_SRC];
    int n_ms = 0;
    for (int s = 0; s < d->n_src; ++s) {
        const ac_tc_src& S = d->src[s];
        AC_REQUIRE(S.base && S.c0 > 0 && S.phases > 0 && S.rows > 0 && S.taps > 0 && S.dilation > 0, "ac_conv_tc: bad source %d", s);
        AC_REQUIRE(((uintptr_t)S.base & 15) == 0 && (S.phase_stride * 2) % 16 == 0 && (S.row_stride * 2) % 16 == 0 &&
                       (S.batch_stride * 2) % 16 == 0, "ac_conv_tc: source %d not 16-byte aligned", s);
        if (S.lo_of >= 0) {
            AC_REQUIRE(S.lo_of < s && d->src[S.lo_of].lo_of < 0, "ac_conv_tc: source %d: bad lo_of", s);
            const ac_tc_src& H = d->src[S.lo_of];
            AC_REQUIRE(H.c0 == S.c0 && H.phases == S.phases && H.taps == S.taps && H.rows == S.rows && H.dilation == S.dilation &&
                           H.shift == S.shift, "ac_conv_tc: source %d is not the lo twin of %d", s, S.lo_of);
            AC_REQUIRE(ms[map_of[S.lo_of]].lo == nullptr, "ac_conv_tc: source %d has two lo twins", S.lo_of);
            ms[map_of[S.lo_of]].lo = &S;
            map_of[s] = map_of[S.lo_of];
        } else {
            map_of[s] = n_ms;
            ms[n_ms].hi = &S;
            ms[n_ms].lo = nullptr;
            ++n_ms;
        }
    }

    const int w_split = d->w_split ? 1 : 0;
    const int f16 = (d->fmt & AC_FMT_A_F16) ? 1 : 0, w_hib = (d->fmt & AC_FMT_W_HIB) ? 1 : 0;
    AC_REQUIRE(!w_hib || f16, "ac_conv_tc: the extra bf16(W) plane only exists for fp16 operands");
    const int w_planes = 1 + w_split + w_hib;
    // N tile: the largest of {256,...,16} that does not over-pad; split weights hold two W tiles per stage -> cap at 128
    int n_tile = 16;
    for (int c : {256, 192, 128, 96, 64, 48, 32, 16})
        if ((w_planes == 1 || c <= 128) && d->n_total % c == 0) { n_tile = c; break; }
    if (d->n_total % n_tile != 0) n_tile = d->n_total >= 128 ? 128 : 16;
    if (d->n_tile_hint > 0) n_tile = d->n_tile_hint;
    // chunked accumulation: the two accumulator stages become partial sums and a third region holds the running sums, so
    // 3 * G * n_tile TMEM columns are needed
    int n_blocks = 0;
    for (int s = 0; s < d->n_src; ++s)
        if (d->src[s].lo_of < 0) n_blocks += d->src[s].taps * (d->src[s].c0 * d->src[s].phases);  // contraction length, for now
    const bool want_flush = d->flush_adds > 0 && 2 * (n_blocks / 16) * (1 + w_split + 1) > 3 * d->flush_adds;  // > 1.5 partials' worth
    if (want_flush && n_tile > 128 && d->n_tile_hint <= 0) n_tile = d->n_total % 128 == 0 ? 128 : (d->n_total % 96 == 0 ? 96 : (d->n_total % 64 == 0 ? 64 : n_tile));
    AC_REQUIRE(n_tile % 16 == 0 && n_tile >= 16 && n_tile <= 256, "ac_conv_tc: n_tile %d", n_tile);
    const int n_tiles = (d->n_total + n_tile - 1) / n_tile;

    int k_total = 0;
    bool any_lo = false;
    int max_halo = 0;
    for (int s = 0; s < n_ms; ++s) {
        k_total += ms[s].hi->taps * ms[s].hi->c0 * ms[s].hi->phases;
        any_lo |= ms[s].lo != nullptr;
        const int halo = (ms[s].hi->taps - 1) * ms[s].hi->dilation;
        if (halo > max_halo) max_halo = halo;
    }
    AC_REQUIRE(k_total == d->k_total, "ac_conv_tc: k_total %d != sum of taps*channels %d", d->k_total, k_total);
    AC_REQUIRE(!(f16 && any_lo) || w_hib, "ac_conv_tc: fp16 operands with a lo plane need the extra bf16(W) plane (AC_FMT_W_HIB)");
    AC_REQUIRE(((uintptr_t)d->w & 15) == 0 && (k_total * 2) % 16 == 0, "ac_conv_tc: weights not 16-byte aligned");

    const int act_mod = d->act_mod > 0 ? d->act_mod : d->n_total;
    AC_REQUIRE(act_mod % 8 == 0 && act_mod <= d->n_total && d->n_total % act_mod == 0, "ac_conv_tc: act_mod %d must divide n_total and be a multiple of 8", act_mod);
    // ---- choose (G, bk, stages) so that everything fits shared memory
    const size_t fixed = 1024 /*align slack*/ + 36 * 8 /*barriers + tmem slot*/ + (size_t)act_mod * 4 * (d->act == AC_ACT_SNAKE ? 3 : 1) + 64;
    const long long m_tiles = (d->m_rows + TILE_M - 1) / TILE_M;
    TcParams p{};
    const int sms = sm_count();
    // The contraction block bk fixes the order in which the (hi, lo) products of a block are accumulated, i.e. the fp32
    // rounding of the result.  It must not depend on the batch size or on the tile grouping G a tuner asks for (a clip's
    // tokens would then depend on its neighbours): the canonical bk is the one the G = 1 search settles on -- a function of
    // the layer shape alone -- and a grouped tiling is accepted only if it fits with that same bk.
    auto search = [&](int g_only, int bk_only, bool use_hint) -> bool {
    // pass 0 insists on deep rings (>= 3 A stages and >= 4 W stages, or resident weights); pass 1 takes anything that fits
    for (int pass = 0; pass < 2; ++pass) {
        for (int bk = d->bk; bk >= 16; bk >>= 1) {
            if (bk_only > 0 && bk != bk_only) continue;
            bool ok = true;
            for (int s = 0; s < n_ms; ++s) ok &= ms[s].hi->c0 % bk == 0;  // a k-block never straddles two phases
            if (!ok) continue;
            const int num_kb = k_total / bk;
            const uint32_t w_kb_bytes = round_up((uint32_t)d->n_total * bk * 2, 1024);
            const size_t w_res_total = (size_t)num_kb * w_kb_bytes * w_planes;
            const bool resident = n_tiles == 1 && n_tile == d->n_total && w_res_total <= 96 * 1024;
            for (int G : {4, 2, 1}) {
                if (g_only > 0 && G != g_only) continue;
                if (use_hint && d->g_hint > 0 && G != d->g_hint) continue;
                if (!(use_hint && d->g_hint > 0) && G > 1) {
                    if (G * n_tile * 2 > 512) continue;                                   // keep two accumulator stages when grouping
                    if ((m_tiles / G) * n_tiles * d->batch < 2LL * sms) continue;         // not enough tiles to fill the chip
                }
                if (G * n_tile > 512 || (want_flush && 3 * G * n_tile > 512)) continue;
                uint32_t a_plane = 0;  // A stage: the largest block over sources
                for (int s = 0; s < n_ms; ++s) {
                    const int R = G * TILE_M + (ms[s].hi->taps - 1) * ms[s].hi->dilation;
                    const int pieces = (R + 255) / 256;
                    const int box_rows = (int)round_up((R + pieces - 1) / pieces, 8);
                    const uint32_t bytes = round_up((uint32_t)pieces * box_rows * bk * 2, 1024);
                    if (bytes > a_plane) a_plane = bytes;
                }
                const uint32_t a_stage = a_plane * (any_lo ? 2 : 1);
                const uint32_t w_plane = round_up((uint32_t)n_tile * bk * 2, 1024);
                const uint32_t w_stage = w_plane * w_planes;
                const size_t budget = SMEM_LIMIT - fixed;
                int a_stages, w_stages;
                if (resident) {
                    if (w_res_total + 2 * (size_t)a_stage > budget) continue;
                    a_stages = (int)((budget - w_res_total) / a_stage);
                    w_stages = 0;
                } else {
                    w_stages = 4;
                    while (w_stages > 2 && (size_t)w_stages * w_stage + (pass == 0 ? 3 : 2) * (size_t)a_stage > budget) --w_stages;
                    if ((size_t)w_stages * w_stage + 2 * (size_t)a_stage > budget) continue;
                    a_stages = (int)((budget - (size_t)w_stages * w_stage) / a_stage);
                    if (pass == 0 && (w_stages < 4 || a_stages < 3)) continue;
                    if (a_stages > 4) {  // spare room: deepen the W ring first, it turns over `taps` times faster
                        const int extra = (int)((budget - (size_t)w_stages * w_stage - 4 * (size_t)a_stage) / w_stage);
                        w_stages = w_stages + extra > MAX_W_STAGES ? MAX_W_STAGES : w_stages + extra;
                        a_stages = (int)((budget - (size_t)w_stages * w_stage) / a_stage);
                    }
                }
                if (a_stages > MAX_A_STAGES) a_stages = MAX_A_STAGES;
                if (a_stages < 2) continue;
                p.G = G; p.bk = bk; p.num_kb = num_kb;
                p.a_stages = a_stages; p.w_stages = w_stages;
                p.a_stage_bytes = a_stage; p.a_plane_bytes = a_plane;
                p.w_stage_bytes = w_stage; p.w_plane_bytes = w_plane;
                p.w_kb_bytes = w_kb_bytes; p.w_res_plane = (uint32_t)((size_t)num_kb * w_kb_bytes);
                p.w_resident = resident ? 1 : 0;
                return true;
            }
        }
    }
    return false;
    };
    bool found = search(1, 0, false);
    if (found) {
        const int bk_canonical = p.bk;
        // a grouped tiling (hinted or heuristic) with the canonical bk, else the G = 1 tiling just found -- unless a hint
        // asked for a specific G that does not fit: then this variant is rejected (ConfigError, nothing launched)
        if (!(d->g_hint == 1) && !search(0, bk_canonical, true)) found = d->g_hint > 0 ? false : search(1, bk_canonical, false);
    }
    AC_REQUIRE(found, "ac_conv_tc: no tiling fits shared memory (n_total %d k_total %d)", d->n_total, k_total);
    const int bk = p.bk;
    p.w_split = w_split; p.f16 = f16; p.w_hib = w_hib;
    p.y_f16 = (d->fmt & AC_FMT_Y_F16) ? 1 : 0; p.ya_f16 = (d->fmt & AC_FMT_YACT_F16) ? 1 : 0; p.res_f16 = (d->fmt & AC_FMT_RES_F16) ? 1 : 0;
    p.acc_stages = 2 * p.G * n_tile <= 512 ? 2 : 1;
    p.flush_blocks = 0; p.n_partials = 1; p.run_col = 0;
    if (want_flush && 3 * p.G * n_tile <= 512) {
        int blocks = 0;  // (chunk, tap) blocks per tile, each bk/16 k-steps x products MMAs per sub-tile
        for (int s = 0; s < n_ms; ++s) blocks += ms[s].hi->taps * (ms[s].hi->c0 * ms[s].hi->phases / p.bk);
        const int per_block = (p.bk / 16) * (1 + w_split + (any_lo ? 1 : 0));
        int fb = d->flush_adds / per_block;
        if (fb < 1) fb = 1;
        if (blocks > fb) {
            p.flush_blocks = fb;
            p.n_partials = (blocks + fb - 1) / fb;
            p.run_col = 2u * p.G * n_tile;
            p.acc_stages = 2;
        }
    }
    uint32_t cols = 32;
    while (cols < (uint32_t)((p.acc_stages + (p.n_partials > 1 ? 1 : 0)) * p.G * n_tile)) cols <<= 1;
    p.tmem_cols = cols;

    TcMaps maps;
    p.n_src = n_ms;
    int kcol = 0;
    for (int s = 0; s < n_ms; ++s) {
        const ac_tc_src& S = *ms[s].hi;
        const int kper = S.c0 * S.phases;  // contraction length per tap
        TcSrc& T = p.src[s];
        T.c0 = S.c0; T.taps = S.taps; T.dil = S.dilation; T.shift = S.shift;
        T.chunks = kper / bk;
        T.kb0 = kcol / bk;
        T.has_lo = ms[s].lo ? 1 : 0;
        kcol += S.taps * kper;
        const int R = p.G * TILE_M + (S.taps - 1) * S.dilation;
        T.pieces = (R + 255) / 256;
        T.box_rows = (int)round_up((R + T.pieces - 1) / T.pieces, 8);
        for (int plane = 0; plane < 1 + T.has_lo; ++plane) {
            const ac_tc_src& Q = plane ? *ms[s].lo : S;
            cuuint64_t gdim[4] = {(cuuint64_t)Q.c0, (cuuint64_t)Q.phases, (cuuint64_t)Q.rows, (cuuint64_t)d->batch};
            cuuint64_t gstr[3] = {(cuuint64_t)Q.phase_stride * 2, (cuuint64_t)Q.row_stride * 2, (cuuint64_t)Q.batch_stride * 2};
            cuuint32_t box[4] = {(cuuint32_t)bk, 1, (cuuint32_t)T.box_rows, 1};
            cuuint32_t est[4] = {1, 1, 1, 1};
            CUresult r = encode(plane ? &maps.a_lo[s] : &maps.a[s], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(Q.base),
                                gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(bk),
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            AC_REQUIRE(r == CUDA_SUCCESS, "ac_conv_tc: cuTensorMapEncodeTiled(A%d) failed: %d", s, (int)r);
        }
        if (!T.has_lo) maps.a_lo[s] = maps.a[s];
    }
    for (int s = n_ms; s < MAX_SRC; ++s) { maps.a[s] = maps.a[0]; maps.a_lo[s] = maps.a[0]; }
    {
        cuuint64_t gdim[2] = {(cuuint64_t)k_total, (cuuint64_t)d->n_total * w_planes};  // W_lo / bf16(W) stacked under W_hi
        cuuint64_t gstr[1] = {(cuuint64_t)k_total * 2};
        cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)n_tile};
        cuuint32_t est[2] = {1, 1};
        CUresult r = encode(&maps.w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(d->w), gdim, gstr, box, est,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(bk), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        AC_REQUIRE(r == CUDA_SUCCESS, "ac_conv_tc: cuTensorMapEncodeTiled(W) failed: %d", (int)r);
    }

    p.w_rows = d->n_total;
    p.n_total = d->n_total;
    p.n_tile = n_tile;
    p.n_tiles = n_tiles;
    p.m_rows = d->m_rows;
    p.m_groups = (d->m_rows + p.G * TILE_M - 1) / (p.G * TILE_M);
    p.batch = d->batch;
    p.bias = d->bias; p.alpha = d->alpha;
    p.res = (const __nv_bfloat16*)d->res; p.y = (__nv_bfloat16*)d->y; p.y_act = (__nv_bfloat16*)d->y_act; p.y32 = d->y32;
    p.y_lo = (__nv_bfloat16*)d->y_lo; p.y_act_lo = (__nv_bfloat16*)d->y_act_lo;
    p.res_lo = (const __nv_bfloat16*)d->res_lo; p.res32 = d->res32;
    AC_REQUIRE((!p.y_lo || p.y) && (!p.y_act_lo || p.y_act), "ac_conv_tc: lo plane without its hi plane");
    p.act = d->act; p.epi = d->epi; p.act_mod = act_mod;
    p.y_bs = d->y_bstride; p.ya_bs = d->y_act_bstride; p.y32_bs = d->y32_bstride; p.res_bs = d->res_bstride;
    p.out_shift = d->out_shift; p.out_valid = d->out_valid;
    {
        auto al = [](const void* ptr, size_t a) { return ((uintptr_t)ptr % a) == 0; };
        const bool strides16 = d->n_total % 16 == 0 && d->out_shift % 16 == 0 && d->y_bstride % 16 == 0 && d->y_act_bstride % 16 == 0 &&
                               d->y32_bstride % 16 == 0 && d->res_bstride % 16 == 0;
        p.wide_ok = strides16 && al(d->y, 32) && al(d->y_lo, 32) && al(d->y_act, 32) && al(d->y_act_lo, 32) && al(d->y32, 32) &&
                    al(d->res, 32) && al(d->res_lo, 32) && al(d->res32, 32);
    }

    const size_t w_area = p.w_resident ? (size_t)p.w_res_plane * w_planes : (size_t)p.w_stages * p.w_stage_bytes;
    const size_t smem = fixed + (size_t)p.a_stages * p.a_stage_bytes + w_area;
    AC_REQUIRE(smem <= (size_t)SMEM_LIMIT, "ac_conv_tc: shared memory %zu", smem);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
        if (e != cudaSuccess) { ac::set_error("ac_conv_tc: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        attr_set = true;
    }
    const long long total_tiles = (long long)p.n_tiles * p.m_groups * p.batch;
    int grid = sms;
    if (d->grid_hint > 0) grid = d->grid_hint;
    if (total_tiles < grid) grid = (int)total_tiles;
    conv_tc_kernel<<<grid, THREADS, smem, (cudaStream_t)stream>>>(maps, p);
    return ac::finish_launch("ac_conv_tc");
}
