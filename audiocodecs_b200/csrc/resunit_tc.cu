// Fused residual unit on tcgen05 (ac_resunit_tc in include/audiocodecs_b200.h): two chained tap-GEMMs per tile with the
// hidden activation kept on chip.
//
//   h   = act1( bias1 + sum_taps A[m + j*dil + shift] . W1^T )                 GEMM1  -> TMEM acc1 [G*128 x Ch]
//   y   = bias2 + h . W2h^T (+ X[m] . W2x^T) (+ res[m])                        GEMM2  -> TMEM acc2 [G*128 x Cout]
//   out = y (raw, hi/lo planes) and/or act2(y) (the CONSUMER's activation)
//
// This is EnCodec's ResnetBlock  shortcut_1x1(x) + conv_k1(ELU(conv_k3(ELU(x))))  (HF/encodec:252-282; X = raw x is the
// conv shortcut), Mimi's (identity skip = res, HF/mimi:412-451) and DAC's ResidualUnit
// x + conv_k1(Snake(conv_k7_dilated(Snake(x))))  (HF/dac:173-207) -- in the reference five to six ATen ops with four
// round trips of the hidden tensor through memory; here the hidden tile goes TMEM -> registers (bias, activation,
// bf16 split) -> shared memory in the UMMA operand layout -> second GEMM, and never touches HBM.
//
// Roles of a CTA: warp 0 = A producer (TMA: the activated input block with its tap halo, loaded once per tile and chunk
// and read by every tap through row-shifted descriptors; then the raw-x blocks of the shortcut), warp 1 = MMA issuer,
// warp 2 = W producer (W1 then W2 blocks through one ring, or both matrices resident), warps 3-18 = epilogue
// (epilogue 1: acc1 -> hidden tile; epilogue 2: acc2 -> global).  The MMA warp issues GEMM1 of tile i+1 right after
// GEMM2 of tile i; acc1 and the hidden tile are double-buffered whenever tensor memory allows (2*G*ch + G*cout <= 512
// columns), so GEMM1 / epilogue 1 of tile i+1 run under GEMM2 / epilogue 2 of tile i and neither the tensor pipe (DAC) nor
// the epilogue warps (EnCodec, Mimi) wait on the other between tiles.
//
// Staged epilogue I/O (io_stage): a lane owns an accumulator ROW, so a direct global load / store of its 32 bytes touches
// 32 different 128-byte lines per warp instruction -- ncu showed the L1 data pipe at 55-84 % of its wavefront peak on DAC's
// 64 / 128-channel units (profiles/r02_dac_fp16_resunit_raw.csv), the top stall long_scoreboard.  When shared memory allows,
// the skip input tile is brought in by TMA (cp.async.bulk.tensor) into a swizzled staging buffer, epilogue 2 reads it there
// and writes the raw / activated outputs into two more, and one thread sends the finished tiles out with TMA stores
// (UTMASTG): coalesced 128-byte lines both ways, rows beyond m_rows clipped by the hardware.  The next tile's skip input is
// requested as soon as every warp has read this one's, and the stores are only waited for when the next tile is about to
// overwrite their buffers, so neither sits on the critical path.
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
constexpr int EPI_WARPS = 16;
constexpr int FIRST_EPI_WARP = 3;
constexpr int XF_WARPS = 4;  // transform warps: raw input block -> activated operand block (raw mode)
constexpr int FIRST_XF_WARP = FIRST_EPI_WARP + EPI_WARPS;
constexpr int THREADS = 32 * (FIRST_EPI_WARP + EPI_WARPS + XF_WARPS);
constexpr int NBARS = 4 * MAX_A_STAGES + 2 * MAX_W_STAGES + 12;
constexpr int SMEM_LIMIT = 227 * 1024;
constexpr int SMEM_HALF = 113 * 1024;

struct RuMaps {
    CUtensorMap a, a_lo, x, x_lo, w1, w2h, w2x;
    CUtensorMap res, res_lo, y, y_lo, ya, ya_lo;   // staged epilogue I/O: [batch][rows][cout] views, box = (bko channels, 128 rows, 1)
};

struct RuParams {
    int cin, taps, dil, shift, chunks1, a_has_lo, a_pieces, a_box_rows;
    int ch, cout, hblocks, xchunks, x_has_lo, x_pieces, x_box_rows, h_split;
    int bk, bkh, G, m_rows, m_groups, batch;  // bk: contraction block of the A / X sources (TMA row = bk*2 bytes), bkh: of the hidden tile
    int a_stages, w_stages, w_resident, w1_split, w2_split;
    uint32_t a_stage_bytes, a_plane_bytes, w_stage_bytes, w_plane_bytes;
    uint32_t w1_kb_bytes, w2h_kb_bytes, w2x_kb_bytes, w1_res_plane, w2h_res_plane, w2x_res_plane, w2h_res_off, w2x_res_off, w_area_bytes;
    uint32_t h_blk_bytes, h_plane_bytes;
    uint32_t tmem_cols, acc2_col, acc1_stride;
    // raw mode: the A source is the RAW input; transform warps apply act0 (the unit's input activation) block by block into a
    // second ring that GEMM1 reads, so the producer layer writes ONE tensor and this kernel reads it once.  x_from_a: the conv
    // shortcut of GEMM2 reads the raw rows of the same staged blocks (no separate X loads).
    int raw, act0, e_split, x_from_a, sc_row_off;
    uint32_t e_stage_bytes, e_plane_bytes;
    const float* alpha0;
    // staged epilogue I/O: planes [res hi][res lo][y hi][y lo][act hi][act lo] (those that exist) of io_plane_bytes each, a
    // plane = cout/bko column blocks of G pieces of [128 rows][bko*2 bytes] in the TMA swizzle of that row width
    int io_stage, bko, io_y_off, io_act_off, io_planes;   // plane index of the y / activated group
    uint32_t io_plane_bytes, io_blk_bytes;
    int f16, w1_hib, w2_hib, y_f16, ya_f16, res_f16;  // formats (AC_FMT_*): hi planes fp16 / extra bf16(W) planes for the lo products
    int dbl;                      // acc1 and the hidden tile are double-buffered: GEMM1 / epilogue 1 of tile i+1 overlap GEMM2 / epilogue 2 of tile i
    int pp;                       // ping-pong (needs dbl): acc2 is double-buffered too and the epilogue warps form two groups of eight that
                                  // take alternate tiles, each running epilogue 1 and epilogue 2 of ITS tile -- two tiles in different
                                  // phases (TMEM loads + activation + shared stores vs skip loads + global stores) share the SM's pipes
    uint32_t acc2_stride;         // columns between the two acc2 buffers
    uint32_t h_stage_bytes;       // one hidden-tile buffer (all k-blocks, hi [+lo] planes)
    const float *bias1, *alpha1, *bias2, *alpha2;
    int act1, act2;
    const __nv_bfloat16 *res, *res_lo;
    __nv_bfloat16 *y, *y_lo, *y_act, *y_act_lo;
    long long res_bs, y_bs, ya_bs;
};

__device__ __forceinline__ void load_block(const CUtensorMap* map, uint8_t* dst, uint64_t* bar, int col, int row0, int b, int pieces,
                                           int box_rows, uint32_t piece_bytes) {
    for (int pc = 0; pc < pieces; ++pc) tma_load_4d(dst + pc * piece_bytes, map, bar, col, 0, row0 + pc * box_rows, b);
}

// one W block against the G sub-tiles of one A block (rows of KS*32 bytes); `a_lo_off` != 0 adds the A_lo * W_hi product
template <int KS>
__device__ __forceinline__ void issue_blocks(bool leader, int G, uint32_t d0, uint32_t dstep, uint32_t a_addr, uint32_t a_lo_off,
                                             uint32_t w_hi, uint32_t w_lo, bool wsplit, uint32_t idesc, uint32_t acc, uint32_t w_hib,
                                             uint32_t idesc_lo) {
    constexpr uint32_t row_bytes = KS * 32;
    const uint64_t desc_base = make_smem_desc(0, row_bytes);
    auto desc = [&](uint32_t addr) { return desc_base | ((addr & 0x3FFFFu) >> 4); };
    const uint64_t bh = desc(w_hi), bl = desc(w_lo), bhb = desc(w_hib);
    for (int g = 0; g < G; ++g, a_addr += TILE_M * row_bytes, d0 += dstep) {
        const uint64_t ad = desc(a_addr);
        if (leader) {
#pragma unroll
            for (int k = 0; k < KS; ++k) umma_bf16(d0, ad + 2 * k, bh + 2 * k, idesc, k == 0 ? acc : 1u);
            if (wsplit) {
#pragma unroll
                for (int k = 0; k < KS; ++k) umma_bf16(d0, ad + 2 * k, bl + 2 * k, idesc, 1u);
            }
            if (a_lo_off) {
                const uint64_t al = desc(a_addr + a_lo_off);  // A_lo (bf16) x bf16(W_hi)
#pragma unroll
                for (int k = 0; k < KS; ++k) umma_bf16(d0, al + 2 * k, bhb + 2 * k, idesc_lo, 1u);
            }
        }
    }
}

template <int KA, int KH>
__device__ __forceinline__ void mma_role(const RuParams& p, uint8_t* a_ring, uint8_t* e_ring, uint8_t* w_area, uint8_t* h_tile, uint64_t* a_full,
                                         uint64_t* a_empty, uint64_t* e_full, uint64_t* e_empty, uint64_t* w_full, uint64_t* w_empty, uint64_t* acc1_full,
                                         uint64_t* h_ready, uint64_t* acc2_full, uint64_t* acc_free, uint64_t* wres_bar,
                                         uint32_t tmem_base, int total_tiles) {
    const bool leader = elect_one();
    int astage = 0, wstage = 0;
    uint32_t aphase = 0, wphase = 0;
    const uint32_t idesc1 = p.f16 ? make_idesc_f16(TILE_M, p.ch) : make_idesc_bf16(TILE_M, p.ch);
    const uint32_t idesc2 = p.f16 ? make_idesc_f16(TILE_M, p.cout) : make_idesc_bf16(TILE_M, p.cout);
    const uint32_t idesc1_lo = make_idesc_bf16(TILE_M, p.ch), idesc2_lo = make_idesc_bf16(TILE_M, p.cout);
    // which plane (0 = W_hi, 1 = W_lo, 2 = bf16 W_hi) the A_lo products read
    const uint32_t hib1 = p.w1_hib ? (uint32_t)(1 + p.w1_split) : 0u, hib2 = p.w2_hib ? (uint32_t)(1 + p.w2_split) : 0u;
    constexpr uint32_t row_bytes = KA * 32;
    const uint32_t a_ring_u = smem_u32(a_ring), e_ring_u = smem_u32(e_ring), w_area_u = smem_u32(w_area), h_u = smem_u32(h_tile);
    int a_base[2] = {0, 0};  // ring stage of chunk 0 of the tiles in flight (raw rows are read again by GEMM2's shortcut)
    if (p.w_resident) { mbar_wait(wres_bar, 0); tc_fence_after(); }
    // next W block: resident address or ring slot (returns hi address; lo = hi + plane).  which: 0 = W1, 1 = W2 hidden part, 2 = W2 x part
    auto w_get = [&](int which, int kb, uint32_t& w_hi, uint32_t& w_lo, uint32_t& w_hib) {
        const uint32_t hq = which == 0 ? hib1 : hib2;
        if (p.w_resident) {
            if (which == 0) { w_hi = w_area_u + kb * p.w1_kb_bytes; w_lo = w_hi + p.w1_res_plane; w_hib = w_hi + hq * p.w1_res_plane; }
            else if (which == 1) { w_hi = w_area_u + p.w2h_res_off + kb * p.w2h_kb_bytes; w_lo = w_hi + p.w2h_res_plane; w_hib = w_hi + hq * p.w2h_res_plane; }
            else { w_hi = w_area_u + p.w2x_res_off + kb * p.w2x_kb_bytes; w_lo = w_hi + p.w2x_res_plane; w_hib = w_hi + hq * p.w2x_res_plane; }
        } else {
            mbar_wait(&w_full[wstage], wphase);
            tc_fence_after();
            w_hi = w_area_u + wstage * p.w_stage_bytes;
            w_lo = w_hi + p.w_plane_bytes;
            w_hib = w_hi + hq * p.w_plane_bytes;
        }
    };
    auto w_done = [&]() {
        if (!p.w_resident) {
            if (leader) umma_commit(&w_empty[wstage]);
            if (++wstage == p.w_stages) { wstage = 0; wphase ^= 1; }
        }
    };
    // GEMM1 of the it-th tile of this CTA: taps over the activated input block -> acc1[it & dbl]
    auto gemm1 = [&](int it) {
        const uint32_t d0 = tmem_base + (p.dbl ? (it & 1) : 0) * p.acc1_stride;
        uint32_t acc = 0;
        a_base[it & 1] = astage;
        for (int cc = 0; cc < p.chunks1; ++cc) {
            mbar_wait(p.raw ? &e_full[astage] : &a_full[astage], aphase);
            tc_fence_after();
            const uint32_t a_hi = p.raw ? e_ring_u + astage * p.e_stage_bytes : a_ring_u + astage * p.a_stage_bytes;
            const uint32_t lo_off = p.raw ? (p.e_split ? p.e_plane_bytes : 0u) : (p.a_has_lo ? p.a_plane_bytes : 0u);
            for (int j = 0; j < p.taps; ++j) {
                uint32_t w_hi, w_lo, w_hib;
                w_get(0, j * p.chunks1 + cc, w_hi, w_lo, w_hib);
                issue_blocks<KA>(leader, p.G, d0, p.ch, a_hi + j * p.dil * row_bytes, lo_off, w_hi, w_lo, p.w1_split != 0, idesc1, acc, w_hib, idesc1_lo);
                acc = 1u;
                w_done();
            }
            if (leader) umma_commit(p.raw ? &e_empty[astage] : &a_empty[astage]);
            if (++astage == p.a_stages) { astage = 0; aphase ^= 1; }
        }
        if (leader) umma_commit(&acc1_full[p.dbl ? (it & 1) : 0]);
        __syncwarp();
    };
    // GEMM2 of the it-th tile: hidden tile (shared memory) [+ raw x blocks of the conv shortcut] -> acc2
    auto gemm2 = [&](int it) {
        const int hb = p.dbl ? (it & 1) : 0;
        const int b2 = p.pp ? (it & 1) : 0;
        const uint32_t acc2 = tmem_base + p.acc2_col + b2 * p.acc2_stride;
        mbar_wait(&h_ready[hb], p.dbl ? ((it >> 1) & 1) : (it & 1));
        mbar_wait(&acc_free[b2], (p.pp ? ((it >> 1) & 1) : (it & 1)) ^ 1);  // epilogue 2 of the previous user of this acc2 buffer has drained it
        tc_fence_after();
        const uint32_t hbase = h_u + hb * p.h_stage_bytes;
        uint32_t acc = 0;
        for (int kb = 0; kb < p.hblocks; ++kb) {
            uint32_t w_hi, w_lo, w_hib;
            w_get(1, kb, w_hi, w_lo, w_hib);
            issue_blocks<KH>(leader, p.G, acc2, p.cout, hbase + kb * p.h_blk_bytes, p.h_split ? p.h_plane_bytes : 0u, w_hi, w_lo, p.w2_split != 0, idesc2, acc, w_hib, idesc2_lo);
            acc = 1u;
            w_done();
        }
        if (p.x_from_a) {
            // conv shortcut from the raw rows of this tile's own staged blocks (already landed: the transform warps waited
            // for them before signalling e_full); the blocks are released here
            for (int cc = 0; cc < p.chunks1; ++cc) {
                int st = a_base[it & 1] + cc;
                if (st >= p.a_stages) st -= p.a_stages;
                uint32_t w_hi, w_lo, w_hib;
                w_get(2, cc, w_hi, w_lo, w_hib);
                issue_blocks<KA>(leader, p.G, acc2, p.cout, a_ring_u + st * p.a_stage_bytes + p.sc_row_off * row_bytes,
                                 p.a_has_lo ? p.a_plane_bytes : 0u, w_hi, w_lo, p.w2_split != 0, idesc2, acc, w_hib, idesc2_lo);
                acc = 1u;
                w_done();
                if (leader) umma_commit(&a_empty[st]);
            }
        } else {
            for (int cc = 0; cc < p.xchunks; ++cc) {
                mbar_wait(&a_full[astage], aphase);
                tc_fence_after();
                uint32_t w_hi, w_lo, w_hib;
                w_get(2, cc, w_hi, w_lo, w_hib);
                issue_blocks<KA>(leader, p.G, acc2, p.cout, a_ring_u + astage * p.a_stage_bytes, p.x_has_lo ? p.a_plane_bytes : 0u, w_hi, w_lo, p.w2_split != 0, idesc2, acc, w_hib, idesc2_lo);
                acc = 1u;
                w_done();
                if (leader) umma_commit(&a_empty[astage]);
                if (++astage == p.a_stages) { astage = 0; aphase ^= 1; }
            }
        }
        if (leader) umma_commit(&acc2_full[b2]);
        __syncwarp();
    };
    // Issue order (the producers follow the same order):  G1(0) | G1(1) G2(0) | G1(2) G2(1) | ...   when double-buffered,
    //                                                     G1(0) G2(0) | G1(1) G2(1) | ...            otherwise.
    const int my_tiles = blockIdx.x < total_tiles ? (total_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    if (p.dbl) {
        if (my_tiles > 0) gemm1(0);
        for (int it = 0; it < my_tiles; ++it) {
            if (it + 1 < my_tiles) gemm1(it + 1);
            gemm2(it);
        }
    } else {
        for (int it = 0; it < my_tiles; ++it) { gemm1(it); gemm2(it); }
    }
}

__device__ __forceinline__ void apply_act(float (&o)[16], int act, const float* alpha_s, const float* ralpha_s, int ch0) {
    if (act == AC_ACT_ELU) {
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = elu_ex2(o[i]);
    } else if (act == AC_ACT_SNAKE) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 a = *reinterpret_cast<const float4*>(alpha_s + ch0 + 4 * q);
            const float4 r = *reinterpret_cast<const float4*>(ralpha_s + ch0 + 4 * q);
            const float al[4] = {a.x, a.y, a.z, a.w}, ra[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float sn = __sinf(al[i] * o[4 * q + i]);
                o[4 * q + i] = fmaf(ra[i], sn * sn, o[4 * q + i]);
            }
        }
    }
}

__device__ __forceinline__ void add_bias16(float (&o)[16], const uint32_t (&v)[16], const float* bias_s) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 b = *reinterpret_cast<const float4*>(bias_s + 4 * q);
        o[4 * q] = __uint_as_float(v[4 * q]) + b.x; o[4 * q + 1] = __uint_as_float(v[4 * q + 1]) + b.y;
        o[4 * q + 2] = __uint_as_float(v[4 * q + 2]) + b.z; o[4 * q + 3] = __uint_as_float(v[4 * q + 3]) + b.w;
    }
}

// RAW = false drops the four transform warps from the launch: 608 threads leave the epilogue a 104-register budget (room
// for the one-item-ahead residual prefetch) instead of 88.
template <bool RAW>
__global__ void __launch_bounds__(RAW ? THREADS : THREADS - 32 * XF_WARPS, 1)
resunit_tc_kernel(const __grid_constant__ RuMaps maps, const __grid_constant__ RuParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align_smem_1024(smem_raw);
    uint8_t* a_ring = smem;
    uint8_t* e_ring = a_ring + (size_t)p.a_stages * p.a_stage_bytes;  // activated blocks (raw mode only)
    uint8_t* w_area = e_ring + (p.raw ? (size_t)p.a_stages * p.e_stage_bytes : 0);
    uint8_t* h_tile = w_area + p.w_area_bytes;
    uint8_t* io_s = h_tile + (size_t)p.h_stage_bytes * (1 + p.dbl);   // staged epilogue I/O planes (1024-aligned: every part before is)
    uint64_t* bars = reinterpret_cast<uint64_t*>(io_s + (p.io_stage ? (size_t)p.io_plane_bytes * p.io_planes : 0));
    uint64_t* a_full = bars;
    uint64_t* a_empty = a_full + MAX_A_STAGES;
    uint64_t* e_full = a_empty + MAX_A_STAGES;
    uint64_t* e_empty = e_full + MAX_A_STAGES;
    uint64_t* w_full = e_empty + MAX_A_STAGES;
    uint64_t* w_empty = w_full + MAX_W_STAGES;
    uint64_t* acc1_full = w_empty + MAX_W_STAGES;  // [2]
    uint64_t* h_ready = acc1_full + 2;             // [2]
    uint64_t* acc2_full = h_ready + 2;             // [2]
    uint64_t* acc_free = acc2_full + 2;            // [2]
    uint64_t* wres_bar = acc_free + 2;
    uint64_t* io_ready = wres_bar + 1;   // this tile's skip input has landed in the staging buffer
    uint64_t* out_free = io_ready + 1;   // the previous tile's stores have read the output staging buffers
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(out_free + 1);
    float* bias1_s = reinterpret_cast<float*>(bars + NBARS);  // 16-byte aligned: the barrier block is 1024-aligned, NBARS is even
    float* alpha1_s = bias1_s + p.ch;
    float* ralpha1_s = alpha1_s + p.ch;
    float* bias2_s = ralpha1_s + p.ch;
    float* alpha2_s = bias2_s + p.cout;
    float* ralpha2_s = alpha2_s + p.cout;
    float* alpha0_s = ralpha2_s + p.cout;   // [cin] raw mode
    float* ralpha0_s = alpha0_s + p.cin;

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform role index
    const int lane = threadIdx.x & 31;
    const int total_tiles = p.m_groups * p.batch;
    const int my_tiles = (int)blockIdx.x < total_tiles ? (total_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const uint32_t row_bytes = p.bk * 2;

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&maps.a);
        prefetch_tensormap(&maps.w1);
        prefetch_tensormap(&maps.w2h);
        for (int i = 0; i < p.a_stages; ++i) {
            mbar_init(&a_full[i], 1);
            mbar_init(&a_empty[i], (p.raw && !p.x_from_a) ? XF_WARPS : 1);  // released by the transform warps, or by an MMA commit
            mbar_init(&e_full[i], XF_WARPS);
            mbar_init(&e_empty[i], 1);
        }
        for (int i = 0; i < p.w_stages; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
        const int epi_n = p.pp ? EPI_WARPS / 2 : EPI_WARPS;   // warps that hand over one tile
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc1_full[i], 1); mbar_init(&h_ready[i], epi_n);
            mbar_init(&acc2_full[i], 1); mbar_init(&acc_free[i], epi_n);
        }
        mbar_init(wres_bar, 1);
        mbar_init(io_ready, 1);
        mbar_init(out_free, 1);
        if (p.io_stage) {
            prefetch_tensormap(&maps.res); prefetch_tensormap(&maps.y); prefetch_tensormap(&maps.ya);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
    for (int i = threadIdx.x; i < p.ch; i += blockDim.x) {
        bias1_s[i] = p.bias1 ? p.bias1[i] : 0.f;
        const float a = p.act1 == AC_ACT_SNAKE ? p.alpha1[i] : 1.f;
        alpha1_s[i] = a;
        ralpha1_s[i] = 1.0f / (a + 1e-9f);
    }
    for (int i = threadIdx.x; i < p.cout; i += blockDim.x) {
        bias2_s[i] = p.bias2 ? p.bias2[i] : 0.f;
        const float a = p.act2 == AC_ACT_SNAKE ? p.alpha2[i] : 1.f;
        alpha2_s[i] = a;
        ralpha2_s[i] = 1.0f / (a + 1e-9f);
    }
    if (p.raw)
        for (int i = threadIdx.x; i < p.cin; i += blockDim.x) {
            const float a = p.act0 == AC_ACT_SNAKE ? p.alpha0[i] : 1.f;
            alpha0_s[i] = a;
            ralpha0_s[i] = 1.0f / (a + 1e-9f);
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
            const uint32_t a_piece = p.a_box_rows * row_bytes, x_piece = p.x_box_rows * row_bytes;
            auto load1 = [&](int it) {  // blocks of GEMM1 of this CTA's it-th tile
                const int tile = blockIdx.x + it * gridDim.x;
                const int mg = tile % p.m_groups, b = tile / p.m_groups, m0 = mg * p.G * TILE_M;
                for (int cc = 0; cc < p.chunks1; ++cc) {
                    mbar_wait(&a_empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&a_full[stage], p.a_pieces * a_piece * (1 + p.a_has_lo));
                    uint8_t* dst = a_ring + (size_t)stage * p.a_stage_bytes;
                    load_block(&maps.a, dst, &a_full[stage], cc * p.bk, m0 + p.shift, b, p.a_pieces, p.a_box_rows, a_piece);
                    if (p.a_has_lo) load_block(&maps.a_lo, dst + p.a_plane_bytes, &a_full[stage], cc * p.bk, m0 + p.shift, b, p.a_pieces, p.a_box_rows, a_piece);
                    if (++stage == p.a_stages) { stage = 0; phase ^= 1; }
                }
            };
            auto load2 = [&](int it) {  // raw-x blocks of GEMM2
                const int tile = blockIdx.x + it * gridDim.x;
                const int mg = tile % p.m_groups, b = tile / p.m_groups, m0 = mg * p.G * TILE_M;
                for (int cc = 0; cc < (p.x_from_a ? 0 : p.xchunks); ++cc) {
                    mbar_wait(&a_empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&a_full[stage], p.x_pieces * x_piece * (1 + p.x_has_lo));
                    uint8_t* dst = a_ring + (size_t)stage * p.a_stage_bytes;
                    load_block(&maps.x, dst, &a_full[stage], cc * p.bk, m0, b, p.x_pieces, p.x_box_rows, x_piece);
                    if (p.x_has_lo) load_block(&maps.x_lo, dst + p.a_plane_bytes, &a_full[stage], cc * p.bk, m0, b, p.x_pieces, p.x_box_rows, x_piece);
                    if (++stage == p.a_stages) { stage = 0; phase ^= 1; }
                }
            };
            if (p.dbl) {
                if (my_tiles > 0) load1(0);
                for (int it = 0; it < my_tiles; ++it) {
                    if (it + 1 < my_tiles) load1(it + 1);
                    load2(it);
                }
            } else {
                for (int it = 0; it < my_tiles; ++it) { load1(it); load2(it); }
            }
        }
    } else if (warp == 2) {
        // ================================================================= W producer
        if (lane == 0) {
            const uint32_t blk1 = p.ch * row_bytes, blk2h = p.cout * p.bkh * 2, blk2x = p.cout * row_bytes;
            const int nkb1 = p.taps * p.chunks1;
            if (p.w_resident) {
                const int pl1 = 1 + p.w1_split + p.w1_hib, pl2 = 1 + p.w2_split + p.w2_hib;
                mbar_arrive_expect_tx(wres_bar, (uint32_t)nkb1 * blk1 * pl1 + ((uint32_t)p.hblocks * blk2h + (uint32_t)p.xchunks * blk2x) * pl2);
                for (int kb = 0; kb < nkb1; ++kb)
                    for (int q = 0; q < pl1; ++q)
                        tma_load_2d(w_area + (size_t)q * p.w1_res_plane + (size_t)kb * p.w1_kb_bytes, &maps.w1, wres_bar, kb * p.bk, q * p.ch);
                for (int kb = 0; kb < p.hblocks; ++kb)
                    for (int q = 0; q < pl2; ++q)
                        tma_load_2d(w_area + p.w2h_res_off + (size_t)q * p.w2h_res_plane + (size_t)kb * p.w2h_kb_bytes, &maps.w2h, wres_bar, kb * p.bkh, q * p.cout);
                for (int cc = 0; cc < p.xchunks; ++cc)
                    for (int q = 0; q < pl2; ++q)
                        tma_load_2d(w_area + p.w2x_res_off + (size_t)q * p.w2x_res_plane + (size_t)cc * p.w2x_kb_bytes, &maps.w2x, wres_bar, p.ch + cc * p.bk, q * p.cout);
            } else {
                int stage = 0;
                uint32_t phase = 0;
                auto put = [&](const CUtensorMap* map, int col, uint32_t blk, int planes, int rows) {
                    mbar_wait(&w_empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&w_full[stage], blk * planes);
                    uint8_t* dst = w_area + (size_t)stage * p.w_stage_bytes;
                    for (int q = 0; q < planes; ++q) tma_load_2d(dst + (size_t)q * p.w_plane_bytes, map, &w_full[stage], col, q * rows);
                    if (++stage == p.w_stages) { stage = 0; phase ^= 1; }
                };
                const int pl1 = 1 + p.w1_split + p.w1_hib, pl2 = 1 + p.w2_split + p.w2_hib;
                auto put1 = [&]() {
                    for (int cc = 0; cc < p.chunks1; ++cc)
                        for (int j = 0; j < p.taps; ++j) put(&maps.w1, (j * p.chunks1 + cc) * p.bk, blk1, pl1, p.ch);
                };
                auto put2 = [&]() {
                    for (int kb = 0; kb < p.hblocks; ++kb) put(&maps.w2h, kb * p.bkh, blk2h, pl2, p.cout);
                    for (int cc = 0; cc < p.xchunks; ++cc) put(&maps.w2x, p.ch + cc * p.bk, blk2x, pl2, p.cout);
                };
                if (p.dbl) {
                    if (my_tiles > 0) put1();
                    for (int it = 0; it < my_tiles; ++it) {
                        if (it + 1 < my_tiles) put1();
                        put2();
                    }
                } else {
                    for (int it = 0; it < my_tiles; ++it) { put1(); put2(); }
                }
            }
        }
    } else if (warp == 1) {
        // ================================================================= MMA issuer
#define AC_RU_ARGS p, a_ring, e_ring, w_area, h_tile, a_full, a_empty, e_full, e_empty, w_full, w_empty, acc1_full, h_ready, acc2_full, acc_free, wres_bar, tmem_base, total_tiles
        const int ka = p.bk / 16, kh = p.bkh / 16;
        if (ka == 4 && kh == 4) mma_role<4, 4>(AC_RU_ARGS);
        else if (ka == 4 && kh == 2) mma_role<4, 2>(AC_RU_ARGS);
        else if (ka == 4 && kh == 1) mma_role<4, 1>(AC_RU_ARGS);
        else if (ka == 2 && kh == 4) mma_role<2, 4>(AC_RU_ARGS);
        else if (ka == 2 && kh == 2) mma_role<2, 2>(AC_RU_ARGS);
        else if (ka == 2 && kh == 1) mma_role<2, 1>(AC_RU_ARGS);
        else if (ka == 1 && kh == 4) mma_role<1, 4>(AC_RU_ARGS);
        else if (ka == 1 && kh == 2) mma_role<1, 2>(AC_RU_ARGS);
        else mma_role<1, 1>(AC_RU_ARGS);
#undef AC_RU_ARGS
    } else if (RAW && warp >= FIRST_XF_WARP) {
        // ================================================================= transform warps (raw mode): block by block,
        // activated = act0(raw_hi [+ raw_lo]) written at the same (swizzled) offsets of the E ring; elementwise, so the
        // operand layout TMA produced is preserved.  Zero-filled (out-of-bounds) rows stay zero: ELU(0) = Snake(0) = 0.
        if (p.raw) {
            const int t = threadIdx.x - 32 * FIRST_XF_WARP;
            const int upr = p.bk / 8;                                   // 16-byte units per block row
            const int xshift = p.bk == 64 ? 0 : (p.bk == 32 ? 1 : 2);
            const int units = p.a_pieces * p.a_box_rows * upr;
            int stage = 0;
            uint32_t phase = 0;
            const int blocks = my_tiles * p.chunks1;
            for (int n = 0, cc = 0; n < blocks; ++n) {
                mbar_wait(&a_full[stage], phase);
                mbar_wait(&e_empty[stage], phase ^ 1);
                const uint8_t* src = a_ring + (size_t)stage * p.a_stage_bytes;
                uint8_t* dst = e_ring + (size_t)stage * p.e_stage_bytes;
                for (int u = t; u < units; u += 32 * XF_WARPS) {
                    const uint4 h = *reinterpret_cast<const uint4*>(src + (size_t)u * 16);
                    const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
                    float v[8];
#pragma unroll
                    for (int i = 0; i < 4; ++i) { const float2 f = unpack16(hw[i], p.f16 != 0); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
                    if (p.a_has_lo) {
                        const uint4 l = *reinterpret_cast<const uint4*>(src + p.a_plane_bytes + (size_t)u * 16);
                        const __nv_bfloat162* lp = reinterpret_cast<const __nv_bfloat162*>(&l);
#pragma unroll
                        for (int i = 0; i < 4; ++i) { v[2 * i] += __low2float(lp[i]); v[2 * i + 1] += __high2float(lp[i]); }
                    }
                    if (p.act0 == AC_ACT_ELU) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] = elu_ex2(v[i]);
                    } else {
                        const int row = u / upr, pu = u - row * upr;
                        const int ch = cc * p.bk + ((pu ^ ((row >> xshift) & (upr - 1))) << 3);  // logical channel of this unit
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const float4 a = *reinterpret_cast<const float4*>(alpha0_s + ch + 4 * q);
                            const float4 r = *reinterpret_cast<const float4*>(ralpha0_s + ch + 4 * q);
                            const float al[4] = {a.x, a.y, a.z, a.w}, ra[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float sn = __sinf(al[i] * v[4 * q + i]);
                                v[4 * q + i] = fmaf(ra[i], sn * sn, v[4 * q + i]);
                            }
                        }
                    }
                    const uint4 q = pack8(v, p.f16 != 0);
                    *reinterpret_cast<uint4*>(dst + (size_t)u * 16) = q;
                    if (p.e_split) *reinterpret_cast<uint4*>(dst + p.e_plane_bytes + (size_t)u * 16) = pack_lo(v, q, p.f16 != 0);
                }
                fence_proxy_async();  // generic-proxy stores of the activated block -> visible to tcgen05.mma
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&e_full[stage]);
                    if (!p.x_from_a) mbar_arrive(&a_empty[stage]);
                }
                if (++stage == p.a_stages) { stage = 0; phase ^= 1; }
                if (++cc == p.chunks1) cc = 0;
            }
        }
    } else {
        // ================================================================= epilogue warps
        const int quarter = warp & 3;
        const int grp = p.pp ? (warp - FIRST_EPI_WARP) >> 3 : 0;           // ping-pong: warps 3-10 take the even tiles, 11-18 the odd ones
        const int nslots = p.pp ? EPI_WARPS / 8 : EPI_WARPS / 4;            // warps per TMEM lane quarter working on one tile
        const int slot = ((warp - FIRST_EPI_WARP) >> 2) & (nslots - 1);
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const int c1 = p.ch / 16, c2 = p.cout / 16;
        const int units_per_row = p.bkh / 8;                       // 16-byte units per hidden-tile row: 8 / 4 / 2
        const int xshift = p.bkh == 64 ? 0 : (p.bkh == 32 ? 1 : 2);  // swizzle phase of a row = (row >> xshift) & (units_per_row - 1)
        const uint32_t h_row_bytes = p.bkh * 2;
        const bool has_res = p.res != nullptr, has_res_lo = p.res_lo != nullptr;
        // ---------------- epilogue 1: acc1 -> bias, activation, bf16 split -> hidden tile in the UMMA operand layout
        auto epi1 = [&](int it) {
            const int ab = p.dbl ? (it & 1) : 0;
            mbar_wait(&acc1_full[ab], p.dbl ? ((it >> 1) & 1) : (it & 1));
            tc_fence_after();
            uint8_t* hbuf = h_tile + (size_t)ab * p.h_stage_bytes;
            const uint32_t acc1 = lane_addr + ab * p.acc1_stride;
            int g = 0, c = slot;
            while (c >= c1) { c -= c1; ++g; }
            for (int item = slot; item < p.G * c1; item += nslots) {
                uint32_t v[16];
                tmem_ld16(acc1 + g * p.ch + c * 16, v);
                const int row = g * TILE_M + quarter * 32 + lane;  // row of the hidden tile
                const int col = c * 16;
                tmem_ld_wait();
                float o[16];
                add_bias16(o, v, bias1_s + col);
                apply_act(o, p.act1, alpha1_s, ralpha1_s, col);
                uint32_t hi[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) hi[i] = pack16(o[2 * i], o[2 * i + 1], p.f16 != 0);
                const int kb = col / p.bkh;
                const int u0 = (col % p.bkh) >> 3;
                const int xr = (row >> xshift) & (units_per_row - 1);
                uint8_t* rowp = hbuf + (size_t)kb * p.h_blk_bytes + (size_t)row * h_row_bytes;
                *reinterpret_cast<uint4*>(rowp + (((u0) ^ xr) << 4)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4*>(rowp + (((u0 + 1) ^ xr) << 4)) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
                if (p.h_split) {
                    uint32_t lo[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float2 hf = unpack16(hi[i], p.f16 != 0);
                        lo[i] = pack_bf16(o[2 * i] - hf.x, o[2 * i + 1] - hf.y);
                    }
                    *reinterpret_cast<uint4*>(rowp + p.h_plane_bytes + (((u0) ^ xr) << 4)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                    *reinterpret_cast<uint4*>(rowp + p.h_plane_bytes + (((u0 + 1) ^ xr) << 4)) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
                }
                c += nslots;
                while (c >= c1) { c -= c1; ++g; }
            }
            tc_fence_before();
            fence_proxy_async();  // generic-proxy stores of the hidden tile -> visible to tcgen05.mma (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(&h_ready[ab]);
        };
        // ---------------- epilogue 2: acc2 -> bias, residual, raw / activated outputs
        // the skip input of an item (32 + 32 bytes per lane) is fetched one item ahead, the first item's before the wait on
        // acc2: its DRAM latency runs under GEMM2 / the previous item's math (ncu: long_scoreboard was the top stall)
        uint4 rnext[4] = {};
        auto fetch_res = [&](int b, int mg, int g, int c) {
            const int m = (mg * p.G + g) * TILE_M + quarter * 32 + lane;
            if (m >= p.m_rows) return;
            const long long off = (long long)b * p.res_bs + (long long)m * p.cout + c * 16;
            const uint4* r = reinterpret_cast<const uint4*>(p.res + off);
            rnext[0] = r[0]; rnext[1] = r[1];
            if (has_res_lo) {
                const uint4* l = reinterpret_cast<const uint4*>(p.res_lo + off);
                rnext[2] = l[0]; rnext[3] = l[1];
            }
        };
        // ---------------- staged variant: skip input and outputs go through swizzled shared-memory tiles and TMA
        const int io_units = p.bko / 8;                                   // 16-byte units per staged row
        const int io_xshift = p.bko == 64 ? 0 : (p.bko == 32 ? 1 : 2);
        const uint32_t io_row_bytes = p.bko * 2, io_piece_bytes = TILE_M * p.bko * 2;
        const bool io_leader = warp == FIRST_EPI_WARP && lane == 0;
        // TMA traffic of one tile: `load` brings the skip tile in (completing on io_ready), else sends the output planes out
        auto io_tma = [&](int it, bool load) {
            const int tile = blockIdx.x + it * gridDim.x;
            const int mg = tile % p.m_groups, b = tile / p.m_groups;
            const int nblk = p.cout / p.bko;
            for (int kb = 0; kb < nblk; ++kb)
                for (int g = 0; g < p.G; ++g) {
                    const int m0 = (mg * p.G + g) * TILE_M;
                    if (m0 >= p.m_rows) continue;
                    uint8_t* piece = io_s + (size_t)kb * p.io_blk_bytes + (size_t)g * io_piece_bytes;
                    if (load) {
                        tma_load_3d(piece, &maps.res, io_ready, kb * p.bko, m0, b);
                        if (has_res_lo) tma_load_3d(piece + p.io_plane_bytes, &maps.res_lo, io_ready, kb * p.bko, m0, b);
                    } else {
                        if (p.y) {
                            uint8_t* yp = piece + (size_t)p.io_y_off * p.io_plane_bytes;
                            tma_store_3d(&maps.y, yp, kb * p.bko, m0, b);
                            if (p.y_lo) tma_store_3d(&maps.y_lo, yp + p.io_plane_bytes, kb * p.bko, m0, b);
                        }
                        if (p.y_act) {
                            uint8_t* ap = piece + (size_t)p.io_act_off * p.io_plane_bytes;
                            tma_store_3d(&maps.ya, ap, kb * p.bko, m0, b);
                            if (p.y_act_lo) tma_store_3d(&maps.ya_lo, ap + p.io_plane_bytes, kb * p.bko, m0, b);
                        }
                    }
                }
        };
        // the skip-input buffer is free (every warp has read the previous tile's): request tile `it`'s
        auto io_open = [&](int it) {
            uint32_t bytes = 0;
            if (has_res) {
                const int tile = blockIdx.x + it * gridDim.x;
                const int mg = tile % p.m_groups;
                int live = 0;
                for (int g = 0; g < p.G; ++g) live += ((mg * p.G + g) * TILE_M < p.m_rows) ? 1 : 0;
                bytes = (uint32_t)live * (uint32_t)(p.cout / p.bko) * io_piece_bytes * (1 + (has_res_lo ? 1 : 0));
            }
            if (bytes) { mbar_arrive_expect_tx(io_ready, bytes); io_tma(it, true); }
            else mbar_arrive(io_ready);
        };
        if (p.io_stage && io_leader && my_tiles > 0) io_open(0);
        auto epi2_staged = [&](int it) {
            if (io_leader) {               // the previous tile's stores have had a whole tile's time to read their buffers
                bulk_wait_read();
                mbar_arrive(out_free);
            }
            mbar_wait(io_ready, it & 1);   // this tile's skip input has landed
            mbar_wait(out_free, it & 1);
            mbar_wait(&acc2_full[0], it & 1);
            tc_fence_after();
            uint8_t* y_s = io_s + (size_t)p.io_y_off * p.io_plane_bytes;
            uint8_t* act_s = io_s + (size_t)p.io_act_off * p.io_plane_bytes;
            int g = 0, c = slot;
            while (c >= c2) { c -= c2; ++g; }
            for (int item = slot; item < p.G * c2; item += nslots) {
                uint32_t v[16];
                tmem_ld16(lane_addr + p.acc2_col + g * p.cout + c * 16, v);
                const int row = quarter * 32 + lane;          // row inside the 128-row piece g
                const int col = c * 16;
                const int kb = col / p.bko, u0 = (col % p.bko) >> 3;
                const int xr = (row >> io_xshift) & (io_units - 1);
                const uint32_t off = (uint32_t)kb * p.io_blk_bytes + (uint32_t)g * io_piece_bytes + (uint32_t)row * io_row_bytes;
                const uint32_t o0 = off + (((u0) ^ xr) << 4), o1 = off + (((u0 + 1) ^ xr) << 4);
                c += nslots;
                while (c >= c2) { c -= c2; ++g; }
                uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0, l0 = r0, l1 = r0;
                if (has_res) {
                    r0 = *reinterpret_cast<const uint4*>(io_s + o0); r1 = *reinterpret_cast<const uint4*>(io_s + o1);
                    if (has_res_lo) {
                        l0 = *reinterpret_cast<const uint4*>(io_s + p.io_plane_bytes + o0);
                        l1 = *reinterpret_cast<const uint4*>(io_s + p.io_plane_bytes + o1);
                    }
                }
                tmem_ld_wait();
                float o[16];
                add_bias16(o, v, bias2_s + col);
                if (has_res) {
                    add_bf16x16(o, r0, r1, p.res_f16 != 0);
                    if (has_res_lo) add_bf16x16(o, l0, l1);
                }
                if (p.y) {
                    float oa[8], ob[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) { oa[i] = o[i]; ob[i] = o[8 + i]; }
                    const uint4 qa = pack8(oa, p.y_f16 != 0), qb = pack8(ob, p.y_f16 != 0);
                    *reinterpret_cast<uint4*>(y_s + o0) = qa; *reinterpret_cast<uint4*>(y_s + o1) = qb;
                    if (p.y_lo) {
                        *reinterpret_cast<uint4*>(y_s + p.io_plane_bytes + o0) = pack_lo(oa, qa, p.y_f16 != 0);
                        *reinterpret_cast<uint4*>(y_s + p.io_plane_bytes + o1) = pack_lo(ob, qb, p.y_f16 != 0);
                    }
                }
                if (p.y_act) {
                    apply_act(o, p.act2, alpha2_s, ralpha2_s, col);
                    float oa[8], ob[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) { oa[i] = o[i]; ob[i] = o[8 + i]; }
                    const uint4 qa = pack8(oa, p.ya_f16 != 0), qb = pack8(ob, p.ya_f16 != 0);
                    *reinterpret_cast<uint4*>(act_s + o0) = qa; *reinterpret_cast<uint4*>(act_s + o1) = qb;
                    if (p.y_act_lo) {
                        *reinterpret_cast<uint4*>(act_s + p.io_plane_bytes + o0) = pack_lo(oa, qa, p.ya_f16 != 0);
                        *reinterpret_cast<uint4*>(act_s + p.io_plane_bytes + o1) = pack_lo(ob, qb, p.ya_f16 != 0);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_free[0]); // acc2 drained: GEMM2 of the next tile may start
            fence_proxy_async();                      // generic-proxy writes of the staged tiles -> visible to the TMA stores
            asm volatile("bar.sync 2, %0;" ::"n"(32 * EPI_WARPS) : "memory");   // every epilogue warp has finished the tile
            if (io_leader) {
                io_tma(it, false);
                bulk_commit();
                if (it + 1 < my_tiles) io_open(it + 1);   // every warp has read this tile's skip input: fetch the next one's
            }
        };
        auto epi2 = [&](int it) {
            if (p.io_stage) { epi2_staged(it); return; }
            const int tile = blockIdx.x + it * gridDim.x;
            const int mg = tile % p.m_groups;
            const int b = tile / p.m_groups;
            int g = 0, c = slot;
            while (c >= c2) { c -= c2; ++g; }
            if (has_res && slot < p.G * c2) fetch_res(b, mg, g, c);
            const int b2 = p.pp ? (it & 1) : 0;
            const uint32_t acc2 = lane_addr + p.acc2_col + b2 * p.acc2_stride;
            mbar_wait(&acc2_full[b2], p.pp ? ((it >> 1) & 1) : (it & 1));
            tc_fence_after();
            for (int item = slot; item < p.G * c2; item += nslots) {
                uint32_t v[16];
                tmem_ld16(acc2 + g * p.cout + c * 16, v);
                const int m = (mg * p.G + g) * TILE_M + quarter * 32 + lane;
                const int col = c * 16;
                const long long flat = (long long)m * p.cout + col;
                c += nslots;
                while (c >= c2) { c -= c2; ++g; }
                const uint4 rcur[4] = {rnext[0], rnext[1], rnext[2], rnext[3]};
                if (has_res && item + nslots < p.G * c2) fetch_res(b, mg, g, c);
                tmem_ld_wait();
                if (m >= p.m_rows) continue;
                float o[16];
                add_bias16(o, v, bias2_s + col);
                if (has_res) {
                    add_bf16x16(o, rcur[0], rcur[1], p.res_f16 != 0);
                    if (has_res_lo) add_bf16x16(o, rcur[2], rcur[3]);
                }
                if (p.y) store_bf16x16(o, p.y + (long long)b * p.y_bs + flat, p.y_lo ? p.y_lo + (long long)b * p.y_bs + flat : nullptr, p.y_f16 != 0);
                if (p.y_act) {
                    apply_act(o, p.act2, alpha2_s, ralpha2_s, col);
                    store_bf16x16(o, p.y_act + (long long)b * p.ya_bs + flat, p.y_act_lo ? p.y_act_lo + (long long)b * p.ya_bs + flat : nullptr, p.ya_f16 != 0);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_free[b2]);
        };
        // same order as the MMA warp: epilogue 1 of tile i+1 runs before epilogue 2 of tile i when double-buffered
        if (p.pp) {
            for (int it = grp; it < my_tiles; it += 2) { epi1(it); epi2(it); }   // this group's tiles; the other group is a phase away
        } else if (p.dbl) {
            if (my_tiles > 0) epi1(0);
            for (int it = 0; it < my_tiles; ++it) {
                if (it + 1 < my_tiles) epi1(it + 1);
                epi2(it);
            }
        } else {
            for (int it = 0; it < my_tiles; ++it) { epi1(it); epi2(it); }
        }
        if (p.io_stage && io_leader) bulk_wait_all();   // the last tile's stores are complete before the CTA exits
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

int encode_act_map(EncodeTiledFn encode, CUtensorMap* map, const void* base, int c0, int rows, int64_t row_stride, int64_t batch_stride,
                   int batch, int bk, int box_rows) {
    cuuint64_t gdim[4] = {(cuuint64_t)c0, 1, (cuuint64_t)rows, (cuuint64_t)batch};
    cuuint64_t gstr[3] = {(cuuint64_t)row_stride * 2, (cuuint64_t)row_stride * 2, (cuuint64_t)batch_stride * 2};
    cuuint32_t box[4] = {(cuuint32_t)bk, 1, (cuuint32_t)box_rows, 1};
    cuuint32_t est[4] = {1, 1, 1, 1};
    return (int)encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       swizzle_for(bk), CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

int encode_w_map(EncodeTiledFn encode, CUtensorMap* map, const void* w, int k_total, int n_rows, int planes, int bk) {
    cuuint64_t gdim[2] = {(cuuint64_t)k_total, (cuuint64_t)n_rows * planes};
    cuuint64_t gstr[1] = {(cuuint64_t)k_total * 2};
    cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)n_rows};
    cuuint32_t est[2] = {1, 1};
    return (int)encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       swizzle_for(bk), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

}  // namespace

extern "C" int ac_resunit_tc(const ac_resunit_tc_desc* d, void* stream) {
    AC_REQUIRE(d && d->a && d->w1 && d->w2, "ac_resunit_tc: null pointer");
    AC_REQUIRE(d->batch > 0 && d->m_rows > 0 && d->cin > 0 && d->taps > 0 && d->dilation > 0, "ac_resunit_tc: bad sizes");
    AC_REQUIRE(d->ch % 16 == 0 && d->ch >= 16 && d->ch <= 256 && d->cout % 16 == 0 && d->cout >= 16 && d->cout <= 256,
               "ac_resunit_tc: hidden %d / output %d channels must be multiples of 16 in [16, 256]", d->ch, d->cout);
    AC_REQUIRE(d->ch + d->cout <= 512, "ac_resunit_tc: accumulators need %d TMEM columns", d->ch + d->cout);
    AC_REQUIRE(d->y || d->y_act, "ac_resunit_tc: no output");
    AC_REQUIRE((!d->y_lo || d->y) && (!d->y_act_lo || d->y_act), "ac_resunit_tc: lo plane without its hi plane");
    AC_REQUIRE((d->act1 != AC_ACT_SNAKE || d->alpha1) && (d->act2 != AC_ACT_SNAKE || d->alpha2), "ac_resunit_tc: snake needs alpha");
    {
        auto al = [](const void* ptr, size_t a) { return ((uintptr_t)ptr % a) == 0; };
        AC_REQUIRE(al(d->a, 16) && al(d->a_lo, 16) && al(d->x, 16) && al(d->x_lo, 16) && al(d->w1, 16) && al(d->w2, 16) &&
                       (d->a_row_stride * 2) % 16 == 0 && (d->a_bstride * 2) % 16 == 0 && (d->x_bstride * 2) % 16 == 0,
                   "ac_resunit_tc: operands must be 16-byte aligned");
        AC_REQUIRE(al(d->y, 32) && al(d->y_lo, 32) && al(d->y_act, 32) && al(d->y_act_lo, 32) && al(d->res, 32) && al(d->res_lo, 32) &&
                       d->y_bstride % 16 == 0 && d->y_act_bstride % 16 == 0 && d->res_bstride % 16 == 0,
                   "ac_resunit_tc: outputs / residual must be 32-byte aligned with strides that are multiples of 16");
    }
    EncodeTiledFn encode = get_encode();
    AC_REQUIRE(encode, "ac_resunit_tc: cuTensorMapEncodeTiled not available");

    const int w1_split = d->w1_split ? 1 : 0, w2_split = d->w2_split ? 1 : 0, h_split = d->h_split ? 1 : 0;
    const int f16 = (d->fmt & AC_FMT_A_F16) ? 1 : 0, w1_hib = (d->fmt & AC_FMT_W_HIB) ? 1 : 0, w2_hib = (d->fmt & AC_FMT_W2_HIB) ? 1 : 0;
    AC_REQUIRE(f16 || !(w1_hib | w2_hib), "ac_resunit_tc: the extra bf16(W) planes only exist for fp16 operands");
    AC_REQUIRE(!f16 || ((!d->a_lo || w1_hib) && (!(h_split || (d->x && d->x_lo)) || w2_hib)),
               "ac_resunit_tc: fp16 operands with a lo plane need the bf16(W) plane of that GEMM");
    const int pl1 = 1 + w1_split + w1_hib, pl2 = 1 + w2_split + w2_hib;
    const int a_has_lo = d->a_lo ? 1 : 0, x_has_lo = (d->x && d->x_lo) ? 1 : 0, has_x = (d->x || d->x_from_a) ? 1 : 0;
    const int any_lo = a_has_lo | x_has_lo;
    const int sms = sm_count();
    const long long m_tiles = (d->m_rows + TILE_M - 1) / TILE_M;
    const int raw = d->act0 != AC_ACT_NONE ? 1 : 0, e_split = (raw && d->e_split) ? 1 : 0, x_from_a = d->x_from_a ? 1 : 0;
    AC_REQUIRE(!x_from_a || (raw && !d->x && d->x_row_off >= 0 && d->x_row_off <= (d->taps - 1) * d->dilation),
               "ac_resunit_tc: x_from_a needs raw mode, no separate x and the raw rows inside the staged block");
    AC_REQUIRE(!raw || d->act0 != AC_ACT_SNAKE || d->alpha0, "ac_resunit_tc: snake input activation needs alpha0");
    const size_t fixed = 1024 /*align slack*/ + NBARS * 8 + (size_t)(d->ch + d->cout) * 12 + (size_t)d->cin * 8 + 64;
    const int halo = (d->taps - 1) * d->dilation;

    RuParams p{};
    // The contraction blocks (bk of the A / X sources, bkh of the hidden tile) fix the order in which the (hi, lo) products are
    // accumulated, i.e. the fp32 rounding of the result.  They must depend on the layer shape alone -- not on the batch size or
    // on the (G, double-buffering) variant a tuner asks for, or a clip's tokens would depend on its neighbours: the canonical
    // pair is the one the un-hinted G = 1 search settles on, and every other tiling is accepted only with that same pair.
    // staged epilogue I/O: planes [raw hi][raw lo][act hi][act lo], each G*128 rows x cout (see the kernel header)
    const int io_res = (d->res ? 1 : 0) + (d->res_lo ? 1 : 0), io_y = (d->y ? 1 : 0) + (d->y_lo ? 1 : 0), io_a = (d->y_act ? 1 : 0) + (d->y_act_lo ? 1 : 0);
    const int io_planes = io_res + io_y + io_a;
    const int bko = d->cout % 64 == 0 ? 64 : (d->cout % 32 == 0 ? 32 : 16);
    const bool want_pp = d->dbl_hint == 2 && !raw;   // ping-pong epilogue groups (the hint 2 = double-buffered + ping-pong)
    bool use_io = d->io_stage >= 0 && !raw && !want_pp;
    auto search = [&](int g_only, int bk_only, int bkh_only, bool use_hint) -> bool {
    // pass 0: double-buffered hidden tile / acc1 with deep rings; pass 1: double-buffered, any rings; pass 2: single-buffered
    for (int pass = 0; pass < 3; ++pass)
        for (int bk = d->bk; bk >= 16; bk >>= 1)
        for (int bkh = 64; bkh >= 16; bkh >>= 1) {
            if (d->cin % bk || d->ch % bkh) continue;
            if ((bk_only > 0 && bk != bk_only) || (bkh_only > 0 && bkh != bkh_only)) continue;
            const int dbl = pass < 2 ? 1 : 0;
            if (use_hint && d->dbl_hint >= 0 && dbl != (d->dbl_hint ? 1 : 0)) continue;
            const int chunks1 = d->cin / bk, hblocks = d->ch / bkh, xchunks = has_x ? d->cin / bk : 0;
            const int nkb1 = d->taps * chunks1;
            const uint32_t w1_kb = round_up((uint32_t)d->ch * bk * 2, 1024), w2h_kb = round_up((uint32_t)d->cout * bkh * 2, 1024),
                           w2x_kb = round_up((uint32_t)d->cout * bk * 2, 1024);
            const size_t w_res_total = (size_t)nkb1 * w1_kb * pl1 + ((size_t)hblocks * w2h_kb + (size_t)xchunks * w2x_kb) * pl2;
            const bool resident = w_res_total <= 64 * 1024;
            for (int G : {4, 2, 1}) {
                if (g_only > 0 && G != g_only) continue;
                if (use_hint && d->g_hint > 0 && G != d->g_hint) continue;
                const int need_cols = G * ((1 + dbl) * d->ch + (1 + (dbl && want_pp ? 1 : 0)) * d->cout);
                if (need_cols > 512) continue;
                if (!(use_hint && d->g_hint > 0) && G > 1 && (m_tiles / G) * d->batch < 2LL * sms) continue;
                const int R = G * TILE_M + halo;
                const int a_pieces = (R + 255) / 256, a_box = (int)round_up((R + a_pieces - 1) / a_pieces, 8);
                const int x_pieces = (G * TILE_M + 255) / 256, x_box = G * TILE_M / x_pieces;
                uint32_t a_plane = round_up((uint32_t)a_pieces * a_box * bk * 2, 1024);
                if (has_x) { const uint32_t xb = (uint32_t)G * TILE_M * bk * 2; if (xb > a_plane) a_plane = xb; }
                const uint32_t a_stage_raw = a_plane * (1 + any_lo);
                const uint32_t e_stage = raw ? a_plane * (1 + e_split) : 0;
                const uint32_t a_stage = a_stage_raw + e_stage;  // ring cost per stage (raw block + activated block)
                const int min_a = x_from_a ? (1 + dbl) * chunks1 : 2;
                const uint32_t h_blk = (uint32_t)G * TILE_M * bkh * 2, h_plane = h_blk * hblocks;
                const size_t h_stage = (size_t)h_plane * (1 + h_split);
                const size_t h_total = h_stage * (1 + dbl);
                uint32_t w_plane = w1_kb > w2h_kb ? w1_kb : w2h_kb;
                if (xchunks && w2x_kb > w_plane) w_plane = w2x_kb;
                const uint32_t w_stage = w_plane * (pl1 > pl2 ? pl1 : pl2);
                uint32_t cols = 32;
                while (cols < (uint32_t)need_cols) cols <<= 1;
                const size_t budget = SMEM_LIMIT - fixed;
                const size_t io_total = use_io ? (size_t)io_planes * G * TILE_M * d->cout * 2 : 0;
                if (h_total + io_total >= budget) continue;
                const size_t rest = budget - h_total - io_total;
                int a_stages, w_stages = 0;
                size_t w_area;
                if (resident) {
                    if (w_res_total + 2 * (size_t)a_stage > rest) continue;
                    w_area = w_res_total;
                    a_stages = (int)((rest - w_res_total) / a_stage);
                } else {
                    w_stages = 4;
                    const int need_a = pass == 0 ? 3 : 2;
                    while (w_stages > 2 && (size_t)w_stages * w_stage + need_a * (size_t)a_stage > rest) --w_stages;
                    if ((size_t)w_stages * w_stage + 2 * (size_t)a_stage > rest) continue;
                    a_stages = (int)((rest - (size_t)w_stages * w_stage) / a_stage);
                    if (pass == 0 && (w_stages < 3 || a_stages < 3)) continue;
                    if (a_stages > 4) {
                        const int extra = (int)((rest - (size_t)w_stages * w_stage - 4 * (size_t)a_stage) / w_stage);
                        w_stages = w_stages + extra > MAX_W_STAGES ? MAX_W_STAGES : w_stages + extra;
                        a_stages = (int)((rest - (size_t)w_stages * w_stage) / a_stage);
                    }
                    w_area = (size_t)w_stages * w_stage;
                }
                if (a_stages > MAX_A_STAGES) a_stages = MAX_A_STAGES;
                if (a_stages < 2 || a_stages < min_a || (pass == 0 && a_stages < (x_from_a ? min_a + 1 : 3))) continue;
                p.bk = bk; p.bkh = bkh; p.G = G; p.chunks1 = chunks1; p.hblocks = hblocks; p.xchunks = xchunks;
                p.a_pieces = a_pieces; p.a_box_rows = a_box; p.x_pieces = x_pieces; p.x_box_rows = x_box;
                p.a_stages = a_stages; p.w_stages = w_stages; p.w_resident = resident ? 1 : 0;
                p.a_stage_bytes = a_stage_raw; p.a_plane_bytes = a_plane; p.e_stage_bytes = e_stage; p.e_plane_bytes = a_plane; p.w_stage_bytes = w_stage; p.w_plane_bytes = w_plane;
                p.w1_kb_bytes = w1_kb; p.w2h_kb_bytes = w2h_kb; p.w2x_kb_bytes = w2x_kb;
                p.w1_res_plane = (uint32_t)nkb1 * w1_kb; p.w2h_res_plane = (uint32_t)hblocks * w2h_kb; p.w2x_res_plane = (uint32_t)xchunks * w2x_kb;
                p.w2h_res_off = (uint32_t)nkb1 * w1_kb * pl1;
                p.w2x_res_off = p.w2h_res_off + (uint32_t)hblocks * w2h_kb * pl2;
                p.w_area_bytes = (uint32_t)w_area;
                p.h_blk_bytes = h_blk; p.h_plane_bytes = h_plane; p.h_stage_bytes = (uint32_t)h_stage;
                p.dbl = dbl; p.pp = (dbl && want_pp) ? 1 : 0; p.acc1_stride = (uint32_t)G * d->ch; p.acc2_stride = (uint32_t)G * d->cout;
                p.tmem_cols = cols; p.acc2_col = (uint32_t)G * d->ch * (1 + dbl);
                return true;
            }
        }
    return false;
    };
    // the canonical contraction blocks are those of the un-staged G = 1 search (staging must not change the arithmetic)
    const bool want_io = use_io;
    use_io = false;
    bool found = search(1, 0, 0, false);
    if (found) {
        const int bk_c = p.bk, bkh_c = p.bkh;
        const bool hinted = d->g_hint > 0 || d->dbl_hint >= 0;
        use_io = want_io;
        bool ok = use_io && search(0, bk_c, bkh_c, true);
        if (!ok) {
            AC_REQUIRE(d->io_stage <= 0, "ac_resunit_tc: staged epilogue I/O does not fit shared memory for this tiling");
            use_io = false;
            if (!search(0, bk_c, bkh_c, true)) found = hinted ? false : search(1, bk_c, bkh_c, false);
        }
    }
    AC_REQUIRE(found, "ac_resunit_tc: no tiling fits shared memory (cin %d ch %d cout %d taps %d)", d->cin, d->ch, d->cout, d->taps);

    RuMaps maps;
    int r = encode_act_map(encode, &maps.a, d->a, d->cin, d->a_rows, d->a_row_stride, d->a_bstride, d->batch, p.bk, p.a_box_rows);
    AC_REQUIRE(r == 0, "ac_resunit_tc: tensor map (A) failed: %d", r);
    maps.a_lo = maps.a;
    if (a_has_lo) {
        r = encode_act_map(encode, &maps.a_lo, d->a_lo, d->cin, d->a_rows, d->a_row_stride, d->a_bstride, d->batch, p.bk, p.a_box_rows);
        AC_REQUIRE(r == 0, "ac_resunit_tc: tensor map (A lo) failed: %d", r);
    }
    maps.x = maps.a; maps.x_lo = maps.a;
    if (has_x && !x_from_a) {
        r = encode_act_map(encode, &maps.x, d->x, d->cin, d->m_rows, d->cin, d->x_bstride, d->batch, p.bk, p.x_box_rows);
        AC_REQUIRE(r == 0, "ac_resunit_tc: tensor map (X) failed: %d", r);
        if (x_has_lo) {
            r = encode_act_map(encode, &maps.x_lo, d->x_lo, d->cin, d->m_rows, d->cin, d->x_bstride, d->batch, p.bk, p.x_box_rows);
            AC_REQUIRE(r == 0, "ac_resunit_tc: tensor map (X lo) failed: %d", r);
        }
    }
    r = encode_w_map(encode, &maps.w1, d->w1, d->taps * d->cin, d->ch, pl1, p.bk);
    AC_REQUIRE(r == 0, "ac_resunit_tc: tensor map (W1) failed: %d", r);
    r = encode_w_map(encode, &maps.w2h, d->w2, d->ch + (has_x ? d->cin : 0), d->cout, pl2, p.bkh);
    AC_REQUIRE(r == 0, "ac_resunit_tc: tensor map (W2 hidden part) failed: %d", r);
    r = encode_w_map(encode, &maps.w2x, d->w2, d->ch + (has_x ? d->cin : 0), d->cout, pl2, p.bk);
    AC_REQUIRE(r == 0, "ac_resunit_tc: tensor map (W2 x part) failed: %d", r);

    p.raw = raw; p.act0 = d->act0; p.e_split = e_split; p.x_from_a = x_from_a; p.sc_row_off = d->x_row_off; p.alpha0 = d->alpha0;
    p.cin = d->cin; p.taps = d->taps; p.dil = d->dilation; p.shift = d->shift; p.a_has_lo = a_has_lo; p.x_has_lo = x_has_lo;
    p.ch = d->ch; p.cout = d->cout; p.h_split = h_split; p.w1_split = w1_split; p.w2_split = w2_split;
    p.io_stage = (found && use_io) ? 1 : 0; p.bko = bko; p.io_y_off = io_res; p.io_act_off = io_res + io_y; p.io_planes = io_planes;
    p.io_blk_bytes = (uint32_t)p.G * TILE_M * bko * 2; p.io_plane_bytes = (uint32_t)p.G * TILE_M * d->cout * 2;
    maps.res = maps.a; maps.res_lo = maps.a; maps.y = maps.a; maps.y_lo = maps.a; maps.ya = maps.a; maps.ya_lo = maps.a;
    if (p.io_stage) {
        auto io_map = [&](CUtensorMap* map, const void* base, int64_t bstride) -> int {
            if (!base) return 0;
            cuuint64_t gdim[3] = {(cuuint64_t)d->cout, (cuuint64_t)d->m_rows, (cuuint64_t)d->batch};
            cuuint64_t gstr[2] = {(cuuint64_t)d->cout * 2, (cuuint64_t)bstride * 2};
            cuuint32_t box[3] = {(cuuint32_t)bko, (cuuint32_t)TILE_M, 1};
            cuuint32_t est[3] = {1, 1, 1};
            return (int)encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               swizzle_for(bko), CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        };
        int e = io_map(&maps.res, d->res, d->res_bstride) | io_map(&maps.res_lo, d->res_lo, d->res_bstride) | io_map(&maps.y, d->y, d->y_bstride) |
                io_map(&maps.y_lo, d->y_lo, d->y_bstride) | io_map(&maps.ya, d->y_act, d->y_act_bstride) | io_map(&maps.ya_lo, d->y_act_lo, d->y_act_bstride);
        AC_REQUIRE(e == 0, "ac_resunit_tc: tensor map (staged epilogue I/O) failed: %d", e);
    }
    p.f16 = f16; p.w1_hib = w1_hib; p.w2_hib = w2_hib;
    p.y_f16 = (d->fmt & AC_FMT_Y_F16) ? 1 : 0; p.ya_f16 = (d->fmt & AC_FMT_YACT_F16) ? 1 : 0; p.res_f16 = (d->fmt & AC_FMT_RES_F16) ? 1 : 0;
    p.m_rows = d->m_rows; p.m_groups = (d->m_rows + p.G * TILE_M - 1) / (p.G * TILE_M); p.batch = d->batch;
    p.bias1 = d->bias1; p.alpha1 = d->alpha1; p.bias2 = d->bias2; p.alpha2 = d->alpha2; p.act1 = d->act1; p.act2 = d->act2;
    p.res = (const __nv_bfloat16*)d->res; p.res_lo = (const __nv_bfloat16*)d->res_lo;
    p.y = (__nv_bfloat16*)d->y; p.y_lo = (__nv_bfloat16*)d->y_lo; p.y_act = (__nv_bfloat16*)d->y_act; p.y_act_lo = (__nv_bfloat16*)d->y_act_lo;
    p.res_bs = d->res_bstride; p.y_bs = d->y_bstride; p.ya_bs = d->y_act_bstride;

    // every CTA owns its SM (tensor memory is allocated in full): ask for more than half the shared memory
    size_t smem = fixed + (size_t)p.a_stages * (p.a_stage_bytes + p.e_stage_bytes) + p.w_area_bytes + (size_t)p.h_stage_bytes * (1 + p.dbl) +
                  (p.io_stage ? (size_t)io_planes * p.io_plane_bytes : 0);
    AC_REQUIRE(smem <= (size_t)SMEM_LIMIT, "ac_resunit_tc: shared memory %zu", smem);
    if (smem <= (size_t)SMEM_HALF + 1024) smem = SMEM_HALF + 2048;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(resunit_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(resunit_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
        if (e != cudaSuccess) { ac::set_error("ac_resunit_tc: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        attr_set = true;
    }
    const long long total_tiles = (long long)p.m_groups * p.batch;
    long long grid = sms;
    if (d->grid_hint > 0) grid = d->grid_hint;
    if (total_tiles < grid) grid = total_tiles;
    if (p.raw) resunit_tc_kernel<true><<<(unsigned)grid, THREADS, smem, (cudaStream_t)stream>>>(maps, p);
    else resunit_tc_kernel<false><<<(unsigned)grid, THREADS - 32 * XF_WARPS, smem, (cudaStream_t)stream>>>(maps, p);
    return ac::finish_launch("ac_resunit_tc");
}
