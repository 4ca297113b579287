// bf16 tensor-core tap-GEMM convolution for sm_100a (ac_conv_tc in include/audiocodecs_b200.h).
//
//   acc[b][m][n] = sum over sources s, taps j, channels k :  A_s[b][m + j*dil_s + shift_s][k] * W[n][kcol(s,j,k)]
//
// A_s are channels-last bf16 activation views [batch][rows][phases*c0] described by 4-D TMA tensor
// maps (c0, phase, row, batch): a stride-s conv with kernel 2s is the 2-tap GEMM over the view whose
// row is s consecutive time steps (no im2col, no padded copy; zero padding is TMA out-of-bounds fill,
// reflect padding lives in the producer-written halo rows of the activation buffer).  W is the packed
// K-major weight matrix [n_total][k_total].  Per 128-row tile the accumulator lives in TMEM; a
// persistent CTA runs three roles:
//   warp 0   : TMA producer   (A box [128 x BK] + W box [n_tile x BK] per k-block into a smem ring)
//   warp 1   : tcgen05.mma issuer (one lane), commits free the ring slot / publish the accumulator
//   warps 2-5: epilogue: tcgen05.ld -> bias (+ residual) -> bf16 / fp32 stores, optional second
//              output with the CONSUMER's activation (ELU / Snake) so no layer ever re-reads raw+act.
// TMEM holds two accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1.
#include <cuda_bf16.h>

#include "common.cuh"
#include "sm100.cuh"

namespace {

using namespace sm100;

constexpr int TILE_M = 128;
constexpr int MAX_STAGES = 8;
constexpr int EPI_WARPS = 8;                 // two warps per TMEM lane quarter, interleaved over 16-column chunks
constexpr int THREADS = 64 + 32 * EPI_WARPS;

constexpr int MAX_SRC = 4;

struct TcSrc {
    int c0;      // innermost tensor-map dim (channels per phase)
    int taps, dil, shift;
    int chunks;  // k-blocks per tap
    int kb0;     // first weight k-block (a lo-plane source re-uses the columns of its hi twin)
    int nb;      // weight tiles per k-block: 2 = W_hi and W_lo (hi-plane source, split weights), 1 = W_hi only
};

struct TcParams {
    TcSrc src[MAX_SRC];
    int n_src, bk, num_kb;
    int n_total, n_tile, n_tiles, m_rows, m_tiles, batch;
    int stages;
    uint32_t a_stage_bytes, b_stage_bytes, tmem_cols;
    const float* bias;
    const float* alpha;
    const __nv_bfloat16* res;
    const __nv_bfloat16* res_lo;  // optional lo plane of the residual
    const float* res32;           // optional fp32 residual (Mimi's fp32 residual stream)
    __nv_bfloat16* y;
    __nv_bfloat16* y_act;
    __nv_bfloat16* y_lo;      // optional lo planes: lo = bf16(v - float(bf16(v)))
    __nv_bfloat16* y_act_lo;
    float* y32;
    int w_rows;               // rows of one weight plane in the B tensor map (n_total); W_lo starts at row w_rows
    int act, epi, act_mod;
    long long y_bs, ya_bs, y32_bs, res_bs, out_shift, out_valid;
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ void add_bf16x8(float (&o)[8], const __nv_bfloat16* ptr) {
    const uint4 r = *reinterpret_cast<const uint4*>(ptr);
    const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&rr[i]);
        o[2 * i] += __low2float(h2);
        o[2 * i + 1] += __high2float(h2);
    }
}

// lo plane of 8 values: bf16(v - float(hi)) where hi is the already-packed bf16 rounding of v
__device__ __forceinline__ uint4 pack_lo(const float (&v)[8], const uint4& hi) {
    const uint32_t h[4] = {hi.x, hi.y, hi.z, hi.w};
    uint32_t r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&h[i]);
        r[i] = pack_bf16(v[2 * i] - __low2float(h2), v[2 * i + 1] - __high2float(h2));
    }
    return make_uint4(r[0], r[1], r[2], r[3]);
}

__global__ void __launch_bounds__(THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap amap0, const __grid_constant__ CUtensorMap amap1,
               const __grid_constant__ CUtensorMap amap2, const __grid_constant__ CUtensorMap amap3,
               const __grid_constant__ CUtensorMap bmap, const TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // dynamic smem base is only guaranteed 16-B aligned: round up to 1024 for the swizzled tiles
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a_ring = smem;
    uint8_t* b_ring = smem + (size_t)p.stages * p.a_stage_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_ring + (size_t)p.stages * p.b_stage_bytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + MAX_STAGES;
    uint64_t* tfull = bars + 2 * MAX_STAGES;
    uint64_t* tempty = bars + 2 * MAX_STAGES + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 4);
    float* bias_s = reinterpret_cast<float*>(tmem_slot + 4);  // [n_total] (<= 8192 floats reserved by the host)

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int total_tiles = p.n_tiles * p.m_tiles * p.batch;

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&amap0);
        if (p.n_src > 1) prefetch_tensormap(&amap1);
        if (p.n_src > 2) prefetch_tensormap(&amap2);
        if (p.n_src > 3) prefetch_tensormap(&amap3);
        prefetch_tensormap(&bmap);
        for (int i = 0; i < p.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
    for (int i = threadIdx.x; i < p.n_total; i += THREADS) bias_s[i] = p.bias ? p.bias[i] : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================================================= TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t a_tx = TILE_M * p.bk * 2, b_tx = p.n_tile * p.bk * 2;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int nt = tile % p.n_tiles;
                const int rest = tile / p.n_tiles;
                const int mt = rest % p.m_tiles;
                const int b = rest / p.m_tiles;
                for (int s = 0; s < p.n_src; ++s) {
                    const TcSrc& S = p.src[s];
                    const CUtensorMap* am = s == 0 ? &amap0 : (s == 1 ? &amap1 : (s == 2 ? &amap2 : &amap3));
                    int kb = S.kb0;
                    for (int j = 0; j < S.taps; ++j) {
                        const int row = mt * TILE_M + j * S.dil + S.shift;
                        for (int cc = 0; cc < S.chunks; ++cc, ++kb) {
                            mbar_wait(&empty[stage], phase ^ 1);
                            mbar_arrive_expect_tx(&full[stage], a_tx + S.nb * b_tx);
                            const int flat = cc * p.bk;
                            uint8_t* bs = b_ring + (size_t)stage * p.b_stage_bytes;
                            tma_load_4d(a_ring + (size_t)stage * p.a_stage_bytes, am, &full[stage], flat % S.c0, flat / S.c0, row, b);
                            tma_load_2d(bs, &bmap, &full[stage], kb * p.bk, nt * p.n_tile);
                            if (S.nb == 2) tma_load_2d(bs + p.b_stage_bytes / 2, &bmap, &full[stage], kb * p.bk, p.w_rows + nt * p.n_tile);
                            if (++stage == p.stages) { stage = 0; phase ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================================================================= MMA issuer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t idesc = make_idesc_bf16(TILE_M, p.n_tile);
            const uint32_t sw = p.bk * 2;
            int it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                mbar_wait(&tempty[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * p.n_tile;
                int kb = 0;
                for (int s = 0; s < p.n_src; ++s) {
                    const int nkb = p.src[s].taps * p.src[s].chunks;
                    const int nb = p.src[s].nb;
                    for (int i = 0; i < nkb; ++i, ++kb) {
                        mbar_wait(&full[stage], phase);
                        tc_fence_after();
                        const uint64_t adesc = make_smem_desc(smem_u32(a_ring + (size_t)stage * p.a_stage_bytes), sw);
                        const uint32_t b_addr = smem_u32(b_ring + (size_t)stage * p.b_stage_bytes);
                        const uint64_t bdesc = make_smem_desc(b_addr, sw);
                        for (int k = 0; k < p.bk / 16; ++k)
                            umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);  // +32 B per K=16 step
                        if (nb == 2) {  // error-compensated weights: A * W_lo into the same accumulator
                            const uint64_t bdesc_lo = make_smem_desc(b_addr + p.b_stage_bytes / 2, sw);
                            for (int k = 0; k < p.bk / 16; ++k) umma_bf16(d_tmem, adesc + 2 * k, bdesc_lo + 2 * k, idesc, 1u);
                        }
                        umma_commit(&empty[stage]);                        // frees the ring slot when the MMAs retire
                        if (kb == p.num_kb - 1) umma_commit(&tfull[as]);   // accumulator complete
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else {
        // ================================================================= epilogue (warps 2..9)
        const int quarter = warp & 3;  // TMEM lane quarter this warp may read
        const int chunk0 = (warp - 2) >> 2;  // this warp takes chunks chunk0, chunk0+2, ...
        const int row_in_tile = quarter * 32 + lane;
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int nt = tile % p.n_tiles;
            const int rest = tile / p.n_tiles;
            const int mt = rest % p.m_tiles;
            const int b = rest / p.m_tiles;
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            mbar_wait(&tfull[as], aphase);
            tc_fence_after();
            const int m = mt * TILE_M + row_in_tile;
            const bool row_ok = m < p.m_rows;
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * p.n_tile;
            for (int c = chunk0; c < p.n_tile / 16; c += EPI_WARPS / 4) {
                uint32_t v[16];
                tmem_ld16(taddr + c * 16, v);
                tmem_ld_wait();
                const int n0 = nt * p.n_tile + c * 16;
                if (!row_ok || n0 >= p.n_total) continue;
                const long long flat = (long long)m * p.n_total + n0 - p.out_shift;
#pragma unroll
                for (int h = 0; h < 2; ++h) {  // two 8-element vectors (16 B of bf16 each)
                    const long long f = flat + h * 8;
                    if (f < 0 || f >= p.out_valid || n0 + h * 8 >= p.n_total) continue;
                    float o[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) o[i] = __uint_as_float(v[h * 8 + i]) + bias_s[n0 + h * 8 + i];
                    if (p.epi == AC_EPI_GELU) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) o[i] = ac::gelu_erf(o[i]);
                    }
                    if (p.res) {
                        add_bf16x8(o, p.res + (long long)b * p.res_bs + f);
                        if (p.res_lo) add_bf16x8(o, p.res_lo + (long long)b * p.res_bs + f);
                    }
                    if (p.res32) {
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
                        uint4 q;
                        q.x = pack_bf16(o[0], o[1]); q.y = pack_bf16(o[2], o[3]);
                        q.z = pack_bf16(o[4], o[5]); q.w = pack_bf16(o[6], o[7]);
                        *reinterpret_cast<uint4*>(p.y + (long long)b * p.y_bs + f) = q;
                        if (p.y_lo) *reinterpret_cast<uint4*>(p.y_lo + (long long)b * p.y_bs + f) = pack_lo(o, q);
                    }
                    if (p.y_act) {
                        float a[8];
                        if (p.act == AC_ACT_ELU) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) a[i] = ac::elu_fast(o[i]);
                        } else if (p.act == AC_ACT_SNAKE) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) a[i] = ac::snake(o[i], __ldg(p.alpha + (n0 + h * 8 + i) % p.act_mod));
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; ++i) a[i] = o[i];
                        }
                        uint4 q;
                        q.x = pack_bf16(a[0], a[1]); q.y = pack_bf16(a[2], a[3]);
                        q.z = pack_bf16(a[4], a[5]); q.w = pack_bf16(a[6], a[7]);
                        *reinterpret_cast<uint4*>(p.y_act + (long long)b * p.ya_bs + f) = q;
                        if (p.y_act_lo) *reinterpret_cast<uint4*>(p.y_act_lo + (long long)b * p.ya_bs + f) = pack_lo(a, q);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[as]);
        }
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
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

CUtensorMapSwizzle swizzle_for(int bk) {
    return bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (bk == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

int sm_count() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    return sms;
}

}  // namespace

extern "C" int ac_conv_tc(const ac_conv_tc_desc* d, void* stream) {
    AC_REQUIRE(d && d->w && d->n_src >= 1 && d->n_src <= MAX_SRC, "ac_conv_tc: bad descriptor");
    AC_REQUIRE(d->bk == 16 || d->bk == 32 || d->bk == 64, "ac_conv_tc: bk %d", d->bk);
    AC_REQUIRE(d->batch > 0 && d->m_rows > 0 && d->n_total >= 16 && d->n_total % 8 == 0 && d->n_total <= 8192,
               "ac_conv_tc: bad sizes (batch %d rows %d n %d)", d->batch, d->m_rows, d->n_total);
    AC_REQUIRE(d->y || d->y_act || d->y32, "ac_conv_tc: no output");
    AC_REQUIRE(d->out_shift % 8 == 0 && d->out_valid % 8 == 0, "ac_conv_tc: out_shift/out_valid must be multiples of 8");
    AC_REQUIRE(d->act != AC_ACT_SNAKE || (d->alpha && d->act_mod > 0), "ac_conv_tc: snake needs alpha/act_mod");
    EncodeTiledFn encode = get_encode();
    AC_REQUIRE(encode, "ac_conv_tc: cuTensorMapEncodeTiled not available");

    TcParams p{};
    p.n_src = d->n_src;
    p.bk = d->bk;
    int k_total = 0, num_kb = 0;
    const int w_split = d->w_split ? 1 : 0;
    CUtensorMap amap[MAX_SRC];
    for (int s = 0; s < d->n_src; ++s) {
        const ac_tc_src& S = d->src[s];
        AC_REQUIRE(S.base && S.c0 > 0 && S.phases > 0 && S.rows > 0 && S.taps > 0, "ac_conv_tc: bad source %d", s);
        const int kper = S.c0 * S.phases;  // contraction length per tap
        // a k-block never straddles two phases: a TMA box whose inner extent is narrower than the swizzle span
        // does not land in the UMMA canonical layout (measured: tests/test_conv_tc_gpu.py history)
        AC_REQUIRE(S.c0 % d->bk == 0, "ac_conv_tc: bk %d must divide c0 %d", d->bk, S.c0);
        AC_REQUIRE(((uintptr_t)S.base & 15) == 0 && (S.phase_stride * 2) % 16 == 0 && (S.row_stride * 2) % 16 == 0 &&
                       (S.batch_stride * 2) % 16 == 0, "ac_conv_tc: source %d not 16-byte aligned", s);
        p.src[s].c0 = S.c0;
        p.src[s].taps = S.taps;
        p.src[s].dil = S.dilation;
        p.src[s].shift = S.shift;
        p.src[s].chunks = kper / d->bk;
        if (S.lo_of >= 0) {
            // lo plane of source lo_of: same weight columns, W_hi only (the A_lo * W_lo term is below fp32 noise)
            AC_REQUIRE(S.lo_of < s && d->src[S.lo_of].c0 == S.c0 && d->src[S.lo_of].phases == S.phases &&
                           d->src[S.lo_of].taps == S.taps, "ac_conv_tc: source %d is not the lo twin of %d", s, S.lo_of);
            p.src[s].kb0 = p.src[S.lo_of].kb0;
            p.src[s].nb = 1;
        } else {
            p.src[s].kb0 = k_total / d->bk;
            p.src[s].nb = 1 + w_split;
            k_total += S.taps * kper;
        }
        num_kb += S.taps * p.src[s].chunks;
        const int box0 = S.c0 < d->bk ? S.c0 : d->bk;
        cuuint64_t gdim[4] = {(cuuint64_t)S.c0, (cuuint64_t)S.phases, (cuuint64_t)S.rows, (cuuint64_t)d->batch};
        cuuint64_t gstr[3] = {(cuuint64_t)S.phase_stride * 2, (cuuint64_t)S.row_stride * 2, (cuuint64_t)S.batch_stride * 2};
        cuuint32_t box[4] = {(cuuint32_t)box0, (cuuint32_t)(d->bk / box0), TILE_M, 1};
        cuuint32_t est[4] = {1, 1, 1, 1};
        CUresult r = encode(&amap[s], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(S.base), gdim, gstr, box, est,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(d->bk), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        AC_REQUIRE(r == CUDA_SUCCESS, "ac_conv_tc: cuTensorMapEncodeTiled(A%d) failed: %d", s, (int)r);
    }
    for (int s = d->n_src; s < MAX_SRC; ++s) amap[s] = amap[0];
    AC_REQUIRE(k_total == d->k_total, "ac_conv_tc: k_total %d != sum of taps*channels %d", d->k_total, k_total);
    p.num_kb = num_kb;
    p.w_rows = d->n_total;

    // N tile: the largest of {256,...,16} that does not over-pad
    int n_tile = 16;
    // with split weights a stage holds two W tiles: cap the tile at 128 columns to keep >= 4 ring stages
    for (int c : {256, 192, 128, 96, 64, 48, 32, 16})
        if ((!w_split || c <= 128) && d->n_total % c == 0) { n_tile = c; break; }
    if (d->n_total % n_tile != 0) n_tile = d->n_total >= 128 ? 128 : 16;
    if (d->n_tile_hint > 0) n_tile = d->n_tile_hint;
    AC_REQUIRE(n_tile % 16 == 0 && n_tile >= 16 && n_tile <= 256, "ac_conv_tc: n_tile %d", n_tile);
    p.n_total = d->n_total;
    p.n_tile = n_tile;
    p.n_tiles = (d->n_total + n_tile - 1) / n_tile;
    p.m_rows = d->m_rows;
    p.m_tiles = (d->m_rows + TILE_M - 1) / TILE_M;
    p.batch = d->batch;

    CUtensorMap bmap;
    {
        AC_REQUIRE(((uintptr_t)d->w & 15) == 0 && (k_total * 2) % 16 == 0, "ac_conv_tc: weights not 16-byte aligned");
        cuuint64_t gdim[2] = {(cuuint64_t)k_total, (cuuint64_t)d->n_total * (1 + w_split)};  // W_lo stacked under W_hi
        cuuint64_t gstr[1] = {(cuuint64_t)k_total * 2};
        cuuint32_t box[2] = {(cuuint32_t)d->bk, (cuuint32_t)n_tile};
        cuuint32_t est[2] = {1, 1};
        CUresult r = encode(&bmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(d->w), gdim, gstr, box, est,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(d->bk), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        AC_REQUIRE(r == CUDA_SUCCESS, "ac_conv_tc: cuTensorMapEncodeTiled(W) failed: %d", (int)r);
    }

    p.a_stage_bytes = (uint32_t)((TILE_M * d->bk * 2 + 1023) & ~1023);
    p.b_stage_bytes = (uint32_t)((n_tile * d->bk * 2 + 1023) & ~1023) * 2;  // room for the W_hi and W_lo tiles
    const size_t fixed = 1024 /*align slack*/ + (2 * MAX_STAGES + 4) * 8 + 16 + (size_t)d->n_total * 4 + 64;
    int stages = (int)((200 * 1024 - fixed) / (p.a_stage_bytes + p.b_stage_bytes));
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (stages > p.num_kb * 2 && p.num_kb * 2 >= 2) stages = p.num_kb * 2;
    AC_REQUIRE(stages >= 2, "ac_conv_tc: tile does not fit shared memory");
    p.stages = stages;
    uint32_t cols = 32;
    while (cols < (uint32_t)(2 * n_tile)) cols <<= 1;
    p.tmem_cols = cols;
    p.bias = d->bias; p.alpha = d->alpha;
    p.res = (const __nv_bfloat16*)d->res; p.y = (__nv_bfloat16*)d->y; p.y_act = (__nv_bfloat16*)d->y_act; p.y32 = d->y32;
    p.y_lo = (__nv_bfloat16*)d->y_lo; p.y_act_lo = (__nv_bfloat16*)d->y_act_lo;
    p.res_lo = (const __nv_bfloat16*)d->res_lo; p.res32 = d->res32;
    AC_REQUIRE((!p.y_lo || p.y) && (!p.y_act_lo || p.y_act), "ac_conv_tc: lo plane without its hi plane");
    p.act = d->act; p.epi = d->epi; p.act_mod = d->act_mod > 0 ? d->act_mod : d->n_total;
    p.y_bs = d->y_bstride; p.ya_bs = d->y_act_bstride; p.y32_bs = d->y32_bstride; p.res_bs = d->res_bstride;
    p.out_shift = d->out_shift; p.out_valid = d->out_valid;

    const size_t smem = fixed + (size_t)stages * (p.a_stage_bytes + p.b_stage_bytes);
    static size_t smem_set = 0;
    if (smem > smem_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024));
        if (e != cudaSuccess) { ac::set_error("ac_conv_tc: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        smem_set = 220 * 1024;
    }
    const long long total_tiles = (long long)p.n_tiles * p.m_tiles * p.batch;
    int grid = sm_count();
    if (d->grid_hint > 0) grid = d->grid_hint;
    if (total_tiles < grid) grid = (int)total_tiles;
    conv_tc_kernel<<<grid, THREADS, smem, (cudaStream_t)stream>>>(amap[0], amap[1], amap[2], amap[3], bmap, p);
    return ac::finish_launch("ac_conv_tc");
}
