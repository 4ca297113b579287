// Residual vector quantizer encode on tcgen05 tensor cores (ac_rvq_encode_tc in include/audiocodecs_b200.h).
//
// Replaces EncodecResidualVectorQuantizer.encode + EncodecEuclideanCodebook.quantize (HF/encodec:364-369,424-438):
//   per stage k:  idx = argmax_c -(|r|^2 - 2 r.E_k[c] + |E_k[c]|^2)   (lowest index wins ties);   r -= E_k[idx]
//
// One persistent CTA owns a tile of 128 frames for ALL stages: the fp32 residual never leaves shared memory.
//   * distance GEMM  r[128 x 128] . E_k^T[128 x 1024]  on tcgen05 with error-compensated bf16 operands
//     (r_hi.E_hi + r_hi.E_lo + r_lo.E_hi, fp32 accumulate in TMEM: ~2^-16 relative; a plain bf16 GEMM matches the
//     fp32 argmin for only 38 % of stage-0 frames).  The codebook (hi/lo bf16 planes, L2-resident)
//     streams through a TMA ring in chunks of 64 codes, continuously across stages and tiles.
//   * TMEM holds two groups of 256 distance columns: the eight epilogue warps scan group g (running top-2 per
//     frame: lane = frame, the two warps of a lane quarter split the columns) while the MMA warp fills group g+1.
//   * the top-2 candidates of a frame are re-scored in exact fp32 (SIMT, reference formula and tie rule), so the
//     emitted code is the fp32 argmin wherever the tensor-core ranking has the true winner in its top 2; the
//     winner is subtracted from the fp32 residual in place and the operand planes are re-split for the next stage.
#include <cuda_bf16.h>

#include "common.cuh"
#include "sm100.cuh"

namespace {

using namespace sm100;

constexpr int D = 128;            // embedding dimension
constexpr int TM = 128;           // frames per tile
constexpr int CHUNK = 64;         // codes per ring stage / per MMA N
constexpr int GROUP = 256;        // codes per TMEM buffer
constexpr int RING = 5;
constexpr int EPI_WARPS = 8;
constexpr int THREADS = 64 + 32 * EPI_WARPS;
constexpr uint32_t PLANE_KB_BYTES = TM * 128;             // one [128 rows x 64] SW128 block of an A plane: 16 KB
constexpr uint32_t A_PLANE_BYTES = 2 * PLANE_KB_BYTES;    // two k-blocks: 32 KB
constexpr uint32_t E_BLOCK_BYTES = CHUNK * 128;           // [64 codes x 64 dims] SW128: 8 KB
constexpr uint32_t RING_STAGE_BYTES = 2 * E_BLOCK_BYTES;  // one k-block of a chunk: E_hi, E_lo = 16 KB
constexpr int XCH = 8;                                     // floats exchanged per (frame, half)

struct RvqParams {
    const float* x;          // [rows][D]
    const float* cb;         // [S][n_codes][D] fp32 (exact re-score + subtract)
    const float* cbn;        // [S][n_codes] |E|^2
    int64_t* codes;
    float* res_out;          // optional [rows][D]
    long long rows;
    int n_codes, stages, code_stride, code_offset, tiles;
    int lo_row0;             // first row of the lo plane in the codebook tensor map
};

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// (score, index) ordering of the reference: larger score first, lower index on ties
__device__ __forceinline__ bool better(float s, int i, float t, int j) { return s > t || (s == t && i < j); }

__device__ __forceinline__ void top2_insert(float s, int c, float& v1, int& i1, float& v2, int& i2) {
    if (better(s, c, v1, i1)) { v2 = v1; i2 = i1; v1 = s; i1 = c; }
    else if (better(s, c, v2, i2)) { v2 = s; i2 = c; }
}

__global__ void __launch_bounds__(THREADS, 1)
rvq_encode_tc_kernel(const __grid_constant__ CUtensorMap emap, const RvqParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a_hi = smem;                                   // 32 KB
    uint8_t* a_lo = a_hi + A_PLANE_BYTES;                   // 32 KB
    uint8_t* ring = a_lo + A_PLANE_BYTES;                   // RING x 16 KB
    float* R = reinterpret_cast<float*>(ring + RING * RING_STAGE_BYTES);  // [D][TM] fp32, dim-major: 64 KB
    float* en_s = R + D * TM;                               // [n_codes <= 1024]
    float* xch = en_s + 1024;                           // [TM][2][XCH] exchange between the two halves of a frame
    uint64_t* bars = reinterpret_cast<uint64_t*>(xch + TM * 2 * XCH);
    uint64_t* full = bars;                  // [RING]
    uint64_t* empty = bars + RING;          // [RING]
    uint64_t* tfull = bars + 2 * RING;      // [2]
    uint64_t* tempty = tfull + 2;           // [2]
    uint64_t* a_ready = tempty + 2;         // operand planes of the next stage are written
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_ready + 1);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const int groups = p.n_codes / GROUP;

    if (threadIdx.x == 0) {
        prefetch_tensormap(&emap);
        for (int i = 0; i < RING; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], EPI_WARPS); }
        mbar_init(a_ready, EPI_WARPS);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
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
                        for (int kb = 0; kb < 2; ++kb) {
                            mbar_wait(&empty[rs], ph ^ 1);
                            mbar_arrive_expect_tx(&full[rs], RING_STAGE_BYTES);
                            uint8_t* dst = ring + (size_t)rs * RING_STAGE_BYTES;
                            const int row = k * p.n_codes + c0;
                            tma_load_2d(dst, &emap, &full[rs], kb * 64, row);
                            tma_load_2d(dst + E_BLOCK_BYTES, &emap, &full[rs], kb * 64, p.lo_row0 + row);
                            if (++rs == RING) { rs = 0; ph ^= 1; }
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
                        for (int kb = 0; kb < 2; ++kb) {
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
                            if (++rs == RING) { rs = 0; ph ^= 1; }
                        }
                    }
                    if (leader) umma_commit(&tfull[buf]);
                    __syncwarp();
                }
            }
    } else {
        // ================================================================= epilogue: 8 warps, thread = (frame, half)
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;            // which 128 of a group's 256 columns / which 64 dims this thread owns
        const int r = quarter * 32 + lane;           // frame within the tile == TMEM lane
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        float* my_x = xch + (r * 2 + half) * XCH;
        const float* peer_x = xch + (r * 2 + (half ^ 1)) * XCH;
        uint8_t* my_hi = a_hi + half * PLANE_KB_BYTES + r * 128;
        uint8_t* my_lo = a_lo + half * PLANE_KB_BYTES + r * 128;
        const int d0 = half * 64;
        uint32_t gcount = 0;

        // writes the operand planes of this thread's 64 dims from the fp32 residual (already in R), returns sum of squares
        auto split_planes = [&]() {
            float ss = 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) { v[j] = R[(d0 + u * 8 + j) * TM + r]; ss = fmaf(v[j], v[j], ss); }
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    hi[j] = pack2(v[2 * j], v[2 * j + 1]);
                    const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&hi[j]);
                    lo[j] = pack2(v[2 * j] - __low2float(h2), v[2 * j + 1] - __high2float(h2));
                }
                const uint32_t off = (uint32_t)((u ^ (r & 7)) << 4);  // SW128: 16-byte unit index XOR (row & 7)
                *reinterpret_cast<uint4*>(my_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4*>(my_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
            return ss;
        };

        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
            const long long row = (long long)tile * TM + r;
            const bool valid = row < p.rows;
            // ---- load the frame's embedding (this thread's 64 dims) into the fp32 residual
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const float4 t = valid ? __ldg(reinterpret_cast<const float4*>(p.x + row * D + d0) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                R[(d0 + 4 * q + 0) * TM + r] = t.x; R[(d0 + 4 * q + 1) * TM + r] = t.y;
                R[(d0 + 4 * q + 2) * TM + r] = t.z; R[(d0 + 4 * q + 3) * TM + r] = t.w;
            }
            float xn_part = split_planes();
            for (int k = 0; k < p.stages; ++k) {
                float* en = en_s;
                for (int i = threadIdx.x - 64; i < p.n_codes; i += 32 * EPI_WARPS) en[i] = __ldg(p.cbn + (size_t)k * p.n_codes + i);
                my_x[0] = xn_part;
                fence_proxy_async();  // operand planes (generic-proxy stores) -> visible to tcgen05 (async proxy)
                epi_bar();            // en / xn parts complete; every thread's plane stores are fenced
                if (lane == 0) mbar_arrive(a_ready);
                const float xn = half == 0 ? xn_part + peer_x[0] : peer_x[0] + xn_part;  // dims 0..63 first, on both halves

                // Candidate selection: per frame the two smallest keys of this thread's 512 codes, where
                //   key = (bits(max(dist, 0)) & ~511) | local_index,   dist = (|r|^2 + |E|^2) - 2 r.E  >= 0.
                // Non-negative floats order like their bit patterns, so ONE integer min/max chain tracks value and index
                // together (3 ops per code instead of compare/select pairs); dropping 9 mantissa bits (2^-14 relative)
                // cannot push the true winner out of the top 2 unless its gap to the runner-up is below 6e-5 -- a near-tie
                // (< 1e-4) by the parity definition -- and the two candidates are re-scored exactly below.
                uint32_t k1 = 0xFFFFFFFFu, k2 = 0xFFFFFFFFu;
                for (int g = 0; g < groups; ++g, ++gcount) {
                    const uint32_t buf = gcount & 1;
                    mbar_wait(&tfull[buf], (gcount >> 1) & 1);
                    tc_fence_after();
#pragma unroll 2
                    for (int cc = 0; cc < 8; ++cc) {
                        const int col = half * 128 + cc * 16;
                        uint32_t v[16];
                        tmem_ld16(lane_addr + buf * GROUP + col, v);
                        const float4* enp = reinterpret_cast<const float4*>(en + g * GROUP + col);
                        float e[16];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 t = enp[q];
                            e[4 * q] = xn + t.x; e[4 * q + 1] = xn + t.y; e[4 * q + 2] = xn + t.z; e[4 * q + 3] = xn + t.w;
                        }
                        const uint32_t lbase = (uint32_t)(g * 128 + cc * 16);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float dist = fmaxf(fmaf(-2.f, __uint_as_float(v[j]), e[j]), 0.f);
                            const uint32_t key = (__float_as_uint(dist) & 0xFFFFFE00u) | (lbase + j);
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
                auto glob = [&](uint32_t k, int h) { const int l = (int)(k & 511u); return (l >> 7) * GROUP + h * 128 + (l & 127); };
                float v1 = -__uint_as_float(k1 & 0xFFFFFE00u), v2 = -__uint_as_float(k2 & 0xFFFFFE00u);  // scores: larger is better
                int i1 = glob(k1, half), i2 = glob(k2, half);
                my_x[1] = v1; my_x[2] = __int_as_float(i1); my_x[3] = v2; my_x[4] = __int_as_float(i2);
                epi_bar();
                top2_insert(peer_x[1], __float_as_int(peer_x[2]), v1, i1, v2, i2);
                top2_insert(peer_x[3], __float_as_int(peer_x[4]), v1, i1, v2, i2);
                // ---- exact fp32 re-score of both candidates: partial dots over this thread's 64 dims
                const float* E = p.cb + (size_t)k * p.n_codes * D;
                const bool two = i2 < p.n_codes;
                const float* e1 = E + (size_t)i1 * D + d0;
                const float* e2 = E + (size_t)(two ? i2 : i1) * D + d0;
                float p1 = 0.f, p2 = 0.f;
#pragma unroll 4
                for (int q = 0; q < 16; ++q) {
                    const float4 a = __ldg(reinterpret_cast<const float4*>(e1) + q);
                    const float4 bq = __ldg(reinterpret_cast<const float4*>(e2) + q);
                    const float r0 = R[(d0 + 4 * q + 0) * TM + r], r1 = R[(d0 + 4 * q + 1) * TM + r];
                    const float r2 = R[(d0 + 4 * q + 2) * TM + r], r3 = R[(d0 + 4 * q + 3) * TM + r];
                    p1 = fmaf(r0, a.x, p1); p1 = fmaf(r1, a.y, p1); p1 = fmaf(r2, a.z, p1); p1 = fmaf(r3, a.w, p1);
                    p2 = fmaf(r0, bq.x, p2); p2 = fmaf(r1, bq.y, p2); p2 = fmaf(r2, bq.z, p2); p2 = fmaf(r3, bq.w, p2);
                }
                my_x[5] = p1; my_x[6] = p2;
                const float en1 = en[i1], en2 = two ? en[i2] : 0.f;  // read before the barrier: en is rewritten for stage k+1 after it
                epi_bar();
                const float dot1 = half == 0 ? p1 + peer_x[5] : peer_x[5] + p1;
                const float dot2 = half == 0 ? p2 + peer_x[6] : peer_x[6] + p2;
                const float s1 = -((xn - 2.f * dot1) + en1);
                const float s2 = two ? -((xn - 2.f * dot2) + en2) : -INFINITY;
                const int sel = (two && better(s2, i2, s1, i1)) ? i2 : i1;
                if (half == 0 && valid) p.codes[row * p.code_stride + p.code_offset + k] = (int64_t)sel;
                // ---- subtract the winner in place (this thread's 64 dims) and re-split the operand planes
                const float* es = E + (size_t)sel * D + d0;
#pragma unroll 4
                for (int q = 0; q < 16; ++q) {
                    const float4 a = __ldg(reinterpret_cast<const float4*>(es) + q);
                    R[(d0 + 4 * q + 0) * TM + r] -= a.x; R[(d0 + 4 * q + 1) * TM + r] -= a.y;
                    R[(d0 + 4 * q + 2) * TM + r] -= a.z; R[(d0 + 4 * q + 3) * TM + r] -= a.w;
                }
                if (k + 1 < p.stages) xn_part = split_planes();  // all MMAs of this stage have retired (last tfull seen)
            }
            if (p.res_out && valid) {
#pragma unroll
                for (int q = 0; q < 16; ++q)
                    reinterpret_cast<float4*>(p.res_out + row * D + d0)[q] =
                        make_float4(R[(d0 + 4 * q + 0) * TM + r], R[(d0 + 4 * q + 1) * TM + r], R[(d0 + 4 * q + 2) * TM + r],
                                    R[(d0 + 4 * q + 3) * TM + r]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

extern "C" int ac_rvq_encode_tc(const float* x, const void* cb_split_bf16, const float* codebooks, const float* cb_norm,
                                int64_t* codes_out, float* residual_out, int64_t rows, int32_t dim, int32_t n_codes,
                                int32_t stages, int32_t stages_total, int32_t code_stride, int32_t code_offset, void* stream) {
    AC_REQUIRE(x && cb_split_bf16 && codebooks && cb_norm && codes_out, "ac_rvq_encode_tc: null pointer");
    AC_REQUIRE(rows > 0 && stages > 0 && stages <= stages_total, "ac_rvq_encode_tc: empty problem");
    AC_REQUIRE(dim == D, "ac_rvq_encode_tc: dim %d (this kernel is built for %d)", dim, D);
    AC_REQUIRE(n_codes % GROUP == 0 && n_codes <= 1024, "ac_rvq_encode_tc: n_codes %d must be a multiple of %d, <= 1024", n_codes, GROUP);
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
        cuuint64_t gdim[2] = {(cuuint64_t)D, (cuuint64_t)2 * stages_total * n_codes};  // hi plane rows, then lo plane rows
        cuuint64_t gstr[1] = {(cuuint64_t)D * 2};
        cuuint32_t box[2] = {64, CHUNK};
        cuuint32_t est[2] = {1, 1};
        CUresult r = encode(&emap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(cb_split_bf16), gdim, gstr, box, est,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        AC_REQUIRE(r == CUDA_SUCCESS, "ac_rvq_encode_tc: cuTensorMapEncodeTiled failed: %d", (int)r);
    }
    RvqParams p{};
    p.x = x; p.cb = codebooks; p.cbn = cb_norm; p.codes = codes_out; p.res_out = residual_out;
    p.rows = rows; p.n_codes = n_codes; p.stages = stages; p.code_stride = code_stride; p.code_offset = code_offset;
    p.tiles = (int)((rows + TM - 1) / TM);
    p.lo_row0 = stages_total * n_codes;
    const size_t smem = 1024 + 2 * A_PLANE_BYTES + RING * RING_STAGE_BYTES + (size_t)D * TM * 4 + 1024 * 4 + TM * 2 * XCH * 4 +
                        (2 * RING + 5) * 8 + 16;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(rvq_encode_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { ac::set_error("ac_rvq_encode_tc: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        configured = true;
    }
    int sms = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = p.tiles < sms ? p.tiles : sms;
    rvq_encode_tc_kernel<<<grid, THREADS, smem, (cudaStream_t)stream>>>(emap, p);
    return ac::finish_launch("ac_rvq_encode_tc");
}
