// fp32 LSTM layer, recurrent part (see ac_lstm_layer_f32 in include/audiocodecs_b200.h).
//
// Persistent kernel: the time loop runs inside one launch (the reference lands on ATen's fused
// RNN op, HF/encodec:242-246).  CTA (ug, bg) owns HU=16 hidden units (their 4 gate rows of W_hh,
// 128 KB, resident in shared memory for the whole sequence) for a slice of BB=16 clips.  Per step
// it stages h[t-1] of its clips (32 KB), computes 16 clips x 16 units x 4 gates dot products of
// length C, applies the cell update (c stays in registers) and publishes h[t].  The only cross-CTA
// dependency is among the C/16 CTAs that share a batch slice, so the step barrier is a per-slice
// monotonic counter in global memory, not a grid-wide sync.  Co-residency of all CTAs is guaranteed
// by a cooperative launch (which fails instead of dead-locking if the grid does not fit).
#include <cooperative_groups.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace {

constexpr int HU = 16;       // hidden units per CTA
constexpr int BB = 16;       // clips per CTA
constexpr int THREADS = HU * BB;

__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(THREADS, 1)
lstm_layer_f32_kernel(const float* __restrict__ pre, const float* __restrict__ w_hh, const float* __restrict__ skip,
                      float* __restrict__ out, int batch, int steps, int C, int* sync_ws,
                      __nv_bfloat16* __restrict__ out_bf, const __nv_bfloat16* __restrict__ skip_bf,
                      __nv_bfloat16* __restrict__ final_bf, long long skip_bs, long long final_bs, int final_act,
                      __nv_bfloat16* __restrict__ out_lo, const __nv_bfloat16* __restrict__ skip_lo,
                      __nv_bfloat16* __restrict__ final_lo) {
    extern __shared__ __align__(16) float smem[];
    float4* Ws = reinterpret_cast<float4*>(smem);   // [C][HU] float4 = (i,f,g,o) rows for (k, unit)
    float* hs = smem + (size_t)C * HU * 4;          // [C][BB] previous hidden state, k-major

    const int ug = blockIdx.x;           // unit group
    const int bg = blockIdx.y;           // batch group
    const int n_ug = gridDim.x;
    const int tid = threadIdx.x;
    const int bl = tid % BB;             // local clip
    const int ul = tid / BB;             // local unit
    const int b = bg * BB + bl;
    const int u = ug * HU + ul;
    const bool live = b < batch;

    // W_hh is [4C][C] row-major (gate-major rows: i,f,g,o blocks of C rows)
    for (int e = tid; e < C * HU; e += THREADS) {
        const int k = e % C;
        const int uu = e / C;
        const int row = ug * HU + uu;
        float4 w;
        w.x = w_hh[(size_t)(0 * C + row) * C + k];
        w.y = w_hh[(size_t)(1 * C + row) * C + k];
        w.z = w_hh[(size_t)(2 * C + row) * C + k];
        w.w = w_hh[(size_t)(3 * C + row) * C + k];
        Ws[(size_t)k * HU + uu] = w;
    }
    __syncthreads();

    int* counter = sync_ws + bg;
    float c_state = 0.f;
    for (int t = 0; t < steps; ++t) {
        float gi = 0.f, gf = 0.f, gg = 0.f, go = 0.f;
        if (live) {
            const float* pp = pre + ((size_t)b * steps + t) * 4 * C + u;
            gi = pp[0]; gf = pp[C]; gg = pp[2 * C]; go = pp[3 * C];
        }
        if (t > 0) {
            // wait until every unit group of this batch slice has published h[t-1]
            if (tid == 0) {
                const int want = n_ug * t;
                while (ld_acquire(counter) < want) { __nanosleep(20); }
            }
            __syncthreads();
            // stage h[t-1] of the slice: hs[k][bl]; reads are coalesced along k
            for (int e = tid; e < C * BB; e += THREADS) {
                const int k = e % C;
                const int bb = e / C;
                const int gb = bg * BB + bb;
                float v = 0.f;
                if (gb < batch) v = __ldcg(out + ((size_t)gb * steps + (t - 1)) * C + k);
                hs[(size_t)k * BB + bb] = v;
            }
            __syncthreads();
#pragma unroll 8
            for (int k = 0; k < C; ++k) {
                const float hv = hs[(size_t)k * BB + bl];
                const float4 w = Ws[(size_t)k * HU + ul];
                gi = fmaf(w.x, hv, gi);
                gf = fmaf(w.y, hv, gf);
                gg = fmaf(w.z, hv, gg);
                go = fmaf(w.w, hv, go);
            }
        }
        const float i_ = ac::sigmoidf_(gi), f_ = ac::sigmoidf_(gf), o_ = ac::sigmoidf_(go);
        c_state = f_ * c_state + i_ * tanhf(gg);
        const float h = o_ * tanhf(c_state);
        if (live) {
            out[((size_t)b * steps + t) * C + u] = h;
            if (out_bf) {
                const __nv_bfloat16 hb = __float2bfloat16(h);
                out_bf[((size_t)b * steps + t) * C + u] = hb;
                if (out_lo) out_lo[((size_t)b * steps + t) * C + u] = __float2bfloat16(h - __bfloat162float(hb));
            }
        }
        __syncthreads();  // all h[t] of this CTA written (and hs reads finished)
        if (tid == 0) {
            __threadfence();
            atomicAdd(counter, 1);
        }
    }
    if (final_bf) {
        // bf16 edge: final = act(h + skip); our own (b,u) column only, h values are still in `out`
        if (live)
            for (int t = 0; t < steps; ++t) {
                float v = out[((size_t)b * steps + t) * C + u];
                if (skip_bf) v += __bfloat162float(skip_bf[(size_t)b * skip_bs + (size_t)t * C + u]);
                if (skip_lo) v += __bfloat162float(skip_lo[(size_t)b * skip_bs + (size_t)t * C + u]);
                if (final_act == AC_ACT_ELU) v = ac::elu1(v);
                const __nv_bfloat16 vb = __float2bfloat16(v);
                final_bf[(size_t)b * final_bs + (size_t)t * C + u] = vb;
                if (final_lo) final_lo[(size_t)b * final_bs + (size_t)t * C + u] = __float2bfloat16(v - __bfloat162float(vb));
            }
    }
    if (skip) {
        // residual connection of EncodecLSTM: out = lstm(x) + x.  Only our own (b,u) column: no sync needed
        // beyond program order, but h[t] is still read by peers for step t+1 -> add after the loop
        // and only after every peer finished the last step.
        if (tid == 0) {
            const int want = n_ug * steps;
            while (ld_acquire(counter) < want) { __nanosleep(20); }
        }
        __syncthreads();
        if (live)
            for (int t = 0; t < steps; ++t) {
                const size_t o = ((size_t)b * steps + t) * C + u;
                out[o] += skip[o];
            }
    }
}

}  // namespace

static int lstm_launch(const float* pre, const float* w_hh, const float* skip, float* out, int32_t batch, int32_t steps,
                       int32_t hidden, int32_t* sync_ws, void* out_bf16, const void* skip_bf16, void* final_bf16,
                       int64_t skip_bs, int64_t final_bs, int32_t final_act, void* stream, void* out_lo = nullptr,
                       const void* skip_lo = nullptr, void* final_lo = nullptr) {
    AC_REQUIRE(pre && w_hh && out && sync_ws, "ac_lstm_layer_f32: null pointer");
    AC_REQUIRE(batch > 0 && steps > 0, "ac_lstm_layer_f32: empty problem");
    AC_REQUIRE(hidden % HU == 0 && hidden <= 1024, "ac_lstm_layer_f32: hidden %d must be a multiple of %d", hidden, HU);
    const int n_ug = hidden / HU;
    const size_t smem = (size_t)hidden * HU * 16 + (size_t)hidden * BB * 4;
    cudaError_t e = cudaFuncSetAttribute(lstm_layer_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { ac::set_error("ac_lstm_layer_f32: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lstm_layer_f32_kernel, THREADS, smem);
    const int max_groups = (sms * per_sm) / n_ug;  // batch groups that can be co-resident
    AC_REQUIRE(max_groups >= 1, "ac_lstm_layer_f32: %d unit groups do not fit on %d SMs", n_ug, sms);
    AC_REQUIRE(max_groups <= 64, "ac_lstm_layer_f32: internal: sync_ws too small");
    cudaStream_t s = (cudaStream_t)stream;
    // the batch is processed in waves of max_groups*BB clips; each wave is one cooperative launch
    for (int b0 = 0; b0 < batch; b0 += max_groups * BB) {
        const int nb = (batch - b0 < max_groups * BB) ? batch - b0 : max_groups * BB;
        const int groups = (nb + BB - 1) / BB;
        e = cudaMemsetAsync(sync_ws, 0, 64 * sizeof(int), s);
        if (e != cudaSuccess) { ac::set_error("ac_lstm_layer_f32: memset: %s", cudaGetErrorString(e)); return (int)e; }
        const float* pre_w = pre + (size_t)b0 * steps * 4 * hidden;
        const float* skip_w = skip ? skip + (size_t)b0 * steps * hidden : nullptr;
        float* out_w = out + (size_t)b0 * steps * hidden;
        int nb_ = nb, steps_ = steps, hid_ = hidden;
        __nv_bfloat16* obf = out_bf16 ? (__nv_bfloat16*)out_bf16 + (size_t)b0 * steps * hidden : nullptr;
        const __nv_bfloat16* sbf = skip_bf16 ? (const __nv_bfloat16*)skip_bf16 + (size_t)b0 * skip_bs : nullptr;
        __nv_bfloat16* fbf = final_bf16 ? (__nv_bfloat16*)final_bf16 + (size_t)b0 * final_bs : nullptr;
        long long sbs = skip_bs, fbs = final_bs;
        int fact = final_act;
        __nv_bfloat16* olo = out_lo ? (__nv_bfloat16*)out_lo + (size_t)b0 * steps * hidden : nullptr;
        const __nv_bfloat16* slo = skip_lo ? (const __nv_bfloat16*)skip_lo + (size_t)b0 * skip_bs : nullptr;
        __nv_bfloat16* flo = final_lo ? (__nv_bfloat16*)final_lo + (size_t)b0 * final_bs : nullptr;
        void* args[] = {(void*)&pre_w, (void*)&w_hh, (void*)&skip_w, (void*)&out_w, &nb_, &steps_, &hid_, (void*)&sync_ws,
                        (void*)&obf, (void*)&sbf, (void*)&fbf, &sbs, &fbs, &fact, (void*)&olo, (void*)&slo, (void*)&flo};
        e = cudaLaunchCooperativeKernel((void*)lstm_layer_f32_kernel, dim3(n_ug, groups), dim3(THREADS), args, smem, s);
        ac::count_launch();
        if (e != cudaSuccess) { ac::set_error("ac_lstm_layer_f32: launch: %s", cudaGetErrorString(e)); return (int)e; }
    }
    return 0;
}

extern "C" int ac_lstm_layer_f32(const float* pre, const float* w_hh, const float* skip, float* out,
                                 int32_t batch, int32_t steps, int32_t hidden, int32_t* sync_ws, void* stream) {
    return lstm_launch(pre, w_hh, skip, out, batch, steps, hidden, sync_ws, nullptr, nullptr, nullptr, 0, 0, 0, stream);
}
