/*
 * audiocodecs_b200 -- C ABI of the B200 (sm_100a) tokenize/detokenize kernels.
 *
 * The reference (lucadellalib/audiocodecs) has no FFI: its boundary for this path is the Python
 * class `audiocodecs.codec.Codec` (R/audiocodecs/codec.py:33-214) whose `_sig_to_toks` /
 * `_toks_to_sig` hooks call third-party PyTorch modules.  This header is the C-ABI a maintainer
 * binds (ctypes) to replace the arithmetic underneath those hooks.  Every
 * entry point names the reference code it replaces (R/ = /root/reference, HF/ =
 * transformers/models, TA/ = torchaudio/functional/functional.py).
 *
 * Conventions: plain pointers to DEVICE memory and sizes; no allocation, no host sync, no global
 * state besides a thread-local last-error string; all work is enqueued on `stream` (a
 * cudaStream_t passed as void*).  Return value 0 = success, otherwise a cudaError_t / negative
 * argument-error code and `ac_last_error()` describes it.
 *
 * Activations are channels-last: x[b][l][c] (row = one time step), fp32 or bf16 as stated.
 */
#ifndef AUDIOCODECS_B200_H
#define AUDIOCODECS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AC_ABI_VERSION 1
#define AC_API __attribute__((visibility("default")))

/* padding modes of the input row index */
#define AC_PAD_ZERO 0      /* DAC (HF/dac:173-262), Mimi convs (HF/mimi configuration pad_mode="constant") */
#define AC_PAD_REFLECT 1   /* EnCodec (HF/encodec:139-162) */
#define AC_PAD_REPLICATE 2 /* Mimi downsample (HF/mimi:1422-1431) */

/* prologue activation applied to the conv input */
#define AC_ACT_NONE 0
#define AC_ACT_ELU 1   /* HF/encodec:268,300 ; HF/mimi:430 */
#define AC_ACT_SNAKE 2 /* HF/dac:85-99 : x + sin^2(alpha x)/(alpha+1e-9), alpha per input channel */
/* epilogue */
#define AC_EPI_NONE 0
#define AC_EPI_TANH 1 /* HF/dac:437 */
#define AC_EPI_GELU 2 /* HF/mimi:852-866 (exact erf GELU) */

/*
 * Generic "tap-GEMM" 1-D convolution, fp32 SIMT (the exact-parity path and the shapes the tensor
 * path does not take: Cin=1, Cout=1).
 *
 *   acc[b][m][n] = bias[n] + sum_{j<taps} sum_{c<cin} act(X[b][pos(m,j)][c]) * W[j][c][n]
 *   pos(m,j)     = m*stride + j*dilation - pad_left      (row index, padded per pad_mode)
 *   flat         = m*n_cols + n - out_shift ;  if 0 <= flat < out_valid:
 *   Y[b][flat]   = epi(acc) (+ R[b][flat])
 *
 * Conv1d        : n_cols = Cout, out_shift = 0, out_valid = Lout*Cout.
 * ConvTranspose : kernel 2s / stride s is the 2-tap conv over n_cols = s*Cout (column = phase*Cout+co)
 *                 whose flat output IS the [Lout][Cout] tensor; out_shift = padding*Cout.
 * Replaces: EncodecConv1d / EncodecConvTranspose1d / EncodecResnetBlock (HF/encodec:82-282),
 *           MimiConv1d / MimiConvTranspose1d (HF/mimi:214-409), DAC convs (HF/dac:173-262,405-472),
 *           and every 1x1 projection / nn.Linear on the path.
 */
typedef struct ac_conv_f32 {
    const float* x;       /* [B][x_rows][cin] , clip stride x_bstride, row stride x_rstride */
    const float* w;       /* [taps][cin][n_cols] */
    const float* bias;    /* [n_cols] or NULL */
    const float* alpha;   /* [cin] snake alpha or NULL */
    const float* res;     /* residual, same flat layout as y, or NULL (may alias y) */
    float* y;
    const int32_t* vlen;  /* optional [B]: input rows >= vlen[b] read as 0 (padding mask, HF/encodec:599-601) */
    int64_t x_bstride, y_bstride, res_bstride; /* elements */
    int32_t x_rstride;
    int32_t batch, x_rows, cin, m_rows, n_cols, taps, stride, dilation, pad_left;
    int32_t pad_mode, reflect_len, act, epi;
    int64_t out_shift, out_valid;
} ac_conv_f32;

AC_API int ac_conv1d_f32(const ac_conv_f32* p, void* stream);

/*
 * One LSTM layer over time (fp32), recurrent part only: gates[t] = pre[b][t][4C] + W_hh h[t-1];
 * gate order i,f,g,o; c,h start at 0.  out[b][t][C] = h[t] (+ skip[b][t][C] if skip != NULL).
 * `pre` already holds W_ih x + b_ih + b_hh (computed with ac_conv1d_f32, taps=1).
 * `sync_ws`: >= 64 zeroed int32 of device scratch for the inter-CTA step barrier.
 * Replaces EncodecLSTM (HF/encodec:236-249).
 */
AC_API int ac_lstm_layer_f32(const float* pre, const float* w_hh, const float* skip, float* out,
                      int32_t batch, int32_t steps, int32_t hidden, int32_t* sync_ws, void* stream);

/*
 * Residual VQ encode, all stages fused (fp32): for k < stages: idx = argmin_c ||r - E_k[c]||^2
 * (reference formula and tie rule: first index wins), r -= E_k[idx].
 *   metric 0: EnCodec  dist = -(|r|^2 - 2 r.E + |E|^2), argmax   (HF/encodec:364-369,424-438)
 *   metric 1: Mimi     cdist(r,E).argmin                          (HF/mimi:1197-1202,1262-1280)
 * x [rows][dim] fp32 (channels-last embeddings); codebooks [stages][codes][dim]; cb_norm [stages][codes]
 * = sum_j E^2; codes_out int64 written at codes_out[row*code_stride + (code_offset + k)].
 * residual_out (optional) [rows][dim] receives the final residual.
 */
AC_API int ac_rvq_encode_f32(const float* x, const float* codebooks, const float* cb_norm, int64_t* codes_out,
                      float* residual_out, int64_t rows, int32_t dim, int32_t codes, int32_t stages,
                      int32_t code_stride, int32_t code_offset, int32_t metric, void* stream);

/*
 * Residual VQ decode: out[row][dim] = sum_{k<stages} E_k[codes[row*code_stride + code_offset + k]]
 * accumulated in stage order from 0.0f.  codes are int64.  Out-of-range codes return -3.
 * Replaces EncodecResidualVectorQuantizer.decode (HF/encodec:440-447,381-383), Mimi (HF/mimi:1282-1293).
 */
AC_API int ac_rvq_decode_f32(const int64_t* codes, const float* codebooks, float* out, int64_t rows, int32_t dim,
                      int32_t n_codes, int32_t stages, int32_t code_stride, int32_t code_offset,
                      int32_t* err_flag, void* stream);

/*
 * Polyphase windowed-sinc resampler (torchaudio.functional.resample, TA:1405-1432):
 * y[b][f*n_phase + i] = sum_j taps[i][j] * xpad[b][f*orig + j], xpad = x left-padded by `width` zeros.
 * Replaces the resample calls at R/audiocodecs/codec.py:59-63,95-99.
 */
AC_API int ac_resample_f32(const float* x, const float* taps, float* y, int32_t batch, int64_t in_len,
                    int64_t out_len, int32_t orig, int32_t n_phase, int32_t n_taps, int32_t width,
                    void* stream);

AC_API int ac_abi_version(void);
AC_API const char* ac_last_error(void);
/* number of kernel launches issued through this library by the calling process (bench: gpu_launches) */
AC_API int64_t ac_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
