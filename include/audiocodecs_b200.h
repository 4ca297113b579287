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

#define AC_ABI_VERSION 3
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
    const void* x;        /* [B][x_rows][cin] fp32 (or bf16 when x_is_bf16), clip stride x_bstride, row stride x_rstride */
    const float* w;       /* [taps][cin][n_cols] */
    const float* bias;    /* [n_cols] or NULL */
    const float* alpha;   /* [cin] snake alpha or NULL */
    const float* res;     /* residual, same flat layout as y, or NULL (may alias y) */
    float* y;             /* fp32 output or NULL */
    const int32_t* vlen;  /* optional [B]: input rows >= vlen[b] read as 0 (padding mask, HF/encodec:599-601) */
    int64_t x_bstride, y_bstride, res_bstride; /* elements */
    int32_t x_rstride;
    int32_t batch, x_rows, cin, m_rows, n_cols, taps, stride, dilation, pad_left;
    int32_t pad_mode, reflect_len, act, epi;
    int64_t out_shift, out_valid;
    /* mixed-precision edges of the bf16 pipeline (first layer: fp32 waveform in, last layer: fp32 waveform out) */
    int32_t x_is_bf16;
    int32_t act2;         /* activation of the y_act_bf16 copy (AC_ACT_NONE / AC_ACT_ELU) */
    void* y_bf16;         /* optional bf16 copy of the output, same flat layout */
    void* y_act_bf16;     /* optional bf16 act2(output) */
    int64_t y_bf16_bstride, y_act_bstride;
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
 * Tensor-core LSTM layer recurrence (hidden = 512): clusters of 16 CTAs, each taking up to 16 clips as two independently
 * pipelined groups of <= 8; W_hh (16-bit) resident in TENSOR memory as the A operand, h exchanged through distributed shared
 * memory by bulk copies, one tcgen05.mma chain per group and step (see csrc/lstm_tc.cu).  Any batch size (waves of clusters).
 * pre [B][T][4*hidden] fp32 holds W_ih x + b_ih + b_hh (ac_conv_tc with y32).  Outputs: bf16 h planes (hi [+lo]) for
 * the next layer's input GEMM, and/or final = act(h + skip) planes in a haloed activation buffer.
 * Replaces EncodecLSTM (HF/encodec:236-249).
 */
typedef struct ac_lstm_tc_desc {
    const float* pre;
    const void* w_hh_bf16;     /* bf16 [4*hidden][hidden], gate order i,f,g,o */
    void* out_hi; void* out_lo;             /* bf16 [B][T][hidden] or NULL */
    const void* skip_hi; const void* skip_lo; /* bf16, clip stride skip_bstride */
    void* final_hi; void* final_lo;         /* bf16, clip stride final_bstride */
    int64_t skip_bstride, final_bstride;
    int32_t final_act, batch, steps, hidden;
    void* dbg;                 /* optional int64 [steps][8] clock samples (profiling aid), else NULL */
    int32_t operand_fp16;      /* 1: w_hh_bf16 holds IEEE fp16 and h[t-1] is fed back as fp16 (11-bit mantissas: |W_hh| and |h| < 1
                                  sit well inside the fp16 range) -- one product at 2^-12 instead of 2^-9 operand rounding */
    int32_t out_fp16, skip_fp16; /* hi planes of out / final, and of skip, hold fp16 instead of bf16 (lo planes: bf16) */
} ac_lstm_tc_desc;
AC_API int ac_lstm_tc(const ac_lstm_tc_desc* d, void* stream);
/* diagnostic: co-resident clusters of `cluster_size` CTAs of the recurrence kernel at `smem_bytes` of dynamic shared memory */
AC_API int ac_lstm_tc_max_clusters(int32_t cluster_size, int32_t smem_bytes);

/*
 * Residual VQ encode on tcgen05 tensor cores (dim 128 or 256, n_codes a multiple of 128 in [256, 2048]), all stages fused,
 * the fp32 residual of a 128-frame tile resident in tensor memory.  Distance GEMM in error-compensated bf16
 * (r_hi.E_hi + r_hi.E_lo + r_lo.E_hi, fp32 accumulate in TMEM), per-frame running top-2, exact fp32 re-score of the two
 * candidates with the reference formula and tie rule (metric 0: EnCodec, metric 1: Mimi cdist), in-place subtract.
 * Same outputs as ac_rvq_encode_f32.  Stages stage0 .. stage0+stages-1 of the stacked codebooks are used:
 * cb_split_bf16 [2][stages_total][n_codes][dim] = bf16(E), bf16(E - bf16(E)); codebooks [stages_total][n_codes][dim] fp32;
 * cb_norm [stages_total][n_codes].  Codes are written at codes_out[row*code_stride + code_offset + k], k < stages.
 * Replaces HF/encodec/modeling_encodec.py:364-369,424-438 and HF/mimi/modeling_mimi.py:1197-1202,1262-1280.
 */
AC_API int ac_rvq_encode_tc(const float* x, const void* cb_split_bf16, const float* codebooks, const float* cb_norm,
                            int64_t* codes_out, float* residual_out, int64_t rows, int32_t dim, int32_t n_codes,
                            int32_t stages, int32_t stage0, int32_t stages_total, int32_t code_stride, int32_t code_offset,
                            int32_t metric, void* stream);

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
/* same sum, also/only written as bf16 into a haloed activation buffer: row r of clip b at
 * out_bf16[b*bstride + r*dim] (rows_per_clip rows per clip). */
AC_API int ac_rvq_decode_bf16(const int64_t* codes, const float* codebooks, void* out_bf16, void* out_lo, int64_t rows,
                      int32_t rows_per_clip, int64_t bstride, int32_t dim, int32_t n_codes, int32_t stages,
                      int32_t code_stride, int32_t code_offset, int32_t* err_flag, int32_t out_f16, void* stream);

/*
 * Polyphase windowed-sinc resampler (torchaudio.functional.resample, TA:1405-1432):
 * y[b][f*n_phase + i] = sum_j taps[i][j] * xpad[b][f*orig + j], xpad = x left-padded by `width` zeros.
 * Replaces the resample calls at R/audiocodecs/codec.py:59-63,95-99.
 */
AC_API int ac_resample_f32(const float* x, const float* taps, float* y, int32_t batch, int64_t in_len,
                    int64_t out_len, int32_t orig, int32_t n_phase, int32_t n_taps, int32_t width,
                    void* stream);


/*
 * bf16 tensor-core tap-GEMM convolution (tcgen05.mma + TMEM accumulators, operands staged by TMA):
 *
 *   acc[b][m][n] = bias[n] + sum_{s<n_src} sum_{j<taps_s} sum_{k<c0_s*phases_s}
 *                      A_s[b][m + j*dilation_s + shift_s][k] * W[n][col(s,j,k)]
 *   flat = m*n_total + n - out_shift ; if 0 <= flat < out_valid:
 *       v = epi(acc) (+ res[b][flat]);  y[b][flat] = bf16(v);  y32[b][flat] = v;  y_act[b][flat] = bf16(act(v))
 *
 * A_s is a VIEW of a channels-last bf16 activation buffer: view row = `phases` consecutive time steps
 * of `c0` channels, so a stride-s / kernel-2s convolution is a 2-tap GEMM over rows of s*c0 values and
 * a transposed convolution is a 2-tap GEMM whose n axis is (phase, cout).  Rows outside [0, rows) read
 * as zero (TMA out-of-bounds fill) -- zero padding costs nothing; reflect padding is materialised by the
 * producer in the buffer's halo rows (ac_pad_halo_bf16).  Two sources let one GEMM take its taps from two
 * tensors (EnCodec ResBlock: shortcut(x) + conv_k1(ELU(h)) is one launch).  `y_act` receives the
 * activation the CONSUMER layer applies (ELU / Snake with the consumer's alpha), so activations are
 * applied once, at production time.  W is bf16 [n_total][k_total], columns ordered source, tap, k.
 * Replaces the same reference code as ac_conv1d_f32 (HF/encodec:82-282, HF/mimi:214-451, HF/dac:173-262,405-472).
 */
/* operand / output formats of the tensor-path launches (`fmt` of ac_conv_tc_desc / ac_resunit_tc_desc).  The hi plane of an
 * activation or weight is bf16 or IEEE fp16 (11 significant bits: one fp16 product is more accurate than the bf16 pair of
 * products, at half the tensor work; values beyond +-65504 saturate); lo planes are always bf16(v - float(hi)). */
#define AC_FMT_A_F16 1     /* hi planes of every A source (and of the hidden tile) and the W_hi / W_lo planes are fp16 */
#define AC_FMT_W_HIB 2     /* with A_F16: an extra bf16(W) plane follows W_hi [W_lo]; the A_lo (bf16) products multiply it */
#define AC_FMT_Y_F16 4     /* y hi plane is written as fp16 */
#define AC_FMT_YACT_F16 8  /* y_act hi plane is written as fp16 */
#define AC_FMT_RES_F16 16  /* res hi plane holds fp16 */
#define AC_FMT_W2_HIB 32   /* ac_resunit_tc: the same as W_HIB for W2 (W_HIB then refers to W1) */

typedef struct ac_tc_src {
    const void* base;            /* bf16: view row 0, phase 0, channel 0 of clip 0 */
    int32_t c0, phases, rows;    /* channels per phase, phases per view row, view rows per clip */
    int64_t phase_stride, row_stride, batch_stride; /* elements */
    int32_t taps, dilation, shift; /* tap j reads view row m + j*dilation + shift */
    int32_t lo_of;               /* -1, or index of the source whose lo plane (x - bf16(x)) this is: it re-uses that
                                    source's weight columns (split-bf16 activations, fp32-accurate products) */
} ac_tc_src;

typedef struct ac_conv_tc_desc {
    ac_tc_src src[4];
    int32_t n_src;
    const void* w;               /* bf16 [n_total][k_total]; with w_split: [2][n_total][k_total] = W_hi then W_lo = bf16(W - W_hi) */
    int32_t w_split;             /* 1: every product also accumulates A*W_lo (error-compensated weights, no HBM cost) */
    int32_t k_total, n_total, bk; /* bk: contraction block 16/32/64, must divide every source's c0 */
    const float* bias;           /* [n_total] or NULL */
    const float* alpha;          /* snake alpha of the consumer, [act_mod] */
    const void* res;             /* bf16 residual in the output's flat layout, or NULL */
    void* y;                     /* bf16 out or NULL */
    void* y_act;                 /* bf16 act(out) or NULL */
    void* y_lo;                  /* optional lo planes of y / y_act (same strides): bf16(v - float(bf16(v))) */
    void* y_act_lo;
    float* y32;                  /* fp32 out or NULL */
    int32_t act, epi, act_mod;   /* act: AC_ACT_* for y_act; epi: AC_EPI_*; channel of column n = n % act_mod */
    int64_t y_bstride, y_act_bstride, y32_bstride, res_bstride; /* elements per clip */
    int64_t out_shift, out_valid;
    int32_t batch, m_rows;
    int32_t n_tile_hint, grid_hint; /* 0 = automatic */
    const void* res_lo;          /* optional lo plane of `res` */
    const float* res32;          /* optional fp32 residual in the output's flat layout (clip stride res_bstride) */
    int32_t g_hint;              /* 128-row sub-tiles per tile (1, 2 or 4) sharing one A block and one W block; 0 = automatic */
    int32_t fmt;                 /* AC_FMT_* flags; 0 = everything bf16 */
    int32_t flush_adds;          /* > 0: chunked accumulation -- at most about this many tcgen05.mma per partial sum, partials added
                                    in fp32 round-to-nearest by the epilogue warps (the tensor-core accumulator truncates every add:
                                    ~2e-8 of shrink per MMA, 2e-5 at K = 4096 with three products).  0 = one accumulator per tile */
} ac_conv_tc_desc;

AC_API int ac_conv_tc(const ac_conv_tc_desc* d, void* stream);

/*
 * Fused residual unit on tcgen05: two chained tap-GEMMs per tile, the hidden activation stays on chip (TMEM -> registers
 * -> shared memory in the UMMA operand layout -> second GEMM).
 *
 *   h[b][m][:]  = act1(bias1 + sum_{j<taps} A[b][m + j*dilation + shift][:] . W1[:, j*cin : (j+1)*cin]^T)      (ch channels)
 *   v[b][m][:]  = bias2 + h[b][m][:] . W2[:, 0:ch]^T  (+ X[b][m][:] . W2[:, ch:ch+cin]^T)  (+ res[b][m][:])      (cout channels)
 *   y = bf16(v) (+ lo plane);   y_act = bf16(act2(v)) (+ lo plane)          act = AC_ACT_{NONE,ELU,SNAKE}
 *
 * A is the ACTIVATED input -- or, in raw mode (act0), the raw input activated on chip -- (channels-last bf16 view, rows outside [0, a_rows) read as zero; `a` points at view row 0 which may
 * be a halo row of the buffer), X the raw input of a convolutional shortcut (EnCodec), res an identity skip (Mimi, DAC).
 * W1 is bf16 [ch][taps*cin] and W2 bf16 [cout][ch (+cin)], each optionally the stacked pair (W_hi, W_lo) (w*_split).
 * h_split: the hidden tile carries a lo plane (adds the h_lo * W2_hi product).  ch, cout multiples of 16, <= 256,
 * ch + cout <= 512.  Replaces EncodecResnetBlock (HF/encodec/modeling_encodec.py:252-282), MimiResnetBlock
 * (HF/mimi/modeling_mimi.py:412-451) and DacResidualUnit (HF/dac/modeling_dac.py:173-207) -- five to six ATen ops each.
 */
typedef struct ac_resunit_tc_desc {
    const void* a;  const void* a_lo;      /* activated input: view row 0 of clip 0 (hi plane, optional lo plane) */
    int64_t a_row_stride, a_bstride;       /* elements */
    int32_t a_rows, cin, taps, dilation, shift;
    const void* x;  const void* x_lo;      /* raw input [batch][m_rows][cin] of the conv shortcut, or NULL */
    int64_t x_bstride;
    const void* w1; const void* w2;
    int32_t w1_split, w2_split, ch, cout, h_split;
    const float* bias1; const float* alpha1; /* [ch] */
    const float* bias2; const float* alpha2; /* [cout] */
    int32_t act1, act2;
    const void* res; const void* res_lo;   /* identity skip [batch][m_rows][cout] or NULL */
    int64_t res_bstride;
    void* y; void* y_lo; void* y_act; void* y_act_lo;
    int64_t y_bstride, y_act_bstride;
    int32_t batch, m_rows, bk, g_hint, grid_hint;
    int32_t dbl_hint;                      /* -1 automatic; 0 / 1: single / double-buffered hidden tile and first accumulator; 2: double-
                                              buffered with the second accumulator doubled too and the epilogue warps in two groups that
                                              take alternate tiles (ping-pong; needs G*(2*ch + 2*cout) <= 512 tensor-memory columns) */
    /* raw mode (act0 != AC_ACT_NONE): `a` is the RAW input and the kernel applies the unit's input activation act0 (ELU, or
       Snake with alpha0[cin]) on chip, block by block, before GEMM1 -- the producer layer then writes one tensor instead of a
       raw and an activated copy.  e_split: the activated operand carries a lo plane.  x_from_a: the conv shortcut reads the raw
       rows of the same staged blocks (x must be NULL; see x_row_off). */
    int32_t act0, e_split, x_from_a;
    const float* alpha0;
    int32_t x_row_off;                     /* x_from_a: view row (m + shift + x_row_off) holds raw x[m]; 0 <= x_row_off <= (taps-1)*dilation */
    int32_t fmt;                           /* AC_FMT_* flags; 0 = everything bf16 (raw mode, act0 != NONE, is bf16 only) */
    int32_t io_stage;                      /* epilogue-2 I/O through shared-memory tiles + TMA loads / stores: 0 = when it fits shared
                                              memory, -1 = never (direct global loads / stores), 1 = required (else an error) */
} ac_resunit_tc_desc;

AC_API int ac_resunit_tc(const ac_resunit_tc_desc* d, void* stream);

/*
 * Fill the halo rows of a channels-last bf16 activation buffer [batch][halo_l + rows + halo_r][ch] from its
 * valid rows: mode AC_PAD_REFLECT (EnCodec, incl. the zero-extension rule for tiny inputs via reflect_len),
 * AC_PAD_REPLICATE (Mimi downsample) or AC_PAD_ZERO.  `data` points at valid row 0 of clip 0.
 */
AC_API int ac_pad_halo_bf16(void* data, int32_t batch, int32_t rows, int32_t ch, int64_t batch_stride,
                            int32_t halo_l, int32_t halo_r, int32_t mode, int32_t reflect_len, void* stream);
/* the same for the (hi, lo) planes of one tensor in ONE launch (data_lo may be NULL): same geometry for both planes */
AC_API int ac_pad_halo2_bf16(void* data, void* data_lo, int32_t batch, int32_t rows, int32_t ch, int64_t batch_stride,
                             int32_t halo_l, int32_t halo_r, int32_t mode, int32_t reflect_len, void* stream);

/*
 * out = act(a + b) elementwise on split-bf16 activations (hi plane + optional lo plane each), `per_clip` elements per
 * clip.  The skip-add + ELU that follows EnCodec's LSTM (HF/encodec:247 `hidden_states + residual`, then the next
 * layer's ELU), kept out of the latency-bound recurrence kernel.
 */
AC_API int ac_add_act_bf16(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, void* out_hi, void* out_lo,
                           int32_t batch, int64_t per_clip, int64_t a_bstride, int64_t b_bstride, int64_t out_bstride,
                           int32_t act, int32_t fmt /* bit 0: a_hi is fp16, bit 1: b_hi, bit 2: out_hi (else bf16) */, void* stream);

/* fp32 [batch][per_clip] -> split-bf16 planes hi = bf16(x), lo = bf16(x - hi) (lo optional): the quantised latents
 * (RVQ decode output, fp32) entering the tensor-core decoder. */
AC_API int ac_f32_to_split_bf16(const float* x, void* out_hi, void* out_lo, int32_t batch, int64_t per_clip, int64_t x_bstride,
                                int64_t out_bstride, int32_t out_f16 /* hi plane as fp16 instead of bf16 */, void* stream);

/*
 * Edge layers of the bf16 pipeline (HBM-bound, SIMT):
 * first: y[b][t][c] = bias[c] + sum_j w[j][c] * x[b][pad(t + j - pad_left)]   (Cin = 1; fp32 waveform in, bf16 out:
 *        raw copy `y` and/or `y_act` = act(y), act = ELU or Snake(alpha));  C in {32, 64, 96}, K <= 8.
 *        y_lo / y_act_lo: optional lo planes bf16(v - float(bf16(v))) of the same layout (the "exact" precision mode
 *        carries every activation as a (hi, lo) pair).
 *        vlen: optional per-clip valid length (padding mask, R/audiocodecs/encodec.py:84-89).
 * last : y[b][t] = epi(bias + sum_j sum_c w[j][c] * x[b][pad(t + j - pad_left)][c])   (Cout = 1; bf16 in, fp32 out).
 * Replace the first/last conv of EncodecEncoder/Decoder (HF/encodec:289,341), MimiEncoder/Decoder (HF/mimi:461,1169),
 * DacEncoder/Decoder (HF/dac:449,434-437).
 */
AC_API int ac_conv_first_bf16(const float* x, const float* w, const float* bias, const float* alpha, const int32_t* vlen,
                              void* y, void* y_act, void* y_lo, void* y_act_lo, int64_t y_bstride, int64_t y_act_bstride,
                              int32_t batch, int32_t T, int32_t C, int32_t K, int32_t pad_left, int32_t pad_mode,
                              int32_t reflect_len, int32_t act, int32_t out_f16, void* stream);
AC_API int ac_conv_last_bf16(const void* x, const float* w, const float* bias, float* y, int64_t x_bstride, int32_t batch,
                             int32_t T, int32_t C, int32_t K, int32_t pad_left, int32_t pad_mode, int32_t reflect_len,
                             int32_t epi, void* stream);

/*
 * Mimi transformer pieces that are not GEMMs (fp32): LayerNorm over the channel axis; RoPE (theta from inv_freq,
 * rotate-half form, positions 0..T-1) + causal sliding-window attention on a fused qkv tensor [B][T][3*H*D]
 * (q | k | v); depthwise ConvTranspose1d(k=4, s=2, groups=C) trimmed right by 2 (the `upsample` layer).
 * Replace HF/mimi/modeling_mimi.py:926-993 (layer), :645-736 (attention), :515-577 (rotary), :1433-1441 (upsample).
 */
AC_API int ac_layernorm_f32(const float* x, const float* w, const float* b, float* y, int64_t rows, int32_t C, float eps,
                            void* stream);
AC_API int ac_attention_f32(const float* qkv, const float* inv_freq, float* out, int32_t batch, int32_t T, int32_t heads,
                            int32_t head_dim, int32_t window, float scaling, void* stream);
AC_API int ac_upsample_dw_f32(const float* x, const float* w, float* y, int32_t batch, int32_t L, int32_t C, void* stream);

/*
 * The same attention (HF/mimi/modeling_mimi.py:645-736, mask :1096-1102) on tcgen05 tensor cores for the bf16 tensor path:
 * Q K^T and P V as split-bf16 products (hi*hi + lo*hi + hi*lo, fp32 accumulate in tensor memory), online softmax in fp32.
 * `rope` is the [T][head_dim] table cos | sin of positions 0..T-1 (ac_rope_table_f32 builds it from inv_freq,
 * HF/mimi:515-558).  Output: fp32 [B][T][H*D] (out32) and/or split-bf16 planes out_hi [+ out_lo] with batch stride
 * out_bstride (elements) -- the activation layout the output-projection GEMM reads.  No window limit.
 */
/* LayerNorm (as ac_layernorm_f32) written straight into split-bf16 planes [B] x out_bstride + [rows_per_clip][C]. */
AC_API int ac_layernorm_split_bf16(const float* x, const float* w, const float* b, void* out_hi, void* out_lo, int32_t batch,
                                   int32_t rows_per_clip, int32_t C, int64_t out_bstride, float eps, int32_t out_f16, void* stream);
AC_API int ac_rope_table_f32(const float* inv_freq, float* table, int32_t T, int32_t half, void* stream);
AC_API int ac_attention_tc(const float* qkv, const float* rope, float* out32, void* out_hi, void* out_lo, int64_t out_bstride,
                           int32_t batch, int32_t T, int32_t heads, int32_t head_dim, int32_t window, float scaling, int32_t out_f16,
                           void* stream);

/*
 * DAC residual VQ (hidden 1024, codebook dim 8), all stages fused, fp32.
 * encode: z [rows][1024]; w_in [S][8][1024], b_in [S][8], codebooks [S][n_codes][8], w_out [S][1024][8], b_out [S][1024];
 *         codes int64 at codes[row*code_stride + k]; zq_out (optional) [rows][1024] = sum_k out_proj_k(.) (the
 *         quantised representation `z` of dac.DAC.encode).
 * decode: from_codes: out[row] = sum_k out_proj_k(codebook_k[code_k]).
 * Replace descript-audio-codec 1.0.0 dac/nn/quantize.py ResidualVectorQuantize.forward / from_codes
 * (twin: HF/dac/modeling_dac.py:122-170,281-369), called at R/audiocodecs/dac.py:96-98,126-128.
 */
AC_API int ac_dac_rvq_encode_f32(const float* z, const float* w_in, const float* b_in, const float* codebooks,
                                 const float* w_out, const float* b_out, int64_t* codes, float* zq_out, int64_t rows,
                                 int32_t hidden, int32_t cb_dim, int32_t n_codes, int32_t stages, int32_t code_stride,
                                 void* stream);
/*
 * Encode from projected latents (bf16 tensor path): proj [rows][ld_proj] = in_proj of ALL stages applied to z (bias included,
 * column k*8+d); cconst [S][8] = -sum_{j<k} W_in_k b_out_j; cross [S][S][8][8] = W_in_k W_out_j (j<k used);
 * cb_normed [S][n_codes][8] = L2-normalised codebooks, cb_norm2 [S][n_codes] their squared norms.  Same score formula,
 * tie-break and straight-through arithmetic as ac_dac_rvq_encode_f32 (dac/nn/quantize.py @1.0.0, HF/dac:122-170), without
 * ever forming the 1024-wide residual.
 */
AC_API int ac_dac_rvq_encode_proj_f32(const float* proj, int32_t ld_proj, const float* cconst, const float* cross,
                                      const float* cb_normed, const float* cb_norm2, const float* codebooks, int64_t* codes,
                                      int64_t rows, int32_t cb_dim, int32_t n_codes, int32_t stages, int32_t stages_total,
                                      int32_t code_stride, void* stream);
AC_API int ac_dac_rvq_decode_f32(const int64_t* codes, const float* codebooks, const float* w_out, const float* b_out,
                                 float* out, int64_t rows, int32_t hidden, int32_t cb_dim, int32_t n_codes, int32_t stages,
                                 int32_t code_stride, int32_t* err_flag, void* stream);

/*
 * Token consumers (the step right after the tokenizer in the reference's downstream recipes):
 * ac_token_histogram: counts[k][v] += #{rows r : toks[r][k] == v} -- the accumulation of CodebookUtil.append
 *   (R/downstream/metrics/codebook_util.py:39-49); counts is int64 [num_codebooks][vocab_size], toks int64 [rows][num_codebooks];
 *   out-of-range tokens set *err_flag and are not counted.
 * ac_multihead_embedding: out[r][k][:] = weight[toks[r][k] + offsets[k]][:] (fp32, dim % 4 == 0) -- MultiHeadEmbedding.forward
 *   (R/downstream/models/multihead.py:52-69); padding_row >= 0: a token equal to vocab_size reads that row (the shared padding
 *   embedding), else -1.
 */
AC_API int ac_token_histogram(const int64_t* toks, int64_t rows, int32_t num_codebooks, int32_t vocab_size, int64_t* counts,
                              int32_t* err_flag, void* stream);
AC_API int ac_multihead_embedding(const int64_t* toks, const float* weight, const int64_t* offsets, float* out, int64_t rows,
                                  int32_t num_codebooks, int32_t dim, int64_t vocab_size, int64_t padding_row,
                                  int64_t num_embeddings, int32_t* err_flag, void* stream);

AC_API int ac_abi_version(void);
AC_API const char* ac_last_error(void);
/* number of kernel launches issued through this library by the calling process (bench: gpu_launches) */
AC_API int64_t ac_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
