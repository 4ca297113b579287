"""Host-side launch helpers: torch tensors (device memory + stream only) -> C-ABI calls.

Activations are channels-last `[B, L, C]` fp32 tensors.  Nothing here computes on the host or
falls back to torch ops; a CPU tensor raises.
"""
import ctypes
import math

import torch

from . import _lib
from ._lib import AcConvF32

PAD_ZERO, PAD_REFLECT, PAD_REPLICATE = 0, 1, 2
ACT_NONE, ACT_ELU, ACT_SNAKE = 0, 1, 2
EPI_NONE, EPI_TANH, EPI_GELU = 0, 1, 2


class Profiler:
    """Per-launch device timing with CUDA events on the launching stream (bench.py roofline leg)."""

    def __init__(self):
        self.records = []

    def begin(self):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def end(self, name, start, flops=0.0, bytes_=0.0, label=None):
        """name = kernel (as ncu lists it); label = which layer family launched it; flops/bytes = ALGORITHMIC work of
        the reference op (2 FLOP per MAC; activation bytes in+out once, 2 B each on the bf16 path -- SURVEY 8d)."""
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        self.records.append((name, start, e, flops, bytes_, label or name))

    def summary(self, by_label=False):
        torch.cuda.synchronize()
        out = {}
        for name, s, e, fl, by, label in self.records:
            d = out.setdefault(label if by_label else name, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "n": 0})
            d["ms"] += s.elapsed_time(e)
            d["flops"] += fl
            d["bytes"] += by
            d["n"] += 1
        return out


_PROFILER = None


def set_profiler(p):
    global _PROFILER
    _PROFILER = p


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("audiocodecs_b200 runs on CUDA (sm_100a) tensors only; there is no CPU fallback. "
                               "Move the codec and its inputs with .to('cuda').")


class ConvSpec:
    """One packed convolution / transposed convolution / linear layer.

    w: [taps, cin, n_cols] fp32 (see packing.py), bias [n_cols] or None.
    geometry: 'causal' (EnCodec/Mimi: left pad K_eff-stride, right pad to ceil(L/stride)),
              'same'   (DAC: symmetric zero padding `padding`, floor length rule),
              'tr'     (transposed, kernel 2*stride; `tr_pad` = torch padding, causal trim when 0)
    """

    def __init__(self, w, bias, *, cout, kernel=1, stride=1, dilation=1, geometry="causal", pad_mode=PAD_ZERO,
                 padding=0, act=ACT_NONE, alpha=None, epi=EPI_NONE, tr_stride=0, tr_pad=0):
        self.w, self.bias, self.alpha = w, bias, alpha
        self.cout, self.kernel, self.stride, self.dilation = cout, kernel, stride, dilation
        self.geometry, self.pad_mode, self.padding = geometry, pad_mode, padding
        self.act, self.epi, self.tr_stride, self.tr_pad = act, epi, tr_stride, tr_pad
        self.taps, self.cin, self.n_cols = w.shape

    def apply(self, fn):
        """module.to()/cuda() support: Codec._apply maps `fn` over the packed tensors."""
        self.w = fn(self.w)
        self.bias = fn(self.bias) if self.bias is not None else None
        self.alpha = fn(self.alpha) if self.alpha is not None else None

    def out_len(self, L):
        if self.geometry == "causal":
            return -(-L // self.stride)
        if self.geometry == "same":
            k_eff = (self.kernel - 1) * self.dilation + 1
            return (L + 2 * self.padding - k_eff) // self.stride + 1
        s = self.tr_stride
        return (L - 1) * s - 2 * self.tr_pad + 2 * s if self.tr_pad else L * s


def conv(spec: ConvSpec, x: torch.Tensor, res: torch.Tensor = None, out: torch.Tensor = None,
         vlen: torch.Tensor = None, y_bf=None, y_act_bf=None, act2=ACT_NONE, want_f32=True) -> torch.Tensor:
    """x [B, L, Cin] (fp32, or the bf16 valid-row view of a tc.Act) -> y [B, Lout, Cout] (+res). `res` may alias
    `out`.  y_bf / y_act_bf: optional tc.Act outputs (bf16 copy, bf16 act2(copy)) for the bf16 pipeline's edges."""
    _need_cuda(x, spec.w, res, out)
    assert x.dtype in (torch.float32, torch.bfloat16) and x.dim() == 3 and x.stride(2) == 1 and x.shape[2] == spec.cin, (x.shape, spec.cin)
    B, L, _ = x.shape
    Lout = spec.out_len(L)
    if Lout <= 0:
        raise ValueError(f"input too short for this layer (L={L})")
    if out is None and want_f32:
        out = torch.empty((B, Lout, spec.cout), device=x.device, dtype=torch.float32)
    assert out is None or (out.is_contiguous() and tuple(out.shape) == (B, Lout, spec.cout))
    p = AcConvF32()
    p.x, p.w, p.bias, p.alpha = x.data_ptr(), spec.w.data_ptr(), _ptr(spec.bias), _ptr(spec.alpha)
    p.res, p.y, p.vlen = _ptr(res), _ptr(out), _ptr(vlen)
    p.x_is_bf16 = int(x.dtype == torch.bfloat16)
    p.act2 = act2
    for o in (y_bf, y_act_bf):
        assert o is None or (o.L == Lout and o.C == spec.cout and o.B == B)
    if y_bf is not None:
        p.y_bf16, p.y_bf16_bstride = y_bf.row_ptr(0), y_bf.bstride
    if y_act_bf is not None:
        p.y_act_bf16, p.y_act_bstride = y_act_bf.row_ptr(0), y_act_bf.bstride
    if res is not None:
        assert res.shape == out.shape and res.is_contiguous()
    p.x_bstride, p.y_bstride = x.stride(0), (out.stride(0) if out is not None else 0)
    p.res_bstride = res.stride(0) if res is not None else 0
    p.x_rstride = x.stride(1)
    p.batch, p.x_rows, p.cin, p.n_cols, p.taps = B, L, spec.cin, spec.n_cols, spec.taps
    p.act, p.epi, p.pad_mode = spec.act, spec.epi, spec.pad_mode
    p.reflect_len = L
    p.out_valid = Lout * spec.cout
    if spec.geometry == "tr":
        p.stride, p.dilation, p.pad_left = 1, 1, 1
        p.m_rows = L + 1 if spec.tr_pad else L
        p.out_shift = spec.tr_pad * spec.cout
        p.pad_mode = PAD_ZERO
    else:
        k_eff = (spec.kernel - 1) * spec.dilation + 1
        p.stride, p.dilation, p.m_rows, p.out_shift = spec.stride, spec.dilation, Lout, 0
        if spec.geometry == "causal":
            p.pad_left = k_eff - spec.stride
            extra = (Lout - 1) * spec.stride + k_eff - p.pad_left - L
            m = max(p.pad_left, extra)
            if spec.pad_mode == PAD_REFLECT and L <= m:
                p.reflect_len = m + 1  # the reference zero-extends tiny inputs first (HF/encodec:148-155)
        else:
            p.pad_left = spec.padding
    t0 = _PROFILER.begin() if _PROFILER else None
    _lib.check(_lib.lib().ac_conv1d_f32(ctypes.byref(p), _stream()), "ac_conv1d_f32")
    if _PROFILER:
        # algorithmic work: 2 FLOP per MAC of the reference's dense op; fp32 activation bytes in+out
        _PROFILER.end("conv1d_f32", t0, 2.0 * B * p.m_rows * spec.n_cols * spec.taps * spec.cin,
                      float(x.element_size() * x.numel() + 4 * spec.w.numel() + B * Lout * spec.cout *
                            (4 * (out is not None) + 2 * (y_bf is not None) + 2 * (y_act_bf is not None))))
    return out


def lstm_layer(pre, w_hh, skip, sync_ws):
    """pre [B,T,4C] -> h [B,T,C] (+skip)."""
    _need_cuda(pre, w_hh, skip)
    B, T, C4 = pre.shape
    C = C4 // 4
    out = torch.empty((B, T, C), device=pre.device, dtype=torch.float32)
    assert pre.is_contiguous() and w_hh.is_contiguous() and (skip is None or skip.is_contiguous())
    t0 = _PROFILER.begin() if _PROFILER else None
    _lib.check(_lib.lib().ac_lstm_layer_f32(_ptr(pre), _ptr(w_hh), _ptr(skip), _ptr(out), B, T, C, _ptr(sync_ws), _stream()),
               "ac_lstm_layer_f32")
    if _PROFILER:
        _PROFILER.end("lstm_layer_f32", t0, 2.0 * B * T * 4 * C * C, 4.0 * (pre.numel() + out.numel()))
    return out


def conv_first_bf16(spec, sig, y=None, y_act=None, act=ACT_NONE, vlen=None, pad_left=None, alpha=None):
    """Cin=1 first layer of the bf16 pipeline: sig [B,T] fp32 -> tc.Act outputs (raw / activated)."""
    _need_cuda(sig, spec.w)
    B, T = sig.shape
    sig = sig.contiguous()
    K, C = spec.taps, spec.cout
    if pad_left is None:
        pad_left = (K - 1) if spec.geometry == "causal" else spec.padding
    reflect_len = pad_left + 1 if (spec.pad_mode == PAD_REFLECT and T <= pad_left) else T
    t0 = _PROFILER.begin() if _PROFILER else None
    _lib.check(_lib.lib().ac_conv_first_bf16(
        _ptr(sig), _ptr(spec.w), _ptr(spec.bias), _ptr(alpha), _ptr(vlen),
        ctypes.c_void_p(y.row_ptr(0)) if y is not None else None, ctypes.c_void_p(y_act.row_ptr(0)) if y_act is not None else None,
        ctypes.c_void_p(y.lo_ptr(0)) if (y is not None and y.lo is not None) else None,
        ctypes.c_void_p(y_act.lo_ptr(0)) if (y_act is not None and y_act.lo is not None) else None,
        y.bstride if y is not None else 0, y_act.bstride if y_act is not None else 0, B, T, C, K, pad_left, spec.pad_mode,
        reflect_len, act, int(_same_f16(y, y_act)), _stream()), "ac_conv_first_bf16")
    if _PROFILER:
        _PROFILER.end("conv_first", t0, 2.0 * B * T * K * C, 4.0 * B * T + 2.0 * B * T * C * ((y is not None) + (y_act is not None)))


def _same_f16(*acts):
    """hi-plane format shared by the given tc.Act outputs of one launch"""
    fm = {a.f16 for a in acts if a is not None}
    assert len(fm) == 1, "the outputs of one launch share the hi-plane format"
    return fm.pop()


def conv_last_bf16(spec, x_act, epi=EPI_NONE, pad_left=None):
    """Cout=1 last layer of the bf16 pipeline: tc.Act [B,T,C] (already activated) -> sig [B,T] fp32."""
    _need_cuda(spec.w)
    B, T, C, K = x_act.B, x_act.L, x_act.C, spec.taps
    if pad_left is None:
        pad_left = (K - 1) if spec.geometry == "causal" else spec.padding
    reflect_len = pad_left + 1 if (spec.pad_mode == PAD_REFLECT and T <= pad_left) else T
    out = torch.empty((B, T), device=x_act.buf.device, dtype=torch.float32)
    t0 = _PROFILER.begin() if _PROFILER else None
    _lib.check(_lib.lib().ac_conv_last_bf16(ctypes.c_void_p(x_act.row_ptr(0)), _ptr(spec.w), _ptr(spec.bias), _ptr(out), x_act.bstride,
                                            B, T, C, K, pad_left, spec.pad_mode, reflect_len, epi, _stream()), "ac_conv_last_bf16")
    if _PROFILER:
        _PROFILER.end("conv_last", t0, 2.0 * B * T * K * C, 2.0 * B * T * C + 4.0 * B * T)
    return out


LSTM_SIDE_STREAM = {}  # current stream id -> high-priority stream the recurrence kernels of that stream run on (overlap mode)


def lstm_tc(pre, w_hh_bf16, out=None, skip=None, final=None, final_act=ACT_NONE, dbg=None):
    """tensor-core LSTM recurrence: pre [B,T,2048] fp32; out / skip / final are tc.Act (bf16 hi[/lo] planes).
    w_hh_bf16: [2048,512] bf16, or fp16 (then h is fed back as fp16 too: 11-bit operands)."""
    from ._lib import AcLstmTcDesc
    _need_cuda(pre, w_hh_bf16)
    B, T, C4 = pre.shape
    d = AcLstmTcDesc()
    d.pre, d.w_hh_bf16 = pre.data_ptr(), w_hh_bf16.data_ptr()
    if out is not None:
        assert out.hl == 0 and out.hr == 0
        d.out_hi, d.out_lo = out.row_ptr(0), out.lo_ptr(0)
    if skip is not None:
        d.skip_hi, d.skip_lo, d.skip_bstride = skip.row_ptr(0), skip.lo_ptr(0), skip.bstride
    if final is not None:
        d.final_hi, d.final_lo, d.final_bstride = final.row_ptr(0), final.lo_ptr(0), final.bstride
    d.final_act, d.batch, d.steps, d.hidden = final_act, B, T, C4 // 4
    d.dbg = _ptr(dbg)
    assert w_hh_bf16.dtype in (torch.bfloat16, torch.float16) and w_hh_bf16.is_contiguous()
    d.operand_fp16 = int(w_hh_bf16.dtype == torch.float16)
    d.out_fp16 = int(_same_f16(out, final))
    d.skip_fp16 = int(skip is not None and skip.f16)
    t0 = _PROFILER.begin() if _PROFILER else None
    cur = torch.cuda.current_stream()
    side = LSTM_SIDE_STREAM.get(cur.cuda_stream)
    if side is not None:
        # the cluster kernel goes to a high-priority stream: when SMs free up at a kernel boundary of the (capped) tap-GEMM
        # kernels its 16-CTA clusters are placed first, and the GEMMs of the other half-batch run beside it
        side.wait_stream(cur)
        _lib.check(_lib.lib().ac_lstm_tc(ctypes.byref(d), ctypes.c_void_p(side.cuda_stream)), "ac_lstm_tc")
        cur.wait_stream(side)
        for t in (pre, w_hh_bf16):
            t.record_stream(side)
        for a in (out, skip, final):
            if a is not None:
                a.buf.record_stream(side)
                if a.lo is not None:
                    a.lo.record_stream(side)
    else:
        _lib.check(_lib.lib().ac_lstm_tc(ctypes.byref(d), _stream()), "ac_lstm_tc")
    if _PROFILER:
        _PROFILER.end("lstm_tc_kernel", t0, 2.0 * B * T * C4 * (C4 // 4), 4.0 * pre.numel() + 2.0 * B * T * (C4 // 4))


def add_act_bf16(a, b, out, act=ACT_NONE):
    """out = act(a + b) on tc.Act tensors (valid rows only; hi [+lo] planes)."""
    assert a.L == b.L == out.L and a.C == b.C == out.C and a.B == b.B == out.B
    vp = lambda v: ctypes.c_void_p(v) if v is not None else None
    t0 = _PROFILER.begin() if _PROFILER else None
    _lib.check(_lib.lib().ac_add_act_bf16(vp(a.row_ptr(0)), vp(a.lo_ptr(0)), vp(b.row_ptr(0)), vp(b.lo_ptr(0)), vp(out.row_ptr(0)),
                                          vp(out.lo_ptr(0)), a.B, a.L * a.C, a.bstride, b.bstride, out.bstride, act,
                                          int(a.f16) | (int(b.f16) << 1) | (int(out.f16) << 2), _stream()),
               "ac_add_act_bf16")
    if _PROFILER:
        _PROFILER.end("add_act_bf16", t0, 0.0, 6.0 * a.B * a.L * a.C)


def f32_to_act(x, out):
    """x [B, L, C] fp32 contiguous -> tc.Act `out` (hi [+lo] planes, valid rows)."""
    _need_cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous() and tuple(x.shape) == (out.B, out.L, out.C)
    vp = lambda v: ctypes.c_void_p(v) if v is not None else None
    t0 = _PROFILER.begin() if _PROFILER else None
    _lib.check(_lib.lib().ac_f32_to_split_bf16(_ptr(x), vp(out.row_ptr(0)), vp(out.lo_ptr(0)), out.B, out.L * out.C, x.stride(0),
                                               out.bstride, int(out.f16), _stream()), "ac_f32_to_split_bf16")
    if _PROFILER:
        _PROFILER.end("f32_to_split_bf16", t0, 0.0, 6.0 * x.numel())


def rvq_decode_bf16(codes, codebooks, stages, out_act, code_offset=0, err_flag=None):
    """codes [B*N, Ktot] int64 -> bf16 rows of the tc.Act `out_act` ([B][hl+N+hr][D])."""
    _need_cuda(codes, codebooks)
    rows = codes.shape[0]
    D = codebooks.shape[2]
    assert codes.dtype == torch.int64 and codes.is_contiguous() and out_act.C == D and rows == out_act.B * out_act.L
    t0 = _PROFILER.begin() if _PROFILER else None
    _lib.check(_lib.lib().ac_rvq_decode_bf16(_ptr(codes), _ptr(codebooks), ctypes.c_void_p(out_act.row_ptr(0)),
                                             ctypes.c_void_p(out_act.lo_ptr(0)) if out_act.lo is not None else None, rows, out_act.L,
                                             out_act.bstride, D, codebooks.shape[1], stages, codes.shape[-1], code_offset,
                                             _ptr(err_flag), int(out_act.f16), _stream()), "ac_rvq_decode_bf16")
    if _PROFILER:
        _PROFILER.end("rvq_decode", t0, 0.0, 8.0 * rows * stages + 4.0 * rows * D * stages + 2.0 * rows * D)


def rvq_encode(x, codebooks, cb_norm, codes_out, stages, code_offset=0, metric=0, residual_out=None):
    """x [rows, D] fp32; codebooks [S, C, D]; writes codes_out[rows, Ktot] int64 columns code_offset..+stages."""
    _need_cuda(x, codebooks, cb_norm, codes_out)
    rows, D = x.shape
    assert x.is_contiguous() and codes_out.dtype == torch.int64 and codes_out.is_contiguous()
    t0 = _PROFILER.begin() if _PROFILER else None
    _lib.check(_lib.lib().ac_rvq_encode_f32(_ptr(x), _ptr(codebooks), _ptr(cb_norm), _ptr(codes_out), _ptr(residual_out),
                                            rows, D, codebooks.shape[1], stages, codes_out.shape[-1], code_offset, metric,
                                            _stream()), "ac_rvq_encode_f32")
    if _PROFILER:
        _PROFILER.end("rvq_encode_f32", t0, 2.0 * rows * D * codebooks.shape[1] * stages, 4.0 * x.numel() + 8.0 * rows * stages)
    return codes_out


def rvq_encode_tc(x, cb_split, codebooks, cb_norm, codes_out, stages, code_offset=0, residual_out=None, stage0=0, metric=0):
    """tensor-core RVQ encode: x [rows, D] fp32 (D = 128 / 256); cb_split [2, S, C, D] bf16 (hi, lo planes); stages
    stage0 .. stage0+stages-1 of the stacked codebooks; metric 0 = EnCodec, 1 = Mimi (cdist)."""
    _need_cuda(x, cb_split, codebooks, cb_norm, codes_out)
    rows, D = x.shape
    assert x.is_contiguous() and codes_out.dtype == torch.int64 and codes_out.is_contiguous()
    assert cb_split.dtype == torch.bfloat16 and cb_split.is_contiguous() and cb_split.shape[1:] == codebooks.shape
    t0 = _PROFILER.begin() if _PROFILER else None
    _lib.check(_lib.lib().ac_rvq_encode_tc(_ptr(x), _ptr(cb_split), _ptr(codebooks), _ptr(cb_norm), _ptr(codes_out), _ptr(residual_out),
                                           rows, D, codebooks.shape[1], stages, stage0, codebooks.shape[0], codes_out.shape[-1],
                                           code_offset, metric, _stream()), "ac_rvq_encode_tc")
    if _PROFILER:
        _PROFILER.end("rvq_encode_tc_kernel", t0, 2.0 * rows * D * codebooks.shape[1] * stages, 4.0 * x.numel() + 8.0 * rows * stages)
    return codes_out


def rvq_decode(codes, codebooks, stages, code_offset=0, err_flag=None):
    """codes [rows, Ktot] int64 -> [rows, D] fp32."""
    _need_cuda(codes, codebooks)
    rows = codes.shape[0]
    D = codebooks.shape[2]
    assert codes.dtype == torch.int64 and codes.is_contiguous()
    out = torch.empty((rows, D), device=codes.device, dtype=torch.float32)
    t0 = _PROFILER.begin() if _PROFILER else None
    _lib.check(_lib.lib().ac_rvq_decode_f32(_ptr(codes), _ptr(codebooks), _ptr(out), rows, D, codebooks.shape[1], stages,
                                            codes.shape[-1], code_offset, _ptr(err_flag), _stream()), "ac_rvq_decode_f32")
    if _PROFILER:
        _PROFILER.end("rvq_decode_f32", t0, 0.0, 8.0 * rows * stages + 4.0 * rows * D * (stages + 1))
    return out


_TAPS_CACHE = {}


def resample_taps(orig_freq, new_freq, device):
    """Windowed-sinc polyphase taps, built once per (orig,new) pair (the reference rebuilds them per
    call, TA/functional.py:1305-1402).  Index arithmetic in fp32 exactly as torchaudio does for fp32
    waveforms; this is setup-time host math on a few hundred floats, not the data path."""
    key = (orig_freq, new_freq, str(device))
    if key not in _TAPS_CACHE:
        g = math.gcd(int(orig_freq), int(new_freq))
        o, n = int(orig_freq) // g, int(new_freq) // g
        base = min(o, n) * 0.99
        width = math.ceil(6 * o / base)
        idx = torch.arange(-width, width + o, dtype=torch.float32)[None] / o
        t = torch.arange(0, -n, -1, dtype=torch.float32)[:, None] / n + idx
        t = (t * base).clamp(-6, 6)
        window = torch.cos(t * math.pi / 6 / 2) ** 2
        t = t * math.pi
        taps = torch.where(t == 0, torch.tensor(1.0), t.sin() / t) * window * (base / o)
        _TAPS_CACHE[key] = (taps.contiguous().to(device), width, o, n)
    return _TAPS_CACHE[key]


def resample(sig, orig_freq, new_freq):
    """sig [B,T] fp32 -> [B, ceil(new*T/orig)] ; identity when the rates match (TA:1473-1474)."""
    if orig_freq == new_freq:
        return sig
    _need_cuda(sig)
    taps, width, o, n = resample_taps(orig_freq, new_freq, sig.device)
    B, T = sig.shape
    sig = sig.contiguous().float()
    out_len = int(torch.ceil(torch.as_tensor(n * T / o)).long())  # TA:1426 (fp32 rounding before ceil)
    out = torch.empty((B, out_len), device=sig.device, dtype=torch.float32)
    _lib.check(_lib.lib().ac_resample_f32(_ptr(sig), _ptr(taps), _ptr(out), B, T, out_len, o, n, taps.shape[1], width,
                                          _stream()), "ac_resample_f32")
    return out


# ---------------------------------------------------------------------------------------------- Mimi transformer pieces
def layernorm(x, w, b, eps=1e-5):
    """x [..., C] fp32 contiguous."""
    _need_cuda(x, w, b)
    assert x.is_contiguous() and x.dtype == torch.float32
    y = torch.empty_like(x)
    C = x.shape[-1]
    t0 = _PROFILER.begin() if _PROFILER else None
    _lib.check(_lib.lib().ac_layernorm_f32(_ptr(x), _ptr(w), _ptr(b), _ptr(y), x.numel() // C, C, eps, _stream()), "ac_layernorm_f32")
    if _PROFILER:
        _PROFILER.end("layernorm_f32_kernel", t0, 0.0, 8.0 * x.numel())
    return y


def attention(qkv, inv_freq, heads, head_dim, window):
    """qkv [B,T,3*H*D] fp32 -> [B,T,H*D]; RoPE + causal sliding-window softmax attention."""
    _need_cuda(qkv, inv_freq)
    B, T, _ = qkv.shape
    assert qkv.is_contiguous() and qkv.shape[2] == 3 * heads * head_dim
    out = torch.empty((B, T, heads * head_dim), device=qkv.device, dtype=torch.float32)
    t0 = _PROFILER.begin() if _PROFILER else None
    _lib.check(_lib.lib().ac_attention_f32(_ptr(qkv), _ptr(inv_freq), _ptr(out), B, T, heads, head_dim, window,
                                           1.0 / math.sqrt(head_dim), _stream()), "ac_attention_f32")
    if _PROFILER:
        _PROFILER.end("attention_f32_kernel", t0, 4.0 * B * heads * head_dim * T * min(T, window) / 2, 4.0 * (qkv.numel() + out.numel()))
    return out


def layernorm_act(x, w, b, out, eps=1e-5):
    """x [B, L, C] fp32 contiguous -> LayerNorm over C written into the tc.Act `out` (hi [+lo] planes)."""
    _need_cuda(x, w, b)
    assert x.is_contiguous() and x.dtype == torch.float32 and tuple(x.shape) == (out.B, out.L, out.C)
    vp = lambda v: ctypes.c_void_p(v) if v is not None else None
    t0 = _PROFILER.begin() if _PROFILER else None
    _lib.check(_lib.lib().ac_layernorm_split_bf16(_ptr(x), _ptr(w), _ptr(b), vp(out.row_ptr(0)), vp(out.lo_ptr(0)), out.B, out.L, out.C,
                                                  out.bstride, eps, int(out.f16), _stream()), "ac_layernorm_split_bf16")
    if _PROFILER:
        _PROFILER.end("layernorm_split_kernel", t0, 0.0, 8.0 * x.numel())


def rope_table(inv_freq, T):
    """[T, 2*len(inv_freq)] fp32: cos | sin of position * inv_freq (HF/mimi:515-558)."""
    _need_cuda(inv_freq)
    half = inv_freq.numel()
    table = torch.empty((T, 2 * half), device=inv_freq.device, dtype=torch.float32)
    _lib.check(_lib.lib().ac_rope_table_f32(_ptr(inv_freq), _ptr(table), T, half, _stream()), "ac_rope_table_f32")
    return table


def attention_tc(qkv, rope, heads, head_dim, window, out_act=None, out32=False):
    """qkv [B,T,3*H*D] fp32 -> tc.Act `out_act` (split bf16) and/or a new fp32 [B,T,H*D]; RoPE + causal sliding-window
    softmax attention on tcgen05 (split-bf16 products, fp32 softmax)."""
    _need_cuda(qkv, rope)
    B, T, _ = qkv.shape
    assert qkv.is_contiguous() and qkv.dtype == torch.float32 and qkv.shape[2] == 3 * heads * head_dim
    assert rope.shape == (T, head_dim) and rope.is_contiguous()
    assert out_act is not None or out32
    o32 = torch.empty((B, T, heads * head_dim), device=qkv.device, dtype=torch.float32) if out32 else None
    vp = lambda v: ctypes.c_void_p(v) if v is not None else None
    hi = lo = None
    bs = 0
    if out_act is not None:
        assert (out_act.B, out_act.L, out_act.C) == (B, T, heads * head_dim)
        hi, lo, bs = out_act.row_ptr(0), out_act.lo_ptr(0), out_act.bstride
    t0 = _PROFILER.begin() if _PROFILER else None
    _lib.check(_lib.lib().ac_attention_tc(_ptr(qkv), _ptr(rope), _ptr(o32), vp(hi), vp(lo), bs, B, T, heads, head_dim, window,
                                          1.0 / math.sqrt(head_dim), int(out_act is not None and out_act.f16), _stream()), "ac_attention_tc")
    if _PROFILER:
        _PROFILER.end("attention_tc_kernel", t0, 4.0 * B * heads * head_dim * T * min(T, window) / 2, 4.0 * qkv.numel() + 4.0 * B * T * heads * head_dim)
    return o32


def upsample_dw(x, w):
    """x [B,L,C] fp32, w [C,4] -> [B,2L,C] (depthwise ConvTranspose1d k4 s2, causal trim)."""
    _need_cuda(x, w)
    B, L, C = x.shape
    y = torch.empty((B, 2 * L, C), device=x.device, dtype=torch.float32)
    _lib.check(_lib.lib().ac_upsample_dw_f32(_ptr(x.contiguous()), _ptr(w), _ptr(y), B, L, C, _stream()), "ac_upsample_dw_f32")
    return y


# ---------------------------------------------------------------------------------------------- DAC RVQ
def dac_rvq_encode(z, w_in, b_in, cb, w_out, b_out, stages, want_zq=False):
    """z [B,N,1024] fp32 -> codes [B,N,stages] int64 (and the quantised sum [B,N,1024])."""
    _need_cuda(z, w_in)
    B, N, H = z.shape
    codes = torch.empty((B, N, stages), device=z.device, dtype=torch.int64)
    zq = torch.empty_like(z) if want_zq else None
    t0 = _PROFILER.begin() if _PROFILER else None
    _lib.check(_lib.lib().ac_dac_rvq_encode_f32(_ptr(z.contiguous()), _ptr(w_in), _ptr(b_in), _ptr(cb), _ptr(w_out), _ptr(b_out),
                                                _ptr(codes), _ptr(zq), B * N, H, cb.shape[2], cb.shape[1], stages, stages,
                                                _stream()), "ac_dac_rvq_encode_f32")
    if _PROFILER:
        _PROFILER.end("dac_rvq_encode_kernel", t0, 2.0 * B * N * stages * (2 * H * 8 + 8 * cb.shape[1]), 4.0 * z.numel())
    return (codes, zq) if want_zq else codes


def dac_rvq_encode_proj(proj, cconst, cross, cb_normed, cb_norm2, cb, stages):
    """proj [B,N,ld] fp32 (all stages' in_proj of z, bias included) -> codes [B,N,stages] int64."""
    _need_cuda(proj, cross)
    B, N, ld = proj.shape
    assert proj.is_contiguous() and proj.dtype == torch.float32
    codes = torch.empty((B, N, stages), device=proj.device, dtype=torch.int64)
    t0 = _PROFILER.begin() if _PROFILER else None
    _lib.check(_lib.lib().ac_dac_rvq_encode_proj_f32(_ptr(proj), ld, _ptr(cconst), _ptr(cross), _ptr(cb_normed), _ptr(cb_norm2), _ptr(cb),
                                                     _ptr(codes), B * N, cb.shape[2], cb.shape[1], stages, cb.shape[0], stages,
                                                     _stream()), "ac_dac_rvq_encode_proj_f32")
    if _PROFILER:
        _PROFILER.end("dac_rvq_encode_proj_kernel", t0, 2.0 * B * N * stages * 8 * cb.shape[1], 4.0 * proj.numel())
    return codes


def dac_rvq_decode(codes, cb, w_out, b_out, err_flag=None):
    """codes [B,N,K] int64 -> z [B,N,1024] fp32 (from_codes)."""
    _need_cuda(codes, cb)
    B, N, K = codes.shape
    codes = codes.to(torch.int64).contiguous()
    out = torch.empty((B, N, w_out.shape[1]), device=codes.device, dtype=torch.float32)
    t0 = _PROFILER.begin() if _PROFILER else None
    _lib.check(_lib.lib().ac_dac_rvq_decode_f32(_ptr(codes), _ptr(cb), _ptr(w_out), _ptr(b_out), _ptr(out), B * N, w_out.shape[1],
                                                cb.shape[2], cb.shape[1], K, K, _ptr(err_flag), _stream()), "ac_dac_rvq_decode_f32")
    if _PROFILER:
        _PROFILER.end("dac_rvq_decode_kernel", t0, 2.0 * B * N * K * 8 * w_out.shape[1], 4.0 * out.numel())
    return out
