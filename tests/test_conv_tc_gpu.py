"""GPU: the tcgen05 tap-GEMM kernel (ac_conv_tc) against a plain PyTorch fp32 reference of the same op on
bf16-rounded operands (floating-point kernel -> torch fp32 reference, tolerance = bf16 output rounding)."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

from audiocodecs_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _run(srcs, w, n_total, bk, batch, m_rows, bias=None, res=None, out_rows=None, out_shift=0, out_valid=None,
         want_y32=False, act=0, alpha=None, act_mod=0, n_tile_hint=0, grid_hint=0):
    """srcs: list of dict(t=[B,R,C] bf16 cuda tensor, origin_row, c0, phases, rows, taps, dilation, shift)."""
    d = _lib.AcConvTcDesc()
    k_total = 0
    for i, s in enumerate(srcs):
        t = s["t"]
        assert t.is_contiguous() and t.dtype == torch.bfloat16
        C = t.shape[2]
        S = d.src[i]
        S.base = t.data_ptr() + s["origin_row"] * C * 2
        S.c0, S.phases, S.rows = s["c0"], s["phases"], s["rows"]
        S.phase_stride, S.row_stride, S.batch_stride = s["c0"], s["c0"] * s["phases"], t.stride(0)
        S.taps, S.dilation, S.shift, S.lo_of = s["taps"], s["dilation"], s["shift"], s.get("lo_of", -1)
        k_total += s["taps"] * s["c0"] * s["phases"]
    d.n_src = len(srcs)
    d.w, d.k_total, d.n_total, d.bk = w.data_ptr(), k_total, n_total, bk
    out_rows = out_rows or m_rows
    n_flat = out_valid if out_valid is not None else out_rows * n_total
    y = torch.full((batch, n_flat), float("nan"), device=DEV, dtype=torch.bfloat16)
    ya = torch.full((batch, n_flat), float("nan"), device=DEV, dtype=torch.bfloat16) if act else None
    y32 = torch.full((batch, n_flat), float("nan"), device=DEV, dtype=torch.float32) if want_y32 else None
    d.bias = bias.data_ptr() if bias is not None else None
    d.alpha = alpha.data_ptr() if alpha is not None else None
    d.res = res.data_ptr() if res is not None else None
    d.y, d.y_act, d.y32 = y.data_ptr(), (ya.data_ptr() if act else None), (y32.data_ptr() if want_y32 else None)
    d.act, d.epi, d.act_mod = act, 0, act_mod
    d.y_bstride = d.y_act_bstride = d.y32_bstride = d.res_bstride = n_flat
    d.out_shift, d.out_valid = out_shift, n_flat
    d.batch, d.m_rows, d.n_tile_hint, d.grid_hint = batch, m_rows, n_tile_hint, grid_hint
    _lib.check(_lib.lib().ac_conv_tc(ctypes.byref(d), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "ac_conv_tc")
    torch.cuda.synchronize()
    return y, ya, y32


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(torch.bfloat16)


def _close(got, ref, what):
    got, ref = got.float().cpu(), ref.float()
    assert torch.isfinite(got).all(), f"{what}: non-finite / unwritten outputs"
    err = (got - ref).abs().max().item()
    tol = 2e-2 * max(1.0, ref.abs().max().item())  # bf16 output rounding (2^-8 rel) + fp32 accumulation order
    assert err <= tol, f"{what}: max err {err} > {tol}"


@pytest.mark.parametrize("B,L,Cin,Cout,bk,n_tile", [(2, 300, 64, 64, 64, 0), (1, 128, 128, 256, 64, 0), (3, 1000, 32, 32, 32, 0),
                                                     (2, 257, 16, 32, 16, 0), (1, 640, 512, 512, 64, 128), (2, 200, 96, 192, 32, 0)])
def test_pointwise_gemm(B, L, Cin, Cout, bk, n_tile):
    """1 tap, stride 1: y[b,l,:] = W x[b,l,:] + bias (+res), fp32 and bf16 outputs."""
    x = _rand((B, L, Cin), 1)
    w = _rand((Cout, Cin), 2, Cin ** -0.5)
    bias = torch.randn(Cout, generator=torch.Generator().manual_seed(3))
    res = _rand((B, L * Cout), 4)
    src = dict(t=x.to(DEV), origin_row=0, c0=Cin, phases=1, rows=L, taps=1, dilation=1, shift=0)
    y, _, y32 = _run([src], w.to(DEV), Cout, bk, B, L, bias=bias.to(DEV), res=res.to(DEV), want_y32=True, n_tile_hint=n_tile)
    ref = x.float() @ w.float().t() + bias + res.float().view(B, L, Cout)
    _close(y32.view(B, L, Cout), ref, "y32")
    _close(y.view(B, L, Cout), ref, "y")


@pytest.mark.parametrize("Cin,Cout,K,dil,bk", [(32, 16, 3, 1, 32), (64, 64, 7, 3, 64), (16, 32, 3, 1, 16), (128, 64, 7, 9, 64)])
def test_stride1_taps_zero_pad(Cin, Cout, K, dil, bk):
    """K-tap dilated conv with zero padding coming from TMA out-of-bounds fill; ELU second output."""
    B, L = 2, 500
    x = _rand((B, L, Cin), 5)
    w = _rand((Cout, Cin, K), 6, (Cin * K) ** -0.5)
    pad = (K - 1) * dil // 2
    wp = w.permute(0, 2, 1).reshape(Cout, K * Cin).contiguous()  # [n][tap][c]
    src = dict(t=x.to(DEV), origin_row=0, c0=Cin, phases=1, rows=L, taps=K, dilation=dil, shift=-pad)
    y, ya, _ = _run([src], wp.to(DEV), Cout, bk, B, L, act=1)
    ref = F.conv1d(x.float().transpose(1, 2), w.float(), padding=pad, dilation=dil).transpose(1, 2)
    _close(y.view(B, L, Cout), ref, "y")
    _close(ya.view(B, L, Cout), F.elu(ref), "y_act")


@pytest.mark.parametrize("Cin,Cout,s,bk", [(64, 128, 4, 64), (128, 256, 5, 64), (256, 512, 8, 64), (32, 64, 2, 32), (16, 32, 2, 16)])
def test_strided_conv_as_two_tap_view(Cin, Cout, s, bk):
    """kernel 2s / stride s causal conv (left pad s, zeros here) through the [L/s][s*Cin] view."""
    B, Lout = 2, 150
    L = Lout * s
    x = _rand((B, L, Cin), 7)
    w = _rand((Cout, Cin, 2 * s), 8, (Cin * 2 * s) ** -0.5)
    # taps j'=0,1 ; within a tap k = phase*Cin + c  <->  kernel index j'*s + phase
    wp = w.permute(0, 2, 1).reshape(Cout, 2 * s * Cin).contiguous()
    src = dict(t=x.to(DEV), origin_row=0, c0=Cin, phases=s, rows=Lout, taps=2, dilation=1, shift=-1)
    y, _, _ = _run([src], wp.to(DEV), Cout, bk, B, Lout)
    ref = F.conv1d(F.pad(x.float().transpose(1, 2), (s, 0)), w.float(), stride=s).transpose(1, 2)
    _close(y.view(B, Lout, Cout), ref, "y")


def test_two_sources_resblock_tail():
    """shortcut(x) + conv_k1(h) as ONE GEMM over two sources (C=32 raw x, C/2=16 hidden)."""
    B, L, C = 2, 700, 32
    x, h = _rand((B, L, C), 9), _rand((B, L, C // 2), 10)
    wsc, w1 = _rand((C, C), 11, C ** -0.5), _rand((C, C // 2), 12, (C // 2) ** -0.5)
    wp = torch.cat([wsc, w1], dim=1).contiguous()
    s0 = dict(t=x.to(DEV), origin_row=0, c0=C, phases=1, rows=L, taps=1, dilation=1, shift=0)
    s1 = dict(t=h.to(DEV), origin_row=0, c0=C // 2, phases=1, rows=L, taps=1, dilation=1, shift=0)
    y, _, _ = _run([s0, s1], wp.to(DEV), C, 16, B, L)
    ref = x.float() @ wsc.float().t() + h.float() @ w1.float().t()
    _close(y.view(B, L, C), ref, "y")


@pytest.mark.parametrize("Cin,Cout,s,pad", [(64, 32, 2, 0), (512, 256, 8, 0), (128, 64, 4, 2), (192, 96, 2, 1)])
def test_transposed_conv_flat_output(Cin, Cout, s, pad):
    """ConvTranspose1d(k=2s, stride s, padding pad) as the 2-tap GEMM over n=(phase,cout) with flat-shifted stores.
    pad=0 -> EnCodec causal trim (L*s outputs); pad=ceil(s/2) -> DAC."""
    B, L = 2, 130
    x = _rand((B, L, Cin), 13)
    w = _rand((Cin, Cout, 2 * s), 14, (2 * Cin) ** -0.5)
    wk = w.float().permute(2, 0, 1)  # [K, Cin, Cout]
    hi = wk[s:].permute(1, 0, 2).reshape(Cin, s * Cout)
    lo = wk[:s].permute(1, 0, 2).reshape(Cin, s * Cout)
    wp = torch.cat([hi.t(), lo.t()], dim=1).contiguous().to(torch.bfloat16)  # [n][tap0: x[q-1] | tap1: x[q]]
    full = F.conv_transpose1d(x.float().transpose(1, 2), w.float(), stride=s, padding=pad)
    Lout = L * s if pad == 0 else full.shape[-1]
    ref = full[..., :Lout].transpose(1, 2)
    m_rows = L + 1 if pad else L
    src = dict(t=x.to(DEV), origin_row=0, c0=Cin, phases=1, rows=L, taps=2, dilation=1, shift=-1)
    y, _, _ = _run([src], wp.to(DEV), s * Cout, 64 if Cin % 64 == 0 else 32, B, m_rows, out_shift=pad * Cout,
                   out_valid=Lout * Cout)
    _close(y.view(B, Lout, Cout), ref, "y")


def test_many_tiles_persistent_schedule():
    """more tiles than CTAs (grid_hint=3) and an N axis split into 4 tiles: exercises the ring/TMEM phase logic."""
    B, L, Cin, Cout = 3, 900, 64, 256
    x = _rand((B, L, Cin), 15)
    w = _rand((Cout, Cin), 16, Cin ** -0.5)
    src = dict(t=x.to(DEV), origin_row=0, c0=Cin, phases=1, rows=L, taps=1, dilation=1, shift=0)
    y, _, _ = _run([src], w.to(DEV), Cout, 64, B, L, n_tile_hint=64, grid_hint=3)
    _close(y.view(B, L, Cout), x.float() @ w.float().t(), "y")


def test_split_precision_near_fp32():
    """hi/lo activation planes + hi/lo weight planes: A_hi*W_hi + A_hi*W_lo + A_lo*W_hi reproduces the fp32 GEMM to
    ~1e-5 relative (plain bf16 operands: ~3e-3), and the lo output plane carries the rounding residual."""
    B, L, Cin, Cout = 2, 300, 128, 128
    g = torch.Generator().manual_seed(21)
    x = torch.randn(B, L, Cin, generator=g)
    w = torch.randn(Cout, Cin, generator=g) * Cin ** -0.5
    x_hi = x.to(torch.bfloat16)
    x_lo = (x - x_hi.float()).to(torch.bfloat16)
    w_hi = w.to(torch.bfloat16)
    w_lo = (w - w_hi.float()).to(torch.bfloat16)
    d = _lib.AcConvTcDesc()
    xh, xl = x_hi.to(DEV), x_lo.to(DEV)
    for i, (t, lo_of) in enumerate(((xh, -1), (xl, 0))):
        S = d.src[i]
        S.base, S.c0, S.phases, S.rows = t.data_ptr(), Cin, 1, L
        S.phase_stride, S.row_stride, S.batch_stride = Cin, Cin, t.stride(0)
        S.taps, S.dilation, S.shift, S.lo_of = 1, 1, 0, lo_of
    wst = torch.stack([w_hi, w_lo]).contiguous().to(DEV)
    y = torch.empty((B, L, Cout), device=DEV, dtype=torch.bfloat16)
    ylo = torch.empty_like(y)
    y32 = torch.empty((B, L, Cout), device=DEV, dtype=torch.float32)
    d.n_src, d.w, d.w_split, d.k_total, d.n_total, d.bk = 2, wst.data_ptr(), 1, Cin, Cout, 64
    d.y, d.y_lo, d.y32 = y.data_ptr(), ylo.data_ptr(), y32.data_ptr()
    d.y_bstride = d.y32_bstride = L * Cout
    d.out_shift, d.out_valid, d.batch, d.m_rows = 0, L * Cout, B, L
    _lib.check(_lib.lib().ac_conv_tc(ctypes.byref(d), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "ac_conv_tc")
    torch.cuda.synchronize()
    ref = x.double() @ w.double().t()
    rel = ((y32.cpu().double() - ref).norm() / ref.norm()).item()
    assert rel < 3e-5, rel
    rel2 = ((y.float().cpu().double() + ylo.float().cpu().double() - ref).norm() / ref.norm()).item()
    assert rel2 < 5e-5, rel2
