"""GPU: the operand formats of the tensor path (tc.Policy): one fp16 plane per operand ("fp16": one product per MAC), and the
(fp16 hi, bf16 lo) activation pair against (fp16 hi, fp16 lo, bf16 hi) weight planes ("exact": three products, ~2^-20).
Floating-point kernels -> plain PyTorch fp32 / fp64 reference of the same op, tolerance = the format's rounding."""
import pytest
import torch
import torch.nn.functional as F

from audiocodecs_b200 import ops, tc
from audiocodecs_b200.tc import Act, Src, TcWeights

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ONE = tc.Policy(True, tc.NEVER, False)           # precision="fp16"
FULL = tc.Policy(True, 0, True, True)            # precision="exact", encoder side
BF = tc.Policy(False, 0, True)                   # split bf16 (round 1's compensated pair), for comparison


def _fill(act, x):
    """x [B, L, C] fp32 -> the planes of `act` (valid rows); returns the value the planes represent"""
    hi = x.to(torch.float16 if act.f16 else torch.bfloat16)
    act.buf[:, act.hl:act.hl + act.L] = hi.to(DEV)
    val = hi.double()
    if act.lo is not None:
        lo = (x - hi.float()).to(torch.bfloat16)
        act.lo[:, act.hl:act.hl + act.L] = lo.to(DEV)
        val = val + lo.double()
    return val


def _wval(W):
    """the weight value the MMAs see, per product group: (what multiplies A_hi, what multiplies A_lo)"""
    if W.f16:
        planes = [p.view(torch.float16).double() for p in W.w[: 1 + W.split].cpu()]
        hi = sum(planes)
        lo_w = W.w[-1].cpu().view(torch.bfloat16).double() if W.hib else planes[0]
        return hi, lo_w
    if W.split:
        return W.w[0].cpu().double() + W.w[1].cpu().double(), W.w[0].cpu().double()
    return W.w.cpu().double(), W.w.cpu().double()


def _to_dev(W):
    W.apply(lambda t: t.to(DEV))
    return W


@pytest.mark.parametrize("pol,tol", [(ONE, 1.5e-3), (FULL, 6e-6), (BF, 3e-5)])
@pytest.mark.parametrize("B,L,Cin,Cout,K,dil", [(2, 300, 64, 64, 1, 1), (1, 200, 32, 16, 3, 1), (2, 150, 128, 256, 7, 3), (1, 130, 16, 32, 3, 1)])
def test_conv_tc_formats(pol, tol, B, L, Cin, Cout, K, dil):
    """K-tap conv (zero padding through TMA out-of-bounds fill), fp32 output + ELU'd 16-bit output in the policy's format."""
    g = torch.Generator().manual_seed(Cin + K)
    x = torch.randn(B, L, Cin, generator=g)
    w = torch.randn(Cout, Cin, K, generator=g) * (Cin * K) ** -0.5
    bias = torch.randn(Cout, generator=g) * 0.1
    a = pol.act(B, L, Cin, DEV)
    W = _to_dev(pol.weights(w.permute(0, 2, 1).reshape(Cout, -1), bias))
    y32 = torch.empty((B, L, Cout), device=DEV, dtype=torch.float32)
    ya = pol.act(B, L, Cout, DEV)
    pad = (K - 1) * dil   # halo-free causal conv: rows before 0 read as zero (TMA out-of-bounds fill)
    _fill(a, x)
    tc.conv_tc(W, [Src(a, taps=K, dilation=dil, shift=-pad)], L, y32=y32, y_act=ya, act=ops.ACT_ELU)
    torch.cuda.synchronize()
    # exact fp64 reference of full-precision operands, and of what the planes hold
    ref = F.conv1d(F.pad(x.double().transpose(1, 2), (pad, 0)), w.double(), bias.double(), dilation=dil).transpose(1, 2)
    err = ((y32.cpu().double() - ref).norm() / ref.norm()).item()
    print(f"{'fp16-one' if pol is ONE else ('exact' if pol is FULL else 'bf16-pair')} conv K={K} C={Cin}->{Cout}: rel err vs fp64 {err:.2e}")
    assert err < tol, err
    ref_act = F.elu(ref)
    got_act = ya.value().cpu().double()
    out_tol = {True: 3e-6, False: 1e-3}[ya.lo is not None] if ya.f16 else (3e-5 if ya.lo is not None else 1e-2)
    assert ((got_act - ref_act).norm() / ref_act.norm()).item() < max(out_tol, 2 * tol)


@pytest.mark.parametrize("pol,tol", [(ONE, 2e-3), (FULL, 5e-6)])
@pytest.mark.parametrize("C,L,B,g,dbl", [(32, 1000, 2, 0, -1), (64, 777, 2, 2, 1), (128, 300, 1, 1, 0)])
def test_resunit_tc_formats(pol, tol, C, L, B, g, dbl):
    """EnCodec residual block in one launch: ye = ELU(shortcut(x) + conv1(ELU(conv3(xe)))), hidden tile in the policy's format."""
    gen = torch.Generator().manual_seed(300 + C)
    x = torch.randn(B, L, C, generator=gen)
    w3 = torch.randn(C // 2, C, 3, generator=gen) * (3 * C) ** -0.5
    w1 = torch.randn(C, C // 2, generator=gen) * (C // 2) ** -0.5
    wsc = torch.randn(C, C, generator=gen) * C ** -0.5
    b3, b1 = torch.randn(C // 2, generator=gen) * 0.1, torch.randn(C, generator=gen) * 0.1
    xa, xe = pol.act(B, L, C, DEV), pol.act(B, L, C, DEV, hl=2)
    xv = _fill(xa, x)
    xev = _fill(xe, F.elu(x))
    xe.buf[:, :2] = 0
    if xe.lo is not None:
        xe.lo[:, :2] = 0
    W1 = _to_dev(pol.weights(w3.permute(0, 2, 1).reshape(C // 2, -1), b3))
    W2 = _to_dev(pol.weights(torch.cat([w1, wsc], dim=1), b1))
    ye = pol.act(B, L, C, DEV)
    try:
        tc.resunit_tc(W1, W2, Src(xe, taps=3, origin=-2, rows=L + 2), L, x=xa, y_act=ye, act1=ops.ACT_ELU, act2=ops.ACT_ELU,
                      h_split=pol.split(C // 2), g_hint=g, dbl_hint=dbl)
    except Exception as e:  # noqa: BLE001
        from audiocodecs_b200 import _lib
        if isinstance(e, _lib.ConfigError):
            pytest.skip("tile grouping does not fit with the canonical contraction blocks")
        raise
    torch.cuda.synchronize()
    xd = x.double()
    h = F.elu(F.conv1d(F.pad(F.elu(xd).transpose(1, 2), (2, 0)), w3.double(), b3.double()))
    ref = F.elu(F.conv1d(h, w1.double()[:, :, None], b1.double()) + F.conv1d(xd.transpose(1, 2), wsc.double()[:, :, None])).transpose(1, 2)
    got = ye.value().cpu().double()
    err = ((got - ref).norm() / ref.norm()).item()
    print(f"{'fp16-one' if pol is ONE else 'exact'} resblock C={C}: rel err vs fp64 {err:.2e}")
    assert torch.isfinite(got).all() and err < tol, err


@pytest.mark.parametrize("K,N", [(4096, 512), (3584, 128), (1280, 256)])
def test_chunked_accumulation(K, N):
    """The tcgen05 accumulator truncates every add (scripts/accum_probe.py): a K-deep contraction shrinks by ~2e-8 per MMA.
    With chunked accumulation (partial sums of ~96 MMAs added in fp32 round-to-nearest by the epilogue warps) the error of
    the "exact" formats no longer grows with K.  Shapes = EnCodec's deepest encoder contractions (down 256->512 k16,
    512->128 k7, down 128->256 k10), positive-mean operands (the worst case: every partial sum has the same sign)."""
    g = torch.Generator().manual_seed(K)
    M = 700
    x = torch.randn(2, M, K, generator=g) + 0.5
    w = (torch.randn(N, K, generator=g) + 0.5) * K ** -0.5
    a = FULL.act(2, M, K, DEV)
    _fill(a, x)
    W = _to_dev(FULL.weights(w, None))
    ref = x.double() @ w.double().t()
    errs = {}
    for flush in (0, 96):
        y32 = torch.empty((2, M, N), device=DEV, dtype=torch.float32)
        saved, tc.FLUSH_ADDS = tc.FLUSH_ADDS, flush
        try:
            tc.conv_tc(W, [Src(a)], M, y32=y32)
        finally:
            tc.FLUSH_ADDS = saved
        torch.cuda.synchronize()
        errs[flush] = ((y32.cpu().double() - ref).norm() / ref.norm()).item()
    print(f"K={K} N={N}: rel err one accumulator {errs[0]:.2e}, chunked {errs[96]:.2e}")
    assert errs[96] < 2.5e-6 and errs[96] < errs[0] / 2


@pytest.mark.parametrize("pol", [ONE, FULL, BF])
@pytest.mark.parametrize("C,L,B,dil,g", [(64, 900, 2, 1, 2), (96, 500, 2, 9, 1), (128, 641, 1, 3, 1), (192, 300, 2, 9, 1)])
def test_staged_epilogue_io_is_bit_identical(pol, C, L, B, dil, g):
    """DAC residual unit y = x + conv1(Snake(conv7_dilated(Snake(x)))) with the skip input and both outputs (raw + activated)
    moved by TMA through shared-memory tiles (io_stage=1: cp.async.bulk.tensor loads / stores, rows beyond L clipped by the
    hardware) against the direct per-lane global loads / stores: the same bits, ragged lengths included."""
    from audiocodecs_b200 import _lib
    gen = torch.Generator().manual_seed(500 + C + dil)
    x = torch.randn(B, L, C, generator=gen)
    al1, al2, al3 = (torch.rand(C, generator=gen) + 0.5 for _ in range(3))
    w7 = torch.randn(C, C, 7, generator=gen) * (7 * C) ** -0.5
    w1 = torch.randn(C, C, generator=gen) * C ** -0.5
    xa, xs = pol.act(B, L, C, DEV, split=pol.f16 is False or pol.full), pol.act(B, L, C, DEV)
    _fill(xa, x)
    _fill(xs, x + torch.sin(al1 * x) ** 2 / (al1 + 1e-9))
    W7 = _to_dev(pol.weights(w7.permute(0, 2, 1).reshape(C, -1), torch.zeros(C)))
    W1 = _to_dev(pol.weights(w1, torch.zeros(C)))
    outs = {}
    for io in (-1, 1):
        y, ys = pol.act(B, L, C, DEV, split=xa.lo is not None, hl=3, hr=5), pol.act(B, L, C, DEV)
        for t in (y.buf, ys.buf) + ((y.lo,) if y.lo is not None else ()) + ((ys.lo,) if ys.lo is not None else ()):
            t.fill_(7.0)   # sentinel: halo rows and the neighbouring clip must stay untouched
        try:
            tc.resunit_tc(W7, W1, Src(xs, taps=7, dilation=dil, shift=-3 * dil), L, res=xa, y=y, y_act=ys, act1=ops.ACT_SNAKE, alpha1=al2.to(DEV),
                          act2=ops.ACT_SNAKE, alpha2=al3.to(DEV), h_split=pol.split(C), g_hint=g, io_stage=io)
        except _lib.ConfigError:
            pytest.skip("staging does not fit shared memory for this tiling")
        torch.cuda.synchronize()
        outs[io] = [t.clone() for t in (y.buf, ys.buf) + ((y.lo,) if y.lo is not None else ()) + ((ys.lo,) if ys.lo is not None else ())]
    for a, b in zip(outs[-1], outs[1]):
        assert torch.equal(a, b)
    assert (outs[1][0][:, :3] == 7.0).all() and (outs[1][0][:, 3 + L:] == 7.0).all()   # halos of y untouched
    ref = x.double()
    assert torch.isfinite(outs[1][1].float()).all()


def test_fp16_saturates_instead_of_inf():
    """values beyond the fp16 range come out as +-65504 in the hi plane (cvt.satfinite), never inf; the bf16 lo plane of a
    pair then carries the remainder (precision degrades gracefully to ~bf16)."""
    x = torch.tensor([[[1.0e5, -2.0e5, 3.0, 7.0e4] * 4] * 130])          # [1, 130, 16]
    w = torch.eye(16)
    for pol in (ONE, FULL):
        a = pol.act(1, 130, 16, DEV)
        y = pol.act(1, 130, 16, DEV)
        y32 = torch.empty((1, 130, 16), device=DEV, dtype=torch.float32)
        xin = x.clamp(-6.0e4, 6.0e4)   # representable inputs, amplified by the layer itself
        _fill(a, xin)
        W2 = _to_dev(pol.weights(w * 2.0, None))
        tc.conv_tc(W2, [Src(a)], 130, y=y, y32=y32)
        torch.cuda.synchronize()
        assert torch.isfinite(y.buf.float()).all() and y.buf.float().abs().max().item() == 65504.0
        if y.lo is not None:
            rel = ((y.value().cpu() - y32.cpu()).abs() / y32.cpu().abs().clamp_min(1.0)).max().item()
            assert rel < 1e-2, rel   # hi (saturated) + lo (bf16 remainder) still represents the value to bf16 precision


def test_producers_write_fp16_planes(encodec_sd):
    """conv_first / rvq_decode / add_act / f32_to_act / LSTM outputs in the fp16 (+ bf16 lo) format."""
    from oracle import encodec_ref
    import audiocodecs_b200 as A
    codec = A.Encodec(24000, 24000, num_codebooks=8, state_dict=encodec_sd, precision="exact").eval().to(DEV)
    sig = torch.randn(2, 3000, generator=torch.Generator().manual_seed(5)) * 0.1
    x, xe = FULL.act(2, 3000, 32, DEV), FULL.act(2, 3000, 32, DEV, hl=2)
    ops.conv_first_bf16(codec._enc[0], sig.to(DEV), y=x, y_act=xe, act=ops.ACT_ELU)
    w, b = encodec_ref.fold_weight_norm(encodec_sd, "encoder.layers.0")
    ref = encodec_ref.causal_conv(sig[:, None], w, b).transpose(1, 2)
    assert x.buf.dtype == torch.float16
    assert ((x.value().cpu() - ref).abs().max() / ref.abs().max()).item() < 2e-6
    assert ((xe.value().cpu() - F.elu(ref)).abs().max() / ref.abs().max()).item() < 2e-6
    f = torch.randn(2, 40, 64, generator=torch.Generator().manual_seed(6))
    a, bb, o = FULL.act(2, 40, 64, DEV), ONE.act(2, 40, 64, DEV), FULL.act(2, 40, 64, DEV)
    ops.f32_to_act(f.to(DEV), a)
    ops.f32_to_act((2 * f).to(DEV), bb)
    ops.add_act_bf16(a, bb, o, ops.ACT_ELU)
    torch.cuda.synchronize()
    assert (a.value().cpu() - f).abs().max().item() < 2e-6 * f.abs().max().item()
    want = F.elu(a.value().cpu() + bb.value().cpu())
    assert (o.value().cpu() - want).abs().max().item() < 3e-6 * want.abs().max().item()
