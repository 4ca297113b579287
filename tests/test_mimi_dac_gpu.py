"""GPU parity of the Mimi and DAC paths (exact fp32 kernels through the C-ABI) vs the oracle / the reference's golden
outputs: codes bit-exact wherever the top-2 relative gap exceeds 1e-4, waveform max-abs <= 1e-3."""
import pytest
import torch

from helpers import WAVE_MAX_ABS_FP32, code_report, make_input
from oracle import dac_ref, mimi_ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


@pytest.mark.parametrize("case", range(4))
def test_mimi_golden_cases(mimi_sd, mimi_golden, dev, case):
    import audiocodecs_b200 as A
    c = mimi_golden["cases"][case]
    codec = A.Mimi(c["sample_rate"], num_codebooks=c["K"], state_dict=mimi_sd, precision="fp32").eval().to(dev)
    sig = make_input(c["seed"], c["B"], c["T"]).to(dev)
    toks = codec.sig_to_toks(sig)
    ref_toks = c["toks"].long()
    assert toks.dtype == torch.int64 and tuple(toks.shape) == tuple(ref_toks.shape)
    eq = toks.cpu() == ref_toks
    assert eq[~c["near_tie"]].all(), f"{(~eq[~c['near_tie']]).sum().item()} code mismatches away from near-ties"
    rec = codec.toks_to_sig(ref_toks.to(dev))
    assert tuple(rec.shape) == tuple(c["rec"].shape)
    err = (rec.cpu() - c["rec"]).abs().max().item()
    assert err <= WAVE_MAX_ABS_FP32 * max(1.0, c["rec"].abs().max().item()), err
    feats = codec.sig_to_feats(sig)
    assert (feats.cpu() - c["feats"].float()).abs().max().item() < 5e-3  # golden feats are stored in fp16


def test_mimi_transformer_pieces(mimi_sd, dev):
    """attention (RoPE + sliding window, T > window) and LayerNorm vs the oracle on a long sequence."""
    import audiocodecs_b200 as A
    codec = A.Mimi(24000, num_codebooks=8, state_dict=mimi_sd, precision="fp32").eval().to(dev)
    h = torch.randn(2, 300, 512, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        ref = mimi_ref.transformer(mimi_sd, "encoder_transformer", h)
    got = codec._run_transformer(codec._enc_tr, h.to(dev)).cpu()
    assert ((got - ref).norm() / ref.norm()).item() < 1e-5


def test_mimi_qfeats_and_errors(mimi_sd, dev):
    import audiocodecs_b200 as A
    codec = A.Mimi(24000, num_codebooks=8, state_dict=mimi_sd, precision="fp32").eval().to(dev)
    toks = torch.randint(0, 2048, (2, 11, 8), generator=torch.Generator().manual_seed(2))
    ref = mimi_ref.toks_to_qfeats(mimi_sd, toks)
    got = codec.toks_to_qfeats(toks.to(dev)).cpu()
    assert (got - ref).abs().max().item() < 1e-5
    bad = A.Mimi(24000, num_codebooks=33, state_dict=mimi_sd, precision="fp32").eval().to(dev)
    with pytest.raises(ValueError):
        bad.sig_to_toks(torch.zeros(1, 4000, device=dev))


@pytest.mark.parametrize("case", range(3))
def test_dac_golden_cases(dac_sd, dac_golden, dev, case):
    import audiocodecs_b200 as A
    c = dac_golden["cases"][case]
    codec = A.DAC(c["sample_rate"], 44100, num_codebooks=c["K"], state_dict=dac_sd, precision="fp32").eval().to(dev)
    sig = make_input(c["seed"], c["B"], c["T"]).to(dev)
    toks = codec.sig_to_toks(sig)
    ref_toks = c["toks"].long()
    assert toks.dtype == torch.int64 and tuple(toks.shape) == tuple(ref_toks.shape)
    eq = toks.cpu() == ref_toks
    assert eq[~c["near_tie"]].all(), f"{(~eq[~c['near_tie']]).sum().item()} code mismatches away from near-ties"
    rec = codec.toks_to_sig(ref_toks.to(dev))
    assert tuple(rec.shape) == tuple(c["rec"].shape)
    assert (rec.cpu() - c["rec"]).abs().max().item() <= WAVE_MAX_ABS_FP32
    rec2 = codec(sig)
    assert tuple(rec2.shape) == tuple(c["rec"].shape)


def test_dac_rvq_kernels_on_oracle_latents(dac_sd, dev):
    import audiocodecs_b200 as A
    from audiocodecs_b200 import ops
    codec = A.DAC(44100, 44100, num_codebooks=9, state_dict=dac_sd, precision="fp32").eval().to(dev)
    z = torch.randn(3, 1024, 50, generator=torch.Generator().manual_seed(9)) * 0.2
    with torch.no_grad():
        codes, gaps, zq = dac_ref.rvq_encode(dac_sd, z, 9, return_gaps=True)
        dec = dac_ref.from_codes(dac_sd, codes)
    got, got_zq = ops.dac_rvq_encode(z.permute(0, 2, 1).contiguous().to(dev), codec.w_in, codec.b_in, codec.codebooks, codec.w_out,
                                     codec.b_out, 9, want_zq=True)
    m_safe, tie, m_all = code_report(got, codes.permute(0, 2, 1), gaps.permute(0, 2, 1))
    assert m_safe == 1.0, (m_safe, tie, m_all)
    if m_all == 1.0:
        assert (got_zq.cpu() - zq.permute(0, 2, 1)).abs().max().item() < 1e-5
    got_dec = codec.toks_to_qfeats(codes.permute(0, 2, 1).contiguous().to(dev)).cpu()
    assert (got_dec - dec.permute(0, 2, 1)).abs().max().item() < 1e-5
