"""GPU parity of the Mimi bf16 tensor-core pipeline (tcgen05 SEANet convs + transformer projections, fp32 residual
stream / LayerNorm / attention) against the fp32 oracle: decoder SI-SNR >= 40 dB on the oracle's codes, bounded
embedding error, end-to-end code-match rate reported."""
import pytest
import torch

from helpers import WAVE_SISNR_BF16_DB, make_input, si_snr_db
from oracle import mimi_ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _codec(sd, dev, **kw):
    import audiocodecs_b200 as A
    return A.Mimi(24000, state_dict=sd, precision="bf16", **kw).eval().to(dev)


@pytest.mark.parametrize("B,T", [(2, 24000), (1, 12345), (3, 1921)])
def test_mimi_embeddings_bf16(mimi_sd, dev, B, T):
    codec = _codec(mimi_sd, dev, num_codebooks=8)
    sig = make_input(51, B, T)
    with torch.no_grad():
        ref = mimi_ref.sig_to_feats(mimi_sd, sig)
    got = codec.sig_to_feats(sig.to(dev)).cpu()
    assert got.shape == ref.shape and torch.isfinite(got).all()
    rel = ((got - ref).norm() / ref.norm()).item()
    print(f"Mimi bf16 embedding rel-err {rel:.3e} (B={B}, T={T})")
    assert rel < 2e-2, rel


@pytest.mark.parametrize("B,N,K", [(2, 13, 8), (1, 7, 32), (2, 2, 1)])
def test_mimi_decoder_waveform_sisnr_bf16(mimi_sd, dev, B, N, K):
    codec = _codec(mimi_sd, dev, num_codebooks=K)
    toks = torch.randint(0, 2048, (B, N, K), generator=torch.Generator().manual_seed(8))
    with torch.no_grad():
        ref = mimi_ref.toks_to_sig(mimi_sd, toks)
    got = codec.toks_to_sig(toks.to(dev)).cpu()
    assert got.shape == ref.shape and torch.isfinite(got).all()
    snr = si_snr_db(ref, got)
    print(f"Mimi bf16 decoder SI-SNR {snr:.1f} dB (B={B}, N={N}, K={K})")
    assert snr >= WAVE_SISNR_BF16_DB, snr


def test_mimi_end_to_end_code_match_report(mimi_sd, dev):
    codec = _codec(mimi_sd, dev, num_codebooks=8)
    sig = make_input(997, 2, 48000)
    with torch.no_grad():
        ref = mimi_ref.sig_to_toks(mimi_sd, sig, 8)
    toks = codec.sig_to_toks(sig.to(dev))
    assert toks.shape == ref.shape and toks.dtype == torch.int64
    per_stage = [(toks.cpu()[..., k] == ref[..., k]).float().mean().item() for k in range(8)]
    print("Mimi bf16 end-to-end code match per stage:", [round(x, 4) for x in per_stage])
    assert per_stage[0] > 0.7 and min(per_stage) > 0.3
    rec = codec(sig.to(dev))
    assert tuple(rec.shape) == (2, 48000) and torch.isfinite(rec).all()
