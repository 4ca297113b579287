"""GPU parity of the DAC bf16 tensor-core pipeline (tcgen05 convs with Snake epilogues) against the fp32 oracle.

Tolerances (BASELINE.json north_star): waveform SI-SNR >= 40 dB in bf16 for the decoder fed the oracle's codes;
the encoder's latent error is bounded and the end-to-end code-match rate is reported (DAC's RVQ runs in fp32 on
the latents; its exact parity on identical latents is tested in test_mimi_dac_gpu.py)."""
import pytest
import torch

from helpers import WAVE_SISNR_BF16_DB, make_input, si_snr_db
from oracle import dac_ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _codec(sd, dev, **kw):
    import audiocodecs_b200 as A
    return A.DAC(44100, 44100, state_dict=sd, precision="bf16", **kw).eval().to(dev)


@pytest.mark.parametrize("B,T", [(2, 22050), (1, 9001), (3, 2048)])
def test_dac_encoder_latents_bf16(dac_sd, dev, B, T):
    codec = _codec(dac_sd, dev, num_codebooks=9)
    sig = make_input(41, B, T)
    with torch.no_grad():
        ref = dac_ref.encoder(dac_sd, sig[:, None]).permute(0, 2, 1)  # [B, N, 1024]
    got = codec.sig_to_feats(sig.to(dev)).cpu()
    assert got.shape == ref.shape and torch.isfinite(got).all()
    rel = ((got - ref).norm() / ref.norm()).item()
    print(f"DAC bf16 encoder latent rel-err {rel:.3e} (B={B}, T={T})")
    assert rel < 2e-2, rel


@pytest.mark.parametrize("B,N,K", [(2, 43, 9), (1, 20, 4), (2, 3, 9)])
def test_dac_decoder_waveform_sisnr_bf16(dac_sd, dev, B, N, K):
    codec = _codec(dac_sd, dev, num_codebooks=K)
    toks = torch.randint(0, 1024, (B, N, K), generator=torch.Generator().manual_seed(6))
    with torch.no_grad():
        ref = dac_ref.toks_to_sig(dac_sd, toks)
    got = codec.toks_to_sig(toks.to(dev)).cpu()
    assert got.shape == ref.shape and torch.isfinite(got).all()
    snr = si_snr_db(ref, got)
    print(f"DAC bf16 decoder SI-SNR {snr:.1f} dB (B={B}, N={N}, K={K})")
    assert snr >= WAVE_SISNR_BF16_DB, snr


def test_dac_end_to_end_code_match_report(dac_sd, dev):
    codec = _codec(dac_sd, dev, num_codebooks=9)
    sig = make_input(998, 2, 44100)
    with torch.no_grad():
        ref = dac_ref.sig_to_toks(dac_sd, sig, 9)
    toks = codec.sig_to_toks(sig.to(dev))
    assert toks.shape == ref.shape and toks.dtype == torch.int64
    per_stage = [(toks.cpu()[..., k] == ref[..., k]).float().mean().item() for k in range(9)]
    print("DAC bf16 end-to-end code match per stage:", [round(x, 4) for x in per_stage])
    assert per_stage[0] > 0.8 and min(per_stage) > 0.4
    rec = codec(sig.to(dev))
    assert tuple(rec.shape) == (2, 44032) and torch.isfinite(rec).all()
