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


def test_mimi_rvq_tensor_core_matches_fp32_kernel(mimi_sd, dev):
    """Mimi's RVQ (dim 256, 2048 codes, cdist metric, semantic + acoustic chains) on the tcgen05 kernel makes the same
    decisions as the exact fp32 SIMT kernel on the same projected embeddings (both re-score their candidates in fp32)."""
    from audiocodecs_b200 import ops
    codec = _codec(mimi_sd, dev, num_codebooks=32)
    sig = make_input(77, 3, 30000).to(dev)
    emb = codec.sig_to_feats(sig)                      # [B, N, 512]
    B, N, _ = emb.shape
    xa = ops.conv(codec._acoustic_in, emb).view(B * N, -1).contiguous()   # [rows, 256]
    t_tc = torch.full((B * N, 31), -1, device=dev, dtype=torch.int64)
    t_f32 = torch.empty_like(t_tc)
    r_tc = torch.empty_like(xa)
    r_f32 = torch.empty_like(xa)
    ops.rvq_encode_tc(xa, codec.cb_split, codec.codebooks, codec.cb_norm, t_tc, 31, stage0=1, metric=1, residual_out=r_tc)
    ops.rvq_encode(xa, codec.codebooks[1:], codec.cb_norm[1:], t_f32, 31, metric=1, residual_out=r_f32)
    agree = (t_tc == t_f32).float().mean().item()
    print(f"Mimi RVQ tensor-core vs fp32 kernel: {agree:.5f} of {t_tc.numel()} decisions agree")
    assert int(t_tc.min()) >= 0 and int(t_tc.max()) < 2048
    assert agree > 0.995, agree
    same = (t_tc == t_f32).all(-1)
    assert torch.allclose(r_tc[same], r_f32[same], atol=1e-5)
