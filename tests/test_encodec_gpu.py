"""GPU parity: the CUDA path (through the C-ABI) vs the oracle and the reference's golden outputs."""
import pytest
import torch

from helpers import NEAR_TIE_REL_GAP, WAVE_MAX_ABS_FP32, code_report, make_input
from oracle import encodec_ref, resample_ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    return torch.device("cuda:0")


def _codec(sd, dev, **kw):
    import audiocodecs_b200 as A
    kw.setdefault("precision", "fp32")  # exact-parity path; the bf16 tensor path is tested in test_encodec_bf16_gpu.py
    return A.Encodec(kw.pop("sample_rate", 24000), 24000, state_dict=sd, **kw).eval().to(dev)


@pytest.mark.parametrize("case", range(5))
def test_golden_cases(encodec_sd, encodec_golden, dev, case):
    """Replays every recorded reference case: codes bit-exact away from near-ties, waveform within 1e-3."""
    c = encodec_golden["cases"][case]
    codec = _codec(encodec_sd, dev, sample_rate=c["sample_rate"], num_codebooks=c["K"])
    sig = make_input(c["seed"], c["B"], c["T"]).to(dev)
    length = None if c["length"] is None else torch.tensor(c["length"], device=dev)
    toks = codec.sig_to_toks(sig, length)
    ref_toks = c["toks"].long()
    assert toks.dtype == torch.int64 and tuple(toks.shape) == tuple(ref_toks.shape)
    safe = ~c["near_tie"]
    eq = toks.cpu() == ref_toks
    assert eq[safe].all(), f"{(~eq[safe]).sum().item()} code mismatches away from near-ties"
    rec = codec.toks_to_sig(ref_toks.to(dev), length)
    assert tuple(rec.shape) == tuple(c["rec"].shape)
    err = (rec.cpu() - c["rec"]).abs().max().item()
    assert err <= WAVE_MAX_ABS_FP32 * max(1.0, c["rec"].abs().max().item()), err
    rec2 = codec(sig, length)  # mode="reconstruct" forward, R/codec.py:45-55
    assert tuple(rec2.shape) == tuple(c["rec"].shape)


def test_rvq_encode_on_oracle_embeddings(encodec_sd, dev):
    """Protocol (1) of SURVEY 8c: RVQ kernel fed the oracle's fp32 embeddings, all 32 stages."""
    from audiocodecs_b200 import ops
    codec = _codec(encodec_sd, dev, num_codebooks=32)
    sig = make_input(7, 3, 24000)
    with torch.no_grad():
        emb = encodec_ref.encoder(encodec_sd, sig[:, None])
        codes, gaps = encodec_ref.rvq_encode(encodec_sd, emb, 32, return_gaps=True)
    x = emb.permute(0, 2, 1).contiguous().to(dev)  # [B,N,128]
    B, N, D = x.shape
    out = torch.empty((B, N, 32), device=dev, dtype=torch.int64)
    ops.rvq_encode(x.view(B * N, D), codec.codebooks, codec.cb_norm, out.view(B * N, 32), 32)
    m_safe, tie, m_all = code_report(out, codes.permute(1, 2, 0), gaps.permute(1, 2, 0))
    assert m_safe == 1.0, (m_safe, tie, m_all)
    assert tie < 0.02


@pytest.mark.parametrize("B,T", [(3, 24000), (5, 31337), (1, 200)])
def test_rvq_encode_tensor_core_on_oracle_embeddings(encodec_sd, dev, B, T):
    """The tcgen05 RVQ kernel (split-bf16 distance GEMM + exact fp32 re-score of the top 2) fed the oracle's fp32
    embeddings, all 32 stages: codes identical to the reference wherever its top-2 gap exceeds 1e-4; ragged tiles."""
    from audiocodecs_b200 import ops
    codec = _codec(encodec_sd, dev, num_codebooks=32, precision="bf16")
    sig = make_input(17 + B, B, T)
    with torch.no_grad():
        emb = encodec_ref.encoder(encodec_sd, sig[:, None])
        codes, gaps = encodec_ref.rvq_encode(encodec_sd, emb, 32, return_gaps=True)
    x = emb.permute(0, 2, 1).contiguous().to(dev)  # [B,N,128]
    B, N, D = x.shape
    out = torch.full((B, N, 32), -1, device=dev, dtype=torch.int64)
    res = torch.empty((B * N, D), device=dev, dtype=torch.float32)
    ops.rvq_encode_tc(x.view(B * N, D), codec.cb_split, codec.codebooks, codec.cb_norm, out.view(B * N, 32), 32, residual_out=res)
    m_safe, tie, m_all = code_report(out, codes.permute(1, 2, 0), gaps.permute(1, 2, 0))
    assert m_safe == 1.0, (m_safe, tie, m_all)
    # same decisions as the fp32 SIMT kernel, near-ties included (both re-score in fp32), and the same final residual
    out2 = torch.empty_like(out)
    res2 = torch.empty_like(res)
    ops.rvq_encode(x.view(B * N, D), codec.codebooks, codec.cb_norm, out2.view(B * N, 32), 32, residual_out=res2)
    agree = (out == out2).float().mean().item()
    assert agree > 0.995, agree
    same_rows = (out == out2).all(-1).view(-1)
    assert torch.allclose(res[same_rows], res2[same_rows], atol=1e-6)


def test_rvq_decode_bit_exact(encodec_sd, dev):
    codec = _codec(encodec_sd, dev, num_codebooks=32)
    toks = torch.randint(0, 1024, (2, 77, 32), generator=torch.Generator().manual_seed(1))
    ref = encodec_ref.toks_to_qfeats(encodec_sd, toks)
    got = codec.toks_to_qfeats(toks.to(dev)).cpu()
    assert torch.equal(got, ref)  # same fp32 adds in the same order


def test_encoder_embeddings_and_lstm(encodec_sd, dev):
    codec = _codec(encodec_sd, dev, num_codebooks=8)
    sig = make_input(11, 2, 16000)
    with torch.no_grad():
        ref = encodec_ref.sig_to_feats(encodec_sd, sig)
    got = codec.sig_to_feats(sig.to(dev)).cpu()
    assert got.shape == ref.shape
    rel = (got - ref).norm() / ref.norm()
    assert rel < 1e-5, rel


def test_batch_ragged_and_large_batch_lstm_waves(encodec_sd, dev):
    """B=70 > one LSTM wave (64 clips) and not a multiple of the 16-clip slice; batch invariance vs B=1."""
    codec = _codec(encodec_sd, dev, num_codebooks=8)
    sig = make_input(21, 70, 3200).to(dev)
    toks = codec.sig_to_toks(sig)
    one = codec.sig_to_toks(sig[69:70])
    assert (toks[69:70] == one).float().mean() > 0.99
    with torch.no_grad():
        ref, gaps, _ = encodec_ref.sig_to_toks(encodec_sd, sig[64:].cpu(), 8, return_gaps=True)
    m_safe, tie, m_all = code_report(toks[64:], ref, gaps)
    assert m_safe >= 0.995, (m_safe, tie, m_all)


@pytest.mark.parametrize("o,n", [(16000, 24000), (24000, 16000), (16000, 44100), (44100, 16000)])
def test_resample_kernel(dev, o, n):
    from audiocodecs_b200 import ops
    x = make_input(5, 3, 4001)
    ref = resample_ref.resample(x, o, n)
    got = ops.resample(x.to(dev), o, n).cpu()
    assert got.shape == ref.shape
    assert (got - ref).abs().max() < 2e-6


def test_full_size_roundtrip_properties(encodec_sd, dev):
    """BASELINE config[1] size (64 x 10 s): size-independent checks -- decode(encode(x)) is deterministic,
    tokens in range, re-encoding a clip alone reproduces its row (batch invariance), and
    decode is linear-free but shift-consistent: tokens of a clip do not depend on its batch neighbours."""
    codec = _codec(encodec_sd, dev, num_codebooks=8)
    sig = make_input(999, 64, 240000).to(dev)
    toks = codec.sig_to_toks(sig)
    assert tuple(toks.shape) == (64, 750, 8) and int(toks.min()) >= 0 and int(toks.max()) < 1024
    again = codec.sig_to_toks(sig)
    assert torch.equal(toks, again)
    solo = codec.sig_to_toks(sig[5:6])
    assert (solo == toks[5:6]).float().mean().item() > 0.995
    rec = codec.toks_to_sig(toks)
    assert tuple(rec.shape) == (64, 240000) and torch.isfinite(rec).all()
    # oracle on one full-length clip (a few seconds of CPU)
    with torch.no_grad():
        ref_toks, gaps, _ = encodec_ref.sig_to_toks(encodec_sd, sig[5:6].cpu(), 8, return_gaps=True)
        ref_rec = encodec_ref.toks_to_sig(encodec_sd, toks[5:6].cpu())
    m_safe, tie, m_all = code_report(toks[5:6], ref_toks, gaps)
    assert m_safe >= 0.999, (m_safe, tie, m_all)
    assert (rec[5:6].cpu() - ref_rec).abs().max().item() <= WAVE_MAX_ABS_FP32


def test_sub_batch_chunking_is_invisible(encodec_sd, dev):
    """a batch larger than `max_chunk_samples` is processed in sub-batches: identical tokens and waveforms"""
    from audiocodecs_b200 import tc
    tc.TUNE, saved = False, tc.TUNE  # per-shape autotuning may pick different (equally valid) tilings for different batch sizes
    tc._TUNED.clear()
    codec = _codec(encodec_sd, dev, num_codebooks=8, precision="bf16")
    sig = make_input(21, 7, 6400).to(dev)
    toks = codec.sig_to_toks(sig)
    rec = codec.toks_to_sig(toks)
    codec.max_chunk_samples = 3 * 6400  # 3 clips per sub-batch -> chunks of 3, 3, 1
    toks2 = codec.sig_to_toks(sig)
    rec2 = codec.toks_to_sig(toks)
    tc.TUNE = saved
    tc._TUNED.clear()
    assert torch.equal(toks, toks2) and torch.equal(rec, rec2)
