"""GPU parity of the bf16 tensor-core pipeline (tcgen05 convs) against the fp32 oracle.

Tolerances (BASELINE.json north_star): waveform SI-SNR >= 40 dB in bf16 for the decoder fed the oracle's
codes; RVQ codes are bit-exact on identical embeddings (tested in test_encodec_gpu.py) -- end to end the
bf16 encoder perturbs the embeddings, so here the code-match rate per stage is REPORTED and bounded
from below, and the embedding error is bounded."""
import pytest
import torch

from helpers import WAVE_SISNR_BF16_DB, code_report, make_input, si_snr_db
from oracle import encodec_ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _codec(sd, dev, **kw):
    import audiocodecs_b200 as A
    return A.Encodec(kw.pop("sample_rate", 24000), 24000, state_dict=sd, precision="bf16", **kw).eval().to(dev)


@pytest.mark.parametrize("B,T", [(2, 24000), (1, 12345), (3, 700)])
def test_encoder_embeddings_bf16(encodec_sd, dev, B, T):
    codec = _codec(encodec_sd, dev, num_codebooks=8)
    sig = make_input(31, B, T)
    with torch.no_grad():
        ref = encodec_ref.sig_to_feats(encodec_sd, sig)
    got = codec.sig_to_feats(sig.to(dev)).cpu()
    assert got.shape == ref.shape and torch.isfinite(got).all()
    rel = ((got - ref).norm() / ref.norm()).item()
    print(f"bf16 encoder embedding rel-err {rel:.3e} (B={B}, T={T})")
    assert rel < 2e-2, rel


@pytest.mark.parametrize("B,N,K", [(2, 75, 8), (1, 39, 32), (2, 3, 2)])
def test_decoder_waveform_sisnr_bf16(encodec_sd, dev, B, N, K):
    """Protocol (2): decoder fed the same codes as the oracle."""
    codec = _codec(encodec_sd, dev, num_codebooks=K)
    toks = torch.randint(0, 1024, (B, N, K), generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        ref = encodec_ref.toks_to_sig(encodec_sd, toks)
    got = codec.toks_to_sig(toks.to(dev)).cpu()
    assert got.shape == ref.shape and torch.isfinite(got).all()
    snr = si_snr_db(ref, got)
    print(f"bf16 decoder SI-SNR {snr:.1f} dB (B={B}, N={N}, K={K})")
    assert snr >= WAVE_SISNR_BF16_DB, snr


def test_end_to_end_code_match_report(encodec_sd, dev):
    codec = _codec(encodec_sd, dev, num_codebooks=8)
    sig = make_input(999, 4, 48000)
    with torch.no_grad():
        ref, gaps, _ = encodec_ref.sig_to_toks(encodec_sd, sig, 8, return_gaps=True)
    toks = codec.sig_to_toks(sig.to(dev))
    assert toks.shape == ref.shape and toks.dtype == torch.int64
    per_stage = [(toks.cpu()[..., k] == ref[..., k]).float().mean().item() for k in range(8)]
    m_safe, tie, m_all = code_report(toks, ref, gaps)
    print("bf16 end-to-end code match per stage:", [round(x, 4) for x in per_stage], f"safe {m_safe:.4f} all {m_all:.4f} near-ties {tie:.4f}")
    # measured (and reproduced by scripts/exact_mode_emulation.py): 0.985 at stage 0 decaying to 0.924 at stage 7, 0.951 overall
    assert per_stage[0] > 0.965 and min(per_stage) > 0.90 and m_safe > 0.93
    rec = codec.toks_to_sig(toks)
    assert tuple(rec.shape) == (4, 48000) and torch.isfinite(rec).all()


@pytest.mark.parametrize("op_dtype,tol", [(torch.float16, 3e-3), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("B,T", [(16, 40), (5, 33), (37, 12), (70, 9), (1, 25), (2, 1), (3, 7), (130, 5)])   # lone clip, one step, odd groups, several waves
def test_lstm_tc_cluster_kernel(encodec_sd, dev, B, T, op_dtype, tol):
    """tcgen05 cluster LSTM (fp16 -- the shipped configuration -- or bf16 W_hh and h operands, fp32 accumulate / cell state)
    vs the oracle's explicit loop."""
    from audiocodecs_b200 import ops
    from audiocodecs_b200.tc import Act
    g = torch.Generator().manual_seed(77)
    C = 512
    pre = torch.randn(B, T, 4 * C, generator=g)
    w_hh = encodec_sd["encoder.layers.13.lstm.weight_hh_l0"]
    skip = torch.randn(B, T, C, generator=g)
    # oracle: explicit recurrence on the given pre-gates
    h = torch.zeros(B, C); c = torch.zeros(B, C); outs = []
    for t in range(T):
        gates = pre[:, t] + h @ w_hh.t()
        i, f, gg, o = gates.split(C, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        outs.append(h)
    ref = torch.stack(outs, 1)
    out = Act(B, T, C, dev, split=True)
    sk = Act(B, T, C, dev, split=True)
    sk.buf.copy_(skip.to(torch.bfloat16)); sk.lo.copy_((skip - skip.to(torch.bfloat16).float()).to(torch.bfloat16))
    fin = Act(B, T, C, dev, hl=3, split=True)
    ops.lstm_tc(pre.to(dev), w_hh.to(op_dtype).to(dev), out=out, skip=sk, final=fin, final_act=ops.ACT_ELU)
    torch.cuda.synchronize()
    got = out.buf.float().cpu() + out.lo.float().cpu()
    assert torch.isfinite(got).all()
    err = (got - ref).abs().max().item()
    print(f"lstm_tc {op_dtype} max|err| {err:.2e} (B={B}, T={T})")
    assert err < tol, err
    ref_fin = torch.nn.functional.elu(ref + skip)
    got_fin = (fin.data().float() + fin.lo[:, 3:3 + T].float()).cpu()
    assert (got_fin - ref_fin).abs().max().item() < 1.5 * tol


def test_host_pipeline_matches_direct_calls(encodec_sd, dev):
    """HostPipeline (copies on their own streams, overlapped with compute) returns exactly what direct calls return"""
    from audiocodecs_b200.hostpipe import HostPipeline
    codec = _codec(encodec_sd, dev, num_codebooks=8)
    batches = [make_input(60 + i, 3, 9600).pin_memory() for i in range(5)]
    direct = [codec.toks_to_sig(codec.sig_to_toks(b.to(dev))).cpu() for b in batches]
    pipe = HostPipeline(codec)
    results = [pipe.submit(b) for b in batches]
    pipe.drain()
    for (out, ev), ref in zip(results, direct):
        ev.synchronize()
        assert torch.equal(out, ref)


def test_cuda_graph_replay_matches_eager(encodec_sd, dev):
    """GraphedCodec: the captured tokenize / detokenize graphs give the eager calls' results bit for bit, also on new inputs
    of the captured shape, and refuse other shapes."""
    import audiocodecs_b200 as A
    codec = A.Encodec(24000, 24000, num_codebooks=8, state_dict=encodec_sd, precision="bf16").eval().to(dev)
    sig = make_input(5, 2, 9600).to(dev)
    g = A.GraphedCodec(codec, sig)
    for seed in (5, 6, 7):
        s = make_input(seed, 2, 9600).to(dev)
        toks = codec.sig_to_toks(s)
        rec = codec.toks_to_sig(toks)
        assert torch.equal(g.sig_to_toks(s), toks)
        assert torch.equal(g.toks_to_sig(toks), rec)
        assert torch.equal(g.reconstruct(s), rec)
    with pytest.raises(ValueError):
        g.sig_to_toks(make_input(1, 1, 9600).to(dev))
