"""GPU tests of the secondary Codec API: `sig_to_qfeats`, DAC `latent=True` features, Mimi's projected
embeddings `embs(latent=False)` -- on the exact-parity fp32 path and on the default tensor path, against the oracle / the
summaries recorded from the live reference (tests/golden/embs_golden.pt)."""
import os

import pytest
import torch
import torch.nn.functional as F

from helpers import make_input
from oracle import dac_ref, encodec_ref, mimi_ref

pytestmark = pytest.mark.gpu
GOLD = torch.load(os.path.join(os.path.dirname(__file__), "golden", "embs_golden.pt"))


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


@pytest.mark.parametrize("precision,tol", [("fp32", 5e-6), ("exact", 1e-4)])
def test_encodec_sig_to_qfeats(encodec_sd, dev, precision, tol):
    """R/audiocodecs/encodec.py:120-127: qfeats = quantizer.decode(encode(sig)) [B, N, 128]"""
    import audiocodecs_b200 as A
    codec = A.Encodec(24000, 24000, num_codebooks=8, state_dict=encodec_sd, precision=precision).eval().to(dev)
    sig = make_input(11, 2, 16000)
    with torch.no_grad():
        toks = encodec_ref.sig_to_toks(encodec_sd, sig, 8)
        ref = encodec_ref.toks_to_qfeats(encodec_sd, toks)
    got = codec.sig_to_qfeats(sig.to(dev)).cpu()
    assert got.shape == ref.shape
    same = (codec.sig_to_toks(sig.to(dev)).cpu() == toks).all(-1)   # frames whose tokens all agree (near-ties may differ)
    assert same.float().mean().item() > 0.98 and _rel(got[same], ref[same]) < tol
    feats = codec.sig_to_feats(sig.to(dev)).cpu()
    assert _rel(feats, encodec_ref.sig_to_feats(encodec_sd, sig)) < (1e-4 if precision == "exact" else 5e-6)


@pytest.mark.parametrize("precision,tol", [("fp32", 5e-6), ("exact", 1e-4)])
def test_dac_latent_feats_and_qfeats(dac_sd, dev, precision, tol):
    """R/audiocodecs/dac.py:103-121: latent=True features = quantizers[0].in_proj(encoder(sig)) [B, N, 8]; qfeats = the
    quantised sum of the encode call [B, N, 1024]."""
    import audiocodecs_b200 as A
    sig = make_input(12, 2, 22050)
    with torch.no_grad():
        z = dac_ref.encoder(dac_sd, sig[:, None])
        ref_lat = F.conv1d(z, dac_sd["quantizer.quantizers.0.in_proj.weight"], dac_sd["quantizer.quantizers.0.in_proj.bias"]).movedim(-1, -2)
        codes, _, zq = dac_ref.rvq_encode(dac_sd, z, 9, return_gaps=True)
    lat = A.DAC(44100, 44100, num_codebooks=9, latent=True, state_dict=dac_sd, precision=precision).eval().to(dev)
    got = lat.sig_to_feats(sig.to(dev)).cpu()
    assert tuple(got.shape) == tuple(ref_lat.shape) and got.shape[-1] == 8 and _rel(got, ref_lat) < 10 * tol
    full = A.DAC(44100, 44100, num_codebooks=9, state_dict=dac_sd, precision=precision).eval().to(dev)
    assert _rel(full.sig_to_feats(sig.to(dev)).cpu(), z.movedim(-1, -2)) < tol
    q = full.sig_to_qfeats(sig.to(dev)).cpu()
    same = (full.sig_to_toks(sig.to(dev)).cpu() == codes.movedim(-1, -2)).all(-1)
    assert same.float().mean().item() > 0.95 and _rel(q[same], zq.movedim(-1, -2)[same]) < 10 * tol


@pytest.mark.parametrize("precision", ["fp32", "exact"])
def test_mimi_projected_embs_and_qfeats(mimi_sd, dev, precision):
    """R/audiocodecs/mimi.py:55-90: embs(latent=False) = output_proj of every code vector [K, 2048, 512] (golden summary
    recorded from the live wrapper); sig_to_qfeats = toks_to_qfeats(sig_to_toks)."""
    import audiocodecs_b200 as A
    codec = A.Mimi(24000, num_codebooks=3, latent=False, state_dict=mimi_sd, precision=precision).eval().to(dev)
    e = codec.embs()
    g = GOLD["mimi_k3_latent0"]
    assert tuple(e.shape) == g["shape"]
    got = e[tuple(g["idx"].t().to(e.device))].float().cpu()
    assert (got - g["vals"]).abs().max().item() <= 1e-5 * max(1.0, g["vals"].abs().max().item())
    assert abs(e.double().sum().item() - g["checksum"]) <= 1e-5 * g["abs_sum"]
    sig = make_input(13, 2, 30000)
    with torch.no_grad():
        toks = mimi_ref.sig_to_toks(mimi_sd, sig, 3)
        ref = mimi_ref.toks_to_qfeats(mimi_sd, toks)
    q = codec.sig_to_qfeats(sig.to(dev)).cpu()
    same = (codec.sig_to_toks(sig.to(dev)).cpu() == toks).all(-1)
    assert same.float().mean().item() > 0.9 and _rel(q[same], ref[same]) < 1e-5
