"""Host-side token augmentation of the Codec interface (R/audiocodecs/codec.py:121-180: `logits`, `resample`, top-k / top-p
sampling) -- properties that pin the reference's semantics without its RNG stream.  CPU only."""
import pytest
import torch


@pytest.fixture(scope="module")
def codec(encodec_sd):
    import audiocodecs_b200 as A
    return A.Encodec(24000, 24000, num_codebooks=4, state_dict=encodec_sd, precision="fp32").eval()


def test_logits_are_negative_pairwise_distances(codec):
    lg = codec.logits()
    e = codec.embs()
    K, C, _ = e.shape
    assert tuple(lg.shape) == (K, C, C)
    assert torch.isinf(lg.diagonal(dim1=-2, dim2=-1)).all() and (lg.diagonal(dim1=-2, dim2=-1) < 0).all()
    i, j = 3, 777
    assert abs(lg[1, i, j].item() + (e[1, i] - e[1, j]).norm().item()) < 1e-3
    lg[0, 0, 1] = 123.0                       # a copy is returned: the cache is untouched
    assert codec.logits()[0, 0, 1] != 123.0


def test_resample_semantics(codec):
    torch.manual_seed(0)
    toks = torch.randint(0, 1024, (3, 17, 4))
    before = toks.clone()
    assert codec.resample(toks, p=0.0) is toks
    all_new = codec.resample(toks, p=1.0)
    assert all_new.shape == toks.shape and all_new.dtype == toks.dtype and (all_new != toks).all()   # never resamples to itself
    some = codec.resample(toks, p=0.3)
    frac = (some != toks).float().mean().item()
    assert 0.15 < frac < 0.45
    assert torch.equal(toks, before)          # the input is never modified
    # top_k = 1: the nearest other code of the same codebook, deterministically
    nearest = codec.resample(toks, p=1.0, top_k=1)
    e = codec.embs()
    for b, n, k in [(0, 0, 0), (2, 16, 3), (1, 5, 2)]:
        d = (e[k] - e[k, toks[b, n, k]]).norm(dim=-1)
        d[toks[b, n, k]] = float("inf")
        assert nearest[b, n, k].item() == d.argmin().item()
    # a tiny nucleus keeps only the most likely code -> same as top_k = 1
    assert (codec.resample(toks, p=1.0, top_p=1e-6) == nearest).all()
    with pytest.raises(NotImplementedError):
        codec.resample(toks, p=0.5, top_k=2, top_p=0.5)
    with pytest.raises(NotImplementedError):
        codec.feats_to_sig(torch.zeros(1, 2, 128))


def test_augmentation_matches_reference_golden(codec):
    """tests/golden/augment_golden.pt (oracle/make_golden_augment.py: the live reference's `logits()` / `resample()` under
    fixed torch seeds): same logits, and -- the torch RNG being consumed in the same order -- the same resampled tokens."""
    import os
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "augment_golden.pt"))
    lg = codec.logits()
    idx = g["logit_idx"]
    got = torch.stack([lg[k, idx[:, 0], idx[:, 1]] for k in range(g["K"])])
    finite = torch.isfinite(g["logit_vals"])
    assert torch.equal(torch.isfinite(got), finite)
    assert (got[finite] - g["logit_vals"][finite]).abs().max().item() < 1e-4
    assert abs(lg[torch.isfinite(lg)].double().sum().item() - g["logit_finite_sum"]) < 1e-6 * abs(g["logit_finite_sum"])
    for c in g["cases"]:
        torch.manual_seed(c["seed"])
        res = codec.resample(g["toks"], p=c["p"], temp=c["temp"], top_k=c["top_k"], top_p=c["top_p"])
        assert torch.equal(res, c["out"]), c
