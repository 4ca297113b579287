"""Token consumers: CodebookUtil and MultiHeadEmbedding against the reference's arithmetic restated with torch ops
(R/downstream/metrics/codebook_util.py:39-83, R/downstream/models/multihead.py:28-69)."""
import math

import pytest
import torch


def _ref_codebook_util(toks_list, K, V):
    """the reference's append / summarize, utterance by utterance"""
    counts = [torch.zeros(V) for _ in range(K)]
    total = 0
    for t in toks_list:
        for k in range(K):
            idx, c = t[..., k].unique(return_counts=True)
            counts[k][idx] += c
        total += t.shape[:2].numel()
    utils, ents = [], []
    for c in counts:
        p = c / total
        valid = p > 0
        pv = p[valid]
        ent = -(pv * pv.log2()).sum()
        n = valid.sum()
        utils.append(n / V if n > 1 else 0)
        ents.append(ent / math.log2(n) if n > 1 else 0.0)
    return {"codebook_util": round(100 * torch.tensor(sum(utils) / K).item(), 2), "norm_entropy": round(100 * torch.tensor(sum(ents) / K).item(), 2)}


def test_constructor_contract_cpu():
    """constructor / state-dict surface of the reference classes, checked without a GPU"""
    import audiocodecs_b200 as A
    m = A.MultiHeadEmbedding(1024, 64, 8, padding_idx=True)
    assert m.weight.shape == (8 * 1024 + 1, 64) and m.padding_idx == 8192 and m.offsets.tolist() == [1024 * k for k in range(8)]
    m2 = A.MultiHeadEmbedding([10, 20, 30], 16, 3)
    assert m2.weight.shape == (60, 16) and m2.offsets.tolist() == [0, 10, 30]
    u = A.CodebookUtil(4, 256)
    assert u.vocab_sizes == [256] * 4 and u.total_toks == 0 and len(u.toks_count_per_codebook) == 4
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 2, 8, dtype=torch.long))          # no CPU fallback
    with pytest.raises(AssertionError):
        u.append(torch.zeros(2, 5, 4, dtype=torch.long))   # "Batch size must be 1", as the reference


@pytest.mark.gpu
def test_codebook_util_matches_reference_arithmetic():
    import audiocodecs_b200 as A
    g = torch.Generator().manual_seed(3)
    K, V = 8, 1024
    utts = [torch.randint(0, V, (1, n, K), generator=g) // (1 + torch.arange(K)) for n in (750, 13, 2000)]   # skewed usage per codebook
    u = A.CodebookUtil(K, V)
    for t in utts:
        u.append(t.cuda())
    ref = _ref_codebook_util(utts, K, V)
    assert u.summarize() == ref and u.summarize("norm_entropy") == ref["norm_entropy"]
    big = torch.randint(0, V, (64, 750, K), generator=g)
    ub = A.CodebookUtil(K, V)
    ub.append_batch(big.cuda())
    assert ub.summarize() == _ref_codebook_util([big], K, V)
    counts = torch.stack([torch.bincount(big[..., k].flatten(), minlength=V) for k in range(K)])
    assert torch.equal(ub._counts.cpu(), counts)
    bad = A.CodebookUtil(K, V)
    bad.append(torch.full((1, 4, K), V).cuda())
    with pytest.raises(IndexError):
        bad.summarize()


@pytest.mark.gpu
@pytest.mark.parametrize("padding", [False, True])
def test_multihead_embedding_matches_torch(padding):
    import audiocodecs_b200 as A
    torch.manual_seed(0)
    K, V, D = 8, 1024, 128
    m = A.MultiHeadEmbedding(V, D, K, padding_idx=padding).cuda()
    toks = torch.randint(0, V, (3, 50, K), generator=torch.Generator().manual_seed(1))
    if padding:
        toks[0, :5] = V       # padding tokens
    # the reference's forward, restated
    idx = toks + m.offsets
    if padding:
        idx[toks == V] = m.padding_idx
    ref = torch.nn.functional.embedding(idx.cuda(), m.weight, m.padding_idx)
    out = m(toks.cuda())
    assert out.shape == (3, 50, K, D) and torch.equal(out, ref)
    out.sum().backward()
    gw = m.weight.grad.clone()
    m.weight.grad = None
    torch.nn.functional.embedding(idx.cuda(), m.weight, m.padding_idx).sum().backward()
    assert torch.allclose(gw, m.weight.grad)
