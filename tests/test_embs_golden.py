"""`embs()` of the three wrappers against summaries recorded from the LIVE reference wrappers (oracle/make_golden_embs.py:
shape, fp64 checksum, 64 sampled entries per configuration).  The configurations whose `embs()` is pure tensor algebra run
on the CPU; Mimi's projected embeddings (latent=False, golden key `mimi_k3_latent0`) go through the conv kernels and are
checked on the GPU (tests/test_feats_gpu.py)."""
import os

import pytest
import torch

GOLD = torch.load(os.path.join(os.path.dirname(__file__), "golden", "embs_golden.pt"))


def _check(e, g):
    assert tuple(e.shape) == g["shape"]
    got = e[tuple(g["idx"].t().to(e.device))].float().cpu()
    assert (got - g["vals"]).abs().max().item() <= 1e-5 * max(1.0, g["vals"].abs().max().item())
    assert abs(e.double().sum().item() - g["checksum"]) <= 1e-5 * g["abs_sum"]


def test_encodec_embs(encodec_sd):
    import audiocodecs_b200 as A
    _check(A.Encodec(24000, 24000, num_codebooks=4, state_dict=encodec_sd).embs(), GOLD["encodec_k4"])


@pytest.mark.parametrize("latent", [False, True])
def test_dac_embs(dac_sd, latent):
    import audiocodecs_b200 as A
    _check(A.DAC(44100, 44100, num_codebooks=3, latent=latent, state_dict=dac_sd).embs(), GOLD[f"dac_k3_latent{int(latent)}"])


def test_mimi_embs_latent(mimi_sd):
    import audiocodecs_b200 as A
    _check(A.Mimi(24000, num_codebooks=3, latent=True, state_dict=mimi_sd).embs(), GOLD["mimi_k3_latent1"])
