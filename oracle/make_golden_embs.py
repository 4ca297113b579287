"""Golden summaries of `embs()` (R/audiocodecs/encodec.py:74-79, dac.py:66-91, mimi.py:55-90) from the LIVE reference
wrappers in the authoring container -- shape, fp64 checksum and 64 sampled entries per configuration -- replayed by
tests/test_embs_golden.py against this repo's `embs()`.

Usage: PYTHONPATH=/root/repo python -m oracle.make_golden_embs
"""
import os
import sys

import torch

from . import weights
from .make_golden import REF
from .make_golden_dac import install_dac_shim


def summarize(e, seed):
    g = torch.Generator().manual_seed(seed)
    idx = torch.stack([torch.randint(0, s, (64,), generator=g) for s in e.shape], dim=1)
    return dict(shape=tuple(e.shape), checksum=e.double().sum().item(), abs_sum=e.double().abs().sum().item(), idx=idx,
                vals=e[tuple(idx.t())].float().clone())


def main():
    sys.path.insert(0, REF)
    from transformers import EncodecConfig, EncodecModel, MimiConfig, MimiModel
    EncodecModel.from_pretrained = classmethod(lambda cls, name, **kw: cls(EncodecConfig()))
    MimiModel.from_pretrained = classmethod(lambda cls, name, **kw: cls(MimiConfig()))
    dac_sd = weights.dac_state_dict(0)
    install_dac_shim(dac_sd)
    import audiocodecs
    out = {}
    with torch.no_grad():
        c = audiocodecs.Encodec(24000, 24000, num_codebooks=4).eval()
        c.model.load_state_dict(weights.encodec_state_dict(0), strict=True)
        out["encodec_k4"] = summarize(c.embs(), 1)
        for latent in (False, True):
            c = audiocodecs.DAC(44100, 44100, num_codebooks=3, latent=latent).eval()
            out[f"dac_k3_latent{int(latent)}"] = summarize(c.embs(), 2)
        for latent in (True, False):
            c = audiocodecs.Mimi(24000, num_codebooks=3, latent=latent).eval()
            c.model.load_state_dict(weights.mimi_state_dict(0), strict=False)
            for m in c.model.modules():
                if hasattr(m, "_embed"):
                    m._embed = None
            out[f"mimi_k3_latent{int(latent)}"] = summarize(c.embs(), 3)
    for k, v in out.items():
        print(k, v["shape"], v["checksum"])
    torch.save(out, os.path.join(weights.GOLDEN_DIR, "embs_golden.pt"))


if __name__ == "__main__":
    main()
