"""Golden vectors for the token-augmentation utilities of the Codec interface (R/audiocodecs/codec.py:121-180), generated
from the LIVE reference in the authoring container: the unmodified `audiocodecs.Encodec` (random-init default architecture
+ our deterministic state dict, as oracle/make_golden.py) answers `logits()` and `resample(...)` under fixed torch seeds.
Recorded: a checksum and sampled entries of the logits, and the resampled tokens for plain / top-k / top-p sampling.
tests/test_token_augmentation.py replays them on any machine (same torch RNG call order => same draws).

Usage: PYTHONPATH=/root/repo python -m oracle.make_golden_augment
"""
import os
import sys

import torch

from . import weights

REF = "/root/reference"
CASES = [dict(p=0.35, temp=1.0, top_k=None, top_p=None, seed=11), dict(p=1.0, temp=0.7, top_k=5, top_p=None, seed=12),
         dict(p=0.5, temp=1.3, top_k=None, top_p=0.8, seed=13)]


def main():
    sys.path.insert(0, REF)
    from transformers import EncodecConfig, EncodecModel
    EncodecModel.from_pretrained = classmethod(lambda cls, name, **kw: cls(EncodecConfig()))
    import audiocodecs
    codec = audiocodecs.Encodec(24000, 24000, num_codebooks=4).eval()
    codec.model.load_state_dict(weights.encodec_state_dict(0), strict=True)
    lg = codec.logits()
    toks = torch.randint(0, 1024, (2, 9, 4), generator=torch.Generator().manual_seed(5))
    idx = torch.randint(0, 1024, (64, 2), generator=torch.Generator().manual_seed(6))
    out = {"K": 4, "toks": toks, "logit_idx": idx, "logit_vals": torch.stack([lg[k, idx[:, 0], idx[:, 1]] for k in range(4)]),
           "logit_finite_sum": lg[torch.isfinite(lg)].double().sum().item(), "cases": []}
    for c in CASES:
        torch.manual_seed(c["seed"])
        res = codec.resample(toks, p=c["p"], temp=c["temp"], top_k=c["top_k"], top_p=c["top_p"])
        out["cases"].append(dict(c, out=res.clone()))
        print(c, "changed", (res != toks).float().mean().item())
    torch.save(out, os.path.join(weights.GOLDEN_DIR, "augment_golden.pt"))


if __name__ == "__main__":
    main()
