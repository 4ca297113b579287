"""Golden vectors for Mimi from the live reference wrapper (see make_golden.py)."""
import os
import sys

import torch

from . import weights
from .make_golden import REF, make_input

MIMI_CASES = [
    dict(name="b2_2s_k8", sample_rate=24000, K=8, B=2, T=48000, seed=2001),
    dict(name="ragged_k32", sample_rate=24000, K=32, B=1, T=31234, seed=2002),
    dict(name="resample16k_k4", sample_rate=16000, K=4, B=2, T=20000, seed=2003),
    dict(name="tiny_k1", sample_rate=24000, K=1, B=1, T=2500, seed=2004),
]


def golden_mimi():
    sys.path.insert(0, REF)
    from transformers import MimiConfig, MimiModel
    MimiModel.from_pretrained = classmethod(lambda cls, name, **kw: cls(MimiConfig()))
    import audiocodecs
    from . import mimi_ref as ref

    sd = weights.mimi_state_dict(0)
    out = {"cases": []}
    for c in MIMI_CASES:
        codec = audiocodecs.Mimi(c["sample_rate"], num_codebooks=c["K"]).eval()
        missing = codec.model.load_state_dict(sd, strict=False)
        assert not missing.unexpected_keys and all("inv_freq" in k for k in missing.missing_keys), missing
        for m in codec.model.modules():
            if hasattr(m, "_embed"):
                m._embed = None  # cached property (HF/mimi:1188-1195)
        sig = make_input(c["seed"], c["B"], c["T"])
        with torch.no_grad():
            toks = codec.sig_to_toks(sig)
            rec = codec.toks_to_sig(toks)
            feats = codec.sig_to_feats(sig)
            qf = codec.toks_to_qfeats(toks)
            o_toks, gaps, emb = ref.sig_to_toks(sd, sig, c["K"], c["sample_rate"], return_gaps=True)
            o_rec = ref.toks_to_sig(sd, toks, c["sample_rate"])
        safe = gaps > 1e-4
        eq = o_toks == toks
        print(f"mimi/{c['name']}: toks {tuple(toks.shape)} match {eq.float().mean().item():.6f} (gap>1e-4: "
              f"{eq[safe].float().mean().item():.6f}, near-ties {(~safe).float().mean().item():.5f}) rec {tuple(rec.shape)} "
              f"std {rec.std().item():.3f} max|d| {(o_rec - rec).abs().max().item():.3e} feats max|d| "
              f"{(emb.movedim(-1, -2) - feats).abs().max().item():.3e}")
        out["cases"].append(dict(c, toks=toks.contiguous().to(torch.int16), rec=rec.contiguous(), feats=feats.contiguous().half(),
                                 qfeats_sum=qf.double().sum().item(), near_tie=(~safe).contiguous()))
    torch.save(out, os.path.join(weights.GOLDEN_DIR, "mimi_golden.pt"))
