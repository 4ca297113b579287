"""Generates tests/golden/*_golden.pt from the LIVE reference (run in the authoring container only).

The reference wrappers are imported unmodified from /root/reference; `from_pretrained` is patched
to build the default-config architecture (no network) and our deterministic
state dict (oracle/weights.py) is loaded with strict=True.  Each case records the seeds/ctor
arguments and the reference's outputs; tests/test_oracle_golden.py replays them through the oracle
on any machine, and the GPU parity tests compare the CUDA path against the same files.

Usage: PYTHONPATH=/root/repo python -m oracle.make_golden [encodec|dac|mimi|all]
"""
import os
import sys

import torch

from . import weights

REF = "/root/reference"


def make_input(seed, B, T):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, T, generator=g) * 0.1


ENCODEC_CASES = [
    # name, sample_rate, K, B, T, length
    dict(name="b2_1s_k8", sample_rate=24000, K=8, B=2, T=24000, seed=999, length=None),
    dict(name="ragged_k32", sample_rate=24000, K=32, B=1, T=12345, seed=1000, length=None),
    dict(name="resample16k_k8", sample_rate=16000, K=8, B=2, T=8000, seed=1001, length=None),
    dict(name="masked_k4", sample_rate=24000, K=4, B=2, T=9600, seed=1002, length=[1.0, 0.55]),
    dict(name="tiny_k2", sample_rate=24000, K=2, B=1, T=700, seed=1003, length=None),
]


def golden_encodec():
    sys.path.insert(0, REF)
    from transformers import EncodecConfig, EncodecModel
    EncodecModel.from_pretrained = classmethod(lambda cls, name, **kw: cls(EncodecConfig()))
    import audiocodecs
    from . import encodec_ref as ref

    sd = weights.encodec_state_dict(0)
    out = {"cases": []}
    for c in ENCODEC_CASES:
        codec = audiocodecs.Encodec(c["sample_rate"], 24000, num_codebooks=c["K"]).eval()
        codec.model.load_state_dict(sd, strict=True)
        sig = make_input(c["seed"], c["B"], c["T"])
        length = None if c["length"] is None else torch.tensor(c["length"])
        with torch.no_grad():
            toks = codec.sig_to_toks(sig, length)
            rec = codec.toks_to_sig(toks, length)
            qf = codec.toks_to_qfeats(toks)
            feats = codec.sig_to_feats(sig, length)
            # oracle on the same input, for the record printed below
            o_toks, gaps, emb = ref.sig_to_toks(sd, sig, c["K"], c["sample_rate"], 24000, length, return_gaps=True)
            o_rec = ref.toks_to_sig(sd, toks, c["sample_rate"], 24000)
        match = (o_toks == toks).float().mean().item()
        safe = gaps > 1e-4
        match_safe = (o_toks == toks)[safe].float().mean().item()
        print(f"encodec/{c['name']}: toks {tuple(toks.shape)} match {match:.6f} (gap>1e-4: {match_safe:.6f}, "
              f"near-ties {(~safe).float().mean().item():.5f}) rec {tuple(rec.shape)} "
              f"max|d| {(o_rec - rec).abs().max().item():.3e} feats max|d| "
              f"{(emb.movedim(-1, -2) - feats).abs().max().item():.3e}")
        out["cases"].append(dict(c, toks=toks.contiguous().to(torch.int16), rec=rec.contiguous(),
                                 qfeats_sum=qf.double().sum().item(), feats=feats.contiguous().half(),
                                 near_tie=(~safe).contiguous()))
    torch.save(out, os.path.join(weights.GOLDEN_DIR, "encodec_golden.pt"))


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    torch.set_num_threads(8)
    if which in ("encodec", "all"):
        golden_encodec()
    if which in ("dac", "all"):
        from .make_golden_dac import golden_dac
        golden_dac()
    if which in ("mimi", "all"):
        from .make_golden_mimi import golden_mimi
        golden_mimi()
