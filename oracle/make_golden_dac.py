"""Golden vectors for DAC-44k: the UNMODIFIED reference wrapper (R/audiocodecs/dac.py) run over an in-memory `dac`
module that adapts `transformers.DacModel` to the descript-audio-codec 1.0.0 call surface the wrapper uses
(`dac.utils.download`, `dac.DAC.load`, `.encode(x, n_quantizers)` 5-tuple, `.quantizer.from_codes`, `.decode`)."""
import os
import sys
import types

import torch

from . import weights
from .make_golden import REF, make_input

DAC_CASES = [
    dict(name="b2_k9", sample_rate=44100, K=9, B=2, T=22050, seed=3001),
    dict(name="ragged_k4", sample_rate=44100, K=4, B=1, T=17777, seed=3002),
    dict(name="resample16k_k9", sample_rate=16000, K=9, B=1, T=6000, seed=3003),
]


def install_dac_shim(sd):
    from transformers import DacConfig, DacModel

    cfg = DacConfig(encoder_hidden_size=64, downsampling_ratios=[2, 4, 8, 8], decoder_hidden_size=1536,
                    upsampling_ratios=[8, 8, 4, 2], n_codebooks=9, codebook_size=1024, codebook_dim=8, hidden_size=1024,
                    sampling_rate=44100)

    class Adapter(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.m = DacModel(cfg)
            self.m.load_state_dict(sd, strict=True)
            self.encoder, self.decoder, self.quantizer = self.m.encoder, self.m.decoder, self.m.quantizer

        def encode(self, x, n_quantizers=None):
            z = self.m.encoder(x)
            zq, codes, latents, commit, cb = self.m.quantizer(z, n_quantizers)
            return zq, codes, latents, commit, cb

        def decode(self, z):
            return self.m.decoder(z)

    dac = types.ModuleType("dac")
    dac.utils = types.SimpleNamespace(download=lambda model_type="44khz": model_type)
    dac.DAC = types.SimpleNamespace(load=lambda path: Adapter())
    sys.modules["dac"] = dac
    return cfg


def golden_dac():
    sys.path.insert(0, REF)
    sd = weights.dac_state_dict(0)
    install_dac_shim(sd)
    import audiocodecs
    from . import dac_ref as ref

    out = {"cases": []}
    for c in DAC_CASES:
        codec = audiocodecs.DAC(c["sample_rate"], 44100, num_codebooks=c["K"]).eval()
        sig = make_input(c["seed"], c["B"], c["T"])
        with torch.no_grad():
            toks = codec.sig_to_toks(sig)
            rec = codec.toks_to_sig(toks)
            feats = codec.sig_to_feats(sig)
            o_toks, gaps, z = ref.sig_to_toks(sd, sig, c["K"], c["sample_rate"], 44100, return_gaps=True)
            o_rec = ref.toks_to_sig(sd, toks, c["sample_rate"], 44100)
        safe = gaps > 1e-4
        eq = o_toks == toks
        print(f"dac/{c['name']}: toks {tuple(toks.shape)} match {eq.float().mean().item():.6f} (gap>1e-4: "
              f"{eq[safe].float().mean().item():.6f}, near-ties {(~safe).float().mean().item():.5f}) rec {tuple(rec.shape)} "
              f"std {rec.std().item():.3f} max|d| {(o_rec - rec).abs().max().item():.3e} feats max|d| "
              f"{(z.movedim(-1, -2) - feats).abs().max().item():.3e} distinct codes {toks.unique().numel()}")
        out["cases"].append(dict(c, toks=toks.contiguous().to(torch.int16), rec=rec.contiguous(), feats=feats.contiguous().half(),
                                 near_tie=(~safe).contiguous()))
    torch.save(out, os.path.join(weights.GOLDEN_DIR, "dac_golden.pt"))
