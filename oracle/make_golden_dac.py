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


DAC_ODD_CASES = [  # the 16 / 24 kHz models (odd stride 5); `DAC(sample_rate)` with its default arguments is case 0
    dict(name="dac16_default_ctor", sample_rate=16000, orig_sample_rate=16000, K=8, B=2, T=8000, seed=3101),
    dict(name="dac16_ragged_k12", sample_rate=16000, orig_sample_rate=16000, K=12, B=1, T=5555, seed=3102),
    dict(name="dac24_k32_from16k", sample_rate=16000, orig_sample_rate=24000, K=32, B=1, T=6000, seed=3103),
]


def install_dac_shim(sd, tag="44khz"):
    from transformers import DacConfig, DacModel

    enc, dec, nq = weights.DAC_ZOO[tag]
    cfg = DacConfig(encoder_hidden_size=64, downsampling_ratios=list(enc), decoder_hidden_size=1536,
                    upsampling_ratios=list(dec), n_codebooks=nq, codebook_size=1024, codebook_dim=8, hidden_size=1024,
                    sampling_rate={"44khz": 44100, "24khz": 24000, "16khz": 16000}[tag])

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

    def download(model_type="44khz"):
        assert model_type == tag, f"the wrapper asked for {model_type}, the shim holds {tag}"
        return model_type

    dac = types.ModuleType("dac")
    dac.utils = types.SimpleNamespace(download=download)
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


def golden_dac_odd():
    """tests/golden/dac_odd_golden.pt: the unmodified wrapper on the 16 / 24 kHz architectures (stride-5 blocks)."""
    sys.path.insert(0, REF)
    import audiocodecs
    from . import dac_ref as ref

    out = {"cases": []}
    for c in DAC_ODD_CASES:
        tag = f"{c['orig_sample_rate'] // 1000}khz"
        sd = weights.dac_state_dict(0, tag=tag)
        install_dac_shim(sd, tag)
        if c["name"] == "dac16_default_ctor":
            codec = audiocodecs.DAC(c["sample_rate"]).eval()  # orig_sample_rate=16000, num_codebooks=8: R/audiocodecs/dac.py:31-38
        else:
            codec = audiocodecs.DAC(c["sample_rate"], c["orig_sample_rate"], num_codebooks=c["K"]).eval()
        sig = make_input(c["seed"], c["B"], c["T"])
        with torch.no_grad():
            toks = codec.sig_to_toks(sig)
            rec = codec.toks_to_sig(toks)
            feats = codec.sig_to_feats(sig)
            o_toks, gaps, z = ref.sig_to_toks(sd, sig, c["K"], c["sample_rate"], c["orig_sample_rate"], return_gaps=True)
            o_rec = ref.toks_to_sig(sd, toks, c["sample_rate"], c["orig_sample_rate"])
        safe = gaps > 1e-4
        eq = o_toks == toks
        print(f"dac/{c['name']}: toks {tuple(toks.shape)} match {eq.float().mean().item():.6f} (gap>1e-4: "
              f"{eq[safe].float().mean().item():.6f}, near-ties {(~safe).float().mean().item():.5f}) rec {tuple(rec.shape)} "
              f"std {rec.std().item():.3f} max|d| {(o_rec - rec).abs().max().item():.3e} feats max|d| "
              f"{(z.movedim(-1, -2) - feats).abs().max().item():.3e} distinct codes {toks.unique().numel()}")
        out["cases"].append(dict(c, toks=toks.contiguous().to(torch.int16), rec=rec.contiguous(), feats=feats.contiguous().half(),
                                 near_tie=(~safe).contiguous()))
    torch.save(out, os.path.join(weights.GOLDEN_DIR, "dac_odd_golden.pt"))


if __name__ == "__main__":
    torch.set_num_threads(8)
    golden_dac_odd()
