"""The UNMODIFIED reference wrappers (`audiocodecs.Encodec / DAC / Mimi`, installed from /root/reference into
baseline/_ref by __graft_entry__.build(): `pip install --no-index --no-deps --target baseline/_ref`) as the comparator of
bench.py: on the host cores (`--impl reference`, `cpu_baseline`) and eagerly on the GPU (`gpu_eager_baseline`).

BASELINE.md section 4: pretrained checkpoints are not reachable offline, so `from_pretrained` is patched to construct the
default-config architecture and the benchmark's deterministic state dict is loaded into it; `descript-audio-codec` is not
installed, so the reference `DAC` wrapper runs over an in-memory `dac` module that adapts `transformers.DacModel` to the call
surface the wrapper uses.  Nothing of the reference is modified or copied: the wrapper classes are imported
from baseline/_ref as installed.  Returns None when baseline/_ref is absent (the caller then falls back to the oracle port
and says so).
"""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

DAC_ZOO = {"44khz": ((2, 4, 8, 8), (8, 8, 4, 2), 9, 44100), "24khz": ((2, 4, 5, 8), (8, 5, 4, 2), 32, 24000),
           "16khz": ((2, 4, 5, 8), (8, 5, 4, 2), 12, 16000)}


def available():
    return os.path.exists(os.path.join(REF_DIR, "audiocodecs", "codec.py"))


def _import_reference():
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import audiocodecs  # the installed, unmodified reference package
    assert os.path.abspath(audiocodecs.__file__).startswith(REF_DIR), audiocodecs.__file__
    return audiocodecs


def _install_dac_shim(sd, tag):
    from transformers import DacConfig, DacModel
    enc, dec, nq, sr = DAC_ZOO[tag]
    cfg = DacConfig(encoder_hidden_size=64, downsampling_ratios=list(enc), decoder_hidden_size=1536, upsampling_ratios=list(dec),
                    n_codebooks=nq, codebook_size=1024, codebook_dim=8, hidden_size=1024, sampling_rate=sr)

    class Adapter(torch.nn.Module):  # descript `dac.DAC` call surface over the HF twin
        def __init__(self):
            super().__init__()
            self.m = DacModel(cfg)
            self.m.load_state_dict(sd, strict=True)
            self.encoder, self.decoder, self.quantizer = self.m.encoder, self.m.decoder, self.m.quantizer

        def encode(self, x, n_quantizers=None):
            return self.m.quantizer(self.m.encoder(x), n_quantizers)  # (z, codes, latents, commitment, codebook)

        def decode(self, z):
            return self.m.decoder(z)

    dac = types.ModuleType("dac")
    dac.utils = types.SimpleNamespace(download=lambda model_type="44khz": model_type)
    dac.DAC = types.SimpleNamespace(load=lambda path: Adapter())
    sys.modules["dac"] = dac


def make_reference(codec, sd, sample_rate, num_codebooks):
    """codec in {"encodec", "encodec32", "dac", "mimi"} -> the reference wrapper instance (eval, CPU) holding `sd`."""
    if not available():
        return None
    if codec in ("encodec", "encodec32"):
        from transformers import EncodecConfig, EncodecModel
        EncodecModel.from_pretrained = classmethod(lambda cls, name, **kw: cls(EncodecConfig()))
        ref = _import_reference().Encodec(sample_rate, 24000, num_codebooks=num_codebooks).eval()
        ref.model.load_state_dict(sd, strict=True)
    elif codec == "mimi":
        from transformers import MimiConfig, MimiModel
        MimiModel.from_pretrained = classmethod(lambda cls, name, **kw: cls(MimiConfig()))
        ref = _import_reference().Mimi(sample_rate, num_codebooks=num_codebooks).eval()
        missing = ref.model.load_state_dict(sd, strict=False)
        assert not missing.unexpected_keys and all("inv_freq" in k for k in missing.missing_keys), missing
        for m in ref.model.modules():
            if hasattr(m, "_embed"):
                m._embed = None  # cached property of MimiEuclideanCodebook
    else:
        _install_dac_shim(sd, "44khz")
        ref = _import_reference().DAC(sample_rate, 44100, num_codebooks=num_codebooks).eval()
    return ref
