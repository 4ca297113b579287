"""Deterministic random-init weights of the named architectures, in the backends' own
state-dict key format (TEST INFRASTRUCTURE; see oracle/__init__.py).

Pretrained checkpoints are not reachable offline, so both the
oracle and the CUDA path are fed these tensors.  Shapes/keys follow
`transformers.EncodecModel(EncodecConfig()).state_dict()` (HF/encodec/modeling_encodec.py:82-447),
`MimiModel(MimiConfig())` and `DacModel(DacConfig 44 kHz)`; the same dicts load with
`load_state_dict(strict=True)` into those classes, which is how `oracle/make_golden.py`
pins the oracle against the live reference.

Codebooks: HF zero-inits them (HF/encodec:484-488) which makes every distance tie.
We use the matched-moment recipe of SURVEY.md section 8c: stage-k codebook = mu_k + sigma_k * randn,
mu/sigma = per-dimension moments of the stage-k residual on a calibration batch.  The moments are a
committed fixture (tests/golden/*_moments.pt, written by oracle/calibrate.py) so the
state dict is a pure function of (seed, fixture) on every machine.
"""
import math
import os

import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _gen(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def _conv_wn(sd, prefix, cout, cin, k, g, transpose=False, gain=1.0):
    """weight-normed conv: original0 = g (per dim-0 slice), original1 = v (HF/encodec:106-111)."""
    shape = (cin, cout, k) if transpose else (cout, cin, k)
    fan_in = cin * k if not transpose else cin * k / max(1, k // 2)  # convT: ~2 taps hit each output
    v = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in) * gain
    norm = v.flatten(1).norm(dim=1).view(-1, 1, 1)
    # g deliberately != ||v|| so that the weight-norm fold is exercised
    gvec = norm * (0.9 + 0.2 * torch.rand(norm.shape, generator=g))
    bound = 1.0 / math.sqrt(cin * k)
    sd[prefix + ".conv.bias"] = (torch.rand(cout, generator=g) * 2 - 1) * bound
    sd[prefix + ".conv.parametrizations.weight.original0"] = gvec
    sd[prefix + ".conv.parametrizations.weight.original1"] = v


def _lstm(sd, prefix, dim, layers, g):
    bound = math.sqrt(6.0 / (dim + 4 * dim))  # xavier-uniform as HF/encodec:470-475
    for l in range(layers):
        for nm in ("ih", "hh"):
            sd[f"{prefix}.lstm.weight_{nm}_l{l}"] = (torch.rand(4 * dim, dim, generator=g) * 2 - 1) * bound
            sd[f"{prefix}.lstm.bias_{nm}_l{l}"] = (torch.rand(4 * dim, generator=g) * 2 - 1) * 0.05


ENCODEC_RATIOS = (8, 5, 4, 2)  # HF/encodec/configuration_encodec.py: upsampling_ratios


def encodec_state_dict(seed: int = 0, num_quantizers: int = 32, codebooks: bool = True):
    """facebook/encodec_24khz architecture (EncodecConfig() defaults)."""
    g = _gen(seed)
    sd = {}
    nf, hidden = 32, 128
    # ---- encoder (HF/encodec:285-313)
    _conv_wn(sd, "encoder.layers.0", nf, 1, 7, g)
    idx, ch = 1, nf
    for r in reversed(ENCODEC_RATIOS):
        _conv_wn(sd, f"encoder.layers.{idx}.block.1", ch // 2, ch, 3, g)
        _conv_wn(sd, f"encoder.layers.{idx}.block.3", ch, ch // 2, 1, g, gain=0.7)
        _conv_wn(sd, f"encoder.layers.{idx}.shortcut", ch, ch, 1, g, gain=0.7)
        _conv_wn(sd, f"encoder.layers.{idx + 2}", ch * 2, ch, 2 * r, g)
        idx += 3
        ch *= 2
    _lstm(sd, f"encoder.layers.{idx}", ch, 2, g)
    _conv_wn(sd, f"encoder.layers.{idx + 2}", hidden, ch, 7, g)
    # ---- decoder (HF/encodec:316-347)
    _conv_wn(sd, "decoder.layers.0", ch, hidden, 7, g)
    _lstm(sd, "decoder.layers.1", ch, 2, g)
    idx = 3
    for r in ENCODEC_RATIOS:
        _conv_wn(sd, f"decoder.layers.{idx}", ch // 2, ch, 2 * r, g, transpose=True)
        ch //= 2
        _conv_wn(sd, f"decoder.layers.{idx + 1}.block.1", ch // 2, ch, 3, g)
        _conv_wn(sd, f"decoder.layers.{idx + 1}.block.3", ch, ch // 2, 1, g, gain=0.7)
        _conv_wn(sd, f"decoder.layers.{idx + 1}.shortcut", ch, ch, 1, g, gain=0.7)
        idx += 3
    _conv_wn(sd, f"decoder.layers.{idx}", 1, ch, 7, g)
    mom = torch.load(os.path.join(GOLDEN_DIR, "encodec_moments.pt")) if codebooks else None
    if mom is not None and "out_scale" in mom:
        # random-init decoders amplify; rescale the last layer so K=8 waveforms have std ~0.1
        # (the parity tolerance max-abs 1e-3 presumes audio-scale outputs)
        p = f"decoder.layers.{idx}.conv."
        sd[p + "parametrizations.weight.original0"] = sd[p + "parametrizations.weight.original0"] * mom["out_scale"]
        sd[p + "bias"] = sd[p + "bias"] * mom["out_scale"]
    # ---- quantizer buffers (HF/encodec:350-361)
    if codebooks:
        mu, sigma = mom["mu"], mom["sigma"]  # [32, 128]
    for k in range(num_quantizers):
        if codebooks:
            e = mu[k][None] + sigma[k][None] * torch.randn(1024, hidden, generator=g)
        else:
            e = torch.zeros(1024, hidden)
        p = f"quantizer.layers.{k}.codebook."
        sd[p + "inited"] = torch.ones(1)
        sd[p + "cluster_size"] = torch.zeros(1024)
        sd[p + "embed"] = e
        sd[p + "embed_avg"] = e.clone()
    return sd
