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


# ---------------------------------------------------------------------------------------------- Mimi
MIMI_RATIOS = (8, 6, 5, 4)  # MimiConfig().upsampling_ratios (kyutai/mimi)


def _conv_plain(sd, prefix, cout, cin, k, g, bias=True, transpose=False, gain=1.0, groups=1):
    """HF Mimi / Dac weights carry no weight-norm: plain `.weight` / `.bias`."""
    shape = (cin, cout // groups, k) if transpose else (cout, cin // groups, k)
    fan_in = (cin // groups) * k if not transpose else cin * k / max(1, k // 2) / groups
    sd[prefix + ".weight"] = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in) * gain
    if bias:
        sd[prefix + ".bias"] = (torch.rand(cout, generator=g) * 2 - 1) / math.sqrt((cin // groups) * k)


def mimi_state_dict(seed: int = 0, codebooks: bool = True):
    """kyutai/mimi architecture (MimiConfig() defaults; HF/mimi/modeling_mimi.py:454-496,1143-1173,996-1140).
    Conv gains are < 1 so that activations stay O(1) (kaiming init without weight-norm blows up)."""
    g = _gen(seed + 101)
    sd = {}
    nf, hid = 64, 512
    _conv_plain(sd, "encoder.layers.0.conv", nf, 1, 7, g)
    idx, ch = 1, nf
    for r in reversed(MIMI_RATIOS):
        _conv_plain(sd, f"encoder.layers.{idx}.block.1.conv", ch // 2, ch, 3, g, gain=0.7)
        _conv_plain(sd, f"encoder.layers.{idx}.block.3.conv", ch, ch // 2, 1, g, gain=0.5)
        _conv_plain(sd, f"encoder.layers.{idx + 2}.conv", ch * 2, ch, 2 * r, g, gain=0.7)
        idx += 3
        ch *= 2
    _conv_plain(sd, f"encoder.layers.{idx + 1}.conv", hid, ch, 3, g, gain=0.7)
    for name in ("encoder_transformer", "decoder_transformer"):
        for l in range(8):
            p = f"{name}.layers.{l}."
            for w in ("q_proj", "k_proj", "v_proj", "o_proj"):
                sd[p + f"self_attn.{w}.weight"] = torch.randn(hid, hid, generator=g) * hid ** -0.5
            sd[p + "mlp.fc1.weight"] = torch.randn(4 * hid, hid, generator=g) * hid ** -0.5
            sd[p + "mlp.fc2.weight"] = torch.randn(hid, 4 * hid, generator=g) * (4 * hid) ** -0.5
            for ln in ("input_layernorm", "post_attention_layernorm"):
                sd[p + ln + ".weight"] = 1.0 + 0.1 * torch.randn(hid, generator=g)
                sd[p + ln + ".bias"] = 0.05 * torch.randn(hid, generator=g)
            # layer-scale: trained checkpoints are O(0.1-1); 0.01 (the init) would hide the attention/MLP arithmetic
            sd[p + "self_attn_layer_scale.scale"] = 0.2 + 0.1 * torch.rand(hid, generator=g)
            sd[p + "mlp_layer_scale.scale"] = 0.2 + 0.1 * torch.rand(hid, generator=g)
    _conv_plain(sd, "downsample.conv", hid, hid, 4, g, bias=False, gain=0.7)
    sd["upsample.conv.weight"] = torch.randn(hid, 1, 4, generator=g) * 0.7
    _conv_plain(sd, "decoder.layers.0.conv", ch, hid, 7, g, gain=0.7)
    idx = 2
    for r in MIMI_RATIOS:
        _conv_plain(sd, f"decoder.layers.{idx}.conv", ch // 2, ch, 2 * r, g, transpose=True, gain=0.7)
        ch //= 2
        _conv_plain(sd, f"decoder.layers.{idx + 1}.block.1.conv", ch // 2, ch, 3, g, gain=0.7)
        _conv_plain(sd, f"decoder.layers.{idx + 1}.block.3.conv", ch, ch // 2, 1, g, gain=0.5)
        idx += 3
    _conv_plain(sd, f"decoder.layers.{idx}.conv", 1, ch, 3, g)
    mom = torch.load(os.path.join(GOLDEN_DIR, "mimi_moments.pt")) if codebooks else None
    if mom is not None and "out_scale" in mom:
        sd[f"decoder.layers.{idx}.conv.weight"] = sd[f"decoder.layers.{idx}.conv.weight"] * mom["out_scale"]
        sd[f"decoder.layers.{idx}.conv.bias"] = sd[f"decoder.layers.{idx}.conv.bias"] * mom["out_scale"]
    for which, n in (("semantic", 1), ("acoustic", 31)):
        q = f"quantizer.{which}_residual_vector_quantizer."
        sd[q + "input_proj.weight"] = torch.randn(256, hid, 1, generator=g) * hid ** -0.5
        sd[q + "output_proj.weight"] = torch.randn(hid, 256, 1, generator=g) * 256 ** -0.5
        for k in range(n):
            if codebooks:
                mu, sigma = mom[which + "_mu"][k], mom[which + "_sigma"][k]
                e = mu[None] + sigma[None] * torch.randn(2048, 256, generator=g)
            else:
                e = torch.zeros(2048, 256)
            sd[q + f"layers.{k}.codebook.initialized"] = torch.ones(1)
            sd[q + f"layers.{k}.codebook.cluster_usage"] = torch.ones(2048)
            sd[q + f"layers.{k}.codebook.embed_sum"] = e
    return sd


# ---------------------------------------------------------------------------------------------- DAC 44.1 kHz
DAC_ENC_RATIOS = (2, 4, 8, 8)  # descript-audio-codec 1.0.0 "44khz": encoder_rates; decoder_rates = (8, 8, 4, 2)
DAC_DEC_RATIOS = (8, 8, 4, 2)


DAC_ZOO = {  # descript-audio-codec 1.0.0 model zoo: tag -> (encoder rates, decoder rates, codebooks)
    "44khz": ((2, 4, 8, 8), (8, 8, 4, 2), 9),
    "24khz": ((2, 4, 5, 8), (8, 5, 4, 2), 32),
    "16khz": ((2, 4, 5, 8), (8, 5, 4, 2), 12),
}


def dac_state_dict(seed: int = 0, n_codebooks: int = None, tag: str = "44khz"):
    """descript architecture `tag` in `transformers.DacModel` key format (HF/dac/modeling_dac.py:405-472,173-262):
    encoder_hidden_size 64, decoder_hidden_size 1536, hidden 1024, codebooks x 1024 x 8.  The 44 kHz dict is unchanged by the
    `tag` argument (same generator seed and draw order); the 16 / 24 kHz models have an ODD stride (5) in both stacks."""
    DAC_ENC_RATIOS, DAC_DEC_RATIOS, zoo_k = DAC_ZOO[tag]
    n_codebooks = n_codebooks or zoo_k
    g = _gen(seed + 202 + {"44khz": 0, "24khz": 1000, "16khz": 2000}[tag])
    sd = {}

    def snake(prefix, c):
        sd[prefix + ".alpha"] = (0.5 + torch.rand(1, c, 1, generator=g))

    def res_unit(prefix, c, gain):
        snake(prefix + ".snake1", c)
        _conv_plain(sd, prefix + ".conv1", c, c, 7, g, gain=gain)
        snake(prefix + ".snake2", c)
        _conv_plain(sd, prefix + ".conv2", c, c, 1, g, gain=0.3)

    _conv_plain(sd, "encoder.conv1", 64, 1, 7, g)
    c = 64
    for i, s in enumerate(DAC_ENC_RATIOS):
        for u in (1, 2, 3):
            res_unit(f"encoder.block.{i}.res_unit{u}", c, 0.5)
        snake(f"encoder.block.{i}.snake1", c)
        _conv_plain(sd, f"encoder.block.{i}.conv1", 2 * c, c, 2 * s, g, gain=0.5)
        c *= 2
    snake("encoder.snake1", c)
    _conv_plain(sd, "encoder.conv2", 1024, c, 3, g, gain=0.5)
    _conv_plain(sd, "decoder.conv1", 1536, 1024, 7, g, gain=0.5)
    c = 1536
    for i, s in enumerate(DAC_DEC_RATIOS):
        snake(f"decoder.block.{i}.snake1", c)
        _conv_plain(sd, f"decoder.block.{i}.conv_t1", c // 2, c, 2 * s, g, transpose=True, gain=0.5)
        c //= 2
        for u in (1, 2, 3):
            res_unit(f"decoder.block.{i}.res_unit{u}", c, 0.5)
    snake("decoder.snake1", c)
    _conv_plain(sd, "decoder.conv2", 1, c, 7, g, gain=0.05)
    for k in range(n_codebooks):
        q = f"quantizer.quantizers.{k}."
        _conv_plain(sd, q + "in_proj", 8, 1024, 1, g, gain=0.7)
        _conv_plain(sd, q + "out_proj", 1024, 8, 1, g, gain=0.7)
        sd[q + "codebook.weight"] = torch.randn(1024, 8, generator=g) * 0.02  # both sides are L2-normalised
    return sd
