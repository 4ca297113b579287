"""Oracle: DAC 44.1 kHz tokenize / detokenize, fp32 CPU restatement (TEST INFRASTRUCTURE).

The reference wrapper (`R/audiocodecs/dac.py:94-100,124-130`) calls the un-vendored third-party package
`descript-audio-codec` (pinned ==1.0.0 in R/downstream/environment.yml:71; `dac.DAC.encode`,
`quantizer.from_codes`, `decode`).  That package is absent here; its published algorithm
(dac/model/dac.py Encoder/Decoder, dac/nn/layers.py Snake1d, dac/nn/quantize.py VectorQuantize /
ResidualVectorQuantize) is restated below and anchored on its in-container twin
`transformers.DacModel` (HF/dac/modeling_dac.py), against which oracle/make_golden.py pins it by running the
UNMODIFIED reference wrapper over an in-memory `dac` shim.
"""
import math

import torch
import torch.nn.functional as F

from .resample_ref import resample

ENC_RATIOS = (2, 4, 8, 8)  # 44 kHz model; the 16 / 24 kHz models use (2, 4, 5, 8) / (8, 5, 4, 2)
DEC_RATIOS = (8, 8, 4, 2)


def _strides(sd, side):
    """strides of the four blocks, read off the state dict: every strided / transposed conv has kernel 2 * stride
    (descript dac/model/dac.py EncoderBlock / DecoderBlock; HF/dac:210-262)."""
    name = "encoder.block.{}.conv1.weight" if side == "enc" else "decoder.block.{}.conv_t1.weight"
    return [sd[name.format(i)].shape[-1] // 2 for i in range(4)]


def snake(x, alpha):
    """Snake1d (HF/dac:85-99): x + sin^2(alpha x) / (alpha + 1e-9)."""
    return x + (alpha + 1e-9).reciprocal() * torch.sin(alpha * x).pow(2)


def res_unit(sd, p, x, dilation):
    """DacResidualUnit (HF/dac:173-207)."""
    y = F.conv1d(snake(x, sd[p + ".snake1.alpha"]), sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], dilation=dilation, padding=3 * dilation)
    y = F.conv1d(snake(y, sd[p + ".snake2.alpha"]), sd[p + ".conv2.weight"], sd[p + ".conv2.bias"])
    return x + y


def encoder(sd, x):
    """DacEncoder (HF/dac:442-472,210-231): [B,1,T] -> [B,1024,N]."""
    x = F.conv1d(x, sd["encoder.conv1.weight"], sd["encoder.conv1.bias"], padding=3)
    for i, s in enumerate(_strides(sd, "enc")):
        p = f"encoder.block.{i}"
        for u, d in ((1, 1), (2, 3), (3, 9)):
            x = res_unit(sd, f"{p}.res_unit{u}", x, d)
        x = F.conv1d(snake(x, sd[p + ".snake1.alpha"]), sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], stride=s, padding=math.ceil(s / 2))
    return F.conv1d(snake(x, sd["encoder.snake1.alpha"]), sd["encoder.conv2.weight"], sd["encoder.conv2.bias"], padding=1)


def decoder(sd, z):
    """DacDecoder (HF/dac:405-439,234-262): [B,1024,N] -> [B,1,512 N], tanh output.  An odd stride s (the 16 / 24 kHz models'
    5: kernel 10, padding 3, no output_padding in descript 1.0.0 / the HF twin) gives (L - 1) s - 2 ceil(s/2) + 2 s = s L - 1."""
    x = F.conv1d(z, sd["decoder.conv1.weight"], sd["decoder.conv1.bias"], padding=3)
    for i, s in enumerate(_strides(sd, "dec")):
        p = f"decoder.block.{i}"
        x = F.conv_transpose1d(snake(x, sd[p + ".snake1.alpha"]), sd[p + ".conv_t1.weight"], sd[p + ".conv_t1.bias"], stride=s, padding=math.ceil(s / 2))
        for u, d in ((1, 1), (2, 3), (3, 9)):
            x = res_unit(sd, f"{p}.res_unit{u}", x, d)
    x = F.conv1d(snake(x, sd["decoder.snake1.alpha"]), sd["decoder.conv2.weight"], sd["decoder.conv2.bias"], padding=3)
    return torch.tanh(x)


def rvq_encode(sd, z, K, return_gaps=False):
    """ResidualVectorQuantize.forward / VectorQuantize.decode_latents (HF/dac:281-343,122-170):
    per stage in_proj -> L2-normalise -> argmax(-dist) -> codebook lookup -> STE value -> out_proj -> subtract."""
    res = z
    zq_sum = 0
    codes, gaps = [], []
    for k in range(K):
        q = f"quantizer.quantizers.{k}."
        z_e = F.conv1d(res, sd[q + "in_proj.weight"], sd[q + "in_proj.bias"])
        B, D, N = z_e.shape
        enc = F.normalize(z_e.permute(0, 2, 1).reshape(B * N, D))
        cb = F.normalize(sd[q + "codebook.weight"])
        dist = -(enc.pow(2).sum(1, keepdim=True) - 2 * enc @ cb.t()) + cb.pow(2).sum(1, keepdim=True).t()
        ind = dist.max(1)[1]
        if return_gaps:
            top2 = dist.topk(2, dim=-1).values
            gaps.append(((top2[:, 0] - top2[:, 1]) / top2[:, 0].abs().clamp_min(1e-30)).view(B, N))
        ind = ind.view(B, N)
        z_q = F.embedding(ind, sd[q + "codebook.weight"]).transpose(1, 2)
        z_q = z_e + (z_q - z_e)  # straight-through arithmetic kept in eval
        out = F.conv1d(z_q, sd[q + "out_proj.weight"], sd[q + "out_proj.bias"])
        zq_sum = zq_sum + out
        res = res - out
        codes.append(ind)
    codes = torch.stack(codes, dim=1)  # [B,K,N]
    if return_gaps:
        return codes, torch.stack(gaps, dim=1), zq_sum
    return codes


def from_codes(sd, codes):
    """ResidualVectorQuantize.from_codes (HF/dac:345-369): sum_k out_proj_k(codebook_k[codes_k]). codes [B,K,N]."""
    z = 0.0
    for k in range(codes.shape[1]):
        q = f"quantizer.quantizers.{k}."
        z_p = F.embedding(codes[:, k], sd[q + "codebook.weight"]).transpose(1, 2)
        z = z + F.conv1d(z_p, sd[q + "out_proj.weight"], sd[q + "out_proj.bias"])
    return z


def sig_to_feats(sd, sig, sample_rate=44100, orig_sample_rate=44100):
    """R/audiocodecs/dac.py:103-112 (latent=False): encoder output [B,N,1024]."""
    return encoder(sd, resample(sig, sample_rate, orig_sample_rate)[:, None]).movedim(-1, -2)


def sig_to_toks(sd, sig, num_codebooks=9, sample_rate=44100, orig_sample_rate=44100, return_gaps=False):
    """Codec.sig_to_toks -> DAC._sig_to_toks (R/dac.py:94-100); `length` is ignored by the reference."""
    z = encoder(sd, resample(sig, sample_rate, orig_sample_rate)[:, None])
    out = rvq_encode(sd, z, num_codebooks, return_gaps)
    if return_gaps:
        return out[0].movedim(-1, -2), out[1].movedim(-1, -2), z
    return out.movedim(-1, -2)


def toks_to_sig(sd, toks, sample_rate=44100, orig_sample_rate=44100):
    """Codec.toks_to_sig -> DAC._toks_to_sig (R/dac.py:124-130)."""
    z = from_codes(sd, toks.long().movedim(-1, -2))
    return resample(decoder(sd, z)[:, 0], orig_sample_rate, sample_rate)
