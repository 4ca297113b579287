"""Oracle: Mimi tokenize / detokenize, fp32 CPU restatement (TEST INFRASTRUCTURE).

Path: `R/audiocodecs/codec.py:57-66,90-100` -> `R/audiocodecs/mimi.py:93-109,144-148` ->
`HF/mimi/modeling_mimi.py` (transformers 5.5.0; the wrapper never enables streaming, mimi.py:105-107).
Pinned against the live reference by oracle/make_golden.py (tests/golden/mimi_golden.pt).
"""
import math

import torch
import torch.nn.functional as F

from .resample_ref import resample

RATIOS = (8, 6, 5, 4)
HID, HEADS, HEAD_DIM, WINDOW = 512, 8, 64, 250


def causal_conv(x, w, b, stride=1, mode="constant"):
    """MimiConv1d.forward, causal branch (HF/mimi:273-283,331-351): left pad K_eff-stride, right pad up to
    ceil(L/stride) frames; pad mode constant-0, or replicate for `downsample` (HF/mimi:1422-1431)."""
    k = w.shape[-1]
    pt = k - stride
    L = x.shape[-1]
    n_frames = math.ceil((L - k + pt) / stride + 1) - 1
    extra = n_frames * stride + k - pt - L
    x = F.pad(x, (pt, extra), mode=mode)
    return F.conv1d(x, w, b, stride=stride)


def causal_convtr(x, w, b, stride, groups=1):
    """MimiConvTranspose1d (HF/mimi:354-409): trim K-stride on the right."""
    y = F.conv_transpose1d(x, w, b, stride=stride, groups=groups)
    return y[..., : y.shape[-1] - (w.shape[-1] - stride)]


def resblock(sd, p, x):
    """MimiResnetBlock with identity shortcut (HF/mimi:412-451)."""
    h = causal_conv(F.elu(x), sd[p + ".block.1.conv.weight"], sd[p + ".block.1.conv.bias"])
    h = causal_conv(F.elu(h), sd[p + ".block.3.conv.weight"], sd[p + ".block.3.conv.bias"])
    return x + h


def encoder(sd, x):
    """MimiEncoder (HF/mimi:454-496): [B,1,T] -> [B,512,T/960]."""
    x = causal_conv(x, sd["encoder.layers.0.conv.weight"], sd["encoder.layers.0.conv.bias"])
    idx = 1
    for r in reversed(RATIOS):
        x = resblock(sd, f"encoder.layers.{idx}", x)
        x = causal_conv(F.elu(x), sd[f"encoder.layers.{idx + 2}.conv.weight"], sd[f"encoder.layers.{idx + 2}.conv.bias"], stride=r)
        idx += 3
    return causal_conv(F.elu(x), sd[f"encoder.layers.{idx + 1}.conv.weight"], sd[f"encoder.layers.{idx + 1}.conv.bias"])


def decoder(sd, x):
    """MimiDecoder (HF/mimi:1143-1173): [B,512,N] -> [B,1,960 N]."""
    x = causal_conv(x, sd["decoder.layers.0.conv.weight"], sd["decoder.layers.0.conv.bias"])
    idx = 2
    for r in RATIOS:
        x = causal_convtr(F.elu(x), sd[f"decoder.layers.{idx}.conv.weight"], sd[f"decoder.layers.{idx}.conv.bias"], r)
        x = resblock(sd, f"decoder.layers.{idx + 1}", x)
        idx += 3
    return causal_conv(F.elu(x), sd[f"decoder.layers.{idx}.conv.weight"], sd[f"decoder.layers.{idx}.conv.bias"])


def rope_tables(T):
    """MimiRotaryEmbedding (HF/mimi:515-577): theta 10000, positions 0..T-1, fp32."""
    inv_freq = 1.0 / (10000.0 ** (torch.arange(0, HEAD_DIM, 2, dtype=torch.int64).float() / HEAD_DIM))
    freqs = torch.arange(T).float()[:, None] * inv_freq[None]
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos(), emb.sin()


def rotate_half(x):
    return torch.cat((-x[..., HEAD_DIM // 2:], x[..., : HEAD_DIM // 2]), dim=-1)


def transformer(sd, name, h):
    """MimiTransformerModel (HF/mimi:996-1140,926-993,645-736): h [B,T,512]; causal + sliding window 250."""
    B, T, _ = h.shape
    cos, sin = rope_tables(T)
    i = torch.arange(T)
    allowed = (i[None, :] <= i[:, None]) & (i[None, :] > i[:, None] - WINDOW)
    mask = torch.zeros(T, T).masked_fill(~allowed, float("-inf"))
    for l in range(8):
        p = f"{name}.layers.{l}."
        x = F.layer_norm(h, (HID,), sd[p + "input_layernorm.weight"], sd[p + "input_layernorm.bias"], 1e-5)
        q = (x @ sd[p + "self_attn.q_proj.weight"].t()).view(B, T, HEADS, HEAD_DIM).transpose(1, 2)
        k = (x @ sd[p + "self_attn.k_proj.weight"].t()).view(B, T, HEADS, HEAD_DIM).transpose(1, 2)
        v = (x @ sd[p + "self_attn.v_proj.weight"].t()).view(B, T, HEADS, HEAD_DIM).transpose(1, 2)
        q = q * cos + rotate_half(q) * sin
        k = k * cos + rotate_half(k) * sin
        att = torch.softmax(q @ k.transpose(2, 3) * (1 / math.sqrt(HEAD_DIM)) + mask, dim=-1, dtype=torch.float32)
        o = (att @ v).transpose(1, 2).reshape(B, T, HID) @ sd[p + "self_attn.o_proj.weight"].t()
        h = h + sd[p + "self_attn_layer_scale.scale"] * o
        x = F.layer_norm(h, (HID,), sd[p + "post_attention_layernorm.weight"], sd[p + "post_attention_layernorm.bias"], 1e-5)
        m = F.gelu(x @ sd[p + "mlp.fc1.weight"].t()) @ sd[p + "mlp.fc2.weight"].t()
        h = h + sd[p + "mlp_layer_scale.scale"] * m
    return h


def codebook(sd, which, k):
    """MimiEuclideanCodebook.embed (HF/mimi:1188-1195): embed_sum / clamp(cluster_usage, 1e-5)."""
    q = f"quantizer.{which}_residual_vector_quantizer.layers.{k}.codebook."
    return sd[q + "embed_sum"] / sd[q + "cluster_usage"].clamp(min=1e-5)[:, None]


def _rvq_encode(sd, which, emb, n, return_gaps):
    """MimiResidualVectorQuantizer.encode (HF/mimi:1262-1280): input_proj, cdist argmin, residual chain."""
    q = f"quantizer.{which}_residual_vector_quantizer."
    res = F.conv1d(emb, sd[q + "input_proj.weight"])
    codes, gaps = [], []
    for k in range(n):
        E = codebook(sd, which, k)
        flat = res.permute(0, 2, 1).reshape(-1, res.shape[1])
        d = torch.cdist(flat[None].float(), E[None].float(), p=2)[0]
        ind = d.argmin(dim=-1)
        if return_gaps:
            top2 = (-d).topk(2, dim=-1).values
            gaps.append(((top2[:, 0] - top2[:, 1]) / top2[:, 0].abs().clamp_min(1e-30)).view(res.shape[0], -1))
        ind = ind.view(res.shape[0], -1)
        res = res - F.embedding(ind, E).permute(0, 2, 1)
        codes.append(ind)
    return codes, gaps


def rvq_encode(sd, emb, K, return_gaps=False):
    """MimiSplitResidualVectorQuantizer.encode (HF/mimi:1311-1338): both quantizers see the same embeddings."""
    if K < 1 or K > 32:
        raise ValueError("The number of quantizers (i.e codebooks) asked should be in [1, 32]")
    c0, g0 = _rvq_encode(sd, "semantic", emb, 1, return_gaps)
    c1, g1 = _rvq_encode(sd, "acoustic", emb, K - 1, return_gaps) if K > 1 else ([], [])
    codes = torch.stack(c0 + c1, dim=1)  # [B,K,N]
    if return_gaps:
        return codes, torch.stack(g0 + g1, dim=1)
    return codes


def rvq_decode(sd, codes):
    """MimiSplitResidualVectorQuantizer.decode (HF/mimi:1340-1349,1282-1293). codes [B,K,N] -> [B,512,N]."""
    out = None
    for which, sl in (("semantic", codes[:, :1]), ("acoustic", codes[:, 1:])):
        if sl.shape[1] == 0:
            continue
        acc = torch.tensor(0.0)
        for k in range(sl.shape[1]):
            acc = acc + F.embedding(sl[:, k], codebook(sd, which, k)).permute(0, 2, 1)
        y = F.conv1d(acc, sd[f"quantizer.{which}_residual_vector_quantizer.output_proj.weight"])
        out = y if out is None else out + y
    return out


def sig_to_feats(sd, sig, sample_rate=24000):
    """R/audiocodecs/mimi.py:112-121: encoder -> transformer -> downsample, [B,N,512]."""
    sig = resample(sig, sample_rate, 24000)
    e = encoder(sd, sig[:, None])
    e = transformer(sd, "encoder_transformer", e.transpose(1, 2)).transpose(1, 2)
    e = causal_conv(e, sd["downsample.conv.weight"], None, stride=2, mode="replicate")
    return e.movedim(-1, -2)


def sig_to_toks(sd, sig, num_codebooks=8, sample_rate=24000, return_gaps=False):
    """Codec.sig_to_toks -> Mimi._sig_to_toks (the padding mask is built but unused, HF/mimi:1469-1472)."""
    emb = sig_to_feats(sd, sig, sample_rate).movedim(-1, -2)
    out = rvq_encode(sd, emb, num_codebooks, return_gaps)
    if return_gaps:
        return out[0].movedim(-1, -2), out[1].movedim(-1, -2), emb
    return out.movedim(-1, -2)


def toks_to_qfeats(sd, toks):
    """R/audiocodecs/mimi.py:151-155."""
    return rvq_decode(sd, toks.movedim(-1, -2)).movedim(-1, -2)


def toks_to_sig(sd, toks, sample_rate=24000):
    """Codec.toks_to_sig -> Mimi._toks_to_sig (R/mimi.py:144-148; HF/mimi:1613-1631)."""
    z = rvq_decode(sd, toks.long().movedim(-1, -2))
    z = causal_convtr(z, sd["upsample.conv.weight"], None, 2, groups=HID)
    z = transformer(sd, "decoder_transformer", z.transpose(1, 2)).transpose(1, 2)
    sig = decoder(sd, z)[:, 0]
    return resample(sig, 24000, sample_rate)
