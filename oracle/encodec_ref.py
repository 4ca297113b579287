"""Oracle: EnCodec-24k tokenize / detokenize, fp32 CPU restatement (TEST INFRASTRUCTURE).

Follows the path `R/audiocodecs/codec.py:57-66,90-100` -> `R/audiocodecs/encodec.py:82-94,130-141`
-> `HF/encodec/modeling_encodec.py` (transformers 5.5.0).  Operates on a state dict in the HF key
format (see oracle/weights.py) so the same tensors feed the live reference, this oracle and the
CUDA weight packer.  Parity: pinned against the live reference by oracle/make_golden.py
(tests/golden/encodec_*.pt) -- see tests/test_oracle_golden.py.
"""
import math

import torch
import torch.nn.functional as F

from .resample_ref import resample

RATIOS = (8, 5, 4, 2)
HOP = 320


def fold_weight_norm(sd, prefix):
    """w = g * v / ||v||, norm over all dims but 0 (HF/encodec:106-111)."""
    g = sd[prefix + ".conv.parametrizations.weight.original0"]
    v = sd[prefix + ".conv.parametrizations.weight.original1"]
    n = v.flatten(1).norm(dim=1).view(-1, 1, 1)
    return g * v / n, sd[prefix + ".conv.bias"]


def reflect_pad(x, left, right):
    """EncodecConv1d._pad1d (HF/encodec:139-155): zero-extend tiny inputs before reflecting."""
    L = x.shape[-1]
    m = max(left, right)
    extra = 0
    if L <= m:
        extra = m - L + 1
        x = F.pad(x, (0, extra))
    y = F.pad(x, (left, right), mode="reflect")
    return y[..., : y.shape[-1] - extra]


def causal_conv(x, w, b, stride=1, dilation=1):
    """EncodecConv1d.forward, causal branch (HF/encodec:116-136,157-176): L_out = ceil(L/stride)."""
    k_eff = (w.shape[-1] - 1) * dilation + 1
    pt = k_eff - stride
    L = x.shape[-1]
    n_frames = math.ceil((L - k_eff + pt) / stride + 1) - 1
    extra = n_frames * stride + k_eff - pt - L
    x = reflect_pad(x, pt, extra)
    return F.conv1d(x, w, b, stride=stride, dilation=dilation)


def causal_convtr(x, w, b, stride):
    """EncodecConvTranspose1d.forward (HF/encodec:206-233): trim K-stride samples on the right."""
    y = F.conv_transpose1d(x, w, b, stride=stride)
    pt = w.shape[-1] - stride
    return y[..., : y.shape[-1] - pt]


def resblock(sd, prefix, x):
    """EncodecResnetBlock (HF/encodec:252-282): shortcut(x) + conv1(ELU(conv3(ELU(x))))."""
    h = causal_conv(F.elu(x), *fold_weight_norm(sd, prefix + ".block.1"))
    h = causal_conv(F.elu(h), *fold_weight_norm(sd, prefix + ".block.3"))
    return causal_conv(x, *fold_weight_norm(sd, prefix + ".shortcut")) + h


def lstm_block(sd, prefix, x, layers=2):
    """EncodecLSTM (HF/encodec:236-249): 2-layer LSTM over time + skip; gates i,f,g,o."""
    B, C, T = x.shape
    inp = x.permute(2, 0, 1)  # [T, B, C]
    cur = inp
    for l in range(layers):
        w_ih, w_hh = sd[f"{prefix}.lstm.weight_ih_l{l}"], sd[f"{prefix}.lstm.weight_hh_l{l}"]
        bias = sd[f"{prefix}.lstm.bias_ih_l{l}"] + sd[f"{prefix}.lstm.bias_hh_l{l}"]
        pre = cur @ w_ih.t() + bias  # [T, B, 4C]
        h = x.new_zeros(B, C)
        c = x.new_zeros(B, C)
        outs = []
        w_hh_t = w_hh.t().contiguous()
        for t in range(T):
            gates = pre[t] + h @ w_hh_t
            i, f, g, o = gates.split(C, dim=1)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
            h = torch.sigmoid(o) * torch.tanh(c)
            outs.append(h)
        cur = torch.stack(outs)
    return (cur + inp).permute(1, 2, 0)


def lstm_block_aten(sd, prefix, x, layers=2):
    """Same block through ATen's fused `nn.LSTM` kernel -- the op the reference itself lands on
    (HF/encodec:242-246).  Used for the timed CPU baseline; tests check it equals `lstm_block`."""
    C = x.shape[1]
    m = torch.nn.LSTM(C, C, layers)
    m.load_state_dict({k.split(".lstm.")[1]: v for k, v in sd.items() if k.startswith(prefix + ".lstm.")})
    inp = x.permute(2, 0, 1)
    with torch.no_grad():
        return (m(inp)[0] + inp).permute(1, 2, 0)


_LSTM = {"loop": lstm_block, "aten": lstm_block_aten}


def encoder(sd, x, lstm="aten"):
    """EncodecEncoder (HF/encodec:285-313). x [B,1,T] -> [B,128,ceil(T/320)]."""
    x = causal_conv(x, *fold_weight_norm(sd, "encoder.layers.0"))
    idx = 1
    for r in reversed(RATIOS):
        x = resblock(sd, f"encoder.layers.{idx}", x)
        x = causal_conv(F.elu(x), *fold_weight_norm(sd, f"encoder.layers.{idx + 2}"), stride=r)
        idx += 3
    x = _LSTM[lstm](sd, f"encoder.layers.{idx}", x)
    return causal_conv(F.elu(x), *fold_weight_norm(sd, f"encoder.layers.{idx + 2}"))


def decoder(sd, z, lstm="aten"):
    """EncodecDecoder (HF/encodec:316-347). z [B,128,N] -> [B,1,320 N]."""
    x = causal_conv(z, *fold_weight_norm(sd, "decoder.layers.0"))
    x = _LSTM[lstm](sd, "decoder.layers.1", x)
    idx = 3
    for r in RATIOS:
        x = causal_convtr(F.elu(x), *fold_weight_norm(sd, f"decoder.layers.{idx}"), stride=r)
        x = resblock(sd, f"decoder.layers.{idx + 1}", x)
        idx += 3
    return causal_conv(F.elu(x), *fold_weight_norm(sd, f"decoder.layers.{idx}"))


def num_quantizers_for(num_codebooks):
    """Wrapper bandwidth = K*75/100 (R/audiocodecs/encodec.py:50); HF validates it and maps back
    to a stage count (HF/encodec:573-578,416-422)."""
    bw = num_codebooks * 75 / 100
    if bw not in (1.5, 3.0, 6.0, 12.0, 24.0):
        raise ValueError(f"This model doesn't support the bandwidth {bw}.")
    return int(max(1, math.floor(bw * 1000 / (10 * 75))))


def rvq_encode(sd, emb, nq, return_gaps=False):
    """EncodecResidualVectorQuantizer.encode + EuclideanCodebook.quantize (HF/encodec:364-369,424-438).

    emb [B,128,N] -> codes [nq,B,N] int64.  With return_gaps also the relative top-2 gap of every
    decision, (d2-d1)/max(|d1|,tiny) on the reference's own distance values (the parity tests
    exclude near-ties below 1e-4, BASELINE.json north_star).
    """
    res = emb
    codes, gaps = [], []
    for k in range(nq):
        E = sd[f"quantizer.layers.{k}.codebook.embed"]
        x = res.permute(0, 2, 1)
        flat = x.reshape(-1, x.shape[-1])
        et = E.t()
        dist = -(flat.pow(2).sum(1, keepdim=True) - 2 * flat @ et + et.pow(2).sum(0, keepdim=True))
        ind = dist.max(dim=-1).indices
        if return_gaps:
            top2 = dist.topk(2, dim=-1).values
            gaps.append(((top2[:, 0] - top2[:, 1]) / top2[:, 0].abs().clamp_min(1e-30)).view(x.shape[:-1]))
        ind = ind.view(x.shape[:-1])
        q = F.embedding(ind, E).permute(0, 2, 1)
        res = res - q
        codes.append(ind)
    codes = torch.stack(codes)
    if return_gaps:
        return codes, torch.stack(gaps)
    return codes


def rvq_decode(sd, codes):
    """EncodecResidualVectorQuantizer.decode (HF/encodec:440-447): sum in stage order. codes [K,B,N]."""
    out = torch.tensor(0.0)
    for k, ind in enumerate(codes):
        out = out + F.embedding(ind, sd[f"quantizer.layers.{k}.codebook.embed"]).permute(0, 2, 1)
    return out


def sig_to_feats(sd, sig, sample_rate=24000, orig_sample_rate=24000):
    """R/audiocodecs/encodec.py:97-117 (normalize=False at 24 kHz): encoder output [B,N,128]."""
    sig = resample(sig, sample_rate, orig_sample_rate)
    return encoder(sd, sig[:, None]).movedim(-1, -2)


def sig_to_toks(sd, sig, num_codebooks=8, sample_rate=24000, orig_sample_rate=24000, length=None,
                return_gaps=False):
    """Codec.sig_to_toks -> Encodec._sig_to_toks (R/codec.py:57-66, R/encodec.py:82-94).

    `length` (relative, max must be 1) zeroes the padded tail like the reference's padding_mask
    (HF/encodec:599-601).
    """
    sig = resample(sig, sample_rate, orig_sample_rate)
    nq = num_quantizers_for(num_codebooks)
    if length is not None:
        abs_lens = sig.shape[-1] * length
        mask = torch.arange(int(abs_lens.max().long()), dtype=length.dtype)[None] < abs_lens[:, None]
        sig = mask * sig
    emb = encoder(sd, sig[:, None])
    out = rvq_encode(sd, emb, nq, return_gaps)
    if return_gaps:
        return out[0].permute(1, 2, 0), out[1].permute(1, 2, 0), emb
    return out.permute(1, 2, 0)  # [B, N, K]


def toks_to_qfeats(sd, toks):
    """R/audiocodecs/encodec.py:144-149: [B,N,K] -> [B,N,128]."""
    return rvq_decode(sd, toks.movedim(-1, 0)).movedim(-1, -2)


def toks_to_sig(sd, toks, sample_rate=24000, orig_sample_rate=24000):
    """Codec.toks_to_sig -> Encodec._toks_to_sig (R/codec.py:90-100, R/encodec.py:130-141)."""
    z = rvq_decode(sd, toks.long().movedim(-1, 0))
    sig = decoder(sd, z)[:, 0]
    return resample(sig, orig_sample_rate, sample_rate)
