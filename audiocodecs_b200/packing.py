"""Weight packing: backend state dicts -> kernel layouts (done once at construction).

The reference re-derives `w = g*v/||v||` on every forward (18x `_weight_norm_interface` per pass,
HF/encodec:106-111); here weight-norm is folded once and weights are re-laid out as the
`[taps][cin][n_cols]` operand of the tap-GEMM kernels (include/audiocodecs_b200.h).
"""
import torch


def fold_weight_norm(sd, prefix):
    """HF parametrized keys (`original0` = g, `original1` = v), descript `weight_g/weight_v`, or a plain weight."""
    if prefix + ".parametrizations.weight.original0" in sd:
        g = sd[prefix + ".parametrizations.weight.original0"].float()
        v = sd[prefix + ".parametrizations.weight.original1"].float()
    elif prefix + ".weight_g" in sd:
        g, v = sd[prefix + ".weight_g"].float(), sd[prefix + ".weight_v"].float()
    else:
        return sd[prefix + ".weight"].float()
    norm = v.flatten(1).norm(dim=1).view(-1, *([1] * (v.dim() - 1)))
    return g * v / norm  # norm over all dims but 0: per out-channel (Conv1d) / per in-channel (ConvTranspose1d)


def pack_conv(w):
    """Conv1d weight [Cout, Cin, K] -> [K, Cin, Cout]."""
    return w.permute(2, 1, 0).contiguous()


def pack_convtr(w, stride):
    """ConvTranspose1d weight [Cin, Cout, 2s] -> 2-tap operand [2, Cin, s*Cout].

    y[q*s + r - pad] = sum_c x[q-1,c] w[c,co,r+s] + x[q,c] w[c,co,r]: tap 0 pairs with the previous input row.
    """
    cin, cout, k = w.shape
    assert k == 2 * stride, "transposed convs on this path all have kernel = 2*stride"
    wk = w.permute(2, 0, 1)  # [K, Cin, Cout]
    hi = wk[stride:].permute(1, 0, 2).reshape(cin, stride * cout)  # taps r+s
    lo = wk[:stride].permute(1, 0, 2).reshape(cin, stride * cout)  # taps r
    return torch.stack([hi, lo]).contiguous()


def pack_linear(w):
    """nn.Linear / 1x1 weight [out, in] -> [1, in, out]."""
    return w.t().contiguous()[None]
