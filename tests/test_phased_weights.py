"""Host logic of the single-channel edge layers run as tap-GEMMs over 16-sample rows (audiocodecs_b200/tc.py:
last_conv_weights_phased / first_conv_weights_phased): the Toeplitz weight matrices, multiplied in plain torch over the
same 16-phase views the kernels read, must reproduce F.conv1d.  CPU only (no kernel is launched)."""
import pytest
import torch
import torch.nn.functional as F

from audiocodecs_b200 import ops, tc

P = tc.LAST_PHASES


def _spec(w_torch, bias, geometry="causal", padding=0):
    # ops.ConvSpec packs Conv1d weights [Cout, Cin, K] as [K, Cin, Cout]
    return ops.ConvSpec(w_torch.permute(2, 1, 0).contiguous(), bias, cout=w_torch.shape[0], geometry=geometry, padding=padding)


@pytest.mark.parametrize("C,taps,pad_left,L", [(32, 7, 6, 320), (96, 7, 3, 512), (64, 3, 2, 1920)])
def test_last_conv_toeplitz_matches_conv1d(C, taps, pad_left, L):
    g = torch.Generator().manual_seed(C)
    w = torch.randn(1, C, taps, generator=g)
    b = torch.randn(1, generator=g)
    x = torch.randn(2, L, C, generator=g)                                   # channels-last activation
    ref = F.conv1d(F.pad(x.permute(0, 2, 1), (pad_left, taps - 1 - pad_left)), w, b)[:, 0]   # [B, L]
    W = tc.last_conv_weights_phased(_spec(w, b))
    Wf = (W.w[0].float() + W.w[1].float()) if W.split else W.w.float()      # hi + lo planes
    hr = (-(pad_left + L)) % P or P
    buf = F.pad(x, (0, 0, pad_left, hr))                                    # [B, pad_left + L + hr, C], zero halos
    rows = buf.shape[1] // P
    view = buf.reshape(2, rows, P * C)                                      # the 16-phase view
    a = torch.cat([view[:, :-1], view[:, 1:]], dim=-1)[:, : L // P]         # taps = 2: view rows n, n+1
    got = (a @ Wf.t() + W.bias).reshape(2, -1)
    assert got.shape == ref.shape
    assert (got - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()   # weights are bf16 hi + lo: ~2^-17 relative


@pytest.mark.parametrize("C,taps,pad_left,T", [(32, 7, 6, 333), (64, 7, 3, 1000), (64, 7, 6, 16)])
def test_first_conv_toeplitz_matches_conv1d(C, taps, pad_left, T):
    g = torch.Generator().manual_seed(T)
    w = torch.randn(C, 1, taps, generator=g)
    b = torch.randn(C, generator=g)
    x = torch.randn(2, T, generator=g)
    ref = F.conv1d(F.pad(x[:, None], (pad_left, taps - 1 - pad_left)), w, b).permute(0, 2, 1)   # [B, T, C]
    W, tv = tc.first_conv_weights_phased(_spec(w, b, geometry="causal" if pad_left == taps - 1 else "same", padding=pad_left), pad_left)
    assert tv == (2 if pad_left == taps - 1 else 3)
    Wf = W.w[0].float() + W.w[1].float()
    R = -(-T // P)
    flat = F.pad(x, (P, (R + 1) * P - T))                                   # one padding row in front (zeros here), zeros behind
    view = flat.reshape(2, R + 2, P)
    a = torch.cat([view[:, k:k + R] for k in range(tv)], dim=-1)            # view rows n-1, n (, n+1) relative to the data
    got = (a @ Wf.t() + W.bias).reshape(2, R * P, C)[:, :T]
    assert (got - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()
