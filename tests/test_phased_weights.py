"""Host logic of the Cout = 1 last layers run as tap-GEMMs over 16-sample rows (audiocodecs_b200/tc.py:
last_conv_weights_phased): the Toeplitz weight matrix, multiplied in plain torch over the same 16-phase view the kernel
reads, must reproduce F.conv1d.  CPU only (no kernel is launched)."""
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
