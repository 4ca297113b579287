"""Oracle: `torchaudio.functional.resample` restated (TA/functional.py:1305-1490).

Called by the reference at R/audiocodecs/codec.py:59-63 and :95-99 with the
torchaudio defaults (lowpass_filter_width=6, rolloff=0.99, sinc_interp_hann).
"""
import math

import torch


def resample_taps(orig_freq: int, new_freq: int, dtype=torch.float32):
    """Polyphase windowed-sinc taps, [new/g, 2*width + orig/g] (TA/functional.py:1305-1402).

    Index arithmetic is carried out in `dtype` (fp32 for fp32 waveforms, TA:1374-1378).
    """
    g = math.gcd(int(orig_freq), int(new_freq))
    o, n = int(orig_freq) // g, int(new_freq) // g
    lowpass_filter_width, rolloff = 6, 0.99
    base = min(o, n) * rolloff
    width = math.ceil(lowpass_filter_width * o / base)
    idx = torch.arange(-width, width + o, dtype=dtype)[None, None] / o
    t = torch.arange(0, -n, -1, dtype=dtype)[:, None, None] / n + idx
    t = t * base
    t = t.clamp(-lowpass_filter_width, lowpass_filter_width)
    window = torch.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    scale = base / o
    taps = torch.where(t == 0, torch.tensor(1.0, dtype=dtype), t.sin() / t)
    taps = taps * window * scale
    return taps[:, 0, :], width, o, n


def resample(sig: torch.Tensor, orig_freq: int, new_freq: int) -> torch.Tensor:
    """sig [B, T] -> [B, ceil(new*T/orig)] (TA/functional.py:1405-1432, 1435-1490)."""
    if orig_freq == new_freq:
        return sig
    taps, width, o, n = resample_taps(orig_freq, new_freq, sig.dtype)
    B, T = sig.shape
    x = torch.nn.functional.pad(sig, (width, width + o))
    y = torch.nn.functional.conv1d(x[:, None], taps[:, None, :], stride=o)  # [B, n, frames]
    y = y.transpose(1, 2).reshape(B, -1)
    # TA:1426 rounds new*T/orig to fp32 (torch.as_tensor of a python float) before the ceil
    target = int(torch.ceil(torch.as_tensor(n * T / o)).long())
    return y[:, :target]
