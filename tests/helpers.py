"""Shared parity helpers (the tolerances of BASELINE.json's north_star are written here once)."""
import torch

NEAR_TIE_REL_GAP = 1e-4      # codes must be bit-exact wherever the top-2 relative distance gap exceeds this
WAVE_MAX_ABS_FP32 = 1e-3     # fp32 path: max-abs waveform error (audio-scale outputs, |x| <~ 1)
WAVE_SISNR_BF16_DB = 40.0    # bf16 path: SI-SNR floor in dB


def make_input(seed, B, T):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, T, generator=g) * 0.1


def si_snr_db(ref, est):
    ref = ref.double().flatten(1)
    est = est.double().flatten(1)
    ref = ref - ref.mean(1, keepdim=True)
    est = est - est.mean(1, keepdim=True)
    s = (est * ref).sum(1, keepdim=True) / ref.pow(2).sum(1, keepdim=True).clamp_min(1e-30) * ref
    n = est - s
    return (10 * torch.log10(s.pow(2).sum(1) / n.pow(2).sum(1).clamp_min(1e-30))).min().item()


def code_report(ours, ref, gaps):
    """(match fraction over safe decisions, near-tie fraction, overall match)."""
    safe = gaps > NEAR_TIE_REL_GAP
    eq = ours.cpu() == ref
    return eq[safe].float().mean().item(), (~safe).float().mean().item(), eq.float().mean().item()


def divergence_report(ours, ref, gaps):
    """Residual quantisation is a chain: once a frame's stage-k code differs, its later stages quantise another residual and
    their codes mean nothing.  So the north_star criterion ("bit-exact wherever the top-2 gap exceeds 1e-4, near-ties
    reported separately") is read per FRAME at its FIRST differing stage: (frames whose first difference is a decision with
    gap > 1e-4 -- violations, frames whose first difference is a near-tie -- excused, frames)."""
    eq = (ours.cpu() == ref).flatten(0, -2)          # [frames, K]
    g = gaps.flatten(0, -2)
    diff = ~eq
    first = torch.where(diff.any(-1), diff.float().argmax(-1), torch.full((eq.shape[0],), -1))
    has = first >= 0
    gap_at = g[torch.arange(eq.shape[0]), first.clamp(min=0)]
    return int((has & (gap_at > NEAR_TIE_REL_GAP)).sum()), int((has & (gap_at <= NEAR_TIE_REL_GAP)).sum()), int(eq.shape[0])
