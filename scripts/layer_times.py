"""Per-launch device times of one full-size step (CUDA events around every C-ABI launch), with the algorithmic
GB/s and TFLOP/s of each launch.  Usage: [AC_PRECISION=exact|fp16|bf16] [AC_TUNE_FUSION=1] python scripts/layer_times.py
[encodec|dac|mimi] [batch] [seconds]      (AC_TUNE_FUSION=1: let the tuner time fused against unfused blocks and print it)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import audiocodecs_b200 as A
from audiocodecs_b200 import ops
from oracle import weights

which = sys.argv[1] if len(sys.argv) > 1 else "encodec"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
secs = float(sys.argv[3]) if len(sys.argv) > 3 else 10.0
smin = int(sys.argv[4]) if len(sys.argv) > 4 else 512
dev = torch.device("cuda:0")
prec = os.environ.get("AC_PRECISION", "exact")
if os.environ.get("AC_TUNE_FUSION"):
    from audiocodecs_b200 import encodec as _enc
    _enc.FUSED_MAX_CH = None
if which in ("encodec", "encodec32"):
    codec, sr = A.Encodec(24000, 24000, num_codebooks=32 if which == "encodec32" else 8, state_dict=weights.encodec_state_dict(0), precision=prec), 24000
elif which == "dac":
    codec, sr = A.DAC(44100, 44100, num_codebooks=9, state_dict=weights.dac_state_dict(0), precision=prec, split_min_ch=smin), 44100
else:
    codec, sr = A.Mimi(24000, num_codebooks=8, state_dict=weights.mimi_state_dict(0), precision=prec), 24000
codec = codec.eval().to(dev)
sig = (torch.randn(B, int(sr * secs), generator=torch.Generator().manual_seed(999)) * 0.1).to(dev)
for _ in range(2):
    rec = codec.toks_to_sig(codec.sig_to_toks(sig))
torch.cuda.synchronize()
prof = ops.Profiler()
ops.set_profiler(prof)
rec = codec.toks_to_sig(codec.sig_to_toks(sig))
torch.cuda.synchronize()
ops.set_profiler(None)
tot = 0.0
for name, s, e, fl, by, label in prof.records:
    ms = s.elapsed_time(e)
    tot += ms
    print(f"{label:16s} {ms:8.3f} ms  {by / ms / 1e6:8.1f} GB/s  {fl / ms / 1e9:8.1f} TFLOP/s  ({by / 1e6:9.1f} MB, {fl / 1e9:9.1f} GFLOP)")
print(f"total {tot:.3f} ms for {B} x {secs} s -> {B * secs / tot * 1e3:.0f} audio-s/s (sum of launches) [{which} {prec}]")
from audiocodecs_b200 import tc
for k, v in tc._TUNED.items():
    print("tuned", k, "->", v)
for k, times in tc.TUNE_LOG:
    print("times", k, {n: round(t, 3) for n, t in times.items()})
