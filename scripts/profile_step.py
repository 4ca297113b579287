"""Two full-size steps for ncu: step 1 warms up, step 2 is the profiled one.
Usage under ncu: python scripts/profile_step.py [steps] [encodec|dac|mimi] [batch]   (see profiles/README.md)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import audiocodecs_b200 as A
from oracle import weights

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
which = sys.argv[2] if len(sys.argv) > 2 else "encodec"
dev = torch.device("cuda:0")
prec = os.environ.get("AC_PRECISION", "exact")
if which == "encodec":
    codec, sr, B = A.Encodec(24000, 24000, num_codebooks=8, state_dict=weights.encodec_state_dict(0), precision=prec), 24000, 64
elif which == "dac":
    codec, sr, B = A.DAC(44100, 44100, num_codebooks=9, state_dict=weights.dac_state_dict(0), precision=prec), 44100, 64
else:
    codec, sr, B = A.Mimi(24000, num_codebooks=8, state_dict=weights.mimi_state_dict(0), precision=prec), 24000, 128
B = int(sys.argv[3]) if len(sys.argv) > 3 else B
codec = codec.eval().to(dev)
sig = (torch.randn(B, sr * 10, generator=torch.Generator().manual_seed(999)) * 0.1).to(dev)
for i in range(steps):
    if i == steps - 1:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()  # with `ncu --profile-from-start off` only the last (tuned) step is captured
    rec = codec.toks_to_sig(codec.sig_to_toks(sig))
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", tuple(rec.shape))
