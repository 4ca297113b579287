"""Two full-size steps (EnCodec-24k, 64 x 10 s, encode+decode) for ncu: step 1 warms up, step 2 is the profiled one.
Usage under ncu: see profiles/README.md."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import audiocodecs_b200 as A
from oracle import weights

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda:0")
codec = A.Encodec(24000, 24000, num_codebooks=8, state_dict=weights.encodec_state_dict(0)).eval().to(dev)
sig = (torch.randn(64, 240000, generator=torch.Generator().manual_seed(999)) * 0.1).to(dev)
for _ in range(steps):
    rec = codec.toks_to_sig(codec.sig_to_toks(sig))
torch.cuda.synchronize()
print("done", tuple(rec.shape))
