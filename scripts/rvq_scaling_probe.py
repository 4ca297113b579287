"""Per-stage time of the tensor-core RVQ kernel vs the number of concurrently running CTAs (one 128-frame tile each):
flat => per-SM bound (tensor pipe / epilogue); growing => a shared resource (L2 -> SM codebook streaming)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import audiocodecs_b200 as A
from audiocodecs_b200 import ops
from oracle import weights

dev = torch.device("cuda:0")
codec = A.Encodec(24000, 24000, num_codebooks=32, state_dict=weights.encodec_state_dict(0)).eval().to(dev)
for ctas in (1, 8, 37, 74, 148, 296):
    rows = 128 * ctas
    x = torch.randn(rows, 128, device=dev) * 0.03
    toks = torch.empty((rows, 32), device=dev, dtype=torch.int64)
    for _ in range(2):
        ops.rvq_encode_tc(x, codec.cb_split, codec.codebooks, codec.cb_norm, toks, 32)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.rvq_encode_tc(x, codec.cb_split, codec.codebooks, codec.cb_norm, toks, 32)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    rounds = -(-ctas // 148)
    print(f"tiles {ctas:4d}: {ms:7.3f} ms  -> {ms * 1e3 / 32 / rounds:6.2f} us per stage per round")
