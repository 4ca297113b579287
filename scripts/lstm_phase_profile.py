"""Per-phase cycle profile of one step of the cluster LSTM kernel (clock64 samples from CTA 0)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audiocodecs_b200 import ops
from audiocodecs_b200.tc import Act

dev = "cuda:0"
B, T, C = 64, 750, 512
pre = torch.randn(B, T, 4 * C, device=dev)
w = (torch.randn(4 * C, C, device=dev) * 0.04).to(torch.bfloat16)
out = Act(B, T, C, dev, split=True)
dbg = torch.zeros(T, 8, dtype=torch.int64, device=dev)
for _ in range(2):
    ops.lstm_tc(pre, w, out=out, dbg=dbg)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.lstm_tc(pre, w, out=out); e1.record(); torch.cuda.synchronize()
print("kernel ms", e0.elapsed_time(e1), "us/step", e0.elapsed_time(e1) * 1e3 / T)
d = dbg.cpu()[100:700].double()
names = ["epi:top->d_full", "epi:ld+act+bar", "epi:cell update", "epi:bar+remote st+fence+bar", "epi:arrive->next top",
         "mma:h_ready->commit", "mma:commit->next h_ready"]
step = (d[1:, 0] - d[:-1, 0]).mean().item()
print("cycles/step", step)
print(names[0], (d[:, 1] - d[:, 0]).mean().item())
print(names[1], (d[:, 2] - d[:, 1]).mean().item())
print(names[2], (d[:, 3] - d[:, 2]).mean().item())
print(names[3], (d[:, 4] - d[:, 3]).mean().item())
print(names[4], (d[1:, 0] - d[:-1, 4]).mean().item())
print(names[5], (d[:, 6] - d[:, 5]).mean().item())
print(names[6], (d[1:, 5] - d[:-1, 6]).mean().item())
print("epi d_full seen - mma commit issued", (d[:, 1] - d[:, 6]).mean().item())
print("mma h_ready seen - epi arrive done(prev step)", (d[1:, 5] - d[:-1, 4]).mean().item())
