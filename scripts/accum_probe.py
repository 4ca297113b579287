"""Does the tcgen05 fp32 accumulator round or truncate?  Pointwise GEMMs of growing contraction length K through ac_conv_tc
with (fp16 hi, bf16 lo) x (fp16 hi, fp16 lo, bf16 hi) operands (operand rounding ~2^-20): relative error vs fp64, and the same
K done as 4 / 16 separate launches whose fp32 outputs are added in torch (round-to-nearest) -- if the error falls with the
number of accumulations per accumulator, the accumulation (not the operands) is what limits precision."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audiocodecs_b200 import tc
from audiocodecs_b200.tc import Src
DEV = "cuda:0"
FULL = tc.Policy(True, 0, True, True)
ONE = tc.Policy(True, tc.NEVER, False)
g = torch.Generator().manual_seed(1)
M, N = 4096, 128
for pol, pname, flush in ((FULL, "3 products", 0), (FULL, "3 products, chunked accumulation", 24), (ONE, "1 product", 0)):
    tc.FLUSH_ADDS = flush
    for K in (64, 256, 1024, 4096):
        for mean in (0.0, 1.0):
            x = torch.randn(1, M, K, generator=g) + mean       # mean != 0: all-positive-ish partial sums (ELU-like common mode)
            w = (torch.randn(N, K, generator=g) + mean) * K ** -0.5
            def run(xs, ws):
                a = pol.act(1, M, xs.shape[2], DEV)
                hi = xs.to(torch.float16); a.buf[:] = hi.to(DEV)
                val = hi.double()
                if a.lo is not None:
                    lo = (xs - hi.float()).to(torch.bfloat16); a.lo[:] = lo.to(DEV); val = val + lo.double()
                W = pol.weights(ws, None); W.apply(lambda t: t.to(DEV))
                wv = W.w[0].cpu().view(torch.float16).double() + (W.w[1].cpu().view(torch.float16).double() if W.split else 0)
                y = torch.empty((1, M, N), device=DEV, dtype=torch.float32)
                tc.conv_tc(W, [Src(a)], M, y32=y)
                torch.cuda.synchronize()
                return y.cpu().double(), val[0] @ wv.t()   # result, fp64 product of what the planes hold
            y, ref_planes = run(x, w)
            ref = x[0].double() @ w.double().t()
            e_all = ((y[0] - ref).norm() / ref.norm()).item()
            e_acc = ((y[0] - ref_planes).norm() / ref_planes.norm()).item()   # accumulation-only error (+ dropped lo*lo)
            bias = ((y[0] - ref_planes) / ref_planes.abs().clamp_min(1e-3)).mean().item()
            out = f"{pname} K={K:5d} mean={mean}: err vs fp64 {e_all:.2e}, vs the planes' own product {e_acc:.2e}, mean signed rel err {bias:+.2e}"
            for S in (4, 16):
                if K // S >= 64:
                    acc = sum(run(x[:, :, i * (K // S):(i + 1) * (K // S)].contiguous(), w[:, i * (K // S):(i + 1) * (K // S)].contiguous())[0] for i in range(S))
                    out += f" | K split in {S}: {((acc[0] - ref).norm() / ref.norm()).item():.2e}"
            print(out, flush=True)
