"""Experiment: the 84 SMs the LSTM recurrences leave idle.  The batch is split in two halves that run on two streams; the
recurrence kernels go to high-priority side streams and every persistent tap-GEMM kernel is capped at GRID_CAP CTAs so that a
32-SM recurrence (two clusters of 16) always fits beside it.  Usage: python scripts/overlap_probe.py [precision] [cap ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import audiocodecs_b200 as A
from audiocodecs_b200 import ops, tc
from oracle import weights

prec = sys.argv[1] if len(sys.argv) > 1 else "fp16"
caps = [int(c) for c in sys.argv[2:]] or [0, 116, 132]
dev = torch.device("cuda:0")
codec = A.Encodec(24000, 24000, num_codebooks=8, state_dict=weights.encodec_state_dict(0), precision=prec).eval().to(dev)
sig = (torch.randn(64, 240000, generator=torch.Generator().manual_seed(999)) * 0.1).to(dev)
ref = codec(sig)
torch.cuda.synchronize()


def timed(fn, steps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


print(f"[{prec}] one stream, 64 clips: {timed(lambda: codec(sig)):.2f} ms/step")
for nsplit in (2, 4):
    streams = [torch.cuda.Stream(priority=0) for _ in range(nsplit)]
    sides = [torch.cuda.Stream(priority=-1) for _ in range(nsplit)]
    parts = sig.chunk(nsplit)
    outs = [None] * nsplit

    def step(use_side):
        cur = torch.cuda.current_stream()
        for i, (s, p) in enumerate(zip(streams, parts)):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                if use_side:
                    ops.LSTM_SIDE_STREAM[s.cuda_stream] = sides[i]
                outs[i] = codec(p)
        for s in streams:
            cur.wait_stream(s)

    for cap in caps:
        for use_side in (False, True):
            tc.GRID_CAP = cap
            ops.LSTM_SIDE_STREAM.clear()
            try:
                ms = [timed(lambda: step(use_side)) for _ in range(3)]
                same = torch.equal(torch.cat(outs), ref)
                print(f"[{prec}] {nsplit} streams x {64 // nsplit} clips, grid cap {cap or 148}, LSTM on priority stream {use_side}: "
                      f"{' / '.join(f'{m:.2f}' for m in ms)} ms/step, output identical: {same}", flush=True)
            except Exception as e:  # noqa: BLE001
                print("failed:", cap, use_side, type(e).__name__, str(e)[:200], flush=True)
    tc.GRID_CAP = 0
    ops.LSTM_SIDE_STREAM.clear()
