"""The reference's `profile()` protocol (R/downstream/profiler.py:52-203, called at R/downstream/test_sr.py:379-391) on this
package's codecs: batch 1, inputs of 1, 2, 4, 8, 16, 32 s at 16 kHz (hparams sample_rate), 20 warm-up passes over all shapes,
then per shape the mean wall-clock latency of 10 `codec(inputs)` calls (mode="reconstruct") with a synchronize on both sides,
the mean peak device memory, the MACs of one call, and the cross-check with torch.utils.benchmark.Timer.  Same result keys as
the reference.  MACs = 0.5 x the algorithmic FLOPs the launch helpers count (2 FLOP per MAC of the reference's dense ops) --
the reference counts them by patching torch.nn.functional, which sees none of this package's kernels.
Usage: python scripts/profile_protocol.py [encodec|dac|mimi ...] [--precision exact] [--graph]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.utils.benchmark import Timer
import audiocodecs_b200 as A
from audiocodecs_b200 import ops
from oracle import weights

args = [a for a in sys.argv[1:] if not a.startswith("--")] or ["encodec", "dac", "mimi"]
precision = sys.argv[sys.argv.index("--precision") + 1] if "--precision" in sys.argv else "exact"
SR = 16000
dev = torch.device("cuda:0")


def profile(model, input_shapes, num_runs=10, num_warmups=20, seed=0):
    torch.manual_seed(seed)
    model = model.to(dev).eval()
    for _ in range(num_warmups):
        for shape in input_shapes:
            with torch.inference_mode():
                model(torch.randn(shape, device=dev))
    res = {"time (s)": [], "memory (GB)": [], "macs (GMACS)": [], "time (s) (PyTorch timer)": []}
    for shape in input_shapes:
        total, peak = 0.0, 0.0
        for _ in range(num_runs):
            torch.cuda.reset_peak_memory_stats()
            x = torch.randn(shape, device=dev)
            with torch.inference_mode():
                torch.cuda.synchronize()
                t0 = time.time()
                model(x)
                torch.cuda.synchronize()
                total += time.time() - t0
                peak += torch.cuda.max_memory_allocated(dev) / 10**9
        res["time (s)"].append(total / num_runs)
        res["memory (GB)"].append(peak / num_runs)
        prof = ops.Profiler()
        ops.set_profiler(prof)
        with torch.inference_mode():
            model(x)
        torch.cuda.synchronize()
        ops.set_profiler(None)
        res["macs (GMACS)"].append(0.5 * sum(r[3] for r in prof.records) / 1e9)
        t = Timer(stmt="with torch.inference_mode():\n    model(x)", globals={"model": model, "x": x, "torch": torch})
        res["time (s) (PyTorch timer)"].append(t.timeit(num_runs).mean)
    n = sum(b.numel() for b in model.buffers()) + sum(sum(t.numel() for t in (getattr(s, "w", None), getattr(s, "bias", None)) if t is not None)
                                                       for s in model._specs)
    res["trainable_params (M)"] = 0.0
    res["total_params (M)"] = n / 1e6
    return res


for name in args:
    if name == "encodec":
        codec = A.Encodec(SR, 24000, num_codebooks=8, state_dict=weights.encodec_state_dict(0), precision=precision)
    elif name == "dac":
        codec = A.DAC(SR, 44100, num_codebooks=9, state_dict=weights.dac_state_dict(0), precision=precision)
    else:
        codec = A.Mimi(SR, num_codebooks=8, state_dict=weights.mimi_state_dict(0), precision=precision)
    r = profile(codec, [(1, SR * s) for s in (1, 2, 4, 8, 16, 32)])
    r = {k: ([round(v, 6) for v in vals] if isinstance(vals, list) else round(vals, 3)) for k, vals in r.items()}
    print(json.dumps({"codec": name, "precision": precision, "input_seconds": [1, 2, 4, 8, 16, 32], "sample_rate": SR, **r}), flush=True)
