"""BASELINE.json configs[4]: EnCodec-24k, 32 codebooks (RVQ-depth stress), batch sweep 1..1024 clips of 10 s on one GPU
(the driver's N>1 runs shard clips).  Prints one JSON line per batch size: device-timed encode+decode
throughput, plus the host-CPU oracle on 1 clip for scale.  Usage: python scripts/batch_sweep.py [max_batch] [K]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import audiocodecs_b200 as A
from oracle import encodec_ref, weights

max_b = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
K = int(sys.argv[2]) if len(sys.argv) > 2 else 32
dev = torch.device("cuda:0")
sd = weights.encodec_state_dict(0)
codec = A.Encodec(24000, 24000, num_codebooks=K, state_dict=sd).eval().to(dev)
T = 240000
b = 1
while b <= max_b:
    sig = (torch.randn(b, T, generator=torch.Generator().manual_seed(b)) * 0.1).to(dev)
    for _ in range(2):
        codec.toks_to_sig(codec.sig_to_toks(sig))
    torch.cuda.synchronize()
    reps = max(2, min(20, 512 // b))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        codec.toks_to_sig(codec.sig_to_toks(sig))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(json.dumps({"batch": b, "K": K, "ms_per_step": round(ms, 3), "audio_s_per_s": round(b * 10 / ms * 1e3, 1),
                      "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 2)}), flush=True)
    del sig
    b *= 2
torch.set_num_threads(os.cpu_count())
sig = torch.randn(1, T, generator=torch.Generator().manual_seed(1)) * 0.1
with torch.no_grad():
    encodec_ref.toks_to_sig(sd, encodec_ref.sig_to_toks(sd, sig, K))
    t0 = time.perf_counter()
    encodec_ref.toks_to_sig(sd, encodec_ref.sig_to_toks(sd, sig, K))
    dt = time.perf_counter() - t0
print(json.dumps({"impl": "cpu oracle (reference's ATen ops)", "batch": 1, "K": K, "cores": os.cpu_count(), "audio_s_per_s": round(10 / dt, 2)}))
