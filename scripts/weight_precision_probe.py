"""Which layers can store their weights as ONE bf16 plane (one tensor-core product fewer per MAC)?
For every candidate regex over state-dict prefixes (`w_single=`): decoder SI-SNR on the oracle's codes (small case, as in
tests/), encoder agreement with the oracle (code match / embedding error), and the full-size step time.
Usage: python scripts/weight_precision_probe.py dac|mimi|encodec [regex ...]      ("-" = no layer, ".*" = every layer)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import audiocodecs_b200 as A
from helpers import make_input, si_snr_db
from oracle import dac_ref, encodec_ref, mimi_ref, weights

which = sys.argv[1]
pats = sys.argv[2:] or ["-", ".*"]
dev = torch.device("cuda:0")
if which == "dac":
    sd, sr, K, V, Bfull, frames = weights.dac_state_dict(0), 44100, 9, 1024, 64, 43
    make = lambda pat: A.DAC(sr, sr, num_codebooks=K, state_dict=sd, precision="bf16", w_single=pat)
    ref_dec = lambda t: dac_ref.toks_to_sig(sd, t)
    ref_enc = lambda s: dac_ref.sig_to_toks(sd, s, K)
elif which == "mimi":
    sd, sr, K, V, Bfull, frames = weights.mimi_state_dict(0), 24000, 8, 2048, 128, 13
    make = lambda pat: A.Mimi(sr, num_codebooks=K, state_dict=sd, precision="bf16", w_single=pat)
    ref_dec = lambda t: mimi_ref.toks_to_sig(sd, t)
    ref_enc = lambda s: mimi_ref.sig_to_toks(sd, s, K)
else:
    sd, sr, K, V, Bfull, frames = weights.encodec_state_dict(0), 24000, 8, 1024, 64, 75
    make = lambda pat: A.Encodec(sr, sr, num_codebooks=K, state_dict=sd, precision="bf16", w_single=pat)
    ref_dec = lambda t: encodec_ref.toks_to_sig(sd, t)
    ref_enc = lambda s: encodec_ref.sig_to_toks(sd, s, K)
toks = torch.randint(0, V, (2, frames, K), generator=torch.Generator().manual_seed(6))
sig_small = make_input(998, 4, 2 * sr)
with torch.no_grad():
    wave_ref, toks_ref = ref_dec(toks), ref_enc(sig_small)
sig_full = (torch.randn(Bfull, sr * 10, generator=torch.Generator().manual_seed(999)) * 0.1).to(dev)
for pat in pats:
    codec = make(None if pat == "-" else pat).eval().to(dev)
    snr = si_snr_db(wave_ref, codec.toks_to_sig(toks.to(dev)).cpu())
    match = (codec.sig_to_toks(sig_small.to(dev)).cpu() == toks_ref).float().mean().item()
    for _ in range(2):
        codec.toks_to_sig(codec.sig_to_toks(sig_full))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        codec.toks_to_sig(codec.sig_to_toks(sig_full))
    e1.record()
    torch.cuda.synchronize()
    print(f"{which} w_single={pat!r:50s} decoder SI-SNR {snr:5.1f} dB  code match {match:.4f}  step {e0.elapsed_time(e1) / 3:8.2f} ms", flush=True)
    del codec
    torch.cuda.empty_cache()
