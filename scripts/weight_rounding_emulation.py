"""CPU emulation (no GPU): how much decoder SI-SNR does rounding the WEIGHTS of a layer group to a single bf16 plane cost,
with everything else in fp32?  The folded weight W = g v / |v| of the selected layers is rounded to bf16 and written back
as (g' = |W_r|, v' = W_r) so the oracle reproduces W_r exactly.  The noise of a group adds (in power) to the measured noise
of the tensor path (EnCodec: 45.1 dB -> 3.1e-5), which predicts the SI-SNR of a candidate `W_SINGLE` policy before spending
GPU time on scripts/weight_precision_probe.py.
Validated against the GPU: EnCodec residual-block k3 convs 43.9 dB predicted / 44.0 measured, 1x1 tails 42.3 / 42.3, the whole
decoder 39.7 / 39.9.
Usage: python scripts/weight_rounding_emulation.py encodec|dac|mimi [regex ...]     (decoder weights; regex over state-dict keys)"""
import math, os, re, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import si_snr_db
from oracle import dac_ref, encodec_ref, mimi_ref, weights

which = sys.argv[1] if len(sys.argv) > 1 else "encodec"
if which == "encodec":
    sd0, dec, base_db = weights.encodec_state_dict(0), encodec_ref.toks_to_sig, 45.1
    toks = torch.randint(0, 1024, (2, 75, 8), generator=torch.Generator().manual_seed(6))
    default = [r"^decoder\.layers\.(0|15)\.", r"^decoder\.layers\.(3|6|9|12)\.conv", r"block\.1", r"block\.3|shortcut", r"^decoder"]
elif which == "dac":   # base = measured SI-SNR of the tensor path with the (hi, lo) pair everywhere
    sd0, dec, base_db = weights.dac_state_dict(0), dac_ref.toks_to_sig, 47.0
    toks = torch.randint(0, 1024, (2, 43, 9), generator=torch.Generator().manual_seed(6))
    default = [r"res_unit.\.conv1", r"res_unit.\.conv2", r"conv_t1", r"^decoder\.conv", r"^decoder"]
else:
    sd0, dec, base_db = weights.mimi_state_dict(0), mimi_ref.toks_to_sig, 51.8
    toks = torch.randint(0, 2048, (2, 13, 8), generator=torch.Generator().manual_seed(8))
    default = [r"^decoder_transformer", r"^decoder\.layers.*block", r"^decoder\.layers\.\d+\.conv", r"^decoder"]
pats = sys.argv[2:] or default
with torch.no_grad():
    ref = dec(sd0, toks)
base_noise = 10 ** (-base_db / 10)
for pat in pats:
    sd = dict(sd0)
    n = 0
    for k in list(sd0):
        if not (k.startswith("decoder") and re.search(pat, k)):
            continue
        if k.endswith("parametrizations.weight.original1"):     # weight-normed (EnCodec): fold, round, write back as (|W_r|, W_r)
            g, v = sd0[k.replace("original1", "original0")].float(), sd0[k].float()
            dims = tuple(range(1, v.dim()))
            wr = (g * v / v.pow(2).sum(dims, keepdim=True).sqrt()).to(torch.bfloat16).float()
            sd[k] = wr
            sd[k.replace("original1", "original0")] = wr.pow(2).sum(dims, keepdim=True).sqrt()
            n += 1
        elif k.endswith(".weight") and sd0[k].dim() >= 2 and "layernorm" not in k and "codebook" not in k:
            sd[k] = sd0[k].float().to(torch.bfloat16).float()
            n += 1
    with torch.no_grad():
        got = dec(sd, toks)
    snr = si_snr_db(ref, got)
    total = -10 * math.log10(base_noise + 10 ** (-snr / 10))
    print(f"{pat!r:45s} {n:2d} layers: weight rounding alone {snr:5.1f} dB -> predicted tensor-path SI-SNR {total:4.1f} dB", flush=True)
