"""Where does the embedding error of the tensor-path EnCodec encoder come from?  Runs the encoder stage by stage on the GPU
(the same calls as Encodec._encoder_tc) and compares every intermediate tensor with the fp32 oracle's (relative L2 error).
Usage: [AC_PRECISION=exact|fp16|bf16] python scripts/encoder_error_probe.py [clips=2] [seconds=4]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import audiocodecs_b200 as A
from audiocodecs_b200 import ops, tc
from audiocodecs_b200.ops import ACT_ELU, PAD_REFLECT
from audiocodecs_b200.tc import Src
from oracle import encodec_ref as R, weights

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 4.0
prec = os.environ.get("AC_PRECISION", "exact")
dev = torch.device("cuda:0")
sd = weights.encodec_state_dict(0)
codec = A.Encodec(24000, 24000, num_codebooks=8, state_dict=sd, precision=prec).eval().to(dev)
sig = torch.randn(B, int(24000 * secs), generator=torch.Generator().manual_seed(77)) * 0.1

# ---- oracle intermediates (fp64 for a clean reference)
sd64 = {k: v.double() for k, v in sd.items()}
ref = {}
with torch.no_grad():
    x = R.causal_conv(sig.double()[:, None], *R.fold_weight_norm(sd64, "encoder.layers.0")); ref["conv_first"] = x
    idx = 1
    for i, r in enumerate(reversed(R.RATIOS)):
        x = R.resblock(sd64, f"encoder.layers.{idx}", x); ref[f"res{i}"] = x
        x = R.causal_conv(F.elu(x), *R.fold_weight_norm(sd64, f"encoder.layers.{idx + 2}"), stride=r); ref[f"down{i}"] = x
        idx += 3
    x = R.lstm_block(sd64, f"encoder.layers.{idx}", x); ref["lstm"] = x
    x = R.causal_conv(F.elu(x), *R.fold_weight_norm(sd64, f"encoder.layers.{idx + 2}")); ref["emb"] = x


def rel(name, got, what="raw"):
    r = ref[name].transpose(1, 2)
    if what == "elu":
        r = F.elu(r)
    d = got.double().cpu() - r
    e = (d.norm() / r.norm()).item()
    alpha = ((d * r).sum() / (r * r).sum()).item()          # multiplicative bias: the part of the error that is a pure scaling
    resid = ((d - alpha * r).norm() / r.norm()).item()
    print(f"{name:12s} {what:4s} rel err {e:.2e}  (scaling part {alpha:+.2e}, rest {resid:.2e})", flush=True)


pol = codec.pol_enc
s = sig.to(dev)
T = s.shape[1]
x = pol.act(B, T, 32, dev)
xe = pol.act(B, T, 32, dev, hl=2)
ops.conv_first_bf16(codec._enc[0], s, y=x, y_act=xe, act=ACT_ELU)
rel("conv_first", x.value()); rel("conv_first", xe.value(), "elu")
L = T
for i, ((Wk3, Wtail), Wdown, r) in enumerate(codec._tenc):
    C = x.C
    Lout = -(-L // r)
    extra = Lout * r - L
    ye = pol.act(B, L, C, dev, hl=r, hr=extra)
    codec._tc_resblock_run(Wk3, Wtail, x, xe, ye, pol)
    rel(f"res{i}", ye.value(), "elu")
    ye.fill_halo(PAD_REFLECT, max(r, extra) + 1 if L <= max(r, extra) else 0)
    last = i == len(codec._tenc) - 1
    x = pol.act(B, Lout, 2 * C, dev)
    xe = None if last else pol.act(B, Lout, 2 * C, dev, hl=2)
    tc.conv_tc(Wdown, [Src(ye, taps=2, origin=-r, phases=r, rows=Lout + 1)], Lout, y=x, y_act=xe, act=ACT_ELU, name="down_tc")
    rel(f"down{i}", x.value())
    L = Lout
le = pol.act(B, L, x.C, dev, hl=6)
codec._tc_run_lstm(codec._tenc_lstm, [n for _, n in codec._enc_lstm], x, le, pol)
rel("lstm", le.value(), "elu")
le.fill_halo(PAD_REFLECT, 7 if L <= 6 else 0)
emb = torch.empty((B, L, 128), device=dev, dtype=torch.float32)
tc.conv_tc(codec._tenc_last, [Src(le, taps=7, origin=-6, rows=L + 6)], L, y32=emb, name="conv_k7_tc")
rel("emb", emb)
# the LSTM alone, fed the ORACLE's input (isolates the recurrence kernel from upstream error)
xin = pol.act(B, L, 512, dev)
ops.f32_to_act(ref["down3"].transpose(1, 2).float().contiguous().to(dev), xin)
le2 = pol.act(B, L, 512, dev, hl=6)
codec._tc_run_lstm(codec._tenc_lstm, [n for _, n in codec._enc_lstm], xin, le2, pol)
rel("lstm", le2.value(), "elu")
print("(last line: LSTM block fed the oracle's input)")
