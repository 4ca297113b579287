"""CPU emulation (no GPU) of the EnCodec encoder in split-bf16 arithmetic: how close do the embeddings and the codes get to
the fp32 oracle when every MMA operand is a (hi, lo) bf16 pair and the products are  A_hi W_hi + A_hi W_lo + A_lo W_hi
(fp32 accumulate)?  Screens a precision policy before it costs GPU time.

    policy "fast"  : what precision="bf16" runs today -- activations with < 128 channels and the LSTM's h / W_hh are single
                     bf16, the LSTM input projection is one product
    policy "exact" : every activation and every weight a (hi, lo) pair, LSTM recurrence in three products as well

Usage: python scripts/exact_mode_emulation.py [seconds=10] [clips=2]"""
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from oracle import encodec_ref, weights


def split(x):
    hi = x.to(torch.bfloat16).float()
    lo = (x - hi).to(torch.bfloat16).float()
    return hi, lo


def make_conv(policy):
    def conv1d(x, w, b=None, stride=1, dilation=1):
        xh, xl = split(x)
        wh, wl = split(w)
        y = F.conv1d(xh, wh, b, stride=stride, dilation=dilation) + F.conv1d(xh, wl, None, stride=stride, dilation=dilation)
        if policy == "exact" or x.shape[1] >= 128:
            y = y + F.conv1d(xl, wh, None, stride=stride, dilation=dilation)
        return y
    return conv1d


def lstm_emul(policy):
    def block(sd, prefix, x, layers=2):
        B, C, T = x.shape
        inp = x.permute(2, 0, 1)
        cur = inp
        for l in range(layers):
            w_ih, w_hh = sd[f"{prefix}.lstm.weight_ih_l{l}"], sd[f"{prefix}.lstm.weight_hh_l{l}"]
            bias = sd[f"{prefix}.lstm.bias_ih_l{l}"] + sd[f"{prefix}.lstm.bias_hh_l{l}"]
            ch, cl = split(cur)
            ih, il = split(w_ih.t().contiguous())
            if policy == "exact":
                pre = ch @ ih + ch @ il + cl @ ih + bias
            else:
                pre = ch @ ih + bias
            hh, hl = split(w_hh.t().contiguous())
            h = x.new_zeros(B, C)
            c = x.new_zeros(B, C)
            outs = []
            for t in range(T):
                a, al = split(h)
                rec = a @ hh + a @ hl + al @ hh if policy == "exact" else a @ hh
                i, f, g, o = (pre[t] + rec).split(C, dim=1)
                c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
                h = torch.sigmoid(o) * torch.tanh(c)
                outs.append(h)
            cur = torch.stack(outs)
        return (cur + inp).permute(1, 2, 0)
    return block


def run(policy, sd, sig):
    shim = types.SimpleNamespace(**{k: getattr(F, k) for k in ("pad", "elu", "embedding", "conv_transpose1d")})
    shim.conv1d = make_conv(policy)
    saved_F, saved_lstm = encodec_ref.F, dict(encodec_ref._LSTM)
    encodec_ref.F = shim
    encodec_ref._LSTM["aten"] = lstm_emul(policy)
    try:
        with torch.no_grad():
            return encodec_ref.encoder(sd, sig[:, None])
    finally:
        encodec_ref.F = saved_F
        encodec_ref._LSTM.update(saved_lstm)


def main():
    secs = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
    clips = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    sd = weights.encodec_state_dict(0)
    sig = torch.randn(clips, int(24000 * secs), generator=torch.Generator().manual_seed(999)) * 0.1
    with torch.no_grad():
        ref_toks, gaps, ref_emb = encodec_ref.sig_to_toks(sd, sig, 32, return_gaps=True)
    safe = gaps > 1e-4
    for policy in ("fast", "exact"):
        emb = run(policy, sd, sig)
        rel = ((emb - ref_emb).norm() / ref_emb.norm()).item()
        with torch.no_grad():
            toks = encodec_ref.rvq_encode(sd, emb, 32).permute(1, 2, 0)
        eq = toks == ref_toks
        per = [eq[..., k].float().mean().item() for k in (0, 3, 7, 15, 31)]
        print(f"{policy:6s} embedding rel-err {rel:.2e}; code match K=8 all {eq[..., :8].float().mean():.4f} safe "
              f"{eq[..., :8][safe[..., :8]].float().mean():.5f}; K=32 all {eq.float().mean():.4f} safe {eq[safe].float().mean():.5f}; "
              f"stages 0/3/7/15/31 {' '.join(f'{p:.4f}' for p in per)}", flush=True)


if __name__ == "__main__":
    main()
