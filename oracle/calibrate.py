"""Writes the codebook-moment fixtures tests/golden/*_moments.pt (run once, here; committed).

Matched-moment recipe, SURVEY.md section 8c: run a fixed calibration batch through the oracle
encoder, and for every RVQ stage record the per-dimension mean/std of the stage's residual, drawing
the stage codebook as mu + sigma*randn before moving on.  Usage: python -m oracle.calibrate [encodec|mimi]
"""
import sys

import torch

from . import weights


def calibrate_encodec(seed=0, calib_seed=1234, clips=4, seconds=5, stages=32):
    import os
    from . import encodec_ref as ref
    sd = weights.encodec_state_dict(seed, codebooks=False)
    g = torch.Generator().manual_seed(calib_seed)
    sig = torch.randn(clips, 24000 * seconds, generator=g) * 0.1
    with torch.no_grad():
        emb = ref.encoder(sd, sig[:, None])  # [B,128,N]
    res = emb.permute(0, 2, 1).reshape(-1, emb.shape[1])
    mus, sigmas = [], []
    for k in range(stages):
        mu, sigma = res.mean(0), res.std(0)
        mus.append(mu)
        sigmas.append(sigma)
        E = mu[None] + sigma[None] * torch.randn(1024, res.shape[1], generator=g)
        d = torch.cdist(res, E)
        res = res - E[d.argmin(1)]
    out = {"mu": torch.stack(mus), "sigma": torch.stack(sigmas)}
    path = os.path.join(weights.GOLDEN_DIR, "encodec_moments.pt")
    torch.save(out, path)
    # second pass: measure the K=8 decoder output scale with the final codebooks
    sd = weights.encodec_state_dict(seed)
    with torch.no_grad():
        toks = ref.sig_to_toks(sd, sig[:1, :48000], 8)
        rec = ref.toks_to_sig(sd, toks)
    out["out_scale"] = torch.tensor(0.1 / float(rec.std()))
    torch.save(out, path)
    print("decoder out std", float(rec.std()), "-> out_scale", float(out["out_scale"]))
    print("encodec moments", out["mu"].shape, "emb std", float(emb.std()), "mean", float(emb.mean()))





def calibrate_mimi(seed=0, calib_seed=1234, clips=4, seconds=8):
    import os
    import torch.nn.functional as F
    from . import mimi_ref as ref
    sd = weights.mimi_state_dict(seed, codebooks=False)
    g = torch.Generator().manual_seed(calib_seed)
    sig = torch.randn(clips, 24000 * seconds, generator=g) * 0.1
    with torch.no_grad():
        emb = ref.sig_to_feats(sd, sig).movedim(-1, -2)  # [B,512,N]
    out = {}
    for which, n in (("semantic", 1), ("acoustic", 31)):
        res = F.conv1d(emb, sd[f"quantizer.{which}_residual_vector_quantizer.input_proj.weight"])
        res = res.permute(0, 2, 1).reshape(-1, 256)
        mus, sigmas = [], []
        for k in range(n):
            mu, sigma = res.mean(0), res.std(0)
            mus.append(mu)
            sigmas.append(sigma)
            E = mu[None] + sigma[None] * torch.randn(2048, 256, generator=g)
            res = res - E[torch.cdist(res, E).argmin(1)]
        out[which + "_mu"], out[which + "_sigma"] = torch.stack(mus), torch.stack(sigmas)
    path = os.path.join(weights.GOLDEN_DIR, "mimi_moments.pt")
    torch.save(out, path)
    sd = weights.mimi_state_dict(seed)
    with torch.no_grad():
        toks = ref.sig_to_toks(sd, sig[:1, :96000], 8)
        rec = ref.toks_to_sig(sd, toks)
    out["out_scale"] = torch.tensor(0.1 / float(rec.std()))
    torch.save(out, path)
    print("mimi: emb std", float(emb.std()), "decoder out std", float(rec.std()), "-> out_scale", float(out["out_scale"]))


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "encodec"
    {"encodec": calibrate_encodec, "mimi": calibrate_mimi}[which]()
