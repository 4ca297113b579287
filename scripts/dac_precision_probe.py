"""DAC bf16 path: SI-SNR / latent error for several split policies (which activations carry a lo plane)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import audiocodecs_b200 as A
from helpers import make_input, si_snr_db
from oracle import dac_ref, weights

sd = weights.dac_state_dict(0)
dev = torch.device("cuda:0")
toks = torch.randint(0, 1024, (2, 43, 9), generator=torch.Generator().manual_seed(6))
sig = make_input(41, 2, 22050)
with torch.no_grad():
    ref_rec = dac_ref.toks_to_sig(sd, toks)
    ref_z = dac_ref.encoder(sd, sig[:, None]).permute(0, 2, 1)
for smin, rmin in ((100000, 64), (100000, 128), (256, 64), (256, 128), (512, 64), (128, 64)):
    codec = A.DAC(44100, 44100, num_codebooks=9, state_dict=sd, precision="bf16", split_min_ch=smin, split_res_min_ch=rmin).eval().to(dev)
    rec = codec.toks_to_sig(toks.to(dev)).cpu()
    z = codec.sig_to_feats(sig.to(dev)).cpu()
    print(f"split_min_ch={smin} split_res_min_ch={rmin}: decoder SI-SNR {si_snr_db(ref_rec, rec):.1f} dB, encoder latent rel-err {((z - ref_z).norm() / ref_z.norm()).item():.3e}", flush=True)
