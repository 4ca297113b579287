"""The reference's own speed metric (R/downstream/test_sr.py:56-59,82-86,264-270): real-time factor of encode and decode at
batch 1 -- RTF = (t_enc + t_dec) / audio_seconds, reported with its inverse -- here with CUDA-event timing of our codecs,
for the three codecs, with the downstream sample rate of 16 kHz (R/downstream/hparams/datasets/librispeech-test.yaml:9)
so that the polyphase FIR resampler (identity at the BASELINE configs) is on the timed path.
Usage: python scripts/rtf_harness.py [seconds]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import audiocodecs_b200 as A
from oracle import weights

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 15.86  # length of R/audiocodecs/example.wav
dev = torch.device("cuda:0")
codecs = {
    "encodec": A.Encodec(16000, 24000, num_codebooks=8, state_dict=weights.encodec_state_dict(0)),
    "dac": A.DAC(16000, 44100, num_codebooks=9, state_dict=weights.dac_state_dict(0)),
    "mimi": A.Mimi(16000, num_codebooks=8, state_dict=weights.mimi_state_dict(0)),
}
sig = (torch.randn(1, int(16000 * secs), generator=torch.Generator().manual_seed(0)) * 0.1).to(dev)
for name, codec in codecs.items():
    codec = codec.eval().to(dev)
    for _ in range(3):
        toks = codec.sig_to_toks(sig)
        codec.toks_to_sig(toks)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    reps = 10
    torch.cuda.synchronize()
    t_enc = t_dec = 0.0
    for _ in range(reps):
        ev[0].record(); toks = codec.sig_to_toks(sig); ev[1].record(); rec = codec.toks_to_sig(toks); ev[2].record()
        torch.cuda.synchronize()
        t_enc += ev[0].elapsed_time(ev[1]); t_dec += ev[1].elapsed_time(ev[2])
    t_enc, t_dec = t_enc / reps / 1e3, t_dec / reps / 1e3
    print(json.dumps({"codec": name, "audio_s": secs, "batch": 1, "sample_rate": 16000, "t_enc_ms": round(t_enc * 1e3, 3),
                      "t_dec_ms": round(t_dec * 1e3, 3), "rtf": round((t_enc + t_dec) / secs, 6), "inverse_rtf": round(secs / (t_enc + t_dec), 1),
                      "toks": list(toks.shape), "rec": list(rec.shape)}), flush=True)
    # the same two calls replayed as CUDA graphs (audiocodecs_b200.GraphedCodec): one driver call per direction
    g = A.GraphedCodec(codec, sig)
    same = bool((g.sig_to_toks(sig) == toks).all()) and bool((g.toks_to_sig(toks) == rec).all())
    torch.cuda.synchronize()
    t_enc = t_dec = 0.0
    for _ in range(reps):
        ev[0].record(); gt = g.sig_to_toks(sig); ev[1].record(); g.toks_to_sig(gt); ev[2].record()
        torch.cuda.synchronize()
        t_enc += ev[0].elapsed_time(ev[1]); t_dec += ev[1].elapsed_time(ev[2])
    t_enc, t_dec = t_enc / reps / 1e3, t_dec / reps / 1e3
    print(json.dumps({"codec": name, "cuda_graphs": True, "identical_to_eager": same, "t_enc_ms": round(t_enc * 1e3, 3),
                      "t_dec_ms": round(t_dec * 1e3, 3), "rtf": round((t_enc + t_dec) / secs, 6),
                      "inverse_rtf": round(secs / (t_enc + t_dec), 1)}), flush=True)
