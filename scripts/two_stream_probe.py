"""Does running consecutive (independent) steps on two CUDA streams raise throughput?  The LSTM recurrences occupy 64 of the
148 SMs for ~4 ms of every EnCodec step; a second stream can fill the other SMs with the next batch's convolutions.
Usage: python scripts/two_stream_probe.py [encodec|mimi|dac] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import audiocodecs_b200 as A
from oracle import weights

which = sys.argv[1] if len(sys.argv) > 1 else "encodec"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
dev = torch.device("cuda:0")
if which == "encodec":
    codec, sr, B = A.Encodec(24000, 24000, num_codebooks=8, state_dict=weights.encodec_state_dict(0)), 24000, 64
elif which == "dac":
    codec, sr, B = A.DAC(44100, 44100, num_codebooks=9, state_dict=weights.dac_state_dict(0), precision="bf16"), 44100, 64
else:
    codec, sr, B = A.Mimi(24000, num_codebooks=8, state_dict=weights.mimi_state_dict(0), precision="bf16"), 24000, 128
codec = codec.eval().to(dev)
sigs = [(torch.randn(B, sr * 10, generator=torch.Generator().manual_seed(s)) * 0.1).to(dev) for s in (1, 2)]
for _ in range(3):
    codec.toks_to_sig(codec.sig_to_toks(sigs[0]))
torch.cuda.synchronize()


def run(nstreams):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for s in streams:
        s.wait_event(e0)
    for i in range(steps):
        with torch.cuda.stream(streams[i % nstreams]):
            codec.toks_to_sig(codec.sig_to_toks(sigs[i % 2]))
    for s in streams:
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def run_split():
    """encode on one stream, decode on another: decode(i) (which starts with its LSTM) overlaps encode(i+1) (which ends with its)"""
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    sa.wait_event(e0); sb.wait_event(e0)
    for i in range(steps):
        with torch.cuda.stream(sa):
            toks = codec.sig_to_toks(sigs[i % 2])
            ev = torch.cuda.Event(); ev.record(sa)
        with torch.cuda.stream(sb):
            sb.wait_event(ev)
            toks.record_stream(sb)
            codec.toks_to_sig(toks)
    torch.cuda.current_stream().wait_stream(sa); torch.cuda.current_stream().wait_stream(sb)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


for rep in range(4):
    ms = run_split()
    print(f"{which}: encode-stream / decode-stream: {ms:.2f} ms/step -> {B * 10 / (ms / 1e3):.0f} audio-s/s", flush=True)
for n in (1, 2, 1, 2):
    ms = run(n)
    print(f"{which}: {n} stream(s): {ms:.2f} ms/step -> {B * 10 / (ms / 1e3):.0f} audio-s/s", flush=True)
