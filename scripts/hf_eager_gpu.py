"""The "existing GPU path" comparator: the reference's backends -- transformers' EncodecModel / MimiModel /
DacModel, which audiocodecs.{Encodec,Mimi} wrap (and DacModel is the in-container twin of descript-audio-codec) -- run
eagerly on the B200 in fp32 (PyTorch defaults: TF32 convolutions), random-init weights of the named architecture, same
synthetic input shapes as bench.py.  Timing: CUDA events, 3 warm-ups, 5 steps.  Not part of the product or the tests.
Usage: python scripts/hf_eager_gpu.py [encodec] [mimi] [dac]"""
import json, sys, time
import torch

which = sys.argv[1:] or ["encodec", "mimi", "dac"]
dev = torch.device("cuda:0")


def timed(fn, steps=5, warmup=3):
    for _ in range(warmup):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


for name in which:
    torch.manual_seed(0)
    try:
        if name == "encodec":
            from transformers import EncodecConfig, EncodecModel
            m, sr, B = EncodecModel(EncodecConfig()).eval().to(dev), 24000, 64
            sig = torch.randn(B, 1, sr * 10, device=dev) * 0.1

            def step():
                codes = m.encode(sig, bandwidth=6.0).audio_codes
                return m.decode(codes, [None])[0]
        elif name == "mimi":
            from transformers import MimiConfig, MimiModel
            m, sr, B = MimiModel(MimiConfig()).eval().to(dev), 24000, 128
            sig = torch.randn(B, 1, sr * 10, device=dev) * 0.1

            def step():
                codes = m.encode(sig, num_quantizers=8).audio_codes
                return m.decode(codes)[0]
        else:
            from transformers import DacConfig, DacModel
            cfg = DacConfig(encoder_hidden_size=64, downsampling_ratios=[2, 4, 8, 8], decoder_hidden_size=1536, n_codebooks=9,
                            codebook_size=1024, codebook_dim=8, sampling_rate=44100)
            m, sr, B = DacModel(cfg).eval().to(dev), 44100, 16   # 16 clips per step: eager fp32 activations of 64 clips need > 100 GB
            sig = torch.randn(B, 1, sr * 10, device=dev) * 0.1

            def step():
                codes = m.encode(sig, n_quantizers=9).audio_codes
                return m.decode(audio_codes=codes).audio_values
        with torch.no_grad():
            ms = timed(step)
        print(json.dumps({"impl": "transformers eager fp32 (TF32 convs) on cuda:0", "codec": name, "batch": B, "clip_seconds": 10,
                          "ms_per_step": round(ms, 2), "audio_s_per_s": round(B * 10 / (ms / 1e3), 1),
                          "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 1)}), flush=True)
    except Exception as e:  # noqa: BLE001 -- a comparator: report and move on
        print(json.dumps({"codec": name, "error": f"{type(e).__name__}: {str(e)[:200]}"}), flush=True)
    m = sig = None
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
