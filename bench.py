#!/usr/bin/env python
"""Headline benchmark: audio-seconds tokenized+detokenized per second (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--codec encodec] [--batch 64]

A step = one pass of `sig_to_toks` -> `toks_to_sig` over one batch of synthetic 10 s clips
(workload = BASELINE.json configs[1]: EnCodec-24k, 8 codebooks, 64 x 10 s mono per GPU).
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for how each field is obtained.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

SECONDS = 10
WORKLOADS = {
    # name: (sample_rate, K, default per-GPU batch, algorithmic GFLOP per audio-second enc+dec -- SURVEY 8d)
    "encodec": dict(sr=24000, K=8, batch=64, gflop_per_s=6.12, desc="EnCodec-24k K=8 64x10s mono encode+decode"),
    # the other BASELINE.json configs: extra bench lines on request (--codec), never the default
    "encodec32": dict(sr=24000, K=32, batch=64, gflop_per_s=6.59, desc="EnCodec-24k K=32 64x10s mono encode+decode"),
    "dac": dict(sr=44100, K=9, batch=64, gflop_per_s=199.8, desc="DAC-44.1k K=9 64x10s mono encode+decode"),
    "mimi": dict(sr=24000, K=8, batch=128, gflop_per_s=11.0, desc="Mimi-24k K=8 128x10s mono encode+decode"),
}


# per-launch DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum, averaged over every launch of the kernel in one
# full-size step) from the committed ncu launch lists profiles/r01h_launches.csv (EnCodec), r01h_mimi_launches.csv,
# r01h_dac_launches.csv; None when not captured
NCU_TRAFFIC = {
    "encodec": {"conv_tc_kernel": 1.194e9, "resunit_tc_kernel": 2.909e9, "lstm_tc_kernel": 4.77e8, "rvq_encode_tc_kernel": 3.3e7},
    "mimi": {"conv_tc_kernel": 8.84e8, "resunit_tc_kernel": 9.416e9, "attention_tc_kernel": 2.38e8},
    "dac": {"conv_tc_kernel": 8.431e9, "resunit_tc_kernel": 2.2226e10},
}
NCU_TRAFFIC["encodec32"] = NCU_TRAFFIC["encodec"]


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sus=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sus=1400.0, src="fallback")


def make_state_dict(codec):
    from oracle import weights  # deterministic random-init weights (no checkpoints offline); not on the timed path
    return {"encodec": weights.encodec_state_dict, "encodec32": weights.encodec_state_dict, "dac": weights.dac_state_dict,
            "mimi": weights.mimi_state_dict}[codec](0)


def make_codec(codec, sd, precision="exact"):
    import audiocodecs_b200 as A
    wl = WORKLOADS[codec]
    if codec in ("encodec", "encodec32"):
        return A.Encodec(wl["sr"], wl["sr"], num_codebooks=wl["K"], state_dict=sd, precision=precision)
    if codec == "dac":
        return A.DAC(wl["sr"], wl["sr"], num_codebooks=wl["K"], state_dict=sd, precision=precision)
    return A.Mimi(wl["sr"], num_codebooks=wl["K"], state_dict=sd, precision=precision)


def oracle_fns(codec):
    """(sig_to_toks, toks_to_sig) of the CPU oracle for this codec (cpu_baseline / --impl reference legs only)."""
    from oracle import dac_ref, encodec_ref, mimi_ref
    mod = {"encodec": encodec_ref, "encodec32": encodec_ref, "dac": dac_ref, "mimi": mimi_ref}[codec]
    return mod.sig_to_toks, mod.toks_to_sig


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                pass
        sm.sort()
        # "under load" = upper half of the samples (the sampler also sees the idle edges)
        load = sm[len(sm) // 2:] if sm else []
        med = load[len(load) // 2] if load else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def run_ours(args, rank, world, local_rank):
    from audiocodecs_b200 import _lib, ops

    wl = WORKLOADS[args.codec]
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    B, T = args.batch or wl["batch"], wl["sr"] * SECONDS
    sd = make_state_dict(args.codec)
    codec = make_codec(args.codec, sd, args.precision).eval().to(dev)
    g = torch.Generator().manual_seed(999 + rank)
    host_sig = (torch.randn(B, T, generator=g) * 0.1).pin_memory()
    sig = host_sig.to(dev)
    with torch.no_grad():
        T_out = codec.toks_to_sig(codec.sig_to_toks(sig[:1])).shape[1]  # DAC returns 512*N samples, not T
    host_out = torch.empty((B, T_out), dtype=torch.float32).pin_memory()
    audio_s_total = B * SECONDS * world

    def step():
        toks = codec.sig_to_toks(sig)
        return codec.toks_to_sig(toks)

    # end to end: every step copies its input from pinned host memory and its result back to pinned host memory; the
    # copies of neighbouring steps run on the copy engines under this step's kernels (audiocodecs_b200.hostpipe)
    from audiocodecs_b200.hostpipe import HostPipeline
    pipe = HostPipeline(codec)
    host_outs = [host_out, torch.empty_like(host_out).pin_memory()]
    e2e_state = {"i": 0}

    def step_e2e():
        i = e2e_state["i"]
        e2e_state["i"] = i + 1
        pipe.submit(host_sig, host_outs[i % 2])

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = t.item()
        barrier()
        return ms

    for _ in range(max(args.warmup, 3)):
        step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count()
    ms = timed(step, args.steps)
    launches = _lib.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    for _ in range(2):
        step_e2e()
    pipe.drain()
    # timed on the device across the three streams: start event before the first input copy is issued, end event after
    # the last output copy; barrier + synchronize on both sides, max over ranks
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(pipe.s_in)
    torch.cuda.current_stream().wait_event(e0)
    for _ in range(args.steps):
        step_e2e()
    pipe.s_out.wait_stream(torch.cuda.current_stream())
    e1.record(pipe.s_out)
    pipe.drain()
    torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_e2e], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms_e2e = t.item()
    barrier()

    # ---- per-kernel device times for the roofline of the dominant kernel (one instrumented step, CUDA events
    # on the launching stream around every C-ABI launch)
    prof = ops.Profiler()
    ops.set_profiler(prof)
    step()
    torch.cuda.synchronize()
    ops.set_profiler(None)
    summary = prof.summary()
    by_label = prof.summary(by_label=True)

    value = audio_s_total * args.steps / (ms / 1e3)
    e2e = audio_s_total * args.steps / (ms_e2e / 1e3)
    peaks = load_peaks()
    name, d = max(summary.items(), key=lambda kv: kv[1]["ms"])
    sec = d["ms"] / 1e3
    tflops = d["flops"] / sec / 1e12 if sec > 0 else 0.0
    gbs = d["bytes"] / sec / 1e9 if sec > 0 else 0.0
    # which roof bounds the dominant kernel: arithmetic intensity of its algorithmic work vs the measured ridge
    ridge = peaks["tf_sus"] * 1e12 / (peaks["hbm"] * 1e9)
    intensity = d["flops"] / d["bytes"] if d["bytes"] else float("inf")
    if intensity < ridge:
        roofline = {"kernel": name, "bound": "hbm", "achieved": round(gbs, 1), "peak": peaks["hbm"], "unit": "GB/s",
                    "frac": round(gbs / peaks["hbm"], 4)}
    else:
        roofline = {"kernel": name, "bound": "tensor", "achieved": round(tflops, 2), "peak": peaks["tf_sus"], "unit": "TFLOP/s",
                    "frac": round(tflops / peaks["tf_sus"], 4)}
    ncu = NCU_TRAFFIC.get(args.codec, {}).get(name)
    roofline.update({"traffic": ncu, "peak_source": peaks["src"] + (" (copy bandwidth)" if roofline["bound"] == "hbm" else " (sustained bf16)"),
                     "flop_per_byte": round(intensity, 1), "ridge_flop_per_byte": round(ridge, 1),
                     "tensor_tflops": round(tflops, 2), "launches_per_step": d["n"], "ms_per_step": round(d["ms"], 3),
                     "share_of_step": round(d["ms"] / sum(v["ms"] for v in summary.values()), 4),
                     "all_kernels_ms": {k: round(v["ms"], 3) for k, v in summary.items()},
                     "by_layer_family_ms": {k: round(v["ms"], 3) for k, v in by_label.items()}})
    out = {
        "metric": "audio_seconds_per_second_encode_decode", "value": round(value, 2), "unit": "audio-s/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": codec.compute_dtype, "data": "synthetic",
        "config": {"workload": wl["desc"], "per_gpu_batch": B, "clip_seconds": SECONDS, "sample_rate": wl["sr"],
                   "num_codebooks": wl["K"], "precision": args.precision, "weights": "random-init (seed 0, matched-moment codebooks)",
                   "l2": "activations per step exceed the 126 MB L2 many times over (inputs_exceed_l2)",
                   "parallelism": f"clip-sharded x{world}, no data-path collective"},
        "e2e": {"value": round(e2e, 2), "unit": "audio-s/s", "h2d_bytes_per_step": B * T * 4 * world,
                "d2h_bytes_per_step": B * T_out * 4 * world, "ms_per_step": round(ms_e2e / args.steps, 3),
                "how": "HostPipeline: pinned host in/out every step, H2D(i+1) and D2H(i-1) on copy streams under the kernels of step i"},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
    }
    if rank == 0 and world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_baseline(args.codec, sd, clips=1, reps=1 if args.codec == "dac" else 3)
    return out


def cpu_baseline(codec, sd, clips, reps, threads=None):
    """The reference's CPU path (oracle port: the same ATen conv/LSTM/matmul ops the reference wrappers reach,
    fp32) on the host cores, on a bounded sample of the workload."""
    enc, dec = oracle_fns(codec)
    wl = WORKLOADS[codec]
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(999)
    sig = torch.randn(clips, wl["sr"] * SECONDS, generator=g) * 0.1
    best = None
    with torch.no_grad():
        for i in range(reps + 1):
            t0 = time.perf_counter()
            x = sig[:, :wl["sr"]] if (i == 0 and codec == "dac") else sig  # DAC: warm up on 1 s (a 10 s clip takes ~10 s)
            toks = enc(sd, x, wl["K"])
            dec(sd, toks)
            dt = time.perf_counter() - t0
            if i > 0:  # first rep is warm-up
                best = dt if best is None else min(best, dt)
    return {"value": round(clips * SECONDS / best, 2), "unit": "audio-s/s", "cores": threads, "kind": "port",
            "sample": f"{clips} x {SECONDS} s clip(s), fp32, best of {reps} after 1 warm-up, torch {torch.__version__} CPU"}


def run_reference(args, rank, world):
    if rank != 0:
        return None
    wl = WORKLOADS[args.codec]
    sd = make_state_dict(args.codec)
    clips = 1 if args.codec == "dac" else 2
    threads = os.cpu_count()
    torch.set_num_threads(threads)
    enc, dec = oracle_fns(args.codec)
    g = torch.Generator().manual_seed(999)
    sig = torch.randn(clips, wl["sr"] * SECONDS, generator=g) * 0.1

    def step():
        with torch.no_grad():
            toks = enc(sd, sig, wl["K"])
            dec(sd, toks)

    for _ in range(max(1, min(args.warmup, 2))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = clips * SECONDS * args.steps / dt
    sample = f"{clips} x {SECONDS} s clips per step (bounded sample of the 64-clip batch), fp32, {threads} threads"
    return {
        "impl": "reference", "metric": "audio_seconds_per_second_encode_decode", "value": round(value, 2), "unit": "audio-s/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "per_gpu_batch": wl["batch"], "clip_seconds": SECONDS, "sample_rate": wl["sr"],
                   "num_codebooks": wl["K"], "weights": "random-init (seed 0, matched-moment codebooks)"},
        "cpu_baseline": {"value": round(value, 2), "unit": "audio-s/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": round(value, 2), "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--codec", default="encodec", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU clips (default: the BASELINE workload's)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--precision", default="exact", choices=["exact", "bf16", "fp32"],
                    help="exact (default): tensor path whose tokens equal the reference's; bf16: fastest tensor path")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        out = run_reference(args, rank, world)
        if out is not None:
            print(json.dumps(out), flush=True)
        return
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    out = run_ours(args, rank, world, local_rank)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
