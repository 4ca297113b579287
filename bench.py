#!/usr/bin/env python
"""Headline benchmark: audio-seconds tokenized+detokenized per second (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--codec encodec] [--batch 64] [--precision exact]

A step = one pass of `sig_to_toks` -> `toks_to_sig` (the reference's mode="reconstruct" forward, R/audiocodecs/codec.py:45-55)
over one batch of synthetic 10 s clips; workload = BASELINE.json configs[1]: EnCodec-24k, 8 codebooks, 64 x 10 s mono per GPU.
Prints ONE JSON line (rank 0).  Besides the contract's keys the line carries
  parity          tokens / waveform of the TIMED configuration against the unmodified reference wrappers run on the host
                  (near-ties, top-2 gap <= 1e-4, classified with the oracle's distances)
  fast_mode       the same workload at precision="fp16" (one fp16 product per MAC: fastest tensor path), with its own parity
  extra_configs   the other BASELINE.json configs (DAC-44.1k 64 x 10 s, Mimi 128 x 10 s, EnCodec K=32), each with value /
                  roofline / parity, so that the driver's 1- and 8-GPU runs carry them
  gpu_eager_baseline  the reference wrappers moved to the same GPU with .to("cuda") (eager PyTorch: the reference's own GPU path)
  with_code_gather    (N > 1) the step with an NCCL all_gather of the int64 codes between the two halves
See DESIGN.md "Measurement" for how each field is obtained.
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
    "encodec": dict(sr=24000, K=8, batch=64, gflop_per_s=6.12, desc="EnCodec-24k K=8 64x10s mono encode+decode", cpu_clips=2),
    # the other BASELINE.json configs: `extra_configs` of the default line, or the main workload with --codec
    "encodec32": dict(sr=24000, K=32, batch=64, gflop_per_s=6.59, desc="EnCodec-24k K=32 64x10s mono encode+decode", cpu_clips=2),
    "dac": dict(sr=44100, K=9, batch=64, gflop_per_s=199.8, desc="DAC-44.1k K=9 64x10s mono encode+decode", cpu_clips=1),
    "mimi": dict(sr=24000, K=8, batch=128, gflop_per_s=11.0, desc="Mimi-24k K=8 128x10s mono encode+decode", cpu_clips=2),
}


# per-launch DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum, averaged over every launch of the kernel in one
# full-size tuned step) from the committed ncu launch lists profiles/r02_<codec>_<precision>_launches.csv; None = not captured
NCU_TRAFFIC = {
    ("encodec", "exact"): {"conv_tc_kernel": 1.282e9, "resunit_tc_kernel": 4.389e9, "lstm_tc_kernel": 4.76e8, "rvq_encode_tc_kernel": 3.3e7},
    ("encodec", "fp16"): {"conv_tc_kernel": 8.92e8, "resunit_tc_kernel": 2.910e9, "lstm_tc_kernel": 4.76e8, "rvq_encode_tc_kernel": 3.3e7},
    ("mimi", "exact"): {"conv_tc_kernel": 7.71e8, "resunit_tc_kernel": 1.0007e10, "attention_tc_kernel": 2.27e8},
    ("mimi", "fp16"): {"conv_tc_kernel": 5.51e8, "resunit_tc_kernel": 6.650e9, "attention_tc_kernel": 2.16e8},
    ("dac", "exact"): {"conv_tc_kernel": 6.023e9, "resunit_tc_kernel": 2.1198e10},
    ("dac", "fp16"): {"conv_tc_kernel": 4.332e9, "resunit_tc_kernel": 1.4572e10},
}
for _p in ("exact", "fp16"):
    NCU_TRAFFIC[("encodec32", _p)] = NCU_TRAFFIC[("encodec", _p)]


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


def oracle_mod(codec):
    """CPU oracle of this codec (checker of the `parity` fields / fallback of the CPU legs only -- never on the timed GPU path)."""
    from oracle import dac_ref, encodec_ref, mimi_ref
    return {"encodec": encodec_ref, "encodec32": encodec_ref, "dac": dac_ref, "mimi": mimi_ref}[codec]


def make_input(rank, B, T):
    g = torch.Generator().manual_seed(999 + rank)
    return torch.randn(B, T, generator=g) * 0.1


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "25"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
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
        sm, mx, reasons, pw = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                pw.append(float(r[2]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                pass
        sm.sort()
        # "under load" = upper half of the samples (the sampler also sees the idle edges)
        load = sm[len(sm) // 2:] if sm else []
        med = load[len(load) // 2] if load else None
        pw.sort()
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": pw[-1] if pw else None}   # board power: the step runs at the 1000 W cap (sw_power_cap), see DESIGN.md


class Ctx:
    def __init__(self, rank, world, local_rank):
        self.rank, self.world, self.local_rank = rank, world, local_rank
        self.dev = torch.device("cuda", local_rank)

    def barrier(self):
        if self.world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        if self.world > 1:
            t = torch.tensor([ms], device=self.dev, dtype=torch.float64)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = t.item()
        return ms

    def timed(self, fn, steps):
        """EXACTLY `steps` calls between two CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks"""
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = self.max_over_ranks(e0.elapsed_time(e1))
        self.barrier()
        return ms


def measure(ctx, name, precision, steps, warmup, batch=0, sample_clocks=False, with_e2e=True, with_gather=False):
    """One workload on this rank's GPU: device-timed value, end-to-end value through HostPipeline, per-kernel roofline.
    Returns (result dict, state) -- state keeps the codec / inputs for the parity leg."""
    from audiocodecs_b200 import _lib, ops, shard

    wl = WORKLOADS[name]
    dev, world = ctx.dev, ctx.world
    B, T = batch or wl["batch"], wl["sr"] * SECONDS
    sd = make_state_dict(name)
    codec = make_codec(name, sd, precision).eval().to(dev)
    host_sig = make_input(ctx.rank, B, T).pin_memory()
    sig = host_sig.to(dev)
    with torch.no_grad():
        T_out = codec(sig[:1]).shape[1]  # DAC returns 512*N samples, not T
    audio_s_total = B * SECONDS * world

    def step():
        return codec(sig)  # mode="reconstruct": sig_to_toks -> toks_to_sig

    for _ in range(max(warmup, 3)):
        step()
    sampler = ClockSampler(ctx.local_rank)
    if sample_clocks and ctx.rank == 0:
        sampler.start()
    l0 = _lib.launch_count()
    ms = ctx.timed(step, steps)
    launches = _lib.launch_count() - l0
    clocks = sampler.stop() if (sample_clocks and ctx.rank == 0) else None
    out = {"value": round(audio_s_total * steps / (ms / 1e3), 2), "unit": "audio-s/s", "ms_per_step": round(ms / steps, 3),
           "precision": precision, "per_gpu_batch": B, "gpu_launches": int(launches)}
    if clocks is not None:
        out["clocks"] = clocks

    if with_e2e:
        # end to end: every step copies its input from pinned host memory and its result back to pinned host memory; the
        # copies of neighbouring steps run on the copy engines under this step's kernels (audiocodecs_b200.hostpipe).  Timed on
        # the device across the three streams: start event before the first input copy is issued, end event after the last
        # output copy; barrier + synchronize on both sides, max over ranks
        from audiocodecs_b200.hostpipe import HostPipeline
        pipe = HostPipeline(codec)
        host_outs = [torch.empty((B, T_out), dtype=torch.float32).pin_memory() for _ in range(2)]
        state = {"i": 0}

        def step_e2e():
            i = state["i"]
            state["i"] = i + 1
            pipe.submit(host_sig, host_outs[i % 2])

        for _ in range(2):
            step_e2e()
        pipe.drain()
        ctx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(pipe.s_in)
        torch.cuda.current_stream().wait_event(e0)
        for _ in range(steps):
            step_e2e()
        pipe.s_out.wait_stream(torch.cuda.current_stream())
        e1.record(pipe.s_out)
        pipe.drain()
        torch.cuda.synchronize()
        ms_e2e = ctx.max_over_ranks(e0.elapsed_time(e1))
        ctx.barrier()
        out["e2e"] = {"value": round(audio_s_total * steps / (ms_e2e / 1e3), 2), "unit": "audio-s/s", "h2d_bytes_per_step": B * T * 4 * world,
                      "d2h_bytes_per_step": B * T_out * 4 * world, "ms_per_step": round(ms_e2e / steps, 3),
                      "how": "HostPipeline: pinned host in/out every step, H2D(i+1) and D2H(i-1) on copy streams under the kernels of step i"}

    if with_gather and world > 1:
        # north_star: "NCCL used only to gather codes": the same step with one all_gather of the int64 codes between the halves
        def step_gather():
            toks = codec.sig_to_toks(sig)
            all_toks = shard.gather_rows(toks, B * world)       # every rank ends up with all B * world clips' tokens
            return codec.toks_to_sig(all_toks[ctx.rank * B:(ctx.rank + 1) * B])

        for _ in range(2):
            step_gather()
        ms_g = ctx.timed(step_gather, steps)
        N, K = codec.sig_to_toks(sig[:1]).shape[1:]
        out["with_code_gather"] = {"value": round(audio_s_total * steps / (ms_g / 1e3), 2), "unit": "audio-s/s",
                                   "ms_per_step": round(ms_g / steps, 3), "gathered_bytes_per_step": int(B * world * N * K * 8),
                                   "collective": "NCCL all_gather of int64 codes [B, N, K], inside the timed region"}

    # ---- per-kernel device times for the roofline of the dominant kernel (one instrumented step, CUDA events on the
    # launching stream around every C-ABI launch)
    prof = ops.Profiler()
    ops.set_profiler(prof)
    step()
    torch.cuda.synchronize()
    ops.set_profiler(None)
    summary = prof.summary()
    by_label = prof.summary(by_label=True)
    peaks = load_peaks()
    kname, d = max(summary.items(), key=lambda kv: kv[1]["ms"])
    sec = d["ms"] / 1e3
    tflops = d["flops"] / sec / 1e12 if sec > 0 else 0.0
    gbs = d["bytes"] / sec / 1e9 if sec > 0 else 0.0
    # which roof bounds the dominant kernel: arithmetic intensity of its algorithmic work vs the measured ridge
    ridge = peaks["tf_sus"] * 1e12 / (peaks["hbm"] * 1e9)
    intensity = d["flops"] / d["bytes"] if d["bytes"] else float("inf")
    if intensity < ridge:
        roofline = {"kernel": kname, "bound": "hbm", "achieved": round(gbs, 1), "peak": peaks["hbm"], "unit": "GB/s",
                    "frac": round(gbs / peaks["hbm"], 4)}
    else:
        roofline = {"kernel": kname, "bound": "tensor", "achieved": round(tflops, 2), "peak": peaks["tf_sus"], "unit": "TFLOP/s",
                    "frac": round(tflops / peaks["tf_sus"], 4)}
    total_ms = sum(v["ms"] for v in summary.values())
    roofline.update({"traffic": NCU_TRAFFIC.get((name, precision), {}).get(kname),
                     "peak_source": peaks["src"] + (" (copy bandwidth)" if roofline["bound"] == "hbm" else " (sustained bf16)"),
                     "flop_per_byte": round(intensity, 1), "ridge_flop_per_byte": round(ridge, 1),
                     "tensor_tflops": round(tflops, 2), "launches_per_step": d["n"], "ms_per_step": round(d["ms"], 3),
                     "share_of_step": round(d["ms"] / total_ms, 4),
                     "whole_step_tensor_frac": round(wl["gflop_per_s"] * 1e9 * B * SECONDS / (total_ms / 1e3) / 1e12 / peaks["tf_sus"], 4),
                     "all_kernels_ms": {k: round(v["ms"], 3) for k, v in summary.items()},
                     "by_layer_family_ms": {k: round(v["ms"], 3) for k, v in by_label.items()}})
    out["roofline"] = roofline
    out["dtype"] = codec.compute_dtype
    return out, dict(codec=codec, sd=sd, host_sig=host_sig, wl=wl, name=name)


def reference_on_cpu(name, sd, sig, threads=None, reps=1):
    """(toks, rec, seconds per pass, kind): the unmodified reference wrapper (baseline/_ref) on the host cores, fp32; falls
    back to the oracle port when baseline/_ref is absent."""
    from baseline import ref_runner
    wl = WORKLOADS[name]
    torch.set_num_threads(threads or os.cpu_count())
    ref = ref_runner.make_reference(name, sd, wl["sr"], wl["K"])
    if ref is not None:
        enc, dec, kind = (lambda x: ref.sig_to_toks(x)), (lambda t: ref.toks_to_sig(t)), "reference"
    else:
        mod = oracle_mod(name)
        enc, dec, kind = (lambda x: mod.sig_to_toks(sd, x, wl["K"])), (lambda t: mod.toks_to_sig(sd, t)), "port"
    best = None
    with torch.no_grad():
        if name == "dac":  # warm-up on 1 s: a 10 s clip takes ~13 s
            dec(enc(sig[:1, :wl["sr"]]))
        else:
            dec(enc(sig))
        for _ in range(reps):
            t0 = time.perf_counter()
            toks = enc(sig)
            rec = dec(toks)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    return toks.contiguous(), rec, best, kind


def parity_and_cpu_baseline(state, reps):
    """Runs the reference on `cpu_clips` clips of the timed batch (this IS the cpu_baseline measurement) and compares the GPU
    path's tokens / waveform for those clips with it.  Near-ties are classified with the oracle's top-2 distance gaps."""
    codec, sd, name, wl = state["codec"], state["sd"], state["name"], state["wl"]
    clips = wl["cpu_clips"]
    sig = state["host_sig"][:clips].clone()
    toks_ref, rec_ref, secs, kind = reference_on_cpu(name, sd, sig, reps=reps)
    threads = torch.get_num_threads()
    cpu = {"value": round(clips * SECONDS / secs, 2), "unit": "audio-s/s", "cores": threads, "kind": kind,
           "sample": f"{clips} x {SECONDS} s clip(s) of the timed batch, fp32, best of {reps} after 1 warm-up, "
                     f"{'unmodified audiocodecs wrappers from baseline/_ref' if kind == 'reference' else 'oracle port'}, torch {torch.__version__} CPU"}
    dev = next(codec.buffers()).device
    full = state["host_sig"].to(dev)
    toks = codec.sig_to_toks(full)[:clips].cpu()          # tokens of these clips INSIDE the full timed batch
    rec = codec.toks_to_sig(toks_ref.to(dev)).cpu()       # decoder fed the reference's tokens
    with torch.no_grad():
        _, gaps, _ = oracle_mod(name).sig_to_toks(sd, sig, wl["K"], return_gaps=True)
    safe = gaps > 1e-4
    eq = toks == toks_ref
    ref64, est = rec_ref.double().flatten(1), rec.double().flatten(1)
    ref64, est = ref64 - ref64.mean(1, keepdim=True), est - est.mean(1, keepdim=True)
    s = (est * ref64).sum(1, keepdim=True) / ref64.pow(2).sum(1, keepdim=True).clamp_min(1e-30) * ref64
    sisnr = (10 * torch.log10(s.pow(2).sum(1) / (est - s).pow(2).sum(1).clamp_min(1e-30))).min().item()
    # residual quantisation is a chain: a frame is judged at its FIRST differing stage (later stages quantise another residual)
    eqf, gf = eq.flatten(0, -2), gaps.flatten(0, -2)
    diff = ~eqf
    first = diff.float().argmax(-1)
    gap_at = gf[torch.arange(eqf.shape[0]), first]
    has = diff.any(-1)
    parity = {"code_match_safe": round(eq[safe].float().mean().item(), 6), "code_match_all": round(eq.float().mean().item(), 6),
              "frames": int(eqf.shape[0]), "frames_first_diff_at_gap_gt_1e-4": int((has & (gap_at > 1e-4)).sum()),
              "frames_first_diff_at_near_tie": int((has & (gap_at <= 1e-4)).sum()),
              "near_tie_frac": round((~safe).float().mean().item(), 6), "sisnr_db": round(sisnr, 2),
              "per_stage_match": [round(eq[..., k].float().mean().item(), 4) for k in range(eq.shape[-1])],
              "decisions": int(eq.numel()), "clips_checked": clips,
              "checker": f"{kind} on CPU (tokens, waveform of its own tokens); near-ties = oracle top-2 relative gap <= 1e-4"}
    return parity, cpu


def gpu_eager_baseline(state, steps=3):
    """The reference's own GPU path: the unmodified wrapper moved to the same GPU with .to('cuda'), eager PyTorch fp32 with
    torch's defaults (cudnn.allow_tf32=True: TF32 convolutions; matmul fp32)."""
    from baseline import ref_runner
    name, sd, wl = state["name"], state["sd"], state["wl"]
    dev = next(state["codec"].buffers()).device
    try:
        ref = ref_runner.make_reference(name, sd, wl["sr"], wl["K"])
        if ref is None:
            return {"unavailable": "baseline/_ref not installed"}
        ref = ref.to(dev)
        B = 16 if name == "dac" else wl["batch"]   # eager fp32 DAC activations of 64 clips need > 100 GB
        sig = state["host_sig"][:B].to(dev)
        with torch.no_grad():
            for _ in range(2):
                ref.toks_to_sig(ref.sig_to_toks(sig))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(steps):
                ref.toks_to_sig(ref.sig_to_toks(sig))
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        return {"value": round(B * SECONDS / (ms / 1e3), 2), "unit": "audio-s/s", "ms_per_step": round(ms, 2), "per_gpu_batch": B,
                "impl": "unmodified audiocodecs wrapper .to('cuda'), eager fp32",
                "tf32": {"cudnn.allow_tf32": bool(torch.backends.cudnn.allow_tf32), "cuda.matmul.allow_tf32": bool(torch.backends.cuda.matmul.allow_tf32)}}
    except Exception as e:  # noqa: BLE001 -- a comparator: report and move on
        return {"unavailable": f"{type(e).__name__}: {str(e)[:160]}"}
    finally:
        torch.cuda.empty_cache()


def config_of(name, B, world):
    wl = WORKLOADS[name]
    return {"workload": wl["desc"], "per_gpu_batch": B, "clip_seconds": SECONDS, "sample_rate": wl["sr"],
            "num_codebooks": wl["K"], "weights": "random-init (seed 0, matched-moment codebooks)",
            "l2": "activations per step exceed the 126 MB L2 many times over (inputs_exceed_l2)",
            "parallelism": f"clip-sharded x{world}, no data-path collective"}


def run_ours(args, rank, world, local_rank):
    ctx = Ctx(rank, world, local_rank)
    torch.cuda.set_device(ctx.dev)
    cpu_legs = rank == 0 and world == 1 and not args.no_cpu
    main, state = measure(ctx, args.codec, args.precision, args.steps, args.warmup, args.batch, sample_clocks=True, with_gather=True)
    B = main["per_gpu_batch"]
    out = {
        "metric": "audio_seconds_per_second_encode_decode", "value": main["value"], "unit": "audio-s/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": main["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": main["dtype"], "data": "synthetic",
        "precision": args.precision, "config": config_of(args.codec, B, world), "e2e": main["e2e"],
        "gpu_launches": main["gpu_launches"], "clocks": main.get("clocks"), "roofline": main["roofline"],
    }
    if "with_code_gather" in main:
        out["with_code_gather"] = main["with_code_gather"]
    if cpu_legs:
        out["parity"], out["cpu_baseline"] = parity_and_cpu_baseline(state, reps=1 if args.codec == "dac" else 3)
        out["gpu_eager_baseline"] = gpu_eager_baseline(state)
    del state
    torch.cuda.empty_cache()
    if args.no_extras:
        return out
    # ---- the fastest tensor path on the same workload
    if args.precision != "fp16":
        try:
            fast, st = measure(ctx, args.codec, "fp16", args.steps, args.warmup, args.batch, with_e2e=False)
            out["fast_mode"] = {k: fast[k] for k in ("value", "unit", "ms_per_step", "precision")}
            out["fast_mode"]["roofline_frac"] = fast["roofline"]["frac"]
            if cpu_legs:
                out["fast_mode"]["parity"], _ = parity_and_cpu_baseline(st, reps=1)
            del st
        except Exception as ex:  # noqa: BLE001 -- a secondary line must not take the headline down
            out["fast_mode"] = {"error": f"{type(ex).__name__}: {str(ex)[:200]}"}
        torch.cuda.empty_cache()
    # ---- the other BASELINE.json configs (each at its own full batch; fewer steps for the 0.3 s DAC step)
    extras = {}
    for name in ("dac", "mimi", "encodec32"):
        if name == args.codec:
            continue
        steps = 3 if name == "dac" else min(args.steps, 10)
        try:
            r, st = measure(ctx, name, args.precision, steps, 3, 0, with_e2e=(name == "dac"))
            e = {"workload": WORKLOADS[name]["desc"], "value": r["value"], "unit": "audio-s/s", "ms_per_step": r["ms_per_step"],
                 "steps": steps, "precision": args.precision, "per_gpu_batch": r["per_gpu_batch"], "gpu_launches": r["gpu_launches"],
                 "roofline": {k: r["roofline"][k] for k in ("kernel", "bound", "achieved", "peak", "unit", "frac", "share_of_step",
                                                             "whole_step_tensor_frac", "all_kernels_ms")}}
            if "e2e" in r:
                e["e2e"] = {k: r["e2e"][k] for k in ("value", "unit", "ms_per_step")}
            if cpu_legs:
                e["parity"], e["cpu_baseline"] = parity_and_cpu_baseline(st, reps=1)
                e["gpu_eager_baseline"] = gpu_eager_baseline(st)
            del st
            torch.cuda.empty_cache()
            if args.precision != "fp16" and name != "encodec32":
                # the one-product fp16 path of this config (DAC: code_match_safe 0.9993, Mimi: see its parity block)
                f, st = measure(ctx, name, "fp16", steps, 3, 0, with_e2e=False)
                e["fast_mode"] = {"value": f["value"], "unit": "audio-s/s", "ms_per_step": f["ms_per_step"], "precision": "fp16",
                                  "roofline_frac": f["roofline"]["frac"], "roofline_kernel": f["roofline"]["kernel"]}
                if cpu_legs:
                    e["fast_mode"]["parity"], _ = parity_and_cpu_baseline(st, reps=1)
                del st
        except Exception as ex:  # noqa: BLE001 -- an extra line must not take the headline down
            e = {"error": f"{type(ex).__name__}: {str(ex)[:200]}"}
        extras[name] = e
        torch.cuda.empty_cache()
    out["extra_configs"] = extras
    return out


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path -- the unmodified wrappers from baseline/_ref -- on
    all host cores, each step a bounded sample (cpu_clips clips) of the workload."""
    if rank != 0:
        return None
    wl = WORKLOADS[args.codec]
    sd = make_state_dict(args.codec)
    clips = wl["cpu_clips"]
    sig = make_input(0, wl["batch"], wl["sr"] * SECONDS)[:clips].clone()  # the first clips of the GPU arm's rank-0 batch
    from baseline import ref_runner
    threads = os.cpu_count()
    torch.set_num_threads(threads)
    ref = ref_runner.make_reference(args.codec, sd, wl["sr"], wl["K"])
    if ref is not None:
        kind = "reference"

        def step():
            with torch.no_grad():
                ref.toks_to_sig(ref.sig_to_toks(sig))
    else:
        kind, mod = "port", oracle_mod(args.codec)

        def step():
            with torch.no_grad():
                mod.toks_to_sig(sd, mod.sig_to_toks(sd, sig, wl["K"]))

    steps = args.steps if args.codec != "dac" else min(args.steps, 3)
    for _ in range(max(1, min(args.warmup, 2 if args.codec != "dac" else 1))):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    value = clips * SECONDS * steps / dt
    sample = (f"{clips} x {SECONDS} s clip(s) per step (a bounded sample of the {wl['batch']}-clip batch), fp32, {threads} threads, "
              f"{'unmodified audiocodecs wrappers from baseline/_ref' if kind == 'reference' else 'oracle port (baseline/_ref absent)'}")
    return {
        "impl": "reference", "metric": "audio_seconds_per_second_encode_decode", "value": round(value, 2), "unit": "audio-s/s",
        "n_gpus": world, "steps": steps, "warmup": args.warmup, "ms_per_step": round(dt / steps * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_of(args.codec, wl["batch"], world), "clips_per_step": clips,
        "cpu_baseline": {"value": round(value, 2), "unit": "audio-s/s", "cores": threads, "kind": kind, "sample": sample},
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
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU legs (parity, cpu_baseline, gpu_eager_baseline)")
    ap.add_argument("--no-extras", action="store_true", help="skip fast_mode and extra_configs")
    ap.add_argument("--precision", default="exact", choices=["exact", "fp16", "bf16", "fp32"],
                    help="exact (default): tensor path whose tokens equal the reference's; fp16: fastest tensor path")
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
