"""BASELINE.json configs[4] on N GPUs: EnCodec-24k K=32, a GLOBAL batch of B clips sharded over the ranks with
audiocodecs_b200.shard (contiguous split, one NCCL all_gather of the int64 codes, no collective on the compute path).
Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/sharded_sweep.py [max_B] [K]
Rank 0 prints one JSON line per global batch size (device time, max over ranks) and a cross-check of the gathered tokens
against its own single-GPU run of the whole batch."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import audiocodecs_b200 as A
from audiocodecs_b200 import shard
from oracle import weights

max_b = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
K = int(sys.argv[2]) if len(sys.argv) > 2 else 32
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
codec = A.Encodec(24000, 24000, num_codebooks=K, state_dict=weights.encodec_state_dict(0)).eval().to(dev)
T = 240000
b = world
while b <= max_b:
    sig = (torch.randn(b, T, generator=torch.Generator().manual_seed(b)) * 0.1).to(dev)   # the global batch, on every rank
    for _ in range(2):
        toks = shard.tokenize_sharded(codec, sig)
        shard.detokenize_sharded(codec, toks, gather=False)
    reps = max(2, min(10, 256 * world // b))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        toks = shard.tokenize_sharded(codec, sig)                    # every rank ends up with all B clips' tokens
        rec = shard.detokenize_sharded(codec, toks, gather=False)    # each rank decodes its own clips
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    line = {"n_gpus": world, "global_batch": b, "K": K, "ms_per_step": round(ms.item(), 3),
            "audio_s_per_s": round(b * 10 / ms.item() * 1e3, 1), "toks": list(toks.shape), "rec_local": list(rec.shape)}
    if rank == 0 and b <= 16:   # tiling depends on the batch shape: near-tied codes may differ between shard and whole-batch runs
        whole = codec.sig_to_toks(sig)
        line["match_vs_single_gpu"] = round((whole == toks).float().mean().item(), 5)
    if rank == 0:
        print(json.dumps(line), flush=True)
    del sig
    b *= 4
if world > 1:
    dist.destroy_process_group()
