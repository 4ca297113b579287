"""Clip sharding across the GPUs of one box.

Every clip is independent on this path, so rank r encodes / decodes clips [start_r, stop_r) with replicated weights and
there is NO collective on the compute path; the only exchange is one all_gather of the int64 codes (and of the
waveforms when asked) after the kernels have run, over `torch.distributed` (NCCL over NVLink on the GPUs; gloo in the
CPU tests of the host logic).
"""
import torch
import torch.distributed as dist

__all__ = ["shard_range", "gather_rows", "tokenize_sharded", "detokenize_sharded"]


def shard_range(n, rank, world):
    """contiguous near-equal split of n clips: the first n % world ranks take one extra clip"""
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def _codec_device(codec, fallback):
    """the device holding the codec's weights (an nn.Module has no .device attribute); test doubles without buffers fall
    back to the input's device"""
    bufs = getattr(codec, "buffers", None)
    if bufs is not None:
        for b in bufs():
            return b.device
    return getattr(codec, "device", None) or fallback


def _world(group):
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def gather_rows(local, n_total, group=None):
    """all_gather of per-rank row blocks with uneven counts: local [n_r, ...] -> [n_total, ...] on every rank.
    Shards are padded to the largest count (all_gather needs equal shapes) and trimmed after the exchange."""
    rank, world = _world(group)
    if world == 1:
        return local
    counts = [shard_range(n_total, r, world) for r in range(world)]
    cap = max(b - a for a, b in counts)
    pad = local.new_zeros((cap,) + tuple(local.shape[1:]))
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad.contiguous(), group=group)
    return torch.cat([p[: b - a] for p, (a, b) in zip(parts, counts)], dim=0)


def tokenize_sharded(codec, sig, length=None, group=None, gather=True):
    """`sig` [B, T] is the GLOBAL batch (same on every rank, host or device): this rank tokenizes its slice on its own
    GPU; with gather=True every rank returns all B clips' tokens [B, N, K] (int64), else only its own slice."""
    rank, world = _world(group)
    a, b = shard_range(sig.shape[0], rank, world)
    dev = _codec_device(codec, sig.device)
    local_len = None if length is None else length[a:b].to(dev)
    toks = codec.sig_to_toks(sig[a:b].to(dev), local_len) if b > a else None
    if toks is None:  # more ranks than clips: learn the token shape from a peer through the padded gather
        shape = [None]
        if gather and world > 1:
            dist.broadcast_object_list(shape, src=0, group=group)
        n, k = shape[0] if shape[0] else (0, getattr(codec, "num_codebooks", 1))
        toks = torch.empty((0, n, k), dtype=torch.int64, device=dev)
    elif gather and world > sig.shape[0]:
        dist.broadcast_object_list([tuple(toks.shape[1:])] if rank == 0 else [None], src=0, group=group)
    return gather_rows(toks, sig.shape[0], group) if gather else toks


def detokenize_sharded(codec, toks, length=None, group=None, gather=True):
    """`toks` [B, N, K] global -> waveforms; same sharding as `tokenize_sharded`."""
    rank, world = _world(group)
    a, b = shard_range(toks.shape[0], rank, world)
    dev = _codec_device(codec, toks.device)
    if b == a:
        raise ValueError("detokenize_sharded needs at least one clip per rank")
    sig = codec.toks_to_sig(toks[a:b].to(dev), None if length is None else length[a:b].to(dev))
    return gather_rows(sig, toks.shape[0], group) if gather else sig
