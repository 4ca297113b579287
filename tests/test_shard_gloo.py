"""CPU, world_size 2 over gloo: the clip-sharding host logic of the N>1 path (audiocodecs_b200/shard.py).  The kernels
need a GPU, so a stand-in codec with the Codec call signature produces deterministic tokens from the samples; what is
under test is the split, the uneven-shard gather and the rank-independence of the result."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from audiocodecs_b200 import shard


class _FakeCodec:
    num_codebooks = 4

    def sig_to_toks(self, sig, length=None):
        frames = sig.reshape(sig.shape[0], -1, 10).sum(-1)  # [B, N]
        return (frames.abs() * 1000).long()[:, :, None] % torch.tensor([7, 11, 13, 17])

    def toks_to_sig(self, toks, length=None):
        return toks.float().sum(-1).repeat_interleave(10, dim=1)


def _worker(rank, world, port, B, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sig = torch.randn(B, 80, generator=torch.Generator().manual_seed(3))
    codec = _FakeCodec()
    toks = shard.tokenize_sharded(codec, sig)
    local = shard.tokenize_sharded(codec, sig, gather=False)
    rec = shard.detokenize_sharded(codec, toks)
    a, b = shard.shard_range(B, rank, world)
    ok = torch.equal(toks, codec.sig_to_toks(sig)) and torch.equal(local, codec.sig_to_toks(sig[a:b])) and \
        torch.equal(rec, codec.toks_to_sig(toks))
    out[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [8, 5, 3])
def test_two_rank_sharded_tokenize_gloo(B):
    world = 2
    port = 29600 + B
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, B, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 64, 1000):
        for world in (1, 2, 4, 8):
            spans = [shard.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
