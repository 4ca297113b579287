"""CUDA-graph replay of the tokenize / detokenize calls for one input shape.

At batch 1 the path is launch-bound (EnCodec: ~60 kernel launches for 16 s of audio, Mimi: ~150), so the reference's own
speed metric -- real-time factor at batch 1, R/downstream/test_sr.py:264-270 -- is dominated by host-side launch cost.
`GraphedCodec` records `sig_to_toks` and `toks_to_sig` once per (batch, samples) shape into two CUDA graphs (the kernels,
their tensor maps and every intermediate buffer, which then lives in the graph's private memory pool) and replays them:
one driver call per direction.  The kernels are the same ones, so the results are bit-identical to the eager calls.
"""
import torch

__all__ = ["GraphedCodec"]


class GraphedCodec:
    def __init__(self, codec, example_sig, length=None, warmup=3):
        """codec: an audiocodecs_b200 codec on a CUDA device (eval mode); example_sig [B, T] fixes the shape."""
        if not example_sig.is_cuda:
            raise RuntimeError("GraphedCodec needs a CUDA tensor: the hot path has no CPU fallback")
        self.codec = codec
        self.length = length
        self._sig = example_sig.clone()
        side = torch.cuda.Stream(device=example_sig.device)
        side.wait_stream(torch.cuda.current_stream(example_sig.device))
        with torch.cuda.stream(side), torch.no_grad():  # per-shape autotuning, lazy function attributes, filter caches
            for _ in range(warmup):
                toks = codec.sig_to_toks(self._sig, length)
                codec.toks_to_sig(toks, length)
        torch.cuda.current_stream(example_sig.device).wait_stream(side)
        self._g_enc, self._g_dec = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.no_grad():
            with torch.cuda.graph(self._g_enc):
                self._toks_out = codec.sig_to_toks(self._sig, length)
            self._toks_in = self._toks_out.contiguous().clone()
            with torch.cuda.graph(self._g_dec):
                self._rec = codec.toks_to_sig(self._toks_in, length)

    def sig_to_toks(self, sig):
        """sig [B, T] (the captured shape) -> toks [B, N, K] int64.  The result is the graph's output buffer: it is
        overwritten by the next call (clone it to keep it)."""
        if tuple(sig.shape) != tuple(self._sig.shape):
            raise ValueError(f"captured for input shape {tuple(self._sig.shape)}, got {tuple(sig.shape)}")
        self._sig.copy_(sig, non_blocking=True)
        self._g_enc.replay()
        return self._toks_out

    def toks_to_sig(self, toks):
        if tuple(toks.shape) != tuple(self._toks_in.shape):
            raise ValueError(f"captured for token shape {tuple(self._toks_in.shape)}, got {tuple(toks.shape)}")
        self._toks_in.copy_(toks, non_blocking=True)
        self._g_dec.replay()
        return self._rec

    def reconstruct(self, sig):
        return self.toks_to_sig(self.sig_to_toks(sig))
