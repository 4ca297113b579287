"""Serving-side plumbing: run the codec over batches that live in HOST memory, overlapping the copies with compute.

The reference's callers do `sig.to(device)` -> codec -> `.cpu()` one batch after the other on one stream, so every
host<->device copy sits between two kernels.  Here the input copy of batch i+1 and the output copy of batch i-1 run on
their own streams (the copy engines) while the kernels of batch i occupy the SMs; CUDA events carry the dependencies and
nothing synchronises the host until the caller asks for a result.  PyTorch is used for streams / events / pinned memory
only; the compute is the same `sig_to_toks` / `toks_to_sig` call as everywhere else.
"""
import torch

__all__ = ["HostPipeline"]


class HostPipeline:
    def __init__(self, codec, fn=None, depth=2):
        """fn(codec, device_batch) -> device tensor (default: reconstruct = toks_to_sig(sig_to_toks(x)))."""
        self.codec = codec
        self.fn = fn or (lambda c, x: c.reconstruct(x))
        self.depth = depth
        self.s_in = torch.cuda.Stream()
        self.s_out = torch.cuda.Stream()
        self._dev_in = [None] * depth
        self._in_ready = [torch.cuda.Event() for _ in range(depth)]
        self._in_free = [torch.cuda.Event() for _ in range(depth)]
        self._out_done = [torch.cuda.Event() for _ in range(depth)]
        self._step = 0

    def submit(self, host_batch, host_out=None):
        """Enqueue one pinned host batch; returns (host_out, event): the result is in `host_out` once `event` has
        completed (`event.synchronize()`).  `host_out` (pinned) is allocated when not given."""
        assert host_batch.is_pinned(), "HostPipeline needs pinned host memory for asynchronous copies"
        slot = self._step % self.depth
        dev = next(self.codec.buffers()).device
        compute = torch.cuda.current_stream(dev)
        if self._dev_in[slot] is None or self._dev_in[slot].shape != host_batch.shape:
            self._dev_in[slot] = torch.empty(host_batch.shape, dtype=host_batch.dtype, device=dev)
        with torch.cuda.stream(self.s_in):
            if self._step >= self.depth:
                self.s_in.wait_event(self._in_free[slot])      # the kernels of batch i-depth have consumed this buffer
            self._dev_in[slot].copy_(host_batch, non_blocking=True)
            self._in_ready[slot].record(self.s_in)
        compute.wait_event(self._in_ready[slot])
        out = self.fn(self.codec, self._dev_in[slot])
        self._in_free[slot].record(compute)
        done = torch.cuda.Event()
        done.record(compute)
        if host_out is None:
            host_out = torch.empty(out.shape, dtype=out.dtype).pin_memory()
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(done)
            host_out.copy_(out, non_blocking=True)
            out.record_stream(self.s_out)
            self._out_done[slot].record(self.s_out)
        self._step += 1
        return host_out, self._out_done[slot]

    def drain(self):
        self.s_in.synchronize()
        self.s_out.synchronize()
        torch.cuda.current_stream().synchronize()
