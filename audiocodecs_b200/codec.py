"""Codec interface: host-side mirror of `audiocodecs.codec.Codec` (R/audiocodecs/codec.py:33-214).

Same constructor arguments, mode dispatch, default `length` and resample-around-the-hooks
behaviour; the resampler is our polyphase FIR kernel instead of torchaudio.
"""
from abc import ABC, abstractmethod

import re

import torch

from . import ops

__all__ = ["Codec"]


class Codec(torch.nn.Module, ABC):
    _MODES = ["encode", "decode", "reconstruct"]
    # Activations of one call stay resident in HBM; a batch larger than this many audio samples (clips x samples at the
    # codec rate) is processed in sub-batches so that any batch size fits the 180 GB of a B200 (the clips are
    # independent, so chunking changes nothing but peak memory).  Sub-classes set it from their per-sample footprint.
    max_chunk_samples = 256 * 240000

    def __init__(self, sample_rate, orig_sample_rate, mode="reconstruct"):
        super().__init__()
        if mode not in self._MODES:
            raise ValueError(f"`mode` ({mode}) must be one of {self._MODES}")  # R/codec.py:38-39
        self.sample_rate = sample_rate
        self.orig_sample_rate = orig_sample_rate
        self.mode = mode

    # R/codec.py:45-55
    def forward(self, input, length=None):
        if self.mode == "encode":
            return self.sig_to_toks(input, length)
        if self.mode == "decode":
            return self.toks_to_sig(input, length)
        toks = self.sig_to_toks(input, length)
        return self.toks_to_sig(toks, length)

    def _prep_sig(self, sig, length):
        sig = ops.resample(sig.float(), self.sample_rate, self.orig_sample_rate)
        # the reference substitutes ones(B) here (R/codec.py:64-65); `None` carries the same meaning to the
        # hooks without spending kernels on an all-true padding mask
        return sig, length

    # bf16 tensor path: layers whose weights are stored as ONE bf16 plane instead of the (hi, lo) pair -- one tensor-core
    # product fewer per MAC where the measured SI-SNR margin allows it (regex over state-dict prefixes; None = no layer)
    W_SINGLE = None

    def _w_split(self, name):
        pat = getattr(self, "w_single", None)
        pat = self.W_SINGLE if pat is None else pat
        return not (pat and re.search(pat, name))

    def _chunks(self, n_clips, samples_per_clip):
        per = max(1, int(self.max_chunk_samples // max(1, samples_per_clip)))
        return [(a, min(n_clips, a + per)) for a in range(0, n_clips, per)]

    def _chunked(self, fn, x, length, samples_per_clip):
        spans = self._chunks(x.shape[0], samples_per_clip)
        if len(spans) <= 1:
            return fn(x, length)
        return torch.cat([fn(x[a:b], None if length is None else length[a:b]) for a, b in spans], dim=0)

    @torch.no_grad()
    def sig_to_toks(self, sig, length=None):  # R/codec.py:57-66
        sig, length = self._prep_sig(sig, length)
        return self._chunked(self._sig_to_toks, sig, length, sig.shape[-1])

    @torch.no_grad()
    def sig_to_feats(self, sig, length=None):  # R/codec.py:68-77
        sig, length = self._prep_sig(sig, length)
        return self._sig_to_feats(sig, length)

    @torch.no_grad()
    def sig_to_qfeats(self, sig, length=None):  # R/codec.py:79-88
        sig, length = self._prep_sig(sig, length)
        return self._sig_to_qfeats(sig, length)

    @torch.no_grad()
    def toks_to_sig(self, toks, length=None):  # R/codec.py:90-100
        sig = self._chunked(self._toks_to_sig, toks, length, toks.shape[1] * self._hop())
        return ops.resample(sig, self.orig_sample_rate, self.sample_rate)

    def _hop(self):
        """codec-rate samples per token frame (used only to size sub-batches)"""
        return 320

    @torch.no_grad()
    def toks_to_qfeats(self, toks, length=None):  # R/codec.py:102-108
        return self._toks_to_qfeats(toks, length)

    @abstractmethod
    def embs(self):
        raise NotImplementedError

    @abstractmethod
    def _sig_to_toks(self, sig, length):
        raise NotImplementedError

    @abstractmethod
    def _sig_to_feats(self, sig, length):
        raise NotImplementedError

    @abstractmethod
    def _sig_to_qfeats(self, sig, length):
        raise NotImplementedError

    @abstractmethod
    def _toks_to_sig(self, toks, length):
        raise NotImplementedError

    def _toks_to_qfeats(self, toks, length):
        raise NotImplementedError

    # ---- plumbing shared by the wrappers
    def _packed(self):
        """objects with .apply(fn) holding packed weights outside the nn.Module buffer registry"""
        return []

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        for spec in self._packed():
            spec.apply(fn)
        return out
