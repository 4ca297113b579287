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
    # toks_to_sig / toks_to_qfeats check the token range on entry and raise IndexError like the reference's embedding lookup
    # (one tiny device reduction + a host sync per call; skipped while a CUDA graph is being captured).  Set to False to
    # drop the sync: out-of-range tokens then decode as code 0 and only set the kernels' error flag (`codec._err`).
    strict_tokens = True

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
        return self.reconstruct(input, length)

    @torch.no_grad()
    def reconstruct(self, sig, length=None):
        """toks_to_sig(sig_to_toks(sig)) (mode="reconstruct", R/codec.py:52-54) without the token range check -- the tokens
        come from this codec's own encoder -- so nothing synchronises the host between the two halves."""
        toks = self.sig_to_toks(sig, length)
        with self._on(toks):
            out = self._chunked(self._toks_to_sig, toks, length, toks.shape[1] * self._hop())
            return ops.resample(out, self.orig_sample_rate, self.sample_rate)

    def _on(self, t):
        """device guard: every launch of a call goes to the device (and its current stream) that holds the input -- the C-ABI
        takes raw pointers and a stream, it has no device of its own.  The packed weights must live on the same device."""
        if not t.is_cuda:
            raise RuntimeError("audiocodecs_b200 runs on CUDA (sm_100a) tensors only; there is no CPU fallback. "
                               "Move the codec and its inputs with .to('cuda').")
        own = next(self.buffers()).device
        if own != t.device:
            raise RuntimeError(f"input on {t.device} but the codec's weights are on {own}: move one of them with .to()")
        return torch.cuda.device(t.device)

    def _check_toks(self, toks):
        if not self.strict_tokens or torch.cuda.is_current_stream_capturing():
            return
        if toks.dtype.is_floating_point or toks.dtype == torch.bool:
            raise TypeError(f"tokens must be an integer tensor, got {toks.dtype}")
        lo, hi = torch.aminmax(toks)
        lo, hi = int(lo), int(hi)
        if lo < 0 or hi >= self.vocab_size:
            raise IndexError(f"token values must lie in [0, {self.vocab_size}), got [{lo}, {hi}]")

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
        with self._on(sig):
            sig, length = self._prep_sig(sig, length)
            return self._chunked(self._sig_to_toks, sig, length, sig.shape[-1])

    @torch.no_grad()
    def sig_to_feats(self, sig, length=None):  # R/codec.py:68-77
        with self._on(sig):
            sig, length = self._prep_sig(sig, length)
            return self._sig_to_feats(sig, length)

    @torch.no_grad()
    def sig_to_qfeats(self, sig, length=None):  # R/codec.py:79-88
        with self._on(sig):
            sig, length = self._prep_sig(sig, length)
            return self._sig_to_qfeats(sig, length)

    @torch.no_grad()
    def toks_to_sig(self, toks, length=None):  # R/codec.py:90-100
        with self._on(toks):
            self._check_toks(toks)
            sig = self._chunked(self._toks_to_sig, toks, length, toks.shape[1] * self._hop())
            return ops.resample(sig, self.orig_sample_rate, self.sample_rate)

    def _hop(self):
        """codec-rate samples per token frame (used only to size sub-batches)"""
        return 320

    @torch.no_grad()
    def toks_to_qfeats(self, toks, length=None):  # R/codec.py:102-108
        with self._on(toks):
            self._check_toks(toks)
            return self._toks_to_qfeats(toks, length)

    def feats_to_sig(self, feats, length=None):  # R/codec.py:109-119 (EnCodec / DAC / Mimi define no `_feats_to_sig`)
        sig = self._feats_to_sig(feats, length)  # NotImplementedError for these three codecs, as in the reference
        with self._on(sig):
            return ops.resample(sig, self.orig_sample_rate, self.sample_rate)

    # ---- token augmentation (R/codec.py:121-180): the step right after the tokenizer in the downstream recipes.  Host-side
    # torch ops over `embs()` -- sampling is not on the hot path.
    @torch.no_grad()
    def logits(self):
        """[K, C, C]: minus the pairwise Euclidean distance between the code vectors of each codebook, -inf on the diagonal
        (a code never resamples to itself).  Cached after the first call, returned as a copy (R/codec.py:150-159)."""
        if getattr(self, "_pair_logits", None) is None:
            e = self.embs()
            lg = torch.cdist(e, e).neg_()
            lg.diagonal(dim1=-2, dim2=-1).fill_(float("-inf"))
            self._pair_logits = lg
        return self._pair_logits.clone()

    def resample(self, toks, p=0.2, temp=1.0, top_k=None, top_p=None):
        """toks [B, N, K] -> a copy in which each token is replaced, with probability p, by a code drawn from
        softmax(-distance / temp) around it (optionally restricted to the top_k nearest / the top_p nucleus)."""
        if p <= 0.0:
            return toks
        if top_k is not None and top_p is not None:
            raise NotImplementedError
        out = toks.clone()
        K = out.shape[-1]
        lg = self.logits().to(out.device)                                   # [K, C, C]
        per_book = out.reshape(-1, K).t()                                   # [K, B*N]
        rows = lg.gather(1, per_book[..., None].expand(-1, -1, lg.shape[-1]))  # [K, B*N, C]: the row of each current code
        probs = (rows.reshape(-1, lg.shape[-1]) / temp).softmax(dim=-1)     # [K*B*N, C]
        if top_k is not None:
            draw = self._sample_top_k(probs, top_k)
        elif top_p is not None:
            draw = self._sample_top_p(probs, top_p)
        else:
            draw = probs.multinomial(num_samples=1)[:, 0]
        draw = draw.reshape(K, -1).t().reshape(out.shape)
        replace = (torch.rand(out.shape) < p).to(out.device)                # drawn on the CPU generator, as the reference does
        out[replace] = draw[replace]
        return out

    def _sample_top_k(self, probs, k):  # R/codec.py:161-168
        top, idx = probs.topk(k, dim=-1)
        pick = (top / top.sum(dim=-1, keepdim=True)).multinomial(num_samples=1)
        return idx.gather(-1, pick)[:, 0]

    def _sample_top_p(self, probs, p):  # R/codec.py:170-180: keep the smallest prefix of the sorted codes whose mass exceeds p
        srt, idx = probs.sort(dim=-1, descending=True)
        before = srt.cumsum(dim=-1) - srt
        srt = srt.masked_fill(before > p, 0.0)
        pick = (srt / srt.sum(dim=-1, keepdim=True)).multinomial(num_samples=1)
        return idx.gather(-1, pick)[:, 0]

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

    def _feats_to_sig(self, feats, length):
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
