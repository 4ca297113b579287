"""Token consumers: host-side mirrors of the two classes the reference's downstream recipes apply right after the tokenizer
(SURVEY 8f-4), over the kernels of csrc/consumers.cu.

  CodebookUtil        R/downstream/metrics/codebook_util.py:28-83 -- same constructor, `append(hyp_toks, lens=None)`,
                      `summarize(field=None)` and the same summary arithmetic (codebook utilisation %, normalised entropy %).
                      The reference counts with K x `unique(return_counts=True)` + a host-side index-add per utterance; here the
                      counts stay on the device (one histogram pass, 8 bytes per token) and only `summarize` reads them back.
                      It does not derive from speechbrain's MetricStats (speechbrain is not a dependency of this package).
  MultiHeadEmbedding  R/downstream/models/multihead.py:28-69 -- an `nn.Embedding` subclass with the reference's constructor
                      (`vocab_size` int or list, `embedding_dim`, `num_codebooks`, `padding_idx`) and state dict; forward =
                      one gather kernel (the offset add, the padding remap and the lookup fused), backward = index-add.
"""
import ctypes
import math

import torch

from . import _lib, ops

__all__ = ["CodebookUtil", "MultiHeadEmbedding"]


class CodebookUtil:
    def __init__(self, num_codebooks, vocab_size):
        self.num_codebooks = num_codebooks
        self.vocab_size = vocab_size
        self.vocab_sizes = [vocab_size] * num_codebooks
        self.clear()

    def clear(self):
        self._counts = None          # int64 [K, V] on the tokens' device
        self._err = None
        self.total_toks = 0
        self.summary = {}

    @property
    def toks_count_per_codebook(self):
        """the reference's attribute: one float count vector per codebook (host)"""
        c = torch.zeros(self.num_codebooks, self.vocab_size) if self._counts is None else self._counts.cpu().float()
        return [c[k] for k in range(self.num_codebooks)]

    @torch.no_grad()
    def append(self, hyp_toks, lens=None):
        assert hyp_toks.ndim == 3
        assert hyp_toks.shape[0] == 1, "Batch size must be 1"   # as the reference (codebook_util.py:41-42)
        self.append_batch(hyp_toks)

    @torch.no_grad()
    def append_batch(self, toks):
        """any [B, N, K] batch at once (the reference's one-utterance restriction is an artefact of its per-call unique())"""
        assert toks.ndim == 3 and toks.shape[-1] == self.num_codebooks
        ops._need_cuda(toks)
        toks = toks.to(torch.int64).contiguous()
        with torch.cuda.device(toks.device):
            if self._counts is None:
                self._counts = torch.zeros((self.num_codebooks, self.vocab_size), dtype=torch.int64, device=toks.device)
                self._err = torch.zeros(1, dtype=torch.int32, device=toks.device)
            rows = toks.shape[0] * toks.shape[1]
            _lib.check(_lib.lib().ac_token_histogram(ops._ptr(toks), rows, self.num_codebooks, self.vocab_size, ops._ptr(self._counts),
                                                     ops._ptr(self._err), ops._stream()), "ac_token_histogram")
        self.total_toks += rows

    def summarize(self, field=None):
        if self._err is not None and int(self._err.item()):
            raise IndexError(f"token values must lie in [0, {self.vocab_size})")
        utils, ents = [], []
        for counts, vocab_size in zip(self.toks_count_per_codebook, self.vocab_sizes):
            probs = counts / self.total_toks
            valid = probs > 0
            p = probs[valid]
            entropy = -(p * p.log2()).sum()
            n_valid = valid.sum()
            if n_valid > 1:
                utils.append(n_valid / vocab_size)
                ents.append(entropy / math.log2(n_valid))
            else:
                utils.append(0)
                ents.append(0.0)
        self.summary = {"codebook_util": round(100 * torch.tensor(sum(utils) / len(utils)).item(), 2),
                        "norm_entropy": round(100 * torch.tensor(sum(ents) / len(ents)).item(), 2)}
        return self.summary[field] if field is not None else self.summary


class _Lookup(torch.autograd.Function):
    @staticmethod
    def forward(ctx, toks, weight, offsets, vocab, padding_row):
        rows, K = toks.numel() // toks.shape[-1], toks.shape[-1]
        D = weight.shape[1]
        out = torch.empty(tuple(toks.shape) + (D,), device=weight.device, dtype=torch.float32)
        err = torch.zeros(1, dtype=torch.int32, device=weight.device)
        with torch.cuda.device(weight.device):
            _lib.check(_lib.lib().ac_multihead_embedding(ops._ptr(toks), ops._ptr(weight), ops._ptr(offsets), ops._ptr(out), rows, K, D,
                                                         vocab, padding_row, weight.shape[0], ops._ptr(err), ops._stream()),
                       "ac_multihead_embedding")
        ctx.save_for_backward(toks, offsets)
        ctx.meta = (weight.shape, vocab, padding_row)
        ctx.err = err
        return out

    @staticmethod
    def backward(ctx, grad):
        toks, offsets = ctx.saved_tensors
        shape, vocab, padding_row = ctx.meta
        idx = toks + offsets
        if padding_row >= 0:
            idx = torch.where(toks == vocab, torch.full_like(idx, padding_row), idx)
        gw = torch.zeros(shape, device=grad.device, dtype=grad.dtype)
        g = grad.reshape(-1, shape[1])
        flat = idx.reshape(-1)
        if padding_row >= 0:  # nn.Embedding(padding_idx) accumulates no gradient into the padding row
            keep = flat != padding_row
            g, flat = g[keep], flat[keep]
        gw.index_add_(0, flat, g)
        return None, gw, None, None, None


class MultiHeadEmbedding(torch.nn.Embedding):
    def __init__(self, vocab_size, embedding_dim, num_codebooks, padding_idx=False, **kwargs):
        if isinstance(vocab_size, (list, tuple)):
            assert len(vocab_size) == num_codebooks, [len(vocab_size), num_codebooks]
            num_embeddings = int(sum(vocab_size))
            offsets = torch.tensor([0] + list(vocab_size[:-1])).cumsum(dim=-1)
        else:
            num_embeddings = vocab_size * num_codebooks
            offsets = torch.arange(0, num_embeddings, vocab_size)
        if padding_idx:
            padding_idx = num_embeddings
            num_embeddings += 1
        else:
            padding_idx = None
        super().__init__(num_embeddings, embedding_dim, padding_idx, **kwargs)
        self.offsets = offsets
        self.vocab_size = vocab_size
        self.num_codebooks = num_codebooks

    def forward(self, input):
        """input [..., K] integer tokens -> [..., K, embedding_dim]"""
        ops._need_cuda(input, self.weight)
        if self.max_norm is not None or self.weight.dtype != torch.float32 or self.embedding_dim % 4:
            raise NotImplementedError("MultiHeadEmbedding kernel: fp32 weights, embedding_dim % 4 == 0, no max_norm renormalisation")
        if isinstance(self.vocab_size, (list, tuple)) and self.padding_idx is not None:
            raise NotImplementedError("padding with per-head vocabulary sizes compares a tensor with a list in the reference")
        toks = input.to(torch.int64).contiguous()
        offsets = self.offsets.to(toks.device)
        vocab = self.vocab_size if not isinstance(self.vocab_size, (list, tuple)) else -1
        return _Lookup.apply(toks, self.weight, offsets, vocab, -1 if self.padding_idx is None else self.padding_idx)
