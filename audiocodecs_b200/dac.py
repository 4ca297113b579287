"""DAC behind the reference's `DAC` wrapper interface (R/audiocodecs/dac.py:30-130).

Same constructor arguments and tensor shapes.  The arithmetic of descript-audio-codec (un-vendored dependency of the
reference; twin: HF/dac/modeling_dac.py) is replaced by the sm_100a kernels of this package.
"""
import math

import torch

from . import ops, packing
from .codec import Codec
from .ops import ACT_NONE, ACT_SNAKE, EPI_NONE, EPI_TANH, PAD_ZERO, ConvSpec

__all__ = ["DAC"]

_ARCH = {  # descript-audio-codec 1.0.0 model zoo: tag -> (encoder rates, decoder rates, n_codebooks)
    "44khz": ((2, 4, 8, 8), (8, 8, 4, 2), 9),
    "24khz": ((2, 4, 5, 8), (8, 5, 4, 2), 32),
    "16khz": ((2, 4, 5, 8), (8, 5, 4, 2), 12),
}


class DAC(Codec):
    """`DAC(sample_rate, orig_sample_rate=16000, mode="reconstruct", num_codebooks=8, latent=False)`; extra keywords
    `state_dict` (transformers.DacModel key format, or descript's weight_g/weight_v format) and `precision`."""

    def __init__(self, sample_rate, orig_sample_rate=16000, mode="reconstruct", num_codebooks=8, latent=False,
                 state_dict=None, precision="fp32"):
        super().__init__(sample_rate, orig_sample_rate, mode)
        if precision not in ("fp32",):
            raise ValueError("DAC currently runs on the exact fp32 path only (precision='fp32')")
        self.num_codebooks = num_codebooks
        self.vocab_size = 1024
        self.latent = latent
        self.precision = precision
        self.compute_dtype = "f32"
        tag = f"{int(orig_sample_rate / 1000)}khz"  # R/audiocodecs/dac.py:55
        if tag not in _ARCH:
            raise ValueError(f"no DAC model for {tag}")
        self._enc_rates, self._dec_rates, self._max_k = _ARCH[tag]
        if state_dict is None:
            try:
                import dac
            except ImportError:
                raise ImportError("`pip install descript-audio-codec` to use this module")
            state_dict = dac.DAC.load(str(dac.utils.download(model_type=tag))).state_dict()
        self._build(state_dict)

    # ------------------------------------------------------------------ packing
    def _conv(self, sd, prefix, stride=1, dilation=1, padding=0, snake=None, epi=EPI_NONE):
        w = packing.fold_weight_norm(sd, prefix)
        cout, _, k = w.shape
        alpha = sd[snake + ".alpha"].float().reshape(-1).contiguous() if snake else None
        spec = ConvSpec(packing.pack_conv(w), sd[prefix + ".bias"].float().clone(), cout=cout, kernel=k, stride=stride,
                        dilation=dilation, geometry="same", pad_mode=PAD_ZERO, padding=padding,
                        act=ACT_SNAKE if snake else ACT_NONE, alpha=alpha, epi=epi)
        self._specs.append(spec)
        return spec

    def _convtr(self, sd, prefix, stride, snake):
        w = packing.fold_weight_norm(sd, prefix)  # [Cin, Cout, 2s]
        if stride % 2:
            raise NotImplementedError("odd-stride transposed convs (24/16 kHz DAC) are not on the fast path yet")
        spec = ConvSpec(packing.pack_convtr(w, stride), sd[prefix + ".bias"].float().repeat(stride), cout=w.shape[1],
                        geometry="tr", tr_stride=stride, tr_pad=math.ceil(stride / 2), act=ACT_SNAKE,
                        alpha=sd[snake + ".alpha"].float().reshape(-1).contiguous())
        self._specs.append(spec)
        return spec

    def _res_unit(self, sd, p, d):
        return (self._conv(sd, p + ".conv1", dilation=d, padding=3 * d, snake=p + ".snake1"),
                self._conv(sd, p + ".conv2", snake=p + ".snake2"))

    def _build(self, sd):
        self._specs = []
        if self.mode != "decode":
            enc = [self._conv(sd, "encoder.conv1", padding=3)]
            for i, s in enumerate(self._enc_rates):
                p = f"encoder.block.{i}"
                for u, d in ((1, 1), (2, 3), (3, 9)):
                    enc.append(self._res_unit(sd, f"{p}.res_unit{u}", d))
                enc.append(self._conv(sd, p + ".conv1", stride=s, padding=math.ceil(s / 2), snake=p + ".snake1"))
            enc.append(self._conv(sd, "encoder.conv2", padding=1, snake="encoder.snake1"))
            self._enc = enc
        if self.mode != "encode":
            dec = [self._conv(sd, "decoder.conv1", padding=3)]
            for i, s in enumerate(self._dec_rates):
                p = f"decoder.block.{i}"
                dec.append(self._convtr(sd, p + ".conv_t1", s, p + ".snake1"))
                for u, d in ((1, 1), (2, 3), (3, 9)):
                    dec.append(self._res_unit(sd, f"{p}.res_unit{u}", d))
            dec.append(self._conv(sd, "decoder.conv2", padding=3, snake="decoder.snake1", epi=EPI_TANH))
            self._dec = dec
        nq = sum(1 for k in sd if k.startswith("quantizer.quantizers.") and k.endswith(".codebook.weight"))
        q = "quantizer.quantizers.{}."
        stack = lambda f: torch.stack([f(q.format(k)) for k in range(nq)]).contiguous()
        self.register_buffer("w_in", stack(lambda p: packing.fold_weight_norm(sd, p + "in_proj")[:, :, 0]), persistent=False)    # [S,8,1024]
        self.register_buffer("b_in", stack(lambda p: sd[p + "in_proj.bias"].float()), persistent=False)
        self.register_buffer("w_out", stack(lambda p: packing.fold_weight_norm(sd, p + "out_proj")[:, :, 0]), persistent=False)  # [S,1024,8]
        self.register_buffer("b_out", stack(lambda p: sd[p + "out_proj.bias"].float()), persistent=False)
        self.register_buffer("codebooks", stack(lambda p: sd[p + "codebook.weight"].float()), persistent=False)                  # [S,1024,8]
        self.register_buffer("_err", torch.zeros(1, dtype=torch.int32), persistent=False)

    def _packed(self):
        return self._specs

    # ------------------------------------------------------------------ pieces
    def _stack(self, layers, x):
        for layer in layers:
            if isinstance(layer, tuple):  # DacResidualUnit: x + conv1x1(snake(conv7(snake(x)))) (HF/dac:173-207)
                x = ops.conv(layer[1], ops.conv(layer[0], x), res=x)
            else:
                x = ops.conv(layer, x)
        return x

    # ------------------------------------------------------------------ Codec hooks
    @torch.no_grad()
    def embs(self):  # R/audiocodecs/dac.py:66-91
        K = self.num_codebooks
        if self.latent:
            return self.codebooks[:K].clone()
        # post-projection embeddings: out_proj_k(codebook_k) -> [K, C, 1024]
        return torch.einsum("kcd,khd->kch", self.codebooks[:K], self.w_out[:K]) + self.b_out[:K, None, :]

    def _sig_to_toks(self, sig, length):  # R/audiocodecs/dac.py:94-100 (`length` is ignored by the reference)
        z = self._stack(self._enc, sig.contiguous()[:, :, None])
        return ops.dac_rvq_encode(z, self.w_in, self.b_in, self.codebooks, self.w_out, self.b_out, self.num_codebooks)

    def _sig_to_feats(self, sig, length):  # R/audiocodecs/dac.py:103-112
        z = self._stack(self._enc, sig.contiguous()[:, :, None])
        if self.latent:
            w = self.w_in[0].t().contiguous()[None]  # [1,1024,8]
            return ops.conv(ConvSpec(w, self.b_in[0], cout=8, geometry="same"), z)
        return z

    def _sig_to_qfeats(self, sig, length):  # R/audiocodecs/dac.py:115-121
        z = self._stack(self._enc, sig.contiguous()[:, :, None])
        return ops.dac_rvq_encode(z, self.w_in, self.b_in, self.codebooks, self.w_out, self.b_out, self.num_codebooks,
                                  want_zq=True)[1]

    def _toks_to_qfeats(self, toks, length):
        K = toks.shape[-1]
        return ops.dac_rvq_decode(toks, self.codebooks[:K], self.w_out[:K], self.b_out[:K], err_flag=self._err)

    def _toks_to_sig(self, toks, length):  # R/audiocodecs/dac.py:124-130
        return self._stack(self._dec, self._toks_to_qfeats(toks, length))[:, :, 0]
