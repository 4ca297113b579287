"""EnCodec-24k behind the reference's `Encodec` wrapper interface (R/audiocodecs/encodec.py:30-149).

Same constructor arguments and tensor shapes; the `transformers.EncodecModel` arithmetic
(HF/encodec/modeling_encodec.py) is replaced by the sm_100a kernels of this package.
"""
import torch

from . import ops, packing
from .codec import Codec
from .ops import ACT_ELU, ACT_NONE, PAD_REFLECT, ConvSpec

__all__ = ["Encodec"]

RATIOS = (8, 5, 4, 2)  # facebook/encodec_24khz upsampling_ratios (HF/encodec/configuration_encodec.py)
_VALID_BW = (1.5, 3.0, 6.0, 12.0, 24.0)


class Encodec(Codec):
    """`Encodec(sample_rate, orig_sample_rate=24000, mode="reconstruct", num_codebooks=8, use_vocos=False)`.

    Extra keyword `state_dict`: weights in `transformers.EncodecModel` key format.  When omitted the
    pretrained checkpoint is fetched exactly like the reference does (`EncodecModel.from_pretrained`,
    R/audiocodecs/encodec.py:51) and only its state dict is kept.
    """

    def __init__(self, sample_rate, orig_sample_rate=24000, mode="reconstruct", num_codebooks=8, use_vocos=False,
                 state_dict=None):
        super().__init__(sample_rate, orig_sample_rate, mode)
        if use_vocos:
            raise NotImplementedError("the Vocos decoder branch (R/audiocodecs/encodec.py:53-66) is outside this build")
        self.num_codebooks = num_codebooks
        self.compute_dtype = "f32"
        self.use_vocos = use_vocos
        self.vocab_size = 1024
        tag = int(orig_sample_rate / 1000)
        if tag != 24:
            raise NotImplementedError("only facebook/encodec_24khz (mono, unchunked) is usable through this wrapper "
                                      "(SURVEY.md section 5: the 48 kHz model is stereo/chunked)")
        self.bandwidth = (num_codebooks * 75) / 100  # R/audiocodecs/encodec.py:50
        if state_dict is None:
            try:
                from transformers import EncodecModel
            except ImportError:
                raise ImportError("`pip install transformers>=4.31.0` to use this module")
            state_dict = EncodecModel.from_pretrained(f"facebook/encodec_{tag}khz").state_dict()
        self._build(state_dict)

    # ------------------------------------------------------------------ packing
    def _conv(self, sd, prefix, stride=1, act=ACT_NONE):
        w = packing.fold_weight_norm(sd, prefix + ".conv")
        cout, _, k = w.shape
        spec = ConvSpec(packing.pack_conv(w), sd[prefix + ".conv.bias"].float().clone(), cout=cout, kernel=k,
                        stride=stride, geometry="causal", pad_mode=PAD_REFLECT, act=act)
        self._specs.append(spec)
        return spec

    def _convtr(self, sd, prefix, stride, act):
        w = packing.fold_weight_norm(sd, prefix + ".conv")  # [Cin, Cout, 2s], g per input channel
        cout = w.shape[1]
        spec = ConvSpec(packing.pack_convtr(w, stride), sd[prefix + ".conv.bias"].float().repeat(stride), cout=cout,
                        geometry="tr", tr_stride=stride, tr_pad=0, act=act)
        self._specs.append(spec)
        return spec

    def _resblock(self, sd, prefix):
        # HF/encodec:252-282: shortcut(x) + conv_k1(ELU(conv_k3(ELU(x))))
        return (self._conv(sd, prefix + ".block.1", act=ACT_ELU), self._conv(sd, prefix + ".block.3", act=ACT_ELU),
                self._conv(sd, prefix + ".shortcut"))

    def _lstm(self, sd, prefix, name):
        layers = []
        for l in range(2):
            w_ih = sd[f"{prefix}.lstm.weight_ih_l{l}"].float()
            bias = (sd[f"{prefix}.lstm.bias_ih_l{l}"] + sd[f"{prefix}.lstm.bias_hh_l{l}"]).float()
            spec = ConvSpec(packing.pack_linear(w_ih), bias, cout=w_ih.shape[0], geometry="causal")
            self._specs.append(spec)
            self.register_buffer(f"{name}_whh{l}", sd[f"{prefix}.lstm.weight_hh_l{l}"].float().contiguous(),
                                 persistent=False)
            layers.append((spec, f"{name}_whh{l}"))
        return layers

    def _build(self, sd):
        self._specs = []
        if self.mode != "decode":  # R/audiocodecs/encodec.py:67-71 drops the unused half
            enc = [self._conv(sd, "encoder.layers.0")]
            idx = 1
            for r in reversed(RATIOS):
                enc.append(self._resblock(sd, f"encoder.layers.{idx}"))
                enc.append(self._conv(sd, f"encoder.layers.{idx + 2}", stride=r, act=ACT_ELU))
                idx += 3
            self._enc = enc
            self._enc_lstm = self._lstm(sd, f"encoder.layers.{idx}", "enc")
            self._enc_last = self._conv(sd, f"encoder.layers.{idx + 2}", act=ACT_ELU)
        if self.mode != "encode":
            self._dec_first = self._conv(sd, "decoder.layers.0")
            self._dec_lstm = self._lstm(sd, "decoder.layers.1", "dec")
            dec = []
            idx = 3
            for r in RATIOS:
                dec.append(self._convtr(sd, f"decoder.layers.{idx}", r, ACT_ELU))
                dec.append(self._resblock(sd, f"decoder.layers.{idx + 1}"))
                idx += 3
            self._dec = dec
            self._dec_last = self._conv(sd, f"decoder.layers.{idx}", act=ACT_ELU)
        nq = sum(1 for k in sd if k.startswith("quantizer.layers.") and k.endswith(".codebook.embed"))
        cb = torch.stack([sd[f"quantizer.layers.{k}.codebook.embed"].float() for k in range(nq)]).contiguous()
        self.register_buffer("codebooks", cb, persistent=False)                 # [32, 1024, 128]
        self.register_buffer("cb_norm", cb.pow(2).sum(-1).contiguous(), persistent=False)  # |E|^2, HF/encodec:367
        self.register_buffer("_sync_ws", torch.zeros(64, dtype=torch.int32), persistent=False)
        self.register_buffer("_err", torch.zeros(1, dtype=torch.int32), persistent=False)

    def _packed(self):
        return self._specs

    # ------------------------------------------------------------------ pieces
    def _num_quantizers(self):
        # HF/encodec:573-578 validates the bandwidth, :416-422 maps it to a stage count
        if self.bandwidth not in _VALID_BW:
            raise ValueError(f"This model doesn't support the bandwidth {self.bandwidth}. Select one of {list(_VALID_BW)}.")
        return int(max(1, (self.bandwidth * 1000) // (10 * 75)))

    def _run_resblock(self, rb, x):
        c3, c1, sc = rb
        h = ops.conv(c3, x)
        y = ops.conv(sc, x)
        return ops.conv(c1, h, res=y, out=y)

    def _run_lstm(self, layers, x):
        (ih0, hh0), (ih1, hh1) = layers
        h0 = ops.lstm_layer(ops.conv(ih0, x), getattr(self, hh0), None, self._sync_ws)
        return ops.lstm_layer(ops.conv(ih1, h0), getattr(self, hh1), x, self._sync_ws)

    def _encoder(self, sig, vlen=None):
        """sig [B,T] -> embeddings [B,N,128] channels-last (HF/encodec:285-313)."""
        x = ops.conv(self._enc[0], sig.contiguous()[:, :, None], vlen=vlen)
        for layer in self._enc[1:]:
            x = self._run_resblock(layer, x) if isinstance(layer, tuple) else ops.conv(layer, x)
        x = self._run_lstm(self._enc_lstm, x)
        return ops.conv(self._enc_last, x)

    def _decoder(self, z):
        """z [B,N,128] -> sig [B, 320N] (HF/encodec:316-347)."""
        x = ops.conv(self._dec_first, z)
        x = self._run_lstm(self._dec_lstm, x)
        for layer in self._dec:
            x = self._run_resblock(layer, x) if isinstance(layer, tuple) else ops.conv(layer, x)
        return ops.conv(self._dec_last, x)[:, :, 0]

    def _vlen(self, sig, length):
        """padding mask of the reference (R/audiocodecs/encodec.py:84-89, HF/encodec:599-601) as a per-clip
        valid sample count consumed by the first conv; None when every clip is full length."""
        if length is None:
            return None
        abs_lens = sig.shape[-1] * length.to(sig.device, torch.float32)
        # sample t is kept iff t < abs_len  <=>  t < ceil(abs_len)
        return torch.ceil(abs_lens).clamp(max=sig.shape[-1]).to(torch.int32)

    # ------------------------------------------------------------------ Codec hooks
    @torch.no_grad()
    def embs(self):  # R/audiocodecs/encodec.py:74-79
        return self.codebooks[: self.num_codebooks].clone()

    def _sig_to_toks(self, sig, length):
        nq = self._num_quantizers()
        emb = self._encoder(sig, self._vlen(sig, length))
        B, N, D = emb.shape
        toks = torch.empty((B, N, nq), device=sig.device, dtype=torch.int64)
        ops.rvq_encode(emb.view(B * N, D), self.codebooks, self.cb_norm, toks.view(B * N, nq), nq)
        return toks  # [B, N, K]

    def _sig_to_feats(self, sig, length):  # R/audiocodecs/encodec.py:97-117 (normalize=False: mask unused)
        return self._encoder(sig)

    def _sig_to_qfeats(self, sig, length):  # R/audiocodecs/encodec.py:120-127
        return self._toks_to_qfeats(self._sig_to_toks(sig, length), length)

    def _toks_to_qfeats(self, toks, length):  # R/audiocodecs/encodec.py:144-149
        B, N, K = toks.shape
        toks = toks.to(torch.int64).contiguous()
        q = ops.rvq_decode(toks.view(B * N, K), self.codebooks, K, err_flag=self._err)
        return q.view(B, N, -1)

    def _toks_to_sig(self, toks, length):  # R/audiocodecs/encodec.py:130-141
        return self._decoder(self._toks_to_qfeats(toks, length))
