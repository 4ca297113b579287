"""DAC behind the reference's `DAC` wrapper interface (R/audiocodecs/dac.py:30-130).

Same constructor arguments and tensor shapes.  The arithmetic of descript-audio-codec (un-vendored dependency of the
reference; twin: HF/dac/modeling_dac.py) is replaced by the sm_100a kernels of this package.
"""
import math
import re

import torch

from . import ops, packing, tc
from .codec import Codec
from .ops import ACT_NONE, ACT_SNAKE, EPI_NONE, EPI_TANH, PAD_ZERO, ConvSpec
from .tc import Act, Src, TcWeights

__all__ = ["DAC"]


class TcAlpha:
    """Snake alpha vector of one activation site (moves with the module like the packed weights)."""

    def __init__(self, t):
        self.t = t

    def apply(self, fn):
        self.t = fn(self.t)

def descript_to_hf_keys(sd):
    """State dict of a descript-audio-codec 1.0.0 `dac.DAC` (nn.Sequential indices: `encoder.block.N...`,
    `decoder.model.N...`, old-style `weight_g / weight_v`; dac/model/dac.py Encoder / EncoderBlock / ResidualUnit / Decoder /
    DecoderBlock) -> the `transformers.DacModel` names this wrapper packs from (HF/dac/modeling_dac.py:173-262,405-472).
    The weight-norm pairs are kept as they are (`packing.fold_weight_norm` folds either spelling); quantizer keys are equal
    in both layouts.  A dict that is already in HF format is returned unchanged."""
    if not any(k.startswith("decoder.model.") or re.match(r"encoder\.block\.\d+\.block\.", k) for k in sd):
        return sd
    unit = {"0": "snake1", "1": "conv1", "2": "snake2", "3": "conv2"}
    out = {}
    for k, v in sd.items():
        p = k.split(".")
        if p[0] == "encoder" and p[1] == "block":
            i = int(p[2])
            if i == 0:
                new = ["encoder", "conv1"] + p[3:]
            elif i == 5:
                new = ["encoder", "snake1"] + p[3:]
            elif i == 6:
                new = ["encoder", "conv2"] + p[3:]
            else:  # EncoderBlock i-1: block.{0,1,2} residual units, block.3 snake, block.4 strided conv
                j = int(p[4])
                if j < 3:
                    new = ["encoder", "block", str(i - 1), f"res_unit{j + 1}", unit[p[6]]] + p[7:]
                else:
                    new = ["encoder", "block", str(i - 1), "snake1" if j == 3 else "conv1"] + p[5:]
        elif p[0] == "decoder" and p[1] == "model":
            i = int(p[2])
            if i == 0:
                new = ["decoder", "conv1"] + p[3:]
            elif i == 5:
                new = ["decoder", "snake1"] + p[3:]
            elif i == 6:
                new = ["decoder", "conv2"] + p[3:]
            else:  # DecoderBlock i-1: block.0 snake, block.1 transposed conv, block.{2,3,4} residual units
                j = int(p[4])
                if j < 2:
                    new = ["decoder", "block", str(i - 1), "snake1" if j == 0 else "conv_t1"] + p[5:]
                else:
                    new = ["decoder", "block", str(i - 1), f"res_unit{j - 1}", unit[p[6]]] + p[7:]
        else:
            new = p
        out[".".join(new)] = v
    return out


_ARCH = {  # descript-audio-codec 1.0.0 model zoo: tag -> (encoder rates, decoder rates, n_codebooks)
    "44khz": ((2, 4, 8, 8), (8, 8, 4, 2), 9),
    "24khz": ((2, 4, 5, 8), (8, 5, 4, 2), 32),
    "16khz": ((2, 4, 5, 8), (8, 5, 4, 2), 12),
}


class DAC(Codec):
    """`DAC(sample_rate, orig_sample_rate=16000, mode="reconstruct", num_codebooks=8, latent=False)`; extra keywords
    `state_dict` (transformers.DacModel key format, or descript's weight_g/weight_v format) and `precision`
    ("exact" / "bf16" / "fp32", see `Encodec`: "exact" is the tensor path whose encoder reproduces the reference's tokens)."""

    max_chunk_samples = 64 * 441000  # ~0.2 KB of live activations per sample on the tensor path

    def _hop(self):
        return math.prod(self._enc_rates)  # 512 (44.1 kHz) / 320 (16, 24 kHz)

    # single-plane weights for the k7 convolutions of the residual units (64 % of the FLOPs): decoder SI-SNR 47.0 -> 45.9 dB,
    # end-to-end code match 99.8 -> 99.4 %, step 329 -> 266 ms (scripts/weight_precision_probe.py; the transposed convs and
    # the k1 convs are the precision-sensitive ones and keep the (hi, lo) pair)
    W_SINGLE = r"res_unit.\.conv1"

    def __init__(self, sample_rate, orig_sample_rate=16000, mode="reconstruct", num_codebooks=8, latent=False,
                 state_dict=None, precision="exact", split_min_ch=512, split_res_min_ch=64, w_single=None):
        super().__init__(sample_rate, orig_sample_rate, mode)
        self.w_single = w_single
        if precision not in ("exact", "fp16", "fp32", "bf16"):
            raise ValueError("precision must be 'exact' (tcgen05 tensor path, split-precision encoder: reference tokens), 'fp16' "
                             "(one fp16 product per MAC: fastest), 'bf16' (error-compensated bf16 products) or 'fp32' (SIMT path)")
        self.tensor_path = precision != "fp32"
        self.exact = precision == "exact"
        if self.tensor_path:
            self.pol_enc, self.pol_dec = tc.policies(precision, split_min_ch)
        # activated tensors (MMA operands) / raw residual-stream tensors with >= this many channels travel as (hi, lo) bf16 planes
        self.split_min_ch = split_min_ch
        self.split_res_min_ch = split_res_min_ch
        self.num_codebooks = num_codebooks
        self.vocab_size = 1024
        self.latent = latent
        self.precision = precision
        self.compute_dtype = {"exact": "f16", "fp16": "f16", "bf16": "bf16", "fp32": "f32"}[precision]
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
        self._build(descript_to_hf_keys(state_dict))

    def _pol(self, name):
        return self.pol_enc if name.startswith("encoder") else self.pol_dec

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
        # padding ceil(s/2), no output_padding (descript 1.0.0, HF/dac:243-249): L s for an even stride, L s - 1 for the odd
        # stride 5 of the 16 / 24 kHz models -- same 2-tap GEMM, one output row fewer
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
        self._tcw = []
        if self.tensor_path:
            self._build_tc(sd)
        nq = sum(1 for k in sd if k.startswith("quantizer.quantizers.") and k.endswith(".codebook.weight"))
        q = "quantizer.quantizers.{}."
        stack = lambda f: torch.stack([f(q.format(k)) for k in range(nq)]).contiguous()
        self.register_buffer("w_in", stack(lambda p: packing.fold_weight_norm(sd, p + "in_proj")[:, :, 0]), persistent=False)    # [S,8,1024]
        self.register_buffer("b_in", stack(lambda p: sd[p + "in_proj.bias"].float()), persistent=False)
        self.register_buffer("w_out", stack(lambda p: packing.fold_weight_norm(sd, p + "out_proj")[:, :, 0]), persistent=False)  # [S,1024,8]
        self.register_buffer("b_out", stack(lambda p: sd[p + "out_proj.bias"].float()), persistent=False)
        self.register_buffer("codebooks", stack(lambda p: sd[p + "codebook.weight"].float()), persistent=False)                  # [S,1024,8]
        self.register_buffer("_err", torch.zeros(1, dtype=torch.int32), persistent=False)
        if self.tensor_path and self.mode != "decode":
            self._build_rvq_proj()

    def _packed(self):
        return self._specs + self._tcw

    # ------------------------------------------------------------------ bf16 tensor path: packing
    def _tcw_conv(self, sd, prefix):
        """Conv1d [Cout,Cin,K] -> [Cout][K*Cin] (column = tap*Cin + c; a stride-s / kernel-2s conv read through the
        s-phase view has exactly this column order)."""
        w = packing.fold_weight_norm(sd, prefix)
        W = self._pol(prefix).weights(w.permute(0, 2, 1).reshape(w.shape[0], -1), sd[prefix + ".bias"], self._w_split(prefix))
        self._tcw.append(W)
        return W

    def _build_rvq_proj(self):
        """Constants of the projected RVQ encode (ops.dac_rvq_encode_proj): every stage's in_proj as one GEMM weight, and the
        8x8 cross terms W_in_k W_out_j that replace the 1024-wide residual update (computed in float64)."""
        S, D, H = self.w_in.shape
        win, wout, bout = self.w_in.double(), self.w_out.double(), self.b_out.double()
        n = -(-S * D // 16) * 16
        w_all = torch.zeros(n, H, dtype=torch.float64)
        w_all[: S * D] = win.reshape(S * D, H)
        b_all = torch.zeros(n)
        b_all[: S * D] = self.b_in.reshape(-1)
        # three products whatever the mode (a tiny GEMM whose rounding lands directly on the 8-dimensional decision)
        self._tproj = TcWeights(w_all.float(), b_all, split=True, f16=self.pol_enc.f16, hib=self.pol_enc.f16)
        self._tcw.append(self._tproj)
        cross = torch.einsum("kdc,jce->kjde", win, wout)                  # [S,S,8,8]
        v = torch.einsum("kdc,jc->kjd", win, bout)                        # [S,S,8]
        lower = torch.tril(torch.ones(S, S, dtype=torch.float64), -1)     # j < k
        cb_n = torch.nn.functional.normalize(self.codebooks, dim=-1)      # eps 1e-12, as F.normalize in the reference
        self.register_buffer("rvq_cross", cross.float().contiguous(), persistent=False)
        self.register_buffer("rvq_cconst", (-(v * lower[:, :, None]).sum(1)).float().contiguous(), persistent=False)
        self.register_buffer("cb_normed", cb_n.contiguous(), persistent=False)
        self.register_buffer("cb_norm2", (cb_n * cb_n).sum(-1).contiguous(), persistent=False)

    def _tcw_convtr(self, sd, prefix, stride):
        w = packing.fold_weight_norm(sd, prefix)     # [Cin, Cout, 2s]
        pk = packing.pack_convtr(w, stride)          # [2, Cin, s*Cout]
        W = self._pol(prefix).weights(pk.permute(2, 0, 1).reshape(pk.shape[2], -1), sd[prefix + ".bias"].float().repeat(stride), self._w_split(prefix))
        self._tcw.append(W)
        return W

    def _alpha(self, sd, name):
        a = TcAlpha(sd[name + ".alpha"].float().reshape(-1).contiguous())
        self._tcw.append(a)
        return a

    def _tc_units(self, sd, p):
        return [(self._alpha(sd, f"{p}.res_unit{u}.snake1"), self._tcw_conv(sd, f"{p}.res_unit{u}.conv1"), d,
                 self._alpha(sd, f"{p}.res_unit{u}.snake2"), self._tcw_conv(sd, f"{p}.res_unit{u}.conv2"))
                for u, d in ((1, 1), (2, 3), (3, 9))]

    def _build_tc(self, sd):
        if self.mode != "decode":
            self._tenc = []
            for i, s in enumerate(self._enc_rates):
                p = f"encoder.block.{i}"
                self._tenc.append((self._tc_units(sd, p), self._alpha(sd, p + ".snake1"), self._tcw_conv(sd, p + ".conv1"), s))
            self._tenc_last = (self._alpha(sd, "encoder.snake1"), self._tcw_conv(sd, "encoder.conv2"))
        if self.mode != "encode":
            self._tdec_first = self._tcw_conv(sd, "decoder.conv1")
            self._tdec = []
            for i, s in enumerate(self._dec_rates):
                p = f"decoder.block.{i}"
                self._tdec.append((self._alpha(sd, p + ".snake1"), self._tcw_convtr(sd, p + ".conv_t1", s), s, self._tc_units(sd, p)))
            self._tdec_last_alpha = self._alpha(sd, "decoder.snake1")
            pd = self.pol_dec  # Cout = 1 as a stride-16 conv with 16 outputs
            self._tdec_last = tc.last_conv_weights_phased(self._dec[-1], split=True if pd.w_split is None else pd.w_split, f16=pd.f16)
            self._tcw.append(self._tdec_last)

    # ------------------------------------------------------------------ bf16 tensor path: execution
    def _split_res(self, pol, C):
        """the raw residual stream (touched by epilogues only, never an MMA operand): with bf16 hi planes it needs the lo
        plane from 64 channels up (8 bits per skip-add is not enough); one fp16 plane (11 bits) is, so "fp16" carries none
        (a third less traffic in the unit kernels); "exact" always carries it"""
        return pol.full or (not pol.f16 and C >= self.split_res_min_ch)

    def _tc_run_units(self, units, x, xs, next_alpha, out_halo=(0, 0), enc=False):
        pol = self.pol_enc if enc else self.pol_dec
        """three DacResidualUnits (HF/dac:173-207): x raw, xs = snake1(x) -> (y raw, ys = next_alpha-snake(y)).  The k7
        conv reads its 7 dilated taps from ONE staged block of xs (zero padding = TMA out-of-bounds fill); the 1x1 conv
        adds the residual and writes the raw stream plus the activation its consumer applies."""
        B, L, C = x.B, x.L, x.C
        dev = x.buf.device
        for i, (a1, W7, d, a2, W1) in enumerate(units):
            last = i == len(units) - 1
            nxt = next_alpha if last else units[i + 1][0]
            y = None if last else pol.act(B, L, C, dev, split=self._split_res(pol, C))
            hl, hr = out_halo if last else (0, 0)
            ys = pol.act(B, L, C, dev, hl=hl, hr=hr)
            a = Src(xs, taps=7, dilation=d, shift=-3 * d)

            def unfused(a=a, x=x, y=y, ys=ys, W7=W7, W1=W1, a2=a2, nxt=nxt):
                hs = pol.act(B, L, C, dev)
                tc.conv_tc(W7, [a], L, y_act=hs, act=ACT_SNAKE, alpha=a2.t, name="res_k7_tc")
                tc.conv_tc(W1, [Src(hs)], L, res=x, y=y, y_act=ys, act=ACT_SNAKE, alpha=nxt.t, name="res_k1_tc")

            def fused(g, dbl, io, a=a, x=x, y=y, ys=ys, W7=W7, W1=W1, a2=a2, nxt=nxt):
                return lambda: tc.resunit_tc(W7, W1, a, L, res=x, y=y, y_act=ys, act1=ACT_SNAKE, alpha1=a2.t, act2=ACT_SNAKE,
                                             alpha2=nxt.t, h_split=pol.split(C), g_hint=g, dbl_hint=dbl, io_stage=io, name="resunit_tc")

            # one fused launch (hidden tensor on chip) when both accumulators fit tensor memory, or two tap-GEMM launches.
            # Encoder: fused whenever it fits -- a rule, because the two forms group the fp32 accumulation differently and a
            # clip's tokens must not depend on the batch it was tuned in (the tuner only picks the bit-identical tile grouping /
            # buffering).  Decoder: the measured-fastest form per layer shape (waveforms agree to ~1e-5 either way)
            variants = [("unfused", unfused)]
            if 2 * C <= 512:
                # io: 1 = skip input / outputs staged in shared memory and moved by TMA, -1 = direct loads / stores (bit-identical)
                fv = [(f"fused_g{g}_d{dbl}_io{io}", fused(g, dbl, io)) for g in (2, 1) for dbl in (2, 1, 0) for io in (-1, 1) if not (dbl == 2 and io == 1)]
                variants = fv if enc else fv + variants
            tc.autotune(("dac_unit", B, L, C, d, last, x.lo is not None, xs.lo is not None, enc, xs.f16, W7.planes, W1.planes), variants)
            x, xs = y, ys
        return xs

    def _encoder_tc(self, sig, proj=False):
        B, T = sig.shape
        dev = sig.device
        C = self._enc[0].cout
        pol = self.pol_enc
        x = pol.act(B, T, C, dev, split=pol.full)   # not "exact": the first layer's outputs stay single planes
        xs = pol.act(B, T, C, dev, split=pol.full)
        ops.conv_first_bf16(self._enc[0], sig, y=x, y_act=xs, act=ACT_SNAKE, alpha=self._tenc[0][0][0][0].t)
        L = T
        for bi, (units, a_down, Wdown, s) in enumerate(self._tenc):
            p = math.ceil(s / 2)
            hr = -(p + L) % s
            ys = self._tc_run_units(units, x, xs, a_down, out_halo=(p, hr), enc=True)
            ys.fill_halo(PAD_ZERO)
            Lout = (L + 2 * p - 2 * s) // s + 1
            C = 2 * C
            nxt = self._tenc[bi + 1][0][0][0] if bi + 1 < len(self._tenc) else self._tenc_last[0]
            last = bi + 1 == len(self._tenc)
            x = None if last else pol.act(B, Lout, C, dev, split=self._split_res(pol, C))
            xs = pol.act(B, Lout, C, dev)
            tc.conv_tc(Wdown, [Src(ys, taps=2, origin=-p, phases=s, rows=(p + L + hr) // s)], Lout, y=x, y_act=xs, act=ACT_SNAKE,
                       alpha=nxt.t, name="down_tc")
            L = Lout
        if proj:  # latents as a split-bf16 activation -> all RVQ stages' in_proj in one GEMM: [B, L, 8S (padded to 16)] fp32
            za = pol.act(B, L, C, dev, split=True)
            tc.conv_tc(self._tenc_last[1], [Src(xs, taps=3, shift=-1)], L, y=za, name="conv_k3_tc")
            P = torch.empty((B, L, self._tproj.n_total), device=dev, dtype=torch.float32)
            tc.conv_tc(self._tproj, [Src(za)], L, y32=P, name="rvq_in_proj_tc")
            return P
        z = torch.empty((B, L, C), device=dev, dtype=torch.float32)
        tc.conv_tc(self._tenc_last[1], [Src(xs, taps=3, shift=-1)], L, y32=z, name="conv_k3_tc")
        return z

    def _decoder_tc(self, zq):
        B, N, C = zq.shape
        dev = zq.device
        pol = self.pol_dec
        z = pol.act(B, N, C, dev)
        ops.f32_to_act(zq.contiguous(), z)
        C = self._tdec_first.n_total
        xs = pol.act(B, N, C, dev)
        tc.conv_tc(self._tdec_first, [Src(z, taps=7, shift=-3)], N, y_act=xs, act=ACT_SNAKE, alpha=self._tdec[0][0].t, name="conv_k7_tc")
        L = N
        for bi, (a_up, Wtr, s, units) in enumerate(self._tdec):
            p = math.ceil(s / 2)
            C = C // 2
            Lout = L * s + s - 2 * p  # (L - 1) s - 2 p + 2 s: L s, or L s - 1 for an odd stride
            x = pol.act(B, Lout, C, dev, split=self._split_res(pol, C))
            us = pol.act(B, Lout, C, dev)
            # transposed conv (k = 2s, stride s, padding p): 2-tap GEMM over n = (phase, cout), flat output shifted by p*C
            tc.conv_tc(Wtr, [Src(xs, taps=2, shift=-1)], L + 1, y=x, y_act=us, act=ACT_SNAKE, alpha=units[0][0].t, act_mod=C,
                       out_rows=Lout, out_ch=C, out_shift=p * C, name="convtr_tc")
            nxt = self._tdec[bi + 1][0] if bi + 1 < len(self._tdec) else self._tdec_last_alpha
            # the last layer (k7, zero padding 3) reads 16-sample view rows: 3 zero rows in front, the rest of one more view
            # row (13 for a length that is a multiple of 16) behind
            xs = self._tc_run_units(units, x, us, nxt, out_halo=(3, tc.last_conv_right_halo(Lout, 3)) if bi + 1 == len(self._tdec) else (0, 0))
            L = Lout
        # last layer (Cout = 1, k7, zero padding 3, tanh) on the tap-GEMM kernel, 16 samples per GEMM row
        xs.fill_halo(PAD_ZERO)
        return tc.conv_last_phased(self._tdec_last, xs, tanh=True)

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

    def _encode_latents(self, sig):
        if self.tensor_path:
            return self._encoder_tc(sig.contiguous())
        return self._stack(self._enc, sig.contiguous()[:, :, None])

    def _sig_to_toks(self, sig, length):  # R/audiocodecs/dac.py:94-100 (`length` is ignored by the reference)
        if self.tensor_path:  # 8-dimensional RVQ chain on the projected latents (no 1024-wide residual)
            P = self._encoder_tc(sig.contiguous(), proj=True)
            return ops.dac_rvq_encode_proj(P, self.rvq_cconst, self.rvq_cross, self.cb_normed, self.cb_norm2, self.codebooks,
                                           self.num_codebooks)
        z = self._encode_latents(sig)
        return ops.dac_rvq_encode(z, self.w_in, self.b_in, self.codebooks, self.w_out, self.b_out, self.num_codebooks)

    def _sig_to_feats(self, sig, length):  # R/audiocodecs/dac.py:103-112
        z = self._encode_latents(sig)
        if self.latent:
            w = self.w_in[0].t().contiguous()[None]  # [1,1024,8]
            return ops.conv(ConvSpec(w, self.b_in[0], cout=8, geometry="same"), z)
        return z

    def _sig_to_qfeats(self, sig, length):  # R/audiocodecs/dac.py:115-121
        z = self._encode_latents(sig)
        return ops.dac_rvq_encode(z, self.w_in, self.b_in, self.codebooks, self.w_out, self.b_out, self.num_codebooks,
                                  want_zq=True)[1]

    def _toks_to_qfeats(self, toks, length):
        K = toks.shape[-1]
        return ops.dac_rvq_decode(toks, self.codebooks[:K], self.w_out[:K], self.b_out[:K], err_flag=self._err)

    def _toks_to_sig(self, toks, length):  # R/audiocodecs/dac.py:124-130
        if self.tensor_path:
            return self._decoder_tc(self._toks_to_qfeats(toks, length))
        return self._stack(self._dec, self._toks_to_qfeats(toks, length))[:, :, 0]
