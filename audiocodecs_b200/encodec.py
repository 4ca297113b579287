"""EnCodec-24k behind the reference's `Encodec` wrapper interface (R/audiocodecs/encodec.py:30-149).

Same constructor arguments and tensor shapes; the `transformers.EncodecModel` arithmetic
(HF/encodec/modeling_encodec.py) is replaced by the sm_100a kernels of this package.
"""
import torch

from . import ops, packing, tc
from .codec import Codec
from .ops import ACT_ELU, ACT_NONE, PAD_REFLECT, ConvSpec
from .tc import Act, Src, TcWeights

__all__ = ["Encodec"]

RATIOS = (8, 5, 4, 2)  # facebook/encodec_24khz upsampling_ratios (HF/encodec/configuration_encodec.py)
SPLIT_MIN_CH = 128     # precision="bf16": activations with >= this many channels travel as (hi, lo) bf16 pairs (DESIGN.md,
                       # precision); precision="exact": every ENCODER activation does (three products per MAC)
RAW_MAX_CH = int(__import__("os").environ.get("AC_RAW_MAX_CH", "0"))   # residual blocks with <= this many channels take RAW x and
                       # apply their input ELU on chip (raw mode of ac_resunit_tc: the producer layer writes one tensor instead of a raw
                       # and an activated copy).  Off: measured twice a net loss.  Round 1 (bf16, 64 channels): producers 25-35 %
                       # faster, blocks 1.8x slower.  Round 2 (all formats, 32 channels): first layer 0.67 -> 0.39 ms (exact) /
                       # 0.40 -> 0.30 ms (fp16) but the block 0.95 -> 1.57 ms / 0.64 -> 0.93 ms -- the transform stage sits between
                       # TMA and MMA on a ring the extra planes shrink to two stages
FUSED_MAX_CH = int(__import__("os").environ.get("AC_FUSED_MAX_CH", "64"))   # residual blocks with <= this many channels run as ONE fused launch (hidden tile on chip), wider ones as two
                       # tap-GEMM launches.  Re-measured in round 2 (after the epilogue diet and the ping-pong groups): at 128 channels fused is
                       # 0.35 vs 0.40 ms (fp16) / 0.76 vs 0.85 ms (exact) -- 0.07 ms of a 17.4 ms step in the bench loop, not worth
                       # re-grouping the encoder's accumulation -- and at 256 it loses (0.65-0.80 vs 0.48 ms exact).  A rule, not a timing: the two forms group
                       # the fp32 accumulation differently, so the choice must not depend on the batch size.  None = let the
                       # tuner time one against the other (AC_TUNE_FUSION=1 python scripts/layer_times.py)
_VALID_BW = (1.5, 3.0, 6.0, 12.0, 24.0)


class Encodec(Codec):
    """`Encodec(sample_rate, orig_sample_rate=24000, mode="reconstruct", num_codebooks=8, use_vocos=False)`.

    Extra keyword `state_dict`: weights in `transformers.EncodecModel` key format.  When omitted the
    pretrained checkpoint is fetched exactly like the reference does (`EncodecModel.from_pretrained`,
    R/audiocodecs/encodec.py:51) and only its state dict is kept.

    `precision`:
      "exact" (default) tcgen05 tensor path whose ENCODER carries every activation as an (fp16 hi, bf16 lo) pair and every
              weight as an fp16 (hi, lo) pair: A_hi W_hi + A_hi W_lo + A_lo bf16(W), fp32 accumulate, ~2^-20 operand
              precision (+ fp16 LSTM recurrence): embeddings within ~1e-5 of the fp32 reference, so the tokens equal the
              reference's wherever its top-2 distance gap exceeds 1e-4.  The decoder is the "fp16" one.
      "fp16"  fastest: ONE fp16 product per MAC (11-bit operands, fp32 accumulate): decoder SI-SNR ~55 dB, ~98.5 % of the
              tokens equal the reference's.
      "bf16"  the error-compensated bf16 path of round 1 (two to three bf16 products per MAC; kept for comparison).
      "fp32"  SIMT fp32 kernels (debugging / parity reference; ~15x slower).
    """

    # every layer keeps the (hi, lo) weight pair: with single-plane weights everywhere the decoder SI-SNR falls to 39.9 dB
    # (residual blocks alone: 44.0 dB, their 1x1 tails: 42.3 dB) and encoder-side rounding shows up as code flips
    # (93.4 -> 90.3 % code match), scripts/weight_precision_probe.py.  Layer names here are `encoder.layers.N`,
    # `decoder.layers.N`, `....block.1`, `....tail`.
    W_SINGLE = None

    def __init__(self, sample_rate, orig_sample_rate=24000, mode="reconstruct", num_codebooks=8, use_vocos=False,
                 state_dict=None, precision="exact", w_single=None):
        super().__init__(sample_rate, orig_sample_rate, mode)
        self.w_single = w_single
        if use_vocos:
            raise NotImplementedError("the Vocos decoder branch (R/audiocodecs/encodec.py:53-66) is outside this build")
        self.num_codebooks = num_codebooks
        if precision not in ("exact", "fp16", "bf16", "fp32"):
            raise ValueError("precision must be 'exact' (tcgen05 tensor path, split-precision encoder: reference tokens), 'fp16' "
                             "(tcgen05 tensor path, one fp16 product per MAC: fastest), 'bf16' (error-compensated bf16 products) "
                             "or 'fp32' (SIMT path)")
        self.precision = precision
        self.tensor_path = precision != "fp32"
        self.exact = precision == "exact"
        if self.tensor_path:
            self.pol_enc, self.pol_dec = tc.policies(precision, SPLIT_MIN_CH)
        self.compute_dtype = {"exact": "f16", "fp16": "f16", "bf16": "bf16", "fp32": "f32"}[precision]
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
            if self.tensor_path:
                # fp16 operands for the recurrence (W_hh in tensor memory, h fed back as fp16): 11-bit mantissas at the
                # cost of one product; |W_hh| <= 1/sqrt(512)-ish and |h| < 1 are far inside the fp16 range.  CPU emulation
                # (scripts/exact_mode_emulation.py): tokens identical to the fp32 oracle's, against 99.94 % with bf16 operands
                whh = sd[f"{prefix}.lstm.weight_hh_l{l}"].float()
                assert whh.abs().max() < 6.0e4, "W_hh outside the fp16 range"
                self.register_buffer(f"{name}_whh{l}_16", whh.to(torch.float16).contiguous(), persistent=False)
            layers.append((spec, f"{name}_whh{l}"))
        return layers

    def _build(self, sd):
        self._specs, self._tcw = [], []
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
        if self.mode != "encode":
            import copy
            self._dec_last_noact = copy.copy(self._dec_last)
            self._dec_last_noact.act = ACT_NONE
            self._specs.append(self._dec_last_noact)
        if self.tensor_path:
            self._build_tc(sd)
        nq = sum(1 for k in sd if k.startswith("quantizer.layers.") and k.endswith(".codebook.embed"))
        cb = torch.stack([sd[f"quantizer.layers.{k}.codebook.embed"].float() for k in range(nq)]).contiguous()
        self.register_buffer("codebooks", cb, persistent=False)                 # [32, 1024, 128]
        self.register_buffer("cb_norm", cb.pow(2).sum(-1).contiguous(), persistent=False)  # |E|^2, HF/encodec:367
        if self.tensor_path:  # operand planes of the tensor-core distance GEMM: bf16(E), bf16(E - bf16(E))
            hi = cb.to(torch.bfloat16)
            self.register_buffer("cb_split", torch.stack([hi, (cb - hi.float()).to(torch.bfloat16)]).contiguous(), persistent=False)
        self.register_buffer("_sync_ws", torch.zeros(64, dtype=torch.int32), persistent=False)
        self.register_buffer("_err", torch.zeros(1, dtype=torch.int32), persistent=False)

    def _packed(self):
        return self._specs + self._tcw

    # ------------------------------------------------------------------ bf16 tensor-path packing
    def _pol(self, prefix):
        return self.pol_enc if prefix.startswith("encoder") else self.pol_dec

    def _tc_conv(self, sd, prefix):
        """Conv1d [Cout,Cin,K] -> bf16 [Cout][K*Cin] (column = tap*Cin + c; a stride-s/kernel-2s conv read through the
        s-phase view has exactly this column order)."""
        w = packing.fold_weight_norm(sd, prefix + ".conv")
        W = self._pol(prefix).weights(w.permute(0, 2, 1).reshape(w.shape[0], -1), sd[prefix + ".conv.bias"], self._w_split(prefix))
        self._tcw.append(W)
        return W

    def _tc_convtr(self, sd, prefix, stride):
        w = packing.fold_weight_norm(sd, prefix + ".conv")  # [Cin, Cout, 2s]
        p = packing.pack_convtr(w, stride)                  # [2, Cin, s*Cout]
        W = self._pol(prefix).weights(p.permute(2, 0, 1).reshape(p.shape[2], -1), sd[prefix + ".conv.bias"].float().repeat(stride), self._w_split(prefix))
        self._tcw.append(W)
        return W

    def _tc_resblock(self, sd, prefix):
        k3 = self._tc_conv(sd, prefix + ".block.1")
        wsc = packing.fold_weight_norm(sd, prefix + ".shortcut.conv")[:, :, 0]
        w1 = packing.fold_weight_norm(sd, prefix + ".block.3.conv")[:, :, 0]
        # second GEMM of the fused unit: columns = [hidden (conv_k1) | raw x (1x1 shortcut)], the two biases summed
        tail = self._pol(prefix).weights(torch.cat([w1, wsc], dim=1), sd[prefix + ".shortcut.conv.bias"] + sd[prefix + ".block.3.conv.bias"],
                                         self._w_split(prefix + ".tail"))
        self._tcw.append(tail)
        return k3, tail

    def _tc_lstm(self, sd, prefix):
        out = []
        for l in range(2):
            # decoder / "bf16" encoder: single bf16 product for the input projection (a split changes the decoder SI-SNR by
            # < 0.1 dB: 45.1 -> 45.1 dB).  "exact" encoder: three products -- with one the embedding error is 1.8e-4 instead of
            # 1.9e-5 and 0.15 % of the safe tokens flip (scripts/exact_mode_emulation.py)
            W = self._pol(prefix).weights(sd[f"{prefix}.lstm.weight_ih_l{l}"], sd[f"{prefix}.lstm.bias_ih_l{l}"] + sd[f"{prefix}.lstm.bias_hh_l{l}"],
                                          False)
            self._tcw.append(W)
            out.append(W)
        return out

    def _build_tc(self, sd):
        if self.mode != "decode":
            idx, self._tenc = 1, []
            for r in reversed(RATIOS):
                self._tenc.append((self._tc_resblock(sd, f"encoder.layers.{idx}"), self._tc_conv(sd, f"encoder.layers.{idx + 2}"), r))
                idx += 3
            self._tenc_lstm = self._tc_lstm(sd, f"encoder.layers.{idx}")
            self._tenc_last = self._tc_conv(sd, f"encoder.layers.{idx + 2}")
        if self.mode != "encode":
            self._tdec_first = self._tc_conv(sd, "decoder.layers.0")
            self._tdec_lstm = self._tc_lstm(sd, "decoder.layers.1")
            idx, self._tdec = 3, []
            for r in RATIOS:
                self._tdec.append((self._tc_convtr(sd, f"decoder.layers.{idx}", r), self._tc_resblock(sd, f"decoder.layers.{idx + 1}"), r))
                idx += 3
            pd = self.pol_dec  # Cout = 1 as a stride-16 conv with 16 outputs
            self._tdec_last = tc.last_conv_weights_phased(self._dec_last, split=True if pd.w_split is None else pd.w_split, f16=pd.f16)
            self._tcw.append(self._tdec_last)

    # ------------------------------------------------------------------ bf16 tensor-path execution
    def _tc_run_lstm(self, Ws, whh, x: Act, final: Act, pol):
        """x raw [B,N,512] -> final = ELU(lstm(x) + x) (HF/encodec:236-249 + the following ELU).  With the "exact" encoder
        policy the input projections read the (hi, lo) pairs of x / h0 (three products)."""
        B, N, C = x.B, x.L, x.C
        dev = x.buf.device
        exact = pol.full
        pre = torch.empty((B, N, 4 * C), device=dev, dtype=torch.float32)
        tc.conv_tc(Ws[0], [Src(x if exact else x.hi_only())], N, y32=pre, name="lstm_ih_tc")
        h0 = pol.act(B, N, C, dev, split=True)   # not "exact": the lo plane matters for the skip-add of the last layer's h only
        ops.lstm_tc(pre, getattr(self, whh[0] + "_16"), out=h0)
        tc.conv_tc(Ws[1], [Src(h0 if exact else h0.hi_only())], N, y32=pre, name="lstm_ih_tc")
        # the skip-add + ELU runs as its own HBM-bound pass: inside the recurrence kernel its loads/stores sat on the
        # per-step critical path (layer 1 took 3.8 ms against 2.4 ms for layer 0)
        ops.lstm_tc(pre, getattr(self, whh[1] + "_16"), out=h0)
        ops.add_act_bf16(h0, x, final, ACT_ELU)

    def _tc_resblock_run(self, Wk3, Wtail, x: Act, xe: Act, ye: Act, pol):
        """x raw, xe = ELU(x) with a 2-row reflect halo -> ye = ELU(shortcut(x) + conv1(ELU(conv3(xe)))).
        Either ONE fused launch with the hidden activation kept on chip (ac_resunit_tc; tile grouping and double
        buffering tuned per shape) or two tap-GEMM launches -- whichever measures faster for this layer shape.
        xe is None in raw mode (C <= RAW_MAX_CH): x itself carries the 2-row halo, the kernel applies the input ELU on
        chip and reads the raw rows of the same staged blocks for the shortcut."""
        B, L, C = x.B, x.L, x.C
        hs = pol.split(C // 2)
        if xe is None:
            x.fill_halo(PAD_REFLECT, 3 if L <= 2 else 0)
            a = Src(x, taps=3, origin=-2, rows=L + 2)

            def raw(g, dbl):
                return lambda: tc.resunit_tc(Wk3, Wtail, a, L, y_act=ye, act1=ACT_ELU, act2=ACT_ELU, h_split=hs, act0=ACT_ELU,
                                             e_split=x.lo is not None, x_from_a=True, g_hint=g, dbl_hint=dbl, name="resblock_tc")

            tc.autotune(("encodec_resblock_raw", B, L, C, x.lo is not None, Wk3.planes, Wtail.planes), [(f"raw_g{g}_d{dbl}", raw(g, dbl)) for g in (4, 2, 1) for dbl in (1, 0)])
            return
        xe.fill_halo(PAD_REFLECT, 3 if L <= 2 else 0)
        a = Src(xe, taps=3, origin=-2, rows=L + 2)

        def unfused():
            he = pol.act(B, L, C // 2, x.buf.device, split=hs)
            tc.conv_tc(Wk3, [a], L, y_act=he, act=ACT_ELU, name="res_k3_tc")
            tc.conv_tc(Wtail, [Src(he), Src(x)], L, y_act=ye, act=ACT_ELU, name="res_tail_tc")

        def fused(g, dbl, io):
            return lambda: tc.resunit_tc(Wk3, Wtail, a, L, x=x, y_act=ye, act1=ACT_ELU, act2=ACT_ELU, h_split=hs, g_hint=g, dbl_hint=dbl,
                                         io_stage=io, name="resblock_tc")

        # io: 1 = outputs staged in shared memory and sent out by TMA stores, -1 = direct stores (bit-identical results)
        variants = [(f"fused_g{g}_d{dbl}_io{io}", fused(g, dbl, io)) for g in (4, 2, 1) for dbl in (2, 1, 0) for io in (-1, 1) if not (dbl == 2 and io == 1)]
        if FUSED_MAX_CH is None:
            variants.append(("unfused", unfused))
        elif C > FUSED_MAX_CH:
            variants = [("unfused", unfused)]
        tc.autotune(("encodec_resblock", B, L, C, x.lo is not None, hs, x.f16, Wk3.planes, Wtail.planes), variants)

    def _encoder_tc(self, sig, vlen=None):
        B, T = sig.shape
        dev = sig.device
        raw = 32 <= RAW_MAX_CH
        pol = self.pol_enc
        x = pol.act(B, T, 32, dev, hl=2 if raw else 0)
        xe = None if raw else pol.act(B, T, 32, dev, hl=2)
        ops.conv_first_bf16(self._enc[0], sig, y=x, y_act=xe, act=ACT_ELU, vlen=vlen)
        L = T
        for i, ((Wk3, Wtail), Wdown, r) in enumerate(self._tenc):
            C = x.C
            Lout = -(-L // r)
            extra = Lout * r - L
            ye = pol.act(B, L, C, dev, hl=r, hr=extra)
            self._tc_resblock_run(Wk3, Wtail, x, xe, ye, pol)
            ye.fill_halo(PAD_REFLECT, max(r, extra) + 1 if L <= max(r, extra) else 0)
            last = i == len(self._tenc) - 1
            raw = not last and 2 * C <= RAW_MAX_CH
            x = pol.act(B, Lout, 2 * C, dev, hl=2 if raw else 0)
            xe = None if (last or raw) else pol.act(B, Lout, 2 * C, dev, hl=2)
            tc.conv_tc(Wdown, [Src(ye, taps=2, origin=-r, phases=r, rows=Lout + 1)], Lout, y=x, y_act=xe, act=ACT_ELU,
                       name="down_tc")
            L = Lout
        le = pol.act(B, L, x.C, dev, hl=6)
        self._tc_run_lstm(self._tenc_lstm, [n for _, n in self._enc_lstm], x, le, pol)
        le.fill_halo(PAD_REFLECT, 7 if L <= 6 else 0)
        emb = torch.empty((B, L, 128), device=dev, dtype=torch.float32)
        tc.conv_tc(self._tenc_last, [Src(le, taps=7, origin=-6, rows=L + 6)], L, y32=emb, name="conv_k7_tc")
        return emb

    def _decoder_tc(self, toks):
        B, N, K = toks.shape
        dev = toks.device
        pol = self.pol_dec
        z = pol.act(B, N, 128, dev, hl=6)
        ops.rvq_decode_bf16(toks.view(B * N, K), self.codebooks, K, z, err_flag=self._err)
        z.fill_halo(PAD_REFLECT, 7 if N <= 6 else 0)
        d0 = pol.act(B, N, 512, dev)
        tc.conv_tc(self._tdec_first, [Src(z, taps=7, origin=-6, rows=N + 6)], N, y=d0, name="conv_k7_tc")
        ye = pol.act(B, N, 512, dev)
        self._tc_run_lstm(self._tdec_lstm, [n for _, n in self._dec_lstm], d0, ye, pol)
        L = N
        for i, (Wtr, (Wk3, Wtail), r) in enumerate(self._tdec):
            C = ye.C // 2
            Lout = L * r
            raw = C <= RAW_MAX_CH
            x = pol.act(B, Lout, C, dev, hl=2 if raw else 0)
            xe = None if raw else pol.act(B, Lout, C, dev, hl=2)
            # transposed conv: 2-tap GEMM over n = (phase, cout); row -1 reads as zero (TMA OOB fill)
            tc.conv_tc(Wtr, [Src(ye, taps=2, shift=-1)], L, y=x, y_act=xe, act=ACT_ELU, act_mod=C, out_rows=Lout, out_ch=C,
                       name="convtr_tc")
            last = i == len(self._tdec) - 1
            ye = pol.act(B, Lout, C, dev, hl=6 if last else 0, hr=10 if last else 0)
            self._tc_resblock_run(Wk3, Wtail, x, xe, ye, pol)
            L = Lout
        # last layer (Cout = 1, k7, causal reflect padding in the 6 left halo rows) on the tap-GEMM kernel, 16 samples per
        # GEMM row; the 10 right halo rows only pad the buffer to whole 16-sample view rows (they meet zero weights)
        ye.fill_halo(PAD_REFLECT, 7 if L <= 6 else 0)
        return tc.conv_last_phased(self._tdec_last, ye)

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
        enc = self._encoder_tc if self.tensor_path else self._encoder
        emb = enc(sig, self._vlen(sig, length))
        B, N, D = emb.shape
        toks = torch.empty((B, N, nq), device=sig.device, dtype=torch.int64)
        if self.tensor_path:
            ops.rvq_encode_tc(emb.view(B * N, D), self.cb_split, self.codebooks, self.cb_norm, toks.view(B * N, nq), nq)
        else:
            ops.rvq_encode(emb.view(B * N, D), self.codebooks, self.cb_norm, toks.view(B * N, nq), nq)
        return toks  # [B, N, K]

    def _sig_to_feats(self, sig, length):  # R/audiocodecs/encodec.py:97-117 (normalize=False: mask unused)
        return (self._encoder_tc if self.tensor_path else self._encoder)(sig)

    def _sig_to_qfeats(self, sig, length):  # R/audiocodecs/encodec.py:120-127
        return self._toks_to_qfeats(self._sig_to_toks(sig, length), length)

    def _toks_to_qfeats(self, toks, length):  # R/audiocodecs/encodec.py:144-149
        B, N, K = toks.shape
        toks = toks.to(torch.int64).contiguous()
        q = ops.rvq_decode(toks.view(B * N, K), self.codebooks, K, err_flag=self._err)
        return q.view(B, N, -1)

    def _toks_to_sig(self, toks, length):  # R/audiocodecs/encodec.py:130-141
        if self.tensor_path:
            return self._decoder_tc(toks.to(torch.int64).contiguous())
        return self._decoder(self._toks_to_qfeats(toks, length))
