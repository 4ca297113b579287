"""Mimi behind the reference's `Mimi` wrapper interface (R/audiocodecs/mimi.py:27-155).

Same constructor arguments and tensor shapes; the `transformers.MimiModel` arithmetic
(HF/mimi/modeling_mimi.py) is replaced by the sm_100a kernels of this package.  The wrapper never enables
Mimi's streaming mode (R/audiocodecs/mimi.py:105-107): whole clip in one call.
"""
import math

import torch

from . import ops, packing, tc
from .codec import Codec
from .ops import ACT_ELU, ACT_NONE, EPI_GELU, PAD_REPLICATE, PAD_ZERO, ConvSpec
from .tc import Act, Src, TcWeights

__all__ = ["Mimi"]

RATIOS = (8, 6, 5, 4)  # MimiConfig().upsampling_ratios
SPLIT_MIN_CH = 128     # precision="bf16": activations with >= this many channels travel as (hi, lo) bf16 pairs (DESIGN.md,
                       # precision); precision="exact": every ENCODER activation does
HID, HEADS, HEAD_DIM, WINDOW, LAYERS = 512, 8, 64, 250, 8


class Mimi(Codec):
    """`Mimi(sample_rate, mode="reconstruct", num_codebooks=8, latent=True)`; extra keywords `state_dict`
    (`transformers.MimiModel` key format; default: fetch kyutai/mimi like the reference, mimi.py:45) and `precision`
    ("exact" / "bf16" / "fp32", see `Encodec`: "exact" is the tensor path whose encoder reproduces the reference's tokens)."""

    def _hop(self):
        return 1920

    # single-plane weights on the decoder side only (codes are untouched): decoder transformer GEMMs and the decoder's residual
    # blocks; decoder SI-SNR 51.8 -> 49.2 dB, step 61.6 -> 58.7 ms at 128 clips (scripts/weight_precision_probe.py)
    W_SINGLE = r"^decoder_transformer|^decoder\.layers.*block"

    def __init__(self, sample_rate, mode="reconstruct", num_codebooks=8, latent=True, state_dict=None, precision="exact",
                 w_single=None):
        super().__init__(sample_rate, 24000, mode)
        self.w_single = w_single
        if precision not in ("exact", "fp16", "fp32", "bf16"):
            raise ValueError("precision must be 'exact' (tcgen05 tensor path, split-precision encoder: reference tokens), 'fp16' "
                             "(one fp16 product per MAC: fastest), 'bf16' (error-compensated bf16 products) or 'fp32' (SIMT path)")
        self.tensor_path = precision != "fp32"
        self.exact = precision == "exact"
        if self.tensor_path:
            self.pol_enc, self.pol_dec = tc.policies(precision, SPLIT_MIN_CH)
        self.num_codebooks = num_codebooks
        self.vocab_size = 2048
        self.latent = latent
        self.precision = precision
        self._rope_cache = {}
        self.compute_dtype = {"exact": "f16", "fp16": "f16", "bf16": "bf16", "fp32": "f32"}[precision]
        if state_dict is None:
            try:
                from transformers import MimiModel
            except ImportError:
                raise ImportError("`pip install transformers>=4.45.1` to use this module")
            state_dict = MimiModel.from_pretrained("kyutai/mimi").state_dict()
        self._build(state_dict)

    # ------------------------------------------------------------------ packing
    def _conv(self, sd, prefix, stride=1, act=ACT_NONE, pad_mode=PAD_ZERO, scale=None, epi=ops.EPI_NONE):
        w = packing.fold_weight_norm(sd, prefix)
        if w.dim() == 2:
            w = w[:, :, None]
        if scale is not None:  # layer-scale folded into the producing projection (HF/mimi:499-511)
            w = w * scale.float().view(-1, 1, 1)
        cout, _, k = w.shape
        bias = sd.get(prefix + ".bias")
        spec = ConvSpec(packing.pack_conv(w), bias.float().clone() if bias is not None else None, cout=cout, kernel=k,
                        stride=stride, geometry="causal", pad_mode=pad_mode, act=act, epi=epi)
        self._specs.append(spec)
        return spec

    def _convtr(self, sd, prefix, stride, act):
        w = packing.fold_weight_norm(sd, prefix)  # [Cin, Cout, 2s]
        spec = ConvSpec(packing.pack_convtr(w, stride), sd[prefix + ".bias"].float().repeat(stride), cout=w.shape[1],
                        geometry="tr", tr_stride=stride, tr_pad=0, act=act)
        self._specs.append(spec)
        return spec

    def _transformer(self, sd, name):
        layers = []
        for l in range(LAYERS):
            p = f"{name}.layers.{l}."
            qkv_w = torch.cat([sd[p + f"self_attn.{w}.weight"].float() for w in ("q_proj", "k_proj", "v_proj")], dim=0)
            sd_l = {"qkv.weight": qkv_w}
            qkv = self._conv(sd_l, "qkv")
            o = self._conv(sd, p + "self_attn.o_proj", scale=sd[p + "self_attn_layer_scale.scale"])
            fc1 = self._conv(sd, p + "mlp.fc1", epi=EPI_GELU)
            fc2 = self._conv(sd, p + "mlp.fc2", scale=sd[p + "mlp_layer_scale.scale"])
            ln = []
            for i, nm in enumerate(("input_layernorm", "post_attention_layernorm")):
                for j, part in enumerate(("weight", "bias")):
                    key = f"{name.split('_')[0]}_l{l}_ln{i}_{part}"
                    self.register_buffer(key, sd[p + nm + "." + part].float().contiguous(), persistent=False)
                    ln.append(key)
            layers.append((ln, qkv, o, fc1, fc2))
        return layers

    def _build(self, sd):
        self._specs = []
        if self.mode != "decode":  # R/audiocodecs/mimi.py:46-51
            enc = [self._conv(sd, "encoder.layers.0.conv")]
            idx = 1
            for r in reversed(RATIOS):
                enc.append((self._conv(sd, f"encoder.layers.{idx}.block.1.conv", act=ACT_ELU),
                            self._conv(sd, f"encoder.layers.{idx}.block.3.conv", act=ACT_ELU)))
                enc.append(self._conv(sd, f"encoder.layers.{idx + 2}.conv", stride=r, act=ACT_ELU))
                idx += 3
            enc.append(self._conv(sd, f"encoder.layers.{idx + 1}.conv", act=ACT_ELU))
            self._enc = enc
            self._enc_tr = self._transformer(sd, "encoder_transformer")
            self._down = self._conv(sd, "downsample.conv", stride=2, pad_mode=PAD_REPLICATE)
        if self.mode != "encode":
            self.register_buffer("up_w", sd["upsample.conv.weight"].float()[:, 0, :].contiguous(), persistent=False)  # [C,4]
            self._dec_tr = self._transformer(sd, "decoder_transformer")
            dec = [self._conv(sd, "decoder.layers.0.conv")]
            idx = 2
            for r in RATIOS:
                dec.append(self._convtr(sd, f"decoder.layers.{idx}.conv", r, ACT_ELU))
                dec.append((self._conv(sd, f"decoder.layers.{idx + 1}.block.1.conv", act=ACT_ELU),
                            self._conv(sd, f"decoder.layers.{idx + 1}.block.3.conv", act=ACT_ELU)))
                idx += 3
            dec.append(self._conv(sd, f"decoder.layers.{idx}.conv", act=ACT_ELU))
            self._dec = dec
        self._tcw = []
        if self.tensor_path:
            self._build_tc(sd)
        # quantizer: E = embed_sum / clamp(cluster_usage, 1e-5) (HF/mimi:1188-1195); semantic (1) then acoustic (31) codebooks
        cbs = []
        for which, n in (("semantic", 1), ("acoustic", 31)):
            q = f"quantizer.{which}_residual_vector_quantizer."
            for k in range(n):
                cbs.append(sd[q + f"layers.{k}.codebook.embed_sum"].float()
                           / sd[q + f"layers.{k}.codebook.cluster_usage"].float().clamp(min=1e-5)[:, None])
            setattr(self, f"_{which}_in", self._conv(sd, q + "input_proj"))
            setattr(self, f"_{which}_out", self._conv(sd, q + "output_proj"))
        cb = torch.stack(cbs).contiguous()
        self.register_buffer("codebooks", cb, persistent=False)                    # [32, 2048, 256]
        self.register_buffer("cb_norm", cb.pow(2).sum(-1).contiguous(), persistent=False)
        if self.tensor_path:  # operand planes of the tensor-core distance GEMM: bf16(E), bf16(E - bf16(E))
            hi = cb.to(torch.bfloat16)
            self.register_buffer("cb_split", torch.stack([hi, (cb - hi.float()).to(torch.bfloat16)]).contiguous(), persistent=False)
        inv_freq = 1.0 / (10000.0 ** (torch.arange(0, HEAD_DIM, 2, dtype=torch.int64).float() / HEAD_DIM))  # HF/mimi:560-562
        self.register_buffer("inv_freq", inv_freq, persistent=False)
        self.register_buffer("_err", torch.zeros(1, dtype=torch.int32), persistent=False)

    def _packed(self):
        return self._specs + self._tcw

    # ------------------------------------------------------------------ bf16 tensor path: packing
    def _pol(self, name):
        """encoder side = everything that shapes the tokens (SEANet encoder, encoder transformer, downsample, quantizer input
        projections); the rest is the decoder side"""
        enc = name.startswith(("encoder", "downsample")) or "input_proj" in name
        return self.pol_enc if enc else self.pol_dec

    def _tw(self, w, bias=None, name=""):
        W = self._pol(name).weights(w, bias, self._w_split(name))
        self._tcw.append(W)
        return W

    def _tw_conv(self, sd, prefix):
        w = packing.fold_weight_norm(sd, prefix)  # [Cout, Cin, K] -> [Cout][K*Cin], column = tap*Cin + c
        return self._tw(w.permute(0, 2, 1).reshape(w.shape[0], -1), sd[prefix + ".bias"], prefix)

    def _tw_convtr(self, sd, prefix, stride):
        pk = packing.pack_convtr(packing.fold_weight_norm(sd, prefix), stride)  # [2, Cin, s*Cout]
        return self._tw(pk.permute(2, 0, 1).reshape(pk.shape[2], -1), sd[prefix + ".bias"].float().repeat(stride), prefix)

    def _tw_transformer(self, sd, name):
        out = []
        for l in range(LAYERS):
            p = f"{name}.layers.{l}."
            qkv = torch.cat([sd[p + f"self_attn.{w}.weight"].float() for w in ("q_proj", "k_proj", "v_proj")], dim=0)
            o = sd[p + "self_attn.o_proj.weight"].float() * sd[p + "self_attn_layer_scale.scale"].float().view(-1, 1)
            fc2 = sd[p + "mlp.fc2.weight"].float() * sd[p + "mlp_layer_scale.scale"].float().view(-1, 1)
            out.append((self._tw(qkv, None, p + "qkv"), self._tw(o, None, p + "o"), self._tw(sd[p + "mlp.fc1.weight"].float(), None, p + "fc1"),
                        self._tw(fc2, None, p + "fc2")))
        return out

    def _build_tc(self, sd):
        q = "quantizer.{}_residual_vector_quantizer.{}"
        proj = lambda which, name: packing.fold_weight_norm(sd, q.format(which, name))[:, :, 0]   # 1x1 conv, no bias: [out, in]
        if self.mode != "decode":
            idx, self._tenc = 1, []
            for r in reversed(RATIOS):
                self._tenc.append((self._tw_conv(sd, f"encoder.layers.{idx}.block.1.conv"), self._tw_conv(sd, f"encoder.layers.{idx}.block.3.conv"),
                                   self._tw_conv(sd, f"encoder.layers.{idx + 2}.conv"), r))
                idx += 3
            self._tenc_last = self._tw_conv(sd, f"encoder.layers.{idx + 1}.conv")
            self._tenc_tr = self._tw_transformer(sd, "encoder_transformer")
            self._tsem_in = self._tw(proj("semantic", "input_proj"), None, "quantizer.semantic.input_proj")
            self._tac_in = self._tw(proj("acoustic", "input_proj"), None, "quantizer.acoustic.input_proj")
            w = packing.fold_weight_norm(sd, "downsample.conv")  # [512, 512, 4], no bias
            self._tdown = self._tw(w.permute(0, 2, 1).reshape(w.shape[0], -1), sd.get("downsample.conv.bias"), "downsample.conv")
        if self.mode != "encode":
            self._tdec_tr = self._tw_transformer(sd, "decoder_transformer")
            # output_proj of the semantic and the acoustic sum as ONE GEMM over the two gathered sources (K = 256 + 256)
            self._tsem_out = self._tw(proj("semantic", "output_proj"), None, "quantizer.semantic.output_proj")
            self._tq_out = self._tw(torch.cat([proj("semantic", "output_proj"), proj("acoustic", "output_proj")], dim=1), None,
                                    "quantizer.output_proj")
            self._tdec_first = self._tw_conv(sd, "decoder.layers.0.conv")
            idx, self._tdec = 2, []
            for r in RATIOS:
                self._tdec.append((self._tw_convtr(sd, f"decoder.layers.{idx}.conv", r), self._tw_conv(sd, f"decoder.layers.{idx + 1}.block.1.conv"),
                                   self._tw_conv(sd, f"decoder.layers.{idx + 1}.block.3.conv"), r))
                idx += 3
            pd = self.pol_dec  # Cout = 1 as a stride-16 conv with 16 outputs
            self._tdec_last = tc.last_conv_weights_phased(self._dec[-1], split=True if pd.w_split is None else pd.w_split, f16=pd.f16)
            self._tcw.append(self._tdec_last)

    # ------------------------------------------------------------------ bf16 tensor path: execution
    def _tc_resblock(self, Wk3, Wk1, x, xe, ye, pol):
        """MimiResnetBlock (HF/mimi:412-451), identity shortcut, causal zero padding (TMA out-of-bounds fill):
        x raw, xe = ELU(x) -> ye = ELU(x + conv_k1(ELU(conv_k3(xe))))."""
        B, L, C = x.B, x.L, x.C
        hs = pol.split(C // 2)
        a = Src(xe, taps=3, shift=-2)

        def unfused():
            he = pol.act(B, L, C // 2, x.buf.device, split=hs)
            tc.conv_tc(Wk3, [a], L, y_act=he, act=ACT_ELU, name="res_k3_tc")
            tc.conv_tc(Wk1, [Src(he)], L, res=x, y_act=ye, act=ACT_ELU, name="res_k1_tc")

        def fused(g, dbl, io):
            return lambda: tc.resunit_tc(Wk3, Wk1, a, L, res=x, y_act=ye, act1=ACT_ELU, act2=ACT_ELU, h_split=hs, g_hint=g, dbl_hint=dbl,
                                         io_stage=io, name="resblock_tc")

        # fused (hidden activation on chip) whenever the accumulators fit tensor memory (measured faster at 64-256 channels),
        # else two launches.  A rule, not a timing: the two forms group the fp32 accumulation differently, and a clip's tokens
        # must not depend on the batch it is tuned in; the tuner only picks the tile grouping / buffering (bit-identical)
        variants = [("unfused", unfused)]
        if C <= 256:
            variants = [(f"fused_g{g}_d{dbl}_io{io}", fused(g, dbl, io)) for g in (4, 2, 1) for dbl in (2, 1, 0) for io in (-1, 1) if not (dbl == 2 and io == 1)]
        tc.autotune(("mimi_resblock", B, L, C, x.lo is not None, hs, x.f16, Wk3.planes, Wk1.planes), variants)

    def _tc_transformer(self, layers, tws, h, pol):
        """fp32 residual stream h [B,T,512]; the four projections of every layer run on tcgen05 (split-bf16 operands,
        fp32 accumulate, residual added in fp32 in the epilogue) and so does the attention (ops.attention_tc); LayerNorm is
        fp32 SIMT."""
        B, T, C = h.shape
        dev = h.device
        xa = pol.act(B, T, C, dev)    # legacy bf16 / exact: (hi, lo) pairs; fp16: one plane
        fa = pol.act(B, T, 4 * C, dev)
        qkv = torch.empty((B, T, 3 * C), device=dev, dtype=torch.float32)
        rope = self._rope_table(T, dev)
        for (ln, *_), (Wqkv, Wo, Wfc1, Wfc2) in zip(layers, tws):
            ops.layernorm_act(h, getattr(self, ln[0]), getattr(self, ln[1]), xa)
            tc.conv_tc(Wqkv, [Src(xa)], T, y32=qkv, name="tr_qkv_tc")
            ops.attention_tc(qkv, rope, HEADS, HEAD_DIM, WINDOW, out_act=xa)
            tc.conv_tc(Wo, [Src(xa)], T, res32=h, y32=h, name="tr_o_tc")
            ops.layernorm_act(h, getattr(self, ln[2]), getattr(self, ln[3]), xa)
            tc.conv_tc(Wfc1, [Src(xa)], T, y=fa, epi=EPI_GELU, name="tr_fc1_tc")
            tc.conv_tc(Wfc2, [Src(fa)], T, res32=h, y32=h, name="tr_fc2_tc")
        return h

    def _rope_table(self, T, dev):
        key = (T, str(dev))
        if key not in self._rope_cache:
            self._rope_cache = {key: ops.rope_table(self.inv_freq, T)}  # one entry: the table of the last sequence length
        return self._rope_cache[key]

    def _encoder_tc(self, sig):
        B, T = sig.shape
        dev = sig.device
        C = self._enc[0].cout
        pol = self.pol_enc
        x = pol.act(B, T, C, dev)
        xe = pol.act(B, T, C, dev)
        ops.conv_first_bf16(self._enc[0], sig, y=x, y_act=xe, act=ACT_ELU)
        L = T
        for i, (Wk3, Wk1, Wdown, r) in enumerate(self._tenc):
            Lout = -(-L // r)
            ye = pol.act(B, L, C, dev, hr=Lout * r - L)
            self._tc_resblock(Wk3, Wk1, x, xe, ye, pol)
            ye.fill_halo(PAD_ZERO)
            C = 2 * C
            last = i == len(self._tenc) - 1
            x = None if last else pol.act(B, Lout, C, dev)
            xe = pol.act(B, Lout, C, dev)
            # kernel 2r / stride r causal conv: 2 taps over the r-phase view, tap 0 = the previous view row (zero for row 0)
            tc.conv_tc(Wdown, [Src(ye, taps=2, shift=-1, phases=r, rows=Lout)], Lout, y=x, y_act=xe, act=ACT_ELU, name="down_tc")
            L = Lout
        h = torch.empty((B, L, self._tenc_last.n_total), device=dev, dtype=torch.float32)
        tc.conv_tc(self._tenc_last, [Src(xe, taps=3, shift=-2)], L, y32=h, name="conv_k3_tc")
        return h

    def _decoder_tc(self, z):
        B, N, C = z.shape
        dev = z.device
        pol = self.pol_dec
        za = pol.act(B, N, C, dev)
        ops.f32_to_act(z.contiguous(), za)
        C = self._tdec_first.n_total
        ye = pol.act(B, N, C, dev)
        tc.conv_tc(self._tdec_first, [Src(za, taps=7, shift=-6)], N, y_act=ye, act=ACT_ELU, name="conv_k7_tc")
        L = N
        for i, (Wtr, Wk3, Wk1, r) in enumerate(self._tdec):
            C = C // 2
            Lout = L * r
            x = pol.act(B, Lout, C, dev)
            xe = pol.act(B, Lout, C, dev)
            tc.conv_tc(Wtr, [Src(ye, taps=2, shift=-1)], L, y=x, y_act=xe, act=ACT_ELU, act_mod=C, out_rows=Lout, out_ch=C, name="convtr_tc")
            last = i == len(self._tdec) - 1   # the last layer reads 16-sample view rows: causal zero halo + pad to whole rows
            pl = self._dec[-1].taps - 1
            ye = pol.act(B, Lout, C, dev, hl=pl if last else 0, hr=(-pl) % 16 if last else 0)
            self._tc_resblock(Wk3, Wk1, x, xe, ye, pol)
            L = Lout
        ye.fill_halo(PAD_ZERO)
        return tc.conv_last_phased(self._tdec_last, ye)

    # ------------------------------------------------------------------ pieces
    def _seanet(self, layers, x):
        for layer in layers:
            if isinstance(layer, tuple):  # MimiResnetBlock, identity shortcut (HF/mimi:412-451)
                h = ops.conv(layer[0], x)
                x = ops.conv(layer[1], h, res=x)
            else:
                x = ops.conv(layer, x)
        return x

    def _run_transformer(self, layers, h):
        """h [B,T,512] (HF/mimi:926-993): h += ls1*o_proj(attn(LN(h))); h += ls2*fc2(gelu(fc1(LN(h))))."""
        for ln, qkv, o, fc1, fc2 in layers:
            x = ops.layernorm(h, getattr(self, ln[0]), getattr(self, ln[1]))
            a = ops.attention(ops.conv(qkv, x), self.inv_freq, HEADS, HEAD_DIM, WINDOW)
            h = ops.conv(o, a, res=h)
            x = ops.layernorm(h, getattr(self, ln[2]), getattr(self, ln[3]))
            h = ops.conv(fc2, ops.conv(fc1, x), res=h)
        return h

    def _embeddings(self, sig, want_act=False):
        """sig [B,T] -> [B,N,512] at 12.5 Hz (HF/mimi:1455-1488)."""
        if self.tensor_path:
            pol = self.pol_enc
            h = self._tc_transformer(self._enc_tr, self._tenc_tr, self._encoder_tc(sig.contiguous()), pol)
            # `downsample` (HF/mimi:1419-1431): k4 s2 causal conv, replicate padding 2 left (+1 right for an odd length), as a
            # 2-tap GEMM over the 2-phase view of the padded split-bf16 copy; fp32 out
            B, L, C = h.shape
            N = -(-L // 2)
            ha = pol.act(B, L, C, h.device, hl=2, hr=2 * N - L)
            ops.f32_to_act(h, ha)
            ha.fill_halo(PAD_REPLICATE)
            emb = torch.empty((B, N, C), device=h.device, dtype=torch.float32)
            ea = pol.act(B, N, C, h.device) if want_act else None   # 16-bit copy for the quantizer's input_proj GEMMs
            tc.conv_tc(self._tdown, [Src(ha, taps=2, origin=-2, phases=2, rows=N + 1)], N, y32=emb, y=ea, name="downsample_tc")
            return (emb, ea) if want_act else emb
        else:
            x = self._seanet(self._enc, sig.contiguous()[:, :, None])
            x = self._run_transformer(self._enc_tr, x)
        return ops.conv(self._down, x)

    def _check_k(self, K):
        if K > 32:
            raise ValueError(f"The number of quantizers (i.e codebooks) asked should be lower than the total number of quantizers 32, but is currently {K}.")
        if K < 1:
            raise ValueError(f"The number of quantizers (i.e codebooks) asked should be higher than the number of semantic quantizers 1, but is currently {K}.")

    # ------------------------------------------------------------------ Codec hooks
    @torch.no_grad()
    def embs(self):  # R/audiocodecs/mimi.py:55-90
        e = self.codebooks[: self.num_codebooks].clone()
        if self.latent:
            return e
        K = self.num_codebooks
        sem = ops.conv(self._semantic_out, e[:1].contiguous())
        if K == 1:
            return sem
        return torch.cat([sem, ops.conv(self._acoustic_out, e[1:].contiguous())])

    def _sig_to_toks(self, sig, length):
        K = self.num_codebooks
        self._check_k(K)
        if self.tensor_path:
            emb, ea = self._embeddings(sig, want_act=True)
            B, N, _ = emb.shape
            toks = torch.empty((B, N, K), device=sig.device, dtype=torch.int64)
            xs = torch.empty((B, N, self._tsem_in.n_total), device=sig.device, dtype=torch.float32)
            tc.conv_tc(self._tsem_in, [Src(ea)], N, y32=xs, name="rvq_in_proj_tc")
            if K > 1:
                xa = torch.empty_like(xs)
                tc.conv_tc(self._tac_in, [Src(ea)], N, y32=xa, name="rvq_in_proj_tc")
        else:
            emb = self._embeddings(sig)
            B, N, _ = emb.shape
            toks = torch.empty((B, N, K), device=sig.device, dtype=torch.int64)
            xs = ops.conv(self._semantic_in, emb)
            xa = ops.conv(self._acoustic_in, emb) if K > 1 else None
        if self.tensor_path:  # tcgen05 distance GEMM + exact fp32 re-score (same decisions as the fp32 kernel)
            ops.rvq_encode_tc(xs.view(B * N, -1), self.cb_split, self.codebooks, self.cb_norm, toks.view(B * N, K), 1, code_offset=0,
                              stage0=0, metric=1)
            if K > 1:
                ops.rvq_encode_tc(xa.view(B * N, -1), self.cb_split, self.codebooks, self.cb_norm, toks.view(B * N, K), K - 1,
                                  code_offset=1, stage0=1, metric=1)
            return toks
        ops.rvq_encode(xs.view(B * N, -1), self.codebooks[:1], self.cb_norm[:1], toks.view(B * N, K), 1, code_offset=0, metric=1)
        if K > 1:
            ops.rvq_encode(xa.view(B * N, -1), self.codebooks[1:], self.cb_norm[1:], toks.view(B * N, K), K - 1, code_offset=1, metric=1)
        return toks

    def _sig_to_feats(self, sig, length):  # R/audiocodecs/mimi.py:112-121
        return self._embeddings(sig)

    def _sig_to_qfeats(self, sig, length):  # R/audiocodecs/mimi.py:124-141
        return self._toks_to_qfeats(self._sig_to_toks(sig, length), length)

    def _toks_to_qfeats(self, toks, length):  # R/audiocodecs/mimi.py:151-155 ; HF/mimi:1340-1349
        B, N, K = toks.shape
        toks = toks.to(torch.int64).contiguous()
        sem = ops.rvq_decode(toks.view(B * N, K), self.codebooks[:1], 1, code_offset=0, err_flag=self._err).view(B, N, -1)
        out = ops.conv(self._semantic_out, sem)
        if K > 1:
            ac = ops.rvq_decode(toks.view(B * N, K), self.codebooks[1:], K - 1, code_offset=1, err_flag=self._err).view(B, N, -1)
            out = ops.conv(self._acoustic_out, ac, res=out, out=out)
        return out

    def _toks_to_sig(self, toks, length):  # R/audiocodecs/mimi.py:144-148 ; HF/mimi:1613-1631
        if self.tensor_path:  # gather-sums straight into split-bf16 operands, both output_proj as one GEMM
            B, N, K = toks.shape
            self._check_k(K)
            toks = toks.to(torch.int64).contiguous()
            dev = toks.device
            D = self.codebooks.shape[2]
            sa = self.pol_dec.act(B, N, D, dev)
            ops.rvq_decode_bf16(toks.view(B * N, K), self.codebooks[:1], 1, sa, code_offset=0, err_flag=self._err)
            z = torch.empty((B, N, self._tq_out.n_total), device=dev, dtype=torch.float32)
            if K > 1:
                aa = self.pol_dec.act(B, N, D, dev)
                ops.rvq_decode_bf16(toks.view(B * N, K), self.codebooks[1:], K - 1, aa, code_offset=1, err_flag=self._err)
                tc.conv_tc(self._tq_out, [Src(sa), Src(aa)], N, y32=z, name="rvq_out_proj_tc")
            else:
                tc.conv_tc(self._tsem_out, [Src(sa)], N, y32=z, name="rvq_out_proj_tc")
        else:
            z = self._toks_to_qfeats(toks, length)
        z = ops.upsample_dw(z, self.up_w)
        if self.tensor_path:
            return self._decoder_tc(self._tc_transformer(self._dec_tr, self._tdec_tr, z, self.pol_dec))
        z = self._run_transformer(self._dec_tr, z)
        return self._seanet(self._dec, z)[:, :, 0]
