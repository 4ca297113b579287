"""Host logic behind the tensor path, checked in plain torch on the CPU (no kernel is launched): the operand packings that
turn transposed / strided convolutions into 2-tap GEMMs over phase views, the (hi, lo) bf16 weight
pair, the per-layer weight policy, and the algebra of the projected DAC RVQ encode."""
import math

import pytest
import torch
import torch.nn.functional as F

from audiocodecs_b200 import packing
from audiocodecs_b200.tc import TcWeights


@pytest.mark.parametrize("cin,cout,s,L,pad", [(8, 4, 2, 11, 0), (6, 3, 5, 7, 0), (4, 6, 8, 5, 4), (4, 2, 4, 9, 2)])
def test_transposed_conv_as_two_tap_gemm(cin, cout, s, L, pad):
    """ConvTranspose1d(k = 2s, stride s) == rows [x[q-1] | x[q]] times the packed [2*Cin, s*Cout] operand, the [L][s*Cout]
    result read as [L*s][Cout]; pad = 0 with the causal right trim (HF/encodec:179-233) or torch padding `pad` (HF/dac:234-262:
    one more GEMM row, flat output shifted by pad*Cout)."""
    g = torch.Generator().manual_seed(cin * 100 + s)
    w = torch.randn(cin, cout, 2 * s, generator=g)
    x = torch.randn(2, L, cin, generator=g)
    pk = packing.pack_convtr(w, s)                                   # [2, Cin, s*Cout]
    W = pk.permute(2, 0, 1).reshape(s * cout, 2 * cin)                  # n = phase*Cout + c, k = tap*Cin + cin (as _tc_convtr)
    rows = L + (1 if pad else 0)
    xp = F.pad(x, (0, 0, 1, 1))                                      # row -1 and row L read as zero (TMA out-of-bounds fill)
    a = torch.cat([xp[:, :rows], xp[:, 1:rows + 1]], dim=-1)            # taps: previous row, current row
    flat = (a @ W.t()).reshape(2, rows * s, cout)
    ref = F.conv_transpose1d(x.permute(0, 2, 1), w, stride=s, padding=pad).permute(0, 2, 1)
    if pad == 0:
        ref = ref[:, : L * s]                                           # causal: drop k - s samples on the right
        got = flat
    else:
        got = flat[:, pad: pad + ref.shape[1]]
    assert got.shape == ref.shape and (got - ref).abs().max().item() < 1e-4


@pytest.mark.parametrize("cin,cout,s,L", [(4, 8, 2, 13), (3, 5, 4, 16), (2, 6, 5, 23)])
def test_strided_conv_as_two_tap_gemm_over_phase_view(cin, cout, s, L):
    """Causal Conv1d(k = 2s, stride s) == a 2-tap GEMM over the view [L/s][s*Cin] of the zero-padded input (tap 0 = previous
    view row), with the weight flattened as column = tap*Cin + c (encodec._tc_conv / mimi._tw_conv)."""
    g = torch.Generator().manual_seed(cin * 10 + s)
    w = torch.randn(cout, cin, 2 * s, generator=g)
    x = torch.randn(2, L, cin, generator=g)
    Lout = -(-L // s)
    ref = F.conv1d(F.pad(x.permute(0, 2, 1), (s, Lout * s - L)), w, stride=s).permute(0, 2, 1)   # left pad k - s, right pad to ceil
    W = w.permute(0, 2, 1).reshape(cout, -1)
    view = F.pad(x, (0, 0, s, Lout * s - L)).reshape(2, Lout + 1, s * cin)                      # one padding view row in front
    a = torch.cat([view[:, :-1], view[:, 1:]], dim=-1)
    got = a @ W.t()
    assert got.shape == ref.shape and (got - ref).abs().max().item() < 1e-4


def test_split_weights_carry_sixteen_bits():
    w = torch.randn(64, 96, generator=torch.Generator().manual_seed(0))
    pair, single = TcWeights(w, None), TcWeights(w, None, split=False)
    assert pair.w.shape == (2, 64, 96) and single.w.shape == (64, 96) and pair.w.dtype == torch.bfloat16
    rel = lambda t: ((t - w).abs().max() / w.abs().max()).item()
    assert rel(pair.w[0].float() + pair.w[1].float()) < 2 ** -15 and 2 ** -10 < rel(single.w.float()) < 2 ** -7


def test_weight_policy_regex(encodec_sd):
    import audiocodecs_b200 as A
    c = A.Encodec(24000, 24000, state_dict=encodec_sd)
    assert c.W_SINGLE is None and c._w_split("encoder.layers.3") and c._w_split("decoder.layers.3")   # EnCodec: pair everywhere
    c.w_single = r"^decoder\.layers\.\d+$"
    assert not c._w_split("decoder.layers.3") and c._w_split("decoder.layers.4.block.1") and c._w_split("encoder.layers.3")
    c.w_single = ".*"
    assert not c._w_split("encoder.layers.0")
    d = A.DAC.W_SINGLE
    import re
    assert re.search(d, "decoder.block.1.res_unit2.conv1") and not re.search(d, "decoder.block.1.res_unit2.conv2")
    assert not re.search(d, "decoder.block.1.conv_t1") and re.search(d, "encoder.block.0.res_unit1.conv1")
    m = A.Mimi.W_SINGLE
    assert re.search(m, "decoder.layers.3.block.1.conv") and re.search(m, "decoder_transformer.layers.0.qkv")
    assert not re.search(m, "encoder.layers.1.block.1.conv") and not re.search(m, "decoder.layers.2.conv")


def test_projected_dac_rvq_algebra(dac_sd):
    """z_e[k] = P[k] + c[k] - sum_{j<k} (W_in,k W_out,j) zq[j] reproduces the oracle's 1024-wide residual chain: same codes
    wherever no near-tie (relative gap < 1e-4) occurred at or before the stage.  Emulates ac_dac_rvq_encode_proj_f32 in fp64."""
    import audiocodecs_b200 as A
    from oracle import dac_ref
    codec = A.DAC(44100, 44100, num_codebooks=9, state_dict=dac_sd, precision="bf16")
    K, B, N = 9, 2, 40
    z = torch.randn(B, 1024, N, generator=torch.Generator().manual_seed(1)) * 2.0
    with torch.no_grad():
        ref, gaps, _ = dac_ref.rvq_encode(dac_sd, z, K, return_gaps=True)          # [B, K, N]
    S = codec.w_in.shape[0]
    P = z.double().permute(0, 2, 1).reshape(B * N, 1024) @ codec.w_in.double().reshape(S * 8, 1024).t() + codec.b_in.double().reshape(-1)
    cross, cconst, cb = codec.rvq_cross.double(), codec.rvq_cconst.double(), codec.codebooks.double()
    cbn = codec.cb_normed.double()
    zq, codes = [], []
    for k in range(K):
        ze = P[:, k * 8:(k + 1) * 8] + cconst[k]
        for j in range(k):
            ze = ze - zq[j] @ cross[k, j].t()
        a = F.normalize(ze, dim=-1)
        score = -((a * a).sum(-1, keepdim=True) - 2 * a @ cbn[k].t()) + codec.cb_norm2[k].double()
        idx = score.argmax(-1)
        codes.append(idx)
        zq.append(ze + (cb[k][idx] - ze))
    got = torch.stack(codes, dim=1).view(B, N, K)
    clear = (gaps.permute(0, 2, 1) > 1e-4).long().cumprod(dim=-1).bool()
    assert (got == ref.permute(0, 2, 1))[clear].all()
    assert clear.float().mean().item() > 0.9


def test_sub_batch_chunking(encodec_sd):
    """Codec._chunks: a batch larger than max_chunk_samples is cut into contiguous sub-batches that cover every clip once."""
    import audiocodecs_b200 as A
    c = A.Encodec(24000, 24000, state_dict=encodec_sd)
    c.max_chunk_samples = 10 * 240000
    assert c._chunks(7, 240000) == [(0, 7)]
    spans = c._chunks(25, 240000)
    assert spans == [(0, 10), (10, 20), (20, 25)]
    assert c._chunks(3, 10 ** 9) == [(0, 1), (1, 2), (2, 3)]          # a clip longer than the budget still runs, alone
    calls = []
    out = c._chunked(lambda x, l: calls.append((x.shape[0], None if l is None else tuple(l.tolist()))) or x * 2,
                     torch.arange(25.0)[:, None], torch.linspace(0.1, 1.0, 25), 240000)
    assert torch.equal(out, torch.arange(25.0)[:, None] * 2) and [n for n, _ in calls] == [10, 10, 5]
    assert calls[2][1] == tuple(torch.linspace(0.1, 1.0, 25)[20:].tolist())


def test_autotune_skips_and_reports(monkeypatch):
    """tc.autotune: variants whose tiling does not fit (ConfigError) are skipped silently, launch errors are skipped but
    recorded, the fastest remaining variant is cached, and the error names the failures when nothing works.  CUDA events are
    replaced by a fake clock so the selection logic runs on the CPU."""
    from audiocodecs_b200 import _lib, tc

    clock = {"t": 0.0}

    class FakeEvent:
        def __init__(self, enable_timing=True):
            self.t = None

        def record(self):
            self.t = clock["t"]

        def synchronize(self):
            pass

        def elapsed_time(self, other):
            return other.t - self.t

    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(tc, "_TUNED", {})
    monkeypatch.setattr(tc, "TUNE_ERRORS", [])
    calls = []

    def variant(name, cost=None, exc=None):
        def fn():
            calls.append(name)
            if exc is not None:
                raise exc
            clock["t"] += cost
            return name
        return name, fn

    vs = [variant("nofit", exc=_lib.ConfigError("no tiling fits")), variant("slow", 5.0), variant("broken", exc=RuntimeError("launch failed")),
          variant("fast", 1.0)]
    assert tc.autotune(("k", 1), vs) == "fast" and tc._TUNED[("k", 1)] == "fast"
    assert [e[1] for e in tc.TUNE_ERRORS] == ["broken"]
    calls.clear()
    assert tc.autotune(("k", 1), vs) == "fast" and calls == ["fast"]            # cached: straight to the winner
    # a near-tie (within 3 %) goes to the variant listed first; a clear win does not
    assert tc.autotune(("k", 4), [variant("fused", 1.02), variant("unfused", 1.0)]) == "fused"
    assert tc.autotune(("k", 5), [variant("fused", 1.10), variant("unfused", 1.0)]) == "unfused"
    monkeypatch.setattr(tc, "TUNE", False)
    assert tc.autotune(("k", 2), vs) == "slow"                                   # no tuning: the first variant that launches
    with pytest.raises(RuntimeError, match="no variant"):
        tc.autotune(("k", 3), vs[:1] + vs[2:3])


def _to_descript(sd):
    """HF DacModel names -> descript-audio-codec 1.0.0 Sequential names with old-style weight_g / weight_v (test double of a
    `dac.DAC.load(...).state_dict()`; the inverse of audiocodecs_b200.dac.descript_to_hf_keys)."""
    import re
    unit = {"snake1": "0", "conv1": "1", "snake2": "2", "conv2": "3"}
    out = {}
    for k, v in sd.items():
        m = re.match(r"(encoder|decoder)\.block\.(\d)\.res_unit(\d)\.(\w+)\.(.+)", k)
        if m:
            side, i, u, part, rest = m.groups()
            j = int(u) - 1 if side == "encoder" else int(u) + 1
            top = f"encoder.block.{int(i) + 1}" if side == "encoder" else f"decoder.model.{int(i) + 1}"
            name = f"{top}.block.{j}.block.{unit[part]}.{rest}"
        else:
            m = re.match(r"(encoder|decoder)\.block\.(\d)\.(snake1|conv1|conv_t1)\.(.+)", k)
            if m:
                side, i, part, rest = m.groups()
                j = {"encoder": {"snake1": 3, "conv1": 4}, "decoder": {"snake1": 0, "conv_t1": 1}}[side][part]
                top = f"encoder.block.{int(i) + 1}" if side == "encoder" else f"decoder.model.{int(i) + 1}"
                name = f"{top}.block.{j}.{rest}"
            else:
                name = k
                for a, b in (("encoder.conv1.", "encoder.block.0."), ("encoder.snake1.", "encoder.block.5."), ("encoder.conv2.", "encoder.block.6."),
                             ("decoder.conv1.", "decoder.model.0."), ("decoder.snake1.", "decoder.model.5."), ("decoder.conv2.", "decoder.model.6.")):
                    if k.startswith(a):
                        name = b + k[len(a):]
        if name.endswith(".weight") and "codebook" not in name:   # weight-normed in descript checkpoints: w = g v / |v|
            g = v.flatten(1).norm(dim=1).view(-1, *([1] * (v.dim() - 1)))
            out[name + "_g"], out[name + "_v"] = g * 1.0, v * 3.0   # v deliberately not normalised
        else:
            out[name] = v
    return out


def test_dac_default_ctor_from_descript_checkpoint():
    """`DAC(sample_rate)` -- orig_sample_rate=16000, the reference's default and its only downstream config --
    builds from a descript-format state dict (Sequential key names, weight_g / weight_v, odd stride 5) and packs exactly what
    the HF-format dict of the same weights packs."""
    import audiocodecs_b200 as A
    from audiocodecs_b200.dac import descript_to_hf_keys
    from oracle import weights
    sd = weights.dac_state_dict(0, tag="16khz")
    dsd = _to_descript(sd)
    assert "encoder.block.0.weight_g" in dsd and "decoder.model.1.block.1.weight_v" in dsd and "encoder.block.1.block.0.block.1.weight_g" in dsd
    back = descript_to_hf_keys(dsd)
    assert {k.replace(".weight_g", ".weight").replace(".weight_v", ".weight") for k in back} == set(sd)
    a = A.DAC(16000, state_dict=dsd, precision="fp32")   # default orig_sample_rate=16000, num_codebooks=8
    b = A.DAC(16000, 16000, num_codebooks=8, state_dict=sd, precision="fp32")
    assert a.orig_sample_rate == 16000 and a.num_codebooks == 8 and a._dec_rates == (8, 5, 4, 2) and a._hop() == 320
    for x, y in zip(a._specs, b._specs):
        assert torch.allclose(x.w, y.w, atol=1e-6) and x.tr_stride == y.tr_stride
    assert a._dec[1].out_len(10) == 80 and a._dec[5].tr_stride == 5 and a._dec[5].out_len(80) == 5 * 80 - 1   # stride 8: 8 L; stride 5: 5 L - 1
    t = A.DAC(24000, 24000, num_codebooks=32, state_dict=weights.dac_state_dict(0, tag="24khz"))  # tensor path packs too
    assert t.precision == "exact" and t._tdec[1][2] == 5
