"""CPU: the oracle restatement reproduces the live reference's recorded outputs (tests/golden/, written by
oracle/make_golden.py from the unmodified /root/reference wrappers)."""
import pytest
import torch

from helpers import make_input
from oracle import encodec_ref, resample_ref


@pytest.mark.parametrize("case", range(5))
def test_encodec_oracle_matches_reference_golden(encodec_sd, encodec_golden, case):
    c = encodec_golden["cases"][case]
    sig = make_input(c["seed"], c["B"], c["T"])
    length = None if c["length"] is None else torch.tensor(c["length"])
    with torch.no_grad():
        toks, gaps, _ = encodec_ref.sig_to_toks(encodec_sd, sig, c["K"], c["sample_rate"], 24000, length, return_gaps=True)
        rec = encodec_ref.toks_to_sig(encodec_sd, c["toks"].long(), c["sample_rate"], 24000)
    ref_toks = c["toks"].long()
    assert toks.shape == ref_toks.shape and toks.dtype == torch.int64
    safe = gaps > 1e-4
    assert (toks == ref_toks)[safe].all(), "oracle codes differ from the reference away from near-ties"
    assert (toks == ref_toks).float().mean() > 0.99
    assert rec.shape == c["rec"].shape
    assert (rec - c["rec"]).abs().max() <= 1e-4 * max(1.0, c["rec"].abs().max().item())


def test_encodec_lstm_loop_equals_aten(encodec_sd):
    x = torch.randn(2, 512, 9, generator=torch.Generator().manual_seed(3))
    a = encodec_ref.lstm_block(encodec_sd, "encoder.layers.13", x)
    b = encodec_ref.lstm_block_aten(encodec_sd, "encoder.layers.13", x)
    assert (a - b).abs().max() < 1e-5


def test_encodec_invalid_num_codebooks():
    with pytest.raises(ValueError):
        encodec_ref.num_quantizers_for(3)


@pytest.mark.parametrize("o,n", [(16000, 24000), (24000, 16000), (16000, 44100), (44100, 16000)])
def test_resample_oracle_matches_torchaudio(o, n):
    torchaudio = pytest.importorskip("torchaudio")
    x = make_input(5, 2, 4001)
    ref = torchaudio.functional.resample(x, o, n)
    got = resample_ref.resample(x, o, n)
    assert got.shape == ref.shape
    assert (got - ref).abs().max() < 1e-6


@pytest.mark.parametrize("case", range(4))
def test_mimi_oracle_matches_reference_golden(mimi_sd, mimi_golden, case):
    from oracle import mimi_ref
    c = mimi_golden["cases"][case]
    sig = make_input(c["seed"], c["B"], c["T"])
    with torch.no_grad():
        toks, gaps, _ = mimi_ref.sig_to_toks(mimi_sd, sig, c["K"], c["sample_rate"], return_gaps=True)
        rec = mimi_ref.toks_to_sig(mimi_sd, c["toks"].long(), c["sample_rate"])
    ref_toks = c["toks"].long()
    assert toks.shape == ref_toks.shape
    assert (toks == ref_toks)[gaps > 1e-4].all()
    assert rec.shape == c["rec"].shape
    assert (rec - c["rec"]).abs().max() <= 1e-4 * max(1.0, c["rec"].abs().max().item())


@pytest.mark.parametrize("case", range(3))
def test_dac_oracle_matches_reference_golden(dac_sd, dac_golden, case):
    from oracle import dac_ref
    c = dac_golden["cases"][case]
    sig = make_input(c["seed"], c["B"], c["T"])
    with torch.no_grad():
        toks, gaps, _ = dac_ref.sig_to_toks(dac_sd, sig, c["K"], c["sample_rate"], 44100, return_gaps=True)
        rec = dac_ref.toks_to_sig(dac_sd, c["toks"].long(), c["sample_rate"], 44100)
    ref_toks = c["toks"].long()
    assert toks.shape == ref_toks.shape
    assert (toks == ref_toks)[gaps > 1e-4].all()
    assert rec.shape == c["rec"].shape
    assert (rec - c["rec"]).abs().max() <= 1e-4


def test_mimi_invalid_num_codebooks(mimi_sd):
    from oracle import mimi_ref
    with pytest.raises(ValueError):
        mimi_ref.rvq_encode(mimi_sd, torch.zeros(1, 512, 2), 33)


@pytest.mark.parametrize("case", range(3))
def test_dac_odd_stride_oracle_matches_reference_golden(case):
    """the 16 / 24 kHz DAC architectures (stride-5 blocks; `DAC(sample_rate)`'s default): tests/golden/dac_odd_golden.pt was
    recorded from the unmodified wrapper over the HF twin (oracle/make_golden_dac.py: golden_dac_odd)."""
    import os
    from oracle import dac_ref, weights
    c = torch.load(os.path.join(weights.GOLDEN_DIR, "dac_odd_golden.pt"))["cases"][case]
    sd = weights.dac_state_dict(0, tag=f"{c['orig_sample_rate'] // 1000}khz")
    sig = make_input(c["seed"], c["B"], c["T"])
    with torch.no_grad():
        toks, gaps, _ = dac_ref.sig_to_toks(sd, sig, c["K"], c["sample_rate"], c["orig_sample_rate"], return_gaps=True)
        rec = dac_ref.toks_to_sig(sd, c["toks"].long(), c["sample_rate"], c["orig_sample_rate"])
    ref_toks = c["toks"].long()
    assert toks.shape == ref_toks.shape
    assert (toks == ref_toks)[gaps > 1e-4].all()
    assert rec.shape == c["rec"].shape
    assert (rec - c["rec"]).abs().max() <= 1e-4
