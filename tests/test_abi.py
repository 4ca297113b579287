"""CPU: the C-ABI library is built in-tree, loads, and exports every symbol include/*.h declares."""
import ctypes
import os

import pytest
import torch

from audiocodecs_b200 import _lib


def test_library_exports_every_declared_symbol():
    names = _lib.declared_symbols()
    assert {"ac_conv1d_f32", "ac_lstm_layer_f32", "ac_rvq_encode_f32", "ac_rvq_decode_f32", "ac_resample_f32"} <= set(names)
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    L = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), n
    assert _lib.lib().ac_abi_version() == 3


def test_struct_mirror_matches_header_size():
    # 7 pointers + 3 int64 + 14 int32 + 2 int64 + 2 int32 + 2 pointers + 2 int64, natural alignment
    assert ctypes.sizeof(_lib.AcConvF32) == 7 * 8 + 3 * 8 + 14 * 4 + 2 * 8 + 2 * 4 + 2 * 8 + 2 * 8
    assert ctypes.sizeof(_lib.AcTcSrc) == 8 + 3 * 4 + 4 + 3 * 8 + 4 * 4



def test_no_cpu_fallback(encodec_sd):
    import audiocodecs_b200 as A
    c = A.Encodec(24000, 24000, num_codebooks=8, state_dict=encodec_sd)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        c.sig_to_toks(torch.zeros(1, 2400))


def test_constructor_contract(encodec_sd):
    import audiocodecs_b200 as A
    with pytest.raises(ValueError):
        A.Encodec(24000, 24000, mode="bogus", state_dict=encodec_sd)
    c = A.Encodec(16000, 24000, mode="encode", num_codebooks=4, state_dict=encodec_sd)
    assert (c.sample_rate, c.orig_sample_rate, c.num_codebooks, c.vocab_size, c.mode) == (16000, 24000, 4, 1024, "encode")
    assert not hasattr(c, "_dec")
    bad = A.Encodec(24000, 24000, num_codebooks=3, state_dict=encodec_sd)
    with pytest.raises(ValueError, match="bandwidth"):
        bad._num_quantizers()


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    """No fallback: with the shared library absent the loader raises and names the build command."""
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "libaudiocodecs_b200.so"))
    with pytest.raises(RuntimeError, match="not built"):
        _lib.lib()


def test_config_error_is_a_runtime_error():
    assert issubclass(_lib.ConfigError, RuntimeError)


def test_library_sass_uses_tcgen05_tmem_tma():
    """the shipped library is real sm_100a code: tcgen05.mma (UTCHMMA), tensor-memory loads / stores (LDTM / STTM), TMA tensor
    loads and stores (UTMALDG / UTMASTG) and bulk DSMEM copies (UBLKCP) are in its SASS; the tap-GEMM kernels address shared
    memory with LDS / STS, not generic loads (the round-2 instruction diet)."""
    import shutil
    import subprocess
    from audiocodecs_b200 import _lib
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    for op in ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS"):
        assert op in sass, op
    # per kernel: no generic LD.E / ST.E in the two tap-GEMM kernels
    cur, generic = None, {}
    for line in sass.splitlines():
        if "Function :" in line:
            cur = line.split("Function :")[1].strip()
        elif cur and ("conv_tc_kernel" in cur or "resunit_tc_kernel" in cur) and (" LD.E" in line or " ST.E" in line):
            generic[cur] = generic.get(cur, 0) + 1
    assert not generic, generic
