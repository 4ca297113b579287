"""ctypes binding of the C-ABI shared library (include/audiocodecs_b200.h).

There is NO fallback: if the library is missing or a symbol is absent this module raises, and every
op in this package goes through it.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AC_LIB_PATH") or os.path.join(_HERE, "lib", "libaudiocodecs_b200.so")   # AC_LIB_PATH: A/B a second build
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "audiocodecs_b200.h")

c_i32, c_i64, c_vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_void_p


class AcConvF32(ctypes.Structure):
    """mirror of `struct ac_conv_f32`"""
    _fields_ = [
        ("x", c_vp), ("w", c_vp), ("bias", c_vp), ("alpha", c_vp), ("res", c_vp), ("y", c_vp), ("vlen", c_vp),
        ("x_bstride", c_i64), ("y_bstride", c_i64), ("res_bstride", c_i64),
        ("x_rstride", c_i32),
        ("batch", c_i32), ("x_rows", c_i32), ("cin", c_i32), ("m_rows", c_i32), ("n_cols", c_i32),
        ("taps", c_i32), ("stride", c_i32), ("dilation", c_i32), ("pad_left", c_i32),
        ("pad_mode", c_i32), ("reflect_len", c_i32), ("act", c_i32), ("epi", c_i32),
        ("out_shift", c_i64), ("out_valid", c_i64),
        ("x_is_bf16", c_i32), ("act2", c_i32), ("y_bf16", c_vp), ("y_act_bf16", c_vp),
        ("y_bf16_bstride", c_i64), ("y_act_bstride", c_i64),
    ]


class AcTcSrc(ctypes.Structure):
    """mirror of `struct ac_tc_src`"""
    _fields_ = [("base", c_vp), ("c0", c_i32), ("phases", c_i32), ("rows", c_i32),
                ("phase_stride", c_i64), ("row_stride", c_i64), ("batch_stride", c_i64),
                ("taps", c_i32), ("dilation", c_i32), ("shift", c_i32), ("lo_of", c_i32)]


class AcConvTcDesc(ctypes.Structure):
    """mirror of `struct ac_conv_tc_desc`"""
    _fields_ = [("src", AcTcSrc * 4), ("n_src", c_i32), ("w", c_vp), ("w_split", c_i32), ("k_total", c_i32), ("n_total", c_i32),
                ("bk", c_i32), ("bias", c_vp), ("alpha", c_vp), ("res", c_vp), ("y", c_vp), ("y_act", c_vp), ("y_lo", c_vp),
                ("y_act_lo", c_vp), ("y32", c_vp),
                ("act", c_i32), ("epi", c_i32), ("act_mod", c_i32),
                ("y_bstride", c_i64), ("y_act_bstride", c_i64), ("y32_bstride", c_i64), ("res_bstride", c_i64),
                ("out_shift", c_i64), ("out_valid", c_i64),
                ("batch", c_i32), ("m_rows", c_i32), ("n_tile_hint", c_i32), ("grid_hint", c_i32),
                ("res_lo", c_vp), ("res32", c_vp), ("g_hint", c_i32), ("fmt", c_i32), ("flush_adds", c_i32)]


class AcResunitTcDesc(ctypes.Structure):
    """mirror of `struct ac_resunit_tc_desc`"""
    _fields_ = [("a", c_vp), ("a_lo", c_vp), ("a_row_stride", c_i64), ("a_bstride", c_i64),
                ("a_rows", c_i32), ("cin", c_i32), ("taps", c_i32), ("dilation", c_i32), ("shift", c_i32),
                ("x", c_vp), ("x_lo", c_vp), ("x_bstride", c_i64), ("w1", c_vp), ("w2", c_vp),
                ("w1_split", c_i32), ("w2_split", c_i32), ("ch", c_i32), ("cout", c_i32), ("h_split", c_i32),
                ("bias1", c_vp), ("alpha1", c_vp), ("bias2", c_vp), ("alpha2", c_vp), ("act1", c_i32), ("act2", c_i32),
                ("res", c_vp), ("res_lo", c_vp), ("res_bstride", c_i64),
                ("y", c_vp), ("y_lo", c_vp), ("y_act", c_vp), ("y_act_lo", c_vp), ("y_bstride", c_i64), ("y_act_bstride", c_i64),
                ("batch", c_i32), ("m_rows", c_i32), ("bk", c_i32), ("g_hint", c_i32), ("grid_hint", c_i32), ("dbl_hint", c_i32),
                ("act0", c_i32), ("e_split", c_i32), ("x_from_a", c_i32), ("alpha0", c_vp), ("x_row_off", c_i32), ("fmt", c_i32), ("io_stage", c_i32)]


class AcLstmTcDesc(ctypes.Structure):
    """mirror of `struct ac_lstm_tc_desc`"""
    _fields_ = [("pre", c_vp), ("w_hh_bf16", c_vp), ("out_hi", c_vp), ("out_lo", c_vp), ("skip_hi", c_vp), ("skip_lo", c_vp),
                ("final_hi", c_vp), ("final_lo", c_vp), ("skip_bstride", c_i64), ("final_bstride", c_i64),
                ("final_act", c_i32), ("batch", c_i32), ("steps", c_i32), ("hidden", c_i32), ("dbg", c_vp), ("operand_fp16", c_i32), ("out_fp16", c_i32), ("skip_fp16", c_i32)]


def declared_symbols():
    """Every `AC_API` entry point the public header declares."""
    with open(HEADER_PATH) as f:
        return re.findall(r"AC_API\s+[\w\s\*]+?\b(ac_\w+)\s*\(", f.read())


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"audiocodecs_b200: CUDA library not built ({LIB_PATH}). Run `python -c 'import __graft_entry__ as g; "
                "g.build()'` or `make -C audiocodecs_b200/csrc`. There is no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name in declared_symbols():
            if not hasattr(L, name):
                raise RuntimeError(f"audiocodecs_b200: {LIB_PATH} does not export {name}")
        L.ac_last_error.restype = ctypes.c_char_p
        L.ac_launch_count.restype = c_i64
        L.ac_abi_version.restype = c_i32
        L.ac_conv1d_f32.argtypes = [ctypes.POINTER(AcConvF32), c_vp]
        L.ac_lstm_layer_f32.argtypes = [c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp]
        L.ac_rvq_encode_tc.argtypes = [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp]
        L.ac_rvq_encode_f32.argtypes = [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp]
        L.ac_rvq_decode_f32.argtypes = [c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp]
        L.ac_resample_f32.argtypes = [c_vp, c_vp, c_vp, c_i32, c_i64, c_i64, c_i32, c_i32, c_i32, c_i32, c_vp]
        L.ac_layernorm_f32.argtypes = [c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, ctypes.c_float, c_vp]
        L.ac_attention_f32.argtypes = [c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, ctypes.c_float, c_vp]
        L.ac_upsample_dw_f32.argtypes = [c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_vp]
        L.ac_layernorm_split_bf16.argtypes = [c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i64, ctypes.c_float, c_i32, c_vp]
        L.ac_rope_table_f32.argtypes = [c_vp, c_vp, c_i32, c_i32, c_vp]
        L.ac_attention_tc.argtypes = [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, ctypes.c_float, c_i32, c_vp]
        L.ac_dac_rvq_encode_f32.argtypes = [c_vp] * 8 + [c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp]
        L.ac_dac_rvq_encode_proj_f32.argtypes = [c_vp, c_i32] + [c_vp] * 6 + [c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp]
        L.ac_dac_rvq_decode_f32.argtypes = [c_vp] * 5 + [c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp]
        L.ac_conv_first_bf16.argtypes = [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i32, c_i32, c_i32, c_i32,
                                         c_i32, c_i32, c_i32, c_i32, c_i32, c_vp]
        L.ac_conv_last_bf16.argtypes = [c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp]
        L.ac_lstm_tc.argtypes = [ctypes.POINTER(AcLstmTcDesc), c_vp]
        L.ac_rvq_decode_bf16.argtypes = [c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_i32, c_vp]
        L.ac_conv_tc.argtypes = [ctypes.POINTER(AcConvTcDesc), c_vp]
        L.ac_resunit_tc.argtypes = [ctypes.POINTER(AcResunitTcDesc), c_vp]
        L.ac_add_act_bf16.argtypes = [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i64, c_i64, c_i64, c_i64, c_i32, c_i32, c_vp]
        L.ac_f32_to_split_bf16.argtypes = [c_vp, c_vp, c_vp, c_i32, c_i64, c_i64, c_i64, c_i32, c_vp]
        L.ac_token_histogram.argtypes = [c_vp, c_i64, c_i32, c_i32, c_vp, c_vp, c_vp]
        L.ac_multihead_embedding.argtypes = [c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_i64, c_i64, c_i64, c_vp, c_vp]
        L.ac_lstm_tc_max_clusters.argtypes = [c_i32, c_i32]
        L.ac_pad_halo_bf16.argtypes = [c_vp, c_i32, c_i32, c_i32, c_i64, c_i32, c_i32, c_i32, c_i32, c_vp]
        L.ac_pad_halo2_bf16.argtypes = [c_vp, c_vp, c_i32, c_i32, c_i32, c_i64, c_i32, c_i32, c_i32, c_i32, c_vp]
        _lib = L
    return _lib


class ConfigError(RuntimeError):
    """an entry point rejected its arguments or found no tiling for the shape (rc = -1); nothing was launched"""


def check(rc, what):
    if rc != 0:
        msg = lib().ac_last_error().decode(errors="replace")
        # rc = -1: argument / configuration check (AC_REQUIRE); rc > 0: a cudaError_t from the launch
        raise (ConfigError if rc == -1 else RuntimeError)(f"audiocodecs_b200: {what} failed (rc={rc}): {msg}")


def launch_count():
    return int(lib().ac_launch_count())
