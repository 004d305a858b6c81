"""CPU checks of the C-ABI boundary: the product library (nvcc, sm_100a) loads in a GPU-less process and
exports every symbol include/s2ag.h declares; the header parser that generates the ctypes prototypes
sees all of them; no compute call is made (there is no GPU here and no CPU fallback)."""
import ctypes
import os
import subprocess

import pytest

from speech2affective_gestures_b200 import _C

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _ensure_built():
    if not os.path.exists(_C.LIB_PATH):
        from speech2affective_gestures_b200 import build
        build.build()
    return _C.LIB_PATH


def test_header_declares_the_documented_entry_points():
    protos = _C.parse_header()
    for name in ("s2ag_linear_fwd", "s2ag_conv_fwd", "s2ag_bn_fwd", "s2ag_bn_bwd", "s2ag_graph_fwd", "s2ag_tcn_block_fwd",
                 "s2ag_tcn_block_bwd", "s2ag_weight_norm_fwd", "s2ag_embedding_fwd", "s2ag_gru_layer_fwd",
                 "s2ag_gru_layer_bwd", "s2ag_reparam_tile_fwd", "s2ag_dhead_fwd", "s2ag_dis_loss", "s2ag_gen_loss",
                 "s2ag_adam_step", "s2ag_attention_fwd", "s2ag_set_engine", "s2ag_set_precision", "s2ag_version"):
        assert name in protos, name
    # plain C types only at the boundary
    for name, (res, argtypes, _) in protos.items():
        for t in argtypes:
            assert t in (ctypes.c_int, ctypes.c_long, ctypes.c_float, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_char_p), \
                (name, t)


def test_library_exports_every_declared_symbol():
    lib_path = _ensure_built()
    cdll = ctypes.CDLL(lib_path)
    missing = [n for n in _C.parse_header() if not hasattr(cdll, n)]
    assert not missing, missing
    cdll.s2ag_version.restype = ctypes.c_int
    assert cdll.s2ag_version() >= 100
    cdll.s2ag_is_device_build.restype = ctypes.c_int
    assert cdll.s2ag_is_device_build() == 1
    # argument validation happens before any launch: callable without a GPU
    cdll.s2ag_set_engine.restype = ctypes.c_int
    cdll.s2ag_last_error.restype = ctypes.c_char_p
    assert cdll.s2ag_set_engine(7) < 0 and b"engine" in cdll.s2ag_last_error()
    assert cdll.s2ag_set_engine(0) == 0
    # scratch registration is host-side bookkeeping: (stream, 16-byte aligned buffer, bytes) or (stream, NULL, 0)
    cdll.s2ag_register_scratch.restype = ctypes.c_int
    cdll.s2ag_register_scratch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long]
    assert cdll.s2ag_register_scratch(ctypes.c_void_p(0x10), ctypes.c_void_p(0x1000), 4096) == 0
    assert cdll.s2ag_register_scratch(ctypes.c_void_p(0x10), ctypes.c_void_p(0x1008), 4096) < 0   # misaligned
    assert cdll.s2ag_register_scratch(ctypes.c_void_p(0x10), None, 4096) < 0                      # NULL with a size
    assert cdll.s2ag_register_scratch(ctypes.c_void_p(0x10), None, 0) == 0                         # unregister


def test_library_contains_sm100a_tensor_core_code():
    lib_path = _ensure_built()
    try:
        sass = subprocess.run(["cuobjdump", "-sass", lib_path], capture_output=True, text=True, timeout=300).stdout
    except (FileNotFoundError, subprocess.TimeoutExpired):
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in sass
    assert "UTCHMMA" in sass and "LDTM" in sass  # tcgen05.mma / tcgen05.ld of the dense-contraction engine
