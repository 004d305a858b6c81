"""ctypes binding of libs2ag_b200.so (the C ABI declared in include/s2ag.h).

The prototypes are parsed from the header itself so the binding cannot drift from the
declaration.  There is NO fallback: if the library is missing, `lib()` raises, and every op in
`ops.py` refuses non-CUDA tensors (the only exception is the kernel-logic emulator that
tests/emu injects explicitly through `_inject_for_tests`).
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.abspath(os.path.join(_HERE, "..", "include", "s2ag.h"))
LIB_PATH = os.path.join(_HERE, "libs2ag_b200.so")

_CT = {
    "int": ctypes.c_int, "long": ctypes.c_long, "float": ctypes.c_float, "uint64_t": ctypes.c_uint64,
    "void": None, "char*": ctypes.c_char_p, "unsignedlonglong": ctypes.c_ulonglong,
}


def _ctype(t):
    t = t.replace("const", "").strip()
    t = re.sub(r"\s+", "", t)
    if t.endswith("*"):
        return ctypes.c_char_p if t == "char*" else ctypes.c_void_p
    return _CT[t]


def parse_header(path=HEADER):
    """-> {name: (restype, [argtypes], [argnames])} for every function declared in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"^\s*#.*$", "", src, flags=re.M)
    protos = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(s2ag_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        argtypes, argnames = [], []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                mm = re.match(r"(.*?)(\w+)$", a, flags=re.S)
                argtypes.append(_ctype(mm.group(1)))
                argnames.append(mm.group(2))
        protos[name] = (_ctype(ret), argtypes, argnames)
    return protos


class S2agError(RuntimeError):
    pass


_lib = None
_emulated = False


def _bind(cdll, strict=True):
    for name, (res, argtypes, _) in parse_header().items():
        if not strict and not hasattr(cdll, name):
            continue
        fn = getattr(cdll, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = argtypes
    return cdll


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise S2agError(
                "libs2ag_b200.so is not built (%s). Run `python -m speech2affective_gestures_b200.build` "
                "(nvcc, sm_100a). There is no CPU or PyTorch fallback for this path." % LIB_PATH)
        _lib = _bind(ctypes.CDLL(LIB_PATH))
        if os.environ.get("S2AG_DEBUG_FLAGS") or os.environ.get("S2AG_DEBUG_FLAGS_FORCE"):
            # bring-up / A-B switches of include/s2ag.h (s2ag_debug_flags); ..._FORCE bits survive later calls
            _lib.s2ag_debug_flags(int(os.environ.get("S2AG_DEBUG_FLAGS", "0"), 0))
    return _lib


def _inject_for_tests(path, strict=True):
    """tests/emu only: route calls to the CPU kernel-logic emulator build.  Refused unless the test harness says so
    (S2AG_ALLOW_EMU=1, set by tests/conftest.py): the product path never runs on the emulator."""
    global _lib, _emulated
    if os.environ.get("S2AG_ALLOW_EMU") != "1":
        raise S2agError("the kernel-logic emulator is test infrastructure (set S2AG_ALLOW_EMU=1 in the test harness)")
    _lib = _bind(ctypes.CDLL(path), strict)
    _emulated = _lib.s2ag_is_device_build() == 0
    return _lib


def is_emulated():
    return _emulated


# S2AG_NVTX=1: one NVTX range per C-ABI call (named after the entry point) for nsys / ncu --nvtx timelines
_NVTX = os.environ.get("S2AG_NVTX") == "1"
_DEBUG_CAPTURE = os.environ.get("S2AG_DEBUG_CAPTURE") == "1"
_dbg_state = {"bad": False}


def call(name, *args):
    l = lib()
    if _DEBUG_CAPTURE and not _dbg_state["bad"]:
        import threading
        st = l.s2ag_stream_capture_status(args[-1])
        if st not in (0, 1):
            _dbg_state["bad"] = True
            print("[s2ag debug] capture already INVALID (%d) before %s (thread %s)" % (
                st, name, threading.current_thread().name), flush=True)
    if _NVTX:
        import torch
        torch.cuda.nvtx.range_push(name)
        try:
            rc = getattr(l, name)(*args)
        finally:
            torch.cuda.nvtx.range_pop()
    else:
        rc = getattr(l, name)(*args)
    if _DEBUG_CAPTURE and not _dbg_state["bad"]:
        st = l.s2ag_stream_capture_status(args[-1])
        if st not in (0, 1):
            _dbg_state["bad"] = True
            print("[s2ag debug] capture became INVALID (%d) inside %s" % (st, name), flush=True)
    if rc != 0:
        raise S2agError("%s failed (%d): %s" % (name, rc, l.s2ag_last_error().decode()))
