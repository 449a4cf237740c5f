"""ctypes binding of libmicrodit_b200.so (the C ABI declared in include/microdit_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails the error is
raised to the caller.  Building is explicit (`python -m micro_diffusion_b200.build`).
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import os

# MD_LIB_PATH: developer knob to load an experimental build of the same ABI (tools/build_variant.py)
_LIB_PATH = Path(os.environ.get("MD_LIB_PATH") or Path(__file__).resolve().parent / "libmicrodit_b200.so")
_lib = None


class MicroditLibraryError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("B", C.c_void_p), ("C", C.c_void_p), ("C2", C.c_void_p),
        ("bias", C.c_void_p), ("res", C.c_void_p), ("gate", C.c_void_p), ("aux", C.c_void_p),
        ("M", C.c_int64), ("N", C.c_int64), ("K", C.c_int64),
        ("lda", C.c_int64), ("ldb", C.c_int64), ("ldc", C.c_int64),
        ("batch", C.c_int64),
        ("strideA", C.c_int64), ("strideB", C.c_int64), ("strideC", C.c_int64), ("strideBias", C.c_int64),
        ("ldgate", C.c_int64), ("rows_per_gate", C.c_int64), ("res_mod", C.c_int64),
        ("layout", C.c_int32), ("epilogue", C.c_int32), ("splits", C.c_int32), ("act", C.c_int32),
        ("alpha", C.c_float), ("sm_limit", C.c_int32),
        ("ldc2", C.c_int64), ("strideC2", C.c_int64), ("row_interleave", C.c_int64),
    ]


def lib_path() -> Path:
    return _LIB_PATH


def load():
    """Load the shared library once; raise loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise MicroditLibraryError(
            f"{_LIB_PATH} not found: build it with `python -m micro_diffusion_b200.build` "
            "(there is no CPU or PyTorch fallback for the MicroDiT hot path)")
    lib = C.CDLL(str(_LIB_PATH))
    lib.md_last_error.restype = C.c_char_p
    lib.md_last_error.argtypes = []
    lib.md_abi_version.restype = C.c_int
    lib.md_gemm_bf16.restype = C.c_int
    lib.md_gemm_bf16.argtypes = [C.POINTER(GemmArgs), C.c_void_p]
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().md_last_error().decode("utf-8", "replace")
        raise MicroditLibraryError(f"{what} failed with code {rc}: {msg}")
