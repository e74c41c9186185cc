"""ctypes binding of libvitad.so (the C ABI declared in include/vitad.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C vit-ad_b200``; there is no
fallback: if the shared object is missing, importing this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libvitad.so")


class VitadError(RuntimeError):
    """Raised when a C-ABI call returns a negative vitad_status."""


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(the CUDA extension is mandatory; there is no CPU path)"
        )
    return C.CDLL(LIB_PATH)


lib = _load()
lib.vitad_last_error.restype = C.c_char_p
lib.vitad_abi_version.restype = C.c_int
lib.vitad_launch_count.restype = C.c_uint64

EPI_BIAS_BF16 = 0
EPI_BIAS_GELU_BF16 = 1
EPI_RESIDUAL_F32 = 2
EPI_QKV = 3
EPI_PATCH_EMBED = 4
EPI_F32 = 5


class LinearArgs(C.Structure):
    _fields_ = [
        ("a", C.c_void_p),
        ("w", C.c_void_p),
        ("bias", C.c_void_p),
        ("m", C.c_int),
        ("n", C.c_int),
        ("k", C.c_int),
        ("lda", C.c_int),
        ("ldw", C.c_int),
        ("epilogue", C.c_int),
        ("block_n", C.c_int),
        ("out", C.c_void_p),
        ("ldo", C.c_int),
        ("resid", C.c_void_p),
        ("q", C.c_void_p),
        ("kmat", C.c_void_p),
        ("vt", C.c_void_p),
        ("tokens", C.c_int),
        ("tokens_pad", C.c_int),
        ("heads", C.c_int),
        ("q_scale", C.c_float),
        ("pos", C.c_void_p),
        ("patches", C.c_int),
        ("prefix", C.c_int),
    ]


lib.vitad_linear_bf16.argtypes = [C.POINTER(LinearArgs), C.c_void_p]
lib.vitad_linear_bf16.restype = C.c_int


def check(rc: int) -> None:
    if rc != 0:
        raise VitadError(f"vitad status {rc}: {lib.vitad_last_error().decode()}")


def launch_count() -> int:
    return int(lib.vitad_launch_count())
