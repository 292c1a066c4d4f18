"""ctypes binding of the C-ABI shared library (`include/b200sv.h`, built from `csrc/` by
`__graft_entry__.build()` into `qadence_b200/lib/libb200sv.so`).

The reference has no FFI (it is pure Python over pyqtorch); this module is the whole binding a
maintainer needs — see INTEGRATION.md.  There is NO CPU fallback: if the library is missing or no
CUDA device is visible, every compute call raises.
"""

from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

LIB_PATH = Path(__file__).resolve().parent / "lib" / "libb200sv.so"

C64, C128 = 0, 1

EXPORTS = [
    "b200sv_abi_version",
    "b200sv_lean_supported",
    "b200sv_device_sm_count",
    "b200sv_init_zero_state",
    "b200sv_plan_workspace_bytes",
    "b200sv_apply_plan",
    "b200sv_adjoint_plan",
    "b200sv_pauli_expectation",
    "b200sv_pauli_apply",
    "b200sv_pauli_diagonal",
    "b200sv_pauli_apply_diag",
    "b200sv_pauli_expectation_diag",
    "b200sv_diag_evolve",
    "b200sv_dot",
    "b200sv_axpy",
    "b200sv_scal",
    "b200sv_lincomb",
    "b200sv_lanczos_step",
    "b200sv_tridiag_expm",
    "b200sv_apply_dense",
    "b200sv_sample",
    "b200sv_bit_reverse",
    "b200sv_permute_bits",
    "b200sv_apply_plan_peers",
    "b200sv_adjoint_plan_peers",
    "b200sv_pauli_expectation_peers",
    "b200sv_pauli_apply_peers",
    "b200sv_ipc_alloc",
    "b200sv_ipc_open",
    "b200sv_ipc_close",
    "b200sv_ipc_free",
    "b200sv_embed_forward",
    "b200sv_embed_backward",
]


class PlanStruct(C.Structure):
    """`b200sv_plan` (include/b200sv.h)."""

    _fields_ = [
        ("sweeps", C.c_void_p),
        ("rounds", C.c_void_p),
        ("gates", C.c_void_p),
        ("consts", C.c_void_p),
        ("sweeps_host", C.c_void_p),
        ("n_sweeps", C.c_int32),
        ("n_exec_total", C.c_int32),
        ("exec_leader", C.c_void_p),
        ("lean_tab_fwd", C.c_void_p),
        ("lean_tab_bwd", C.c_void_p),
        ("tma_host", C.c_void_p),
        ("exec_mode", C.c_void_p),
    ]


class B200svError(RuntimeError):
    pass


_lib: C.CDLL | None = None


def library_path() -> Path:
    return Path(os.environ.get("B200SV_LIB", str(LIB_PATH)))


def load() -> C.CDLL:
    """Load the shared library and declare every prototype; raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not path.exists():
        raise B200svError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). The b200sv backend has no CPU fallback."
        )
    lib = C.CDLL(str(path))
    vp, i32, i64, u64, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double
    protos = {
        "b200sv_abi_version": [],
        "b200sv_lean_supported": [i32, i32, i32, i32],
        "b200sv_device_sm_count": [],
        "b200sv_init_zero_state": [vp, i32, i32, i64, vp],
        "b200sv_plan_workspace_bytes": [vp, i32, i64, i32],
        "b200sv_apply_plan": [vp, i32, i32, i64, vp, vp, i32, i32, u64, vp, vp],
        "b200sv_adjoint_plan": [vp, vp, i32, i32, i64, vp, vp, vp, i32, i32, u64, vp, vp],
        "b200sv_pauli_expectation": [vp, i32, i32, i64, vp, i32, i32, vp, i64, vp, u64, vp],
        "b200sv_pauli_apply": [vp, vp, vp, i32, i32, i64, vp, i32, i32, vp, i64, f64, f64, f64, f64, u64, vp],
        "b200sv_pauli_diagonal": [vp, i32, i32, vp, i64, i32, i64, vp, u64, vp],
        "b200sv_pauli_apply_diag": [vp, vp, vp, i32, i32, i64, vp, i32, vp, i64, vp, i64, vp, f64, f64, f64, f64, vp],
        "b200sv_pauli_expectation_diag": [vp, i32, i32, i64, vp, i32, vp, i64, vp, i64, vp, vp],
        "b200sv_diag_evolve": [vp, i32, i32, i64, vp, i32, i32, vp, i64, vp, i64, f64, u64, vp],
        "b200sv_dot": [vp, vp, i32, i32, i64, vp, vp],
        "b200sv_axpy": [vp, vp, vp, i32, i32, i64, vp],
        "b200sv_scal": [vp, vp, i32, i32, i64, vp],
        "b200sv_lincomb": [vp, vp, i32, vp, i32, i32, i64, vp],
        "b200sv_lanczos_step": [vp, vp, vp, vp, vp, vp, vp, i32, i32, i64, vp],
        "b200sv_tridiag_expm": [vp, vp, vp, i32, i64, vp, vp, vp],
        "b200sv_apply_dense": [vp, vp, i32, i32, i64, vp, vp, i32, vp],
        "b200sv_sample": [vp, i32, i32, i64, vp, vp, vp, i64, vp, vp],
        "b200sv_bit_reverse": [vp, vp, i32, i32, i64, vp],
        "b200sv_permute_bits": [vp, vp, i32, i32, i64, vp, vp],
        "b200sv_apply_plan_peers": [vp, i32, i32, i32, i32, vp, vp, i32, i32, vp, vp],
        "b200sv_adjoint_plan_peers": [vp, vp, i32, i32, i32, i32, vp, vp, vp, i32, i32, vp, vp],
        "b200sv_pauli_expectation_peers": [vp, i32, i32, i32, i32, vp, i32, i32, vp, vp, vp],
        "b200sv_pauli_apply_peers": [vp, i32, i32, vp, i32, i32, vp, i32, i32, vp, f64, f64, vp],
        "b200sv_ipc_alloc": [i64, vp, vp],
        "b200sv_ipc_open": [vp, vp],
        "b200sv_ipc_close": [vp],
        "b200sv_ipc_free": [vp],
        "b200sv_embed_forward": [vp, vp, i32, vp, vp, vp, i64, vp, vp],
        "b200sv_embed_backward": [vp, vp, i32, vp, vp, vp, i64, vp, vp, vp],
    }
    for name, args in protos.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int64 if name == "b200sv_plan_workspace_bytes" else C.c_int
    if lib.b200sv_abi_version() != 2:
        raise B200svError("libb200sv.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    if rc > 0:
        raise B200svError(f"{what}: CUDA error {rc}")
    names = {-1: "invalid argument", -2: "unsupported", -3: "no CUDA device"}
    raise B200svError(f"{what}: {names.get(rc, rc)}")


def require_device() -> int:
    """Number of SMs of the current CUDA device; raises when there is none (no CPU fallback)."""
    sms = load().b200sv_device_sm_count()
    if sms <= 0:
        raise B200svError(
            "b200sv needs a CUDA device (sm_100a): none is visible and there is no CPU fallback"
        )
    return sms


def host_ptr(a: np.ndarray) -> int:
    return a.ctypes.data
