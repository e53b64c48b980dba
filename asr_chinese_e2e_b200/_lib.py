"""Loader (and in-tree builder) of libctcb200.so, the C-ABI CUDA library (include/ctcb200.h).

There is exactly one implementation of the hot path: hand-written sm_100a kernels.  If the
shared library is missing or cannot be loaded this module raises -- there is no CPU or
PyTorch fallback (BASELINE.json north_star).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
SO_PATH = os.path.join(_HERE, "libctcb200.so")
SOURCES = ["ctcb200.cu", "head.cu"]
HEADERS = ["ptx.cuh", "layout.h", "stream_kernels.cuh", "sweep_direct.cuh", "lattice_kernel.cuh", "lattice_lin.cuh", "ce_kernel.cuh",
           "decode_kernel.cuh", "head_kernels.cuh", "gemm_tf32x3.cuh", "internal.h", os.path.join("..", "..", "include", "ctcb200.h")]

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-diag-suppress", "128"]


def needs_build() -> bool:
    if not os.path.exists(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    return any(os.path.getmtime(os.path.join(_CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a -> asr_chinese_e2e_b200/libctcb200.so (in-tree)."""
    if not force and not needs_build():
        return SO_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    bdir = os.path.join(_HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    # one object per translation unit, compiled in parallel; then one link into the in-tree shared library
    procs, objs = [], []
    for src in SOURCES:
        obj = os.path.join(bdir, os.path.splitext(src)[0] + ".o")
        objs.append(obj)
        newest = max(os.path.getmtime(os.path.join(_CSRC, f)) for f in [src] + HEADERS)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > newest:
            continue
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, os.path.join(_CSRC, src)]
        procs.append((cmd, subprocess.Popen(cmd)))
    for cmd, p in procs:
        if p.wait() != 0:
            raise subprocess.CalledProcessError(p.returncode, cmd)
    subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", SO_PATH] + objs)
    return SO_PATH


class CtcB200Error(RuntimeError):
    pass


_lib = None

_i, _i64, _sz, _f, _p = ctypes.c_int, ctypes.c_int64, ctypes.c_size_t, ctypes.c_float, ctypes.c_void_p

_FWD_ARGS = [_p, _p, _i64, _i64, _p, _p, _i, _i, _i, _i, _i, _i, _p, _p, _p, _sz, _p, _p]
SIGNATURES = {
    "ctcb200_version": (_i, []),
    "ctcb200_strerror": (ctypes.c_char_p, [_i]),
    "ctcb200_set_option": (_i, [ctypes.c_char_p, _i]),
    "ctcb200_get_option": (_i, [ctypes.c_char_p, ctypes.POINTER(_i)]),
    "ctcb200_workspace_bytes": (_i, [_i, _i, _i, _i, ctypes.POINTER(_sz)]),
    "ctcb200_forward": (_i, _FWD_ARGS),
    "ctcb200_loss_only": (_i, _FWD_ARGS),
    "ctcb200_loss_grad": (_i, [_p, _p, _i64, _i64, _p, _p, _i, _i, _i, _i, _i, _i, _i, _f, _p, _p, _p, _p, _sz, _p, _p]),
    "ctcb200_loss_grad_stages": (_i, [_i, _p, _p, _i64, _i64, _p, _p, _i, _i, _i, _i, _i, _i, _i, _f, _p, _p, _p, _p, _sz, _p]),
    "ctcb200_backward": (_i, [_p, _p, _i64, _i64, _p, _i64, _i, _f, _i, _i, _i, _i, _i, _i, _p, _p, _sz, _p]),
    "ctcb200_rescale_grad": (_i, [_p, _p, _i64, _p, _p, _i, _i, _i, _p]),
    "ctcb200_greedy_decode": (_i, [_p, _i64, _i64, _i, _i, _i, _i, _i, _p, _sz, _p, _p, _p, _p]),
    "ctcb200_edit_distance": (_i, [_p, _i64, _p, _i64, _i, _i, _i, _i, _p, _p, _p]),
    "ctcb200_ce_workspace_bytes": (_i, [_i64, ctypes.POINTER(_sz)]),
    "ctcb200_ce_loss_grad": (_i, [_p, _p, _i64, _i, _i, _f, _f, _p, _p, _p, _sz, _p]),
    "ctcb200_head_workspace_bytes": (_i, [_i, _i, _i, _i, _i, _i, ctypes.POINTER(_sz)]),
    "ctcb200_head_loss": (_i, [_p, _p, _p, _p, _i64, _i64, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _sz, _p]),
    "ctcb200_head_loss_grad": (_i, [_p, _p, _p, _p, _i64, _i64, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _f, _p, _p,
                                    _p, _i64, _p, _sz, _p]),
    "ctcb200_head_param_grads_workspace_bytes": (_i, [_i, _i, ctypes.POINTER(_sz)]),
    "ctcb200_head_param_grads": (_i, [_p, _i64, _p, _p, _i, _i, _i, _i, _p, _p, _p, _sz, _p]),
    "ctcb200_read_status": (_i, [_p, ctypes.POINTER(_i), _p]),
    "ctcb200_read_lattice_stats": (_i, [_p, ctypes.POINTER(_i), _p]),
}


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise CtcB200Error(
                f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU fallback for the CTC hot path.")
        try:
            handle = ctypes.CDLL(SO_PATH)
        except OSError as e:  # pragma: no cover
            raise CtcB200Error(f"cannot load {SO_PATH}: {e}") from e
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def strerror(code: int) -> str:
    return lib().ctcb200_strerror(int(code)).decode()


def check(code: int, what: str) -> None:
    if code != 0:
        raise CtcB200Error(f"{what} failed: [{code}] {strerror(code)}")


def head_workspace_bytes(B: int, T: int, V: int, K: int, Umax: int, precision: int) -> int:
    out = _sz(0)
    check(lib().ctcb200_head_workspace_bytes(B, T, V, K, Umax, precision, ctypes.byref(out)), "ctcb200_head_workspace_bytes")
    return int(out.value)


def head_param_grads_workspace_bytes(V: int, K: int) -> int:
    out = _sz(0)
    check(lib().ctcb200_head_param_grads_workspace_bytes(V, K, ctypes.byref(out)), "ctcb200_head_param_grads_workspace_bytes")
    return int(out.value)


def set_option(name: str, value: int) -> None:
    """Developer tunable of the library; the CTCB200_* environment variables are read once,
    when the library is loaded -- afterwards this is the way to change one."""
    check(lib().ctcb200_set_option(name.encode(), int(value)), f"ctcb200_set_option({name})")


def get_option(name: str) -> int:
    out = _i(0)
    check(lib().ctcb200_get_option(name.encode(), ctypes.byref(out)), f"ctcb200_get_option({name})")
    return int(out.value)


def workspace_bytes(B: int, T: int, V: int, Umax: int) -> int:
    out = _sz(0)
    check(lib().ctcb200_workspace_bytes(B, T, V, Umax, ctypes.byref(out)), "ctcb200_workspace_bytes")
    return int(out.value)
