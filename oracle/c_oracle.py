"""ctypes binding of the plain-C float64 oracle (oracle/ctc_oracle.c).  TEST INFRASTRUCTURE."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libctc_oracle.so")


def build(force=False):
    src = os.path.join(_HERE, "ctc_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return _SO


_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.ctc_oracle_f64.restype = ctypes.c_int
    return _lib


def ctc_c_f64(logits, targets, input_lengths, target_lengths, blank=0, reduction="mean",
              zero_infinity=False, want_grad=True, nthreads=None):
    """Same contract as oracle.ctc_f64.ctc_f64 (grad_output = 1). numpy in, numpy out."""
    lib = _load()
    x = np.ascontiguousarray(logits, dtype=np.float32)
    B, T, V = x.shape
    tg = np.ascontiguousarray(targets, dtype=np.int64)
    il = np.ascontiguousarray(input_lengths, dtype=np.int64)
    tl = np.ascontiguousarray(target_lengths, dtype=np.int64)
    stride = tg.shape[1] if tg.ndim == 2 else 0
    denom = np.maximum(tl, 1).astype(np.float64)
    scale = np.ones(B) if reduction in ("none", "sum") else 1.0 / (B * denom)
    scale = np.ascontiguousarray(scale, dtype=np.float64)
    nll = np.empty(B, dtype=np.float64)
    grad = np.empty((B, T, V), dtype=np.float64) if want_grad else None
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p) if a is not None else None
    rc = lib.ctc_oracle_f64(p(x), p(tg), ctypes.c_int64(stride), p(il), p(tl), B, T, V,
                            int(blank), int(bool(zero_infinity)), p(scale), p(nll), p(grad),
                            int(nthreads or os.cpu_count() or 1))
    if rc != 0:
        raise MemoryError("ctc_oracle_f64 failed")
    if reduction == "none":
        loss = nll
    elif reduction == "sum":
        loss = nll.sum()
    else:
        loss = (nll / denom).mean()
    return loss, nll, grad
