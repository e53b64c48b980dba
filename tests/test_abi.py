"""CPU: the C-ABI shared library builds for sm_100a, loads, exports every symbol that
include/ctcb200.h declares, and validates arguments before touching CUDA."""
import ctypes
import os
import re

import pytest

from asr_chinese_e2e_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    _lib.build()
    return _lib.lib()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "ctcb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ctcb200_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    names = declared_symbols()
    assert {"ctcb200_forward", "ctcb200_backward", "ctcb200_loss_only", "ctcb200_workspace_bytes",
            "ctcb200_strerror", "ctcb200_version", "ctcb200_read_status", "ctcb200_read_lattice_stats"} <= set(names)
    for n in names:
        assert hasattr(lib, n), n
    assert set(_lib.SIGNATURES) == set(names)
    assert lib.ctcb200_version() == 100


def test_workspace_bytes_is_host_arithmetic(lib):
    # frames of 4+16*NS floats (NS=4 up to U=63) for lp_lab and gam, stored half-lattice stages of
    # {32 lane exponents, 8 rows of 32*NS doubles}, 1 int argmax per frame, 1 double per 8 frames
    b = _lib.workspace_bytes(256, 400, 4234, 50)
    want = 256 * 400 * (2 * 68 + 1) * 4 + 256 * 50 * (128 + 8 * 128 * 8) + 256 * 50 * 8
    assert want <= b <= want + 64 * 1024
    assert _lib.workspace_bytes(64, 1500, 4234, 120) > 64 * 1500 * (2 * 132 + 256) * 4
    out = ctypes.c_size_t(0)
    assert lib.ctcb200_workspace_bytes(1, 1, 5, 256, ctypes.byref(out)) == -4      # Umax > 255
    assert lib.ctcb200_workspace_bytes(1, 0, 5, 3, ctypes.byref(out)) == -2        # T < 1
    assert lib.ctcb200_workspace_bytes(1, 1, 5, 3, None) == -1


def test_argument_validation_precedes_cuda(lib):
    ws = 256 * 4096
    f = lib.ctcb200_forward
    ok = dict(logits=4096, targets=4096, ts=3, tn=3, il=4096, tl=4096, B=1, T=4, V=5, U=3, blank=0, zi=0,
              nll=4096, sums=None, ws=4096 * 256, wsb=1 << 30, stream=None)

    def call(**kw):
        a = dict(ok, **kw)
        return f(a["logits"], a["targets"], a["ts"], a["tn"], a["il"], a["tl"], a["B"], a["T"], a["V"], a["U"],
                 a["blank"], a["zi"], a["nll"], a["sums"], a["ws"], a["wsb"], a["stream"], None)

    assert call(logits=None) == -1
    assert call(V=1) == -2
    assert call(blank=5) == -3
    assert call(U=300) == -4
    assert call(logits=4100) == -5
    assert call(wsb=16) == -7
    assert lib.ctcb200_backward(4096, 4096, 3, 3, 4096, 0, 9, 1.0, 1, 4, 5, 3, 0, 0, 4096, 4096 * 256, 1 << 30,
                                None) == -6
    assert b"Umax" in lib.ctcb200_strerror(-4)
    del ws


def test_sass_is_blackwell_native():
    """The built library carries sm_100a SASS with bulk-TMA (UBLKCP) in all three kernels."""
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    _lib.build()
    out = subprocess.run(["cuobjdump", "-sass", _lib.SO_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    cur, has = None, {}
    for line in out.splitlines():
        if "Function :" in line:
            cur = line.split("Function :")[1].strip()
        elif "UBLKCP" in line and cur:
            has[cur] = True
    for k in ("k1_lse_gather", "k2_lattice", "k3_grad"):
        assert any(k in f for f in has), k
