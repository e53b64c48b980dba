"""Per-stage device timing of the two-sweep loss+grad call, for bench.py's roofline line.

``time_stages``: one un-chunked ``ctcb200_loss_grad`` call on the current stream, bracketed by CUDA events on that
stream; the library's own ``sweep_done`` event (recorded right after the fused sweep kernel) splits the
call into  [k0_prep + the fused sweep kernel]  and  [k2_lattice + k3p_patch].
``time_sweep_kernel``: the dominant kernel by itself -- the stage-split entry point launches k0_prep (stage 8) and the
fused sweep (stage 16) separately with an event in between, so the timed interval holds the sweep kernel (k1p_sweep<FUSED>, or k1_lse_gather<FUSED> for odd V / T) alone.
"""
from __future__ import annotations

import torch

from . import _lib
from .ctc import _RED, _prepare


def time_stages(logits, targets, input_lengths, target_lengths, blank=0, reduction="mean",
                zero_infinity=False, iters=10, warmup=3):
    """Returns dict(sweep_ms=[...], rest_ms=[...]) with one entry per timed iteration, and lattice_stats =
    [utterances that ran the log-space recursion, of which: underflowed in the linear domain] of the last call."""
    x, tg, stride, il, tl, B, T, V, umax = _prepare(logits.detach(), targets, input_lengths, target_lengths,
                                                    blank, None)
    L = _lib.lib()
    ws_bytes = _lib.workspace_bytes(B, T, V, umax)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
    nll = torch.empty(B, device=x.device)
    sums = torch.zeros(4, device=x.device)
    grad = torch.empty_like(x)
    import ctypes
    out = {"sweep_ms": [], "rest_ms": [], "lattice_stats": [0, 0]}
    with torch.cuda.device(x.device):
        st = torch.cuda.current_stream()
        for i in range(warmup + iters):
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e1.record(st)                                   # materialise the handle; re-recorded by the library
            e0.record(st)
            _lib.check(L.ctcb200_loss_grad(x.data_ptr(), tg.data_ptr(), stride, tg.numel(), il.data_ptr(),
                                           tl.data_ptr(), B, T, V, umax, int(blank), int(bool(zero_infinity)),
                                           _RED[reduction], 1.0 / max(B, 1), nll.data_ptr(), sums.data_ptr(),
                                           grad.data_ptr(), ws.data_ptr(), ws_bytes, st.cuda_stream, e1.cuda_event),
                       "ctcb200_loss_grad")
            e2.record(st)
            e2.synchronize()
            if i >= warmup:
                out["sweep_ms"].append(e0.elapsed_time(e1))
                out["rest_ms"].append(e1.elapsed_time(e2))
        stats = (ctypes.c_int * 2)()
        _lib.check(L.ctcb200_read_lattice_stats(ws.data_ptr(), stats, st.cuda_stream), "ctcb200_read_lattice_stats")
        out["lattice_stats"] = [int(stats[0]), int(stats[1])]
    return out


def time_sweep_kernel(logits, targets, input_lengths, target_lengths, blank=0, reduction="mean", zero_infinity=False,
                      iters=10, warmup=3):
    """Average duration (ms) of the fused sweep kernel alone, CUDA events on the launching stream."""
    x, tg, stride, il, tl, B, T, V, umax = _prepare(logits.detach(), targets, input_lengths, target_lengths, blank, None)
    L = _lib.lib()
    ws_bytes = _lib.workspace_bytes(B, T, V, umax)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
    nll = torch.empty(B, device=x.device)
    sums = torch.zeros(4, device=x.device)
    grad = torch.empty_like(x)
    out = []
    with torch.cuda.device(x.device):
        st = torch.cuda.current_stream()

        def call(stages):
            _lib.check(L.ctcb200_loss_grad_stages(stages, x.data_ptr(), tg.data_ptr(), stride, tg.numel(), il.data_ptr(),
                                                  tl.data_ptr(), B, T, V, umax, int(blank), int(bool(zero_infinity)),
                                                  _RED[reduction], 1.0 / max(B, 1), nll.data_ptr(), sums.data_ptr(),
                                                  grad.data_ptr(), ws.data_ptr(), ws_bytes, st.cuda_stream),
                       "ctcb200_loss_grad_stages")
        for i in range(warmup + iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            call(8)
            e0.record(st)
            call(16)
            e1.record(st)
            call(2 | 4)
            e1.synchronize()
            if i >= warmup:
                out.append(e0.elapsed_time(e1))
        torch.cuda.synchronize()
    return sum(out) / len(out)
