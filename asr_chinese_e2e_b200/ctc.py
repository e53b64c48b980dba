"""Python boundary of the B200 CTC op: ``ctc_loss_b200`` / ``CTCLossB200``.

Mirrors ``torch.nn.functional.ctc_loss`` (blank / reduction / zero_infinity semantics, padded
2-D or concatenated 1-D targets, int32/int64 lengths) but takes the reference's batch-major
``[B, T, V]`` LOGITS (Predictor/Utils/loss.py:10 "pred: N x T x C", the convention of
``cal_performance``), i.e. it is

    F.ctc_loss(F.log_softmax(logits, -1).transpose(0, 1), targets, input_lengths,
               target_lengths, blank, reduction, zero_infinity)

differentiable w.r.t. ``logits`` only.  It plugs in next to ``cal_performance`` inside
``TransformerOffical.cal_metrics`` (Predictor/Models/transformer_official.py:83-94); see
``asr_chinese_e2e_b200.joint``.  All arithmetic runs in libctcb200.so (sm_100a CUDA); torch
provides device memory, the current stream and autograd plumbing only.
"""
from __future__ import annotations

import os

import torch

from . import _lib

_RED = {"none": 0, "mean": 1, "sum": 2}
_DEBUG = bool(int(os.environ.get("CTCB200_DEBUG", "0")))


def _as_i64_cuda(x, device):
    if not torch.is_tensor(x):
        x = torch.as_tensor(x)
    return x.to(device=device, dtype=torch.int64, non_blocking=True).contiguous()


def _prepare(logits, targets, input_lengths, target_lengths, blank, max_target_length):
    if not (torch.is_tensor(logits) and logits.is_cuda):
        raise _lib.CtcB200Error("ctc_loss_b200 needs CUDA logits: the hot path has no CPU fallback")
    if logits.dtype != torch.float32:
        raise _lib.CtcB200Error(f"logits must be float32 (got {logits.dtype}); the path computes in f32")
    if logits.dim() != 3:
        raise ValueError("logits must be [B, T, V] (batch-major, as in the reference's loss API)")
    x = logits.contiguous()
    if x.data_ptr() % 16:
        x = x.clone()
    B, T, V = x.shape
    dev = x.device
    il = _as_i64_cuda(input_lengths, dev)
    tl = _as_i64_cuda(target_lengths, dev)
    tg = _as_i64_cuda(targets, dev)
    if il.numel() != B or tl.numel() != B:
        raise ValueError("input_lengths and target_lengths must have B elements")
    if tg.dim() == 2:
        if tg.shape[0] != B:
            raise ValueError("2-D targets must be [B, Umax]")
        umax, stride = tg.shape[1], tg.shape[1]
        if umax > 255:   # wider padding than the lattice kernel supports: use the true maximum
            umax = int(max_target_length if max_target_length is not None else tl.max().item())
        if stride == 0:
            tg, stride = tg.new_zeros(B, 1), 1
    elif tg.dim() == 1:
        stride = 0
        if max_target_length is not None:
            umax = int(max_target_length)
        elif torch.is_tensor(target_lengths) and not target_lengths.is_cuda:
            umax = int(target_lengths.max()) if B else 0
        else:
            umax = int(tl.max().item()) if B else 0    # one host sync; pass max_target_length to avoid it
        if tg.numel() == 0:
            tg = tg.new_zeros(1)
    else:
        raise ValueError("targets must be [B, Umax] or 1-D concatenated")
    return x, tg, stride, il, tl, B, T, V, int(umax)


def _check_status(ws, stream):
    import ctypes
    st = ctypes.c_int(0)
    _lib.check(_lib.lib().ctcb200_read_status(ws.data_ptr(), ctypes.byref(st), stream), "ctcb200_read_status")
    if st.value:
        raise _lib.CtcB200Error(f"invalid CTC inputs, device status word = {st.value} "
                                "(1: input length, 2: target length, 4: label)")


class _CTCLossB200Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, targets, input_lengths, target_lengths, blank, reduction, zero_infinity,
                inv_batch, max_target_length):
        x, tg, stride, il, tl, B, T, V, umax = _prepare(logits, targets, input_lengths, target_lengths,
                                                        blank, max_target_length)
        L = _lib.lib()
        need_grad = ctx.needs_input_grad[0]
        with torch.cuda.device(x.device):
            stream = torch.cuda.current_stream().cuda_stream
            ws_bytes = _lib.workspace_bytes(B, T, V, umax)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
            nll = torch.empty(B, dtype=torch.float32, device=x.device)
            sums = torch.empty(3, dtype=torch.float32, device=x.device)
            fn = L.ctcb200_forward if need_grad else L.ctcb200_loss_only
            _lib.check(fn(x.data_ptr(), tg.data_ptr(), stride, tg.numel(), il.data_ptr(), tl.data_ptr(),
                          B, T, V, umax, int(blank), int(bool(zero_infinity)), nll.data_ptr(),
                          sums.data_ptr(), ws.data_ptr(), ws_bytes, stream),
                       "ctcb200_forward" if need_grad else "ctcb200_loss_only")
            if _DEBUG:
                _check_status(ws, stream)
        if B == 0:
            nll.zero_(); sums.zero_()
        ctx.inv_batch = float(inv_batch) if inv_batch is not None else (1.0 / max(B, 1))
        ctx.cfg = (stride, B, T, V, umax, int(blank), int(bool(zero_infinity)), _RED[reduction], ws_bytes)
        if need_grad:
            ctx.save_for_backward(x, tg, ws)
        if reduction == "none":
            return nll
        if reduction == "sum":
            return sums[1]
        return sums[0] * ctx.inv_batch

    @staticmethod
    def backward(ctx, grad_out):
        x, tg, ws = ctx.saved_tensors
        stride, B, T, V, umax, blank, zi, red, ws_bytes = ctx.cfg
        go = grad_out.to(dtype=torch.float32).contiguous()
        grad = torch.empty_like(x)
        with torch.cuda.device(x.device):
            stream = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().ctcb200_backward(
                x.data_ptr(), tg.data_ptr(), stride, tg.numel(), go.data_ptr(), 1 if red == 0 else 0, red,
                ctx.inv_batch, B, T, V, umax, blank, zi, grad.data_ptr(), ws.data_ptr(), ws_bytes, stream),
                "ctcb200_backward")
        return grad, None, None, None, None, None, None, None, None


def ctc_loss_b200(logits, targets, input_lengths, target_lengths, blank: int = 0,
                  reduction: str = "mean", zero_infinity: bool = False, *, inv_batch=None,
                  max_target_length=None):
    """CTC loss on batch-major logits; same flags and results as ``F.ctc_loss`` (see module doc).

    inv_batch: 1/(global batch) for 'mean' when the batch is sharded over ranks (default 1/B).
    max_target_length: upper bound on target_lengths for 1-D targets (avoids one host sync).
    """
    if reduction not in _RED:
        raise ValueError(f"reduction must be one of {list(_RED)}")
    return _CTCLossB200Fn.apply(logits, targets, input_lengths, target_lengths, blank, reduction,
                                zero_infinity, inv_batch, max_target_length)


class CTCLossB200(torch.nn.Module):
    """``nn.CTCLoss``-shaped module over ``ctc_loss_b200`` (batch-major logits in, not log-probs)."""

    def __init__(self, blank: int = 0, reduction: str = "mean", zero_infinity: bool = False):
        super().__init__()
        if reduction not in _RED:
            raise ValueError(f"reduction must be one of {list(_RED)}")
        self.blank, self.reduction, self.zero_infinity = blank, reduction, zero_infinity

    def forward(self, logits, targets, input_lengths, target_lengths):
        return ctc_loss_b200(logits, targets, input_lengths, target_lengths, self.blank, self.reduction,
                             self.zero_infinity)
