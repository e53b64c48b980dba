"""Attention-branch cross-entropy on the B200 sweep kernels: ``attention_ce_b200``.

Same contract as the reference's ``cal_loss`` (Predictor/Utils/loss.py:26-51) / ``calculate_loss``
(:54-76): ``pred`` logits ``[N, T, C]`` (or ``[rows, C]``), ``gold`` int64 with PAD = 0 ignored,
optional label smoothing, scalar mean over the non-pad tokens.  One sweep computes the loss and
(when ``pred`` requires grad) the gradient, speculatively for an upstream gradient of 1; ``weight``
(e.g. ``1 - ctc_weight``) is folded in so that the joint loss needs no rescaling sweep.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib

IGNORE_ID = 0


class _CEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, gold, smoothing, weight, ignore_index):
        if not (pred.is_cuda and pred.dtype == torch.float32):
            raise _lib.CtcB200Error("attention_ce_b200 needs CUDA float32 logits: there is no CPU fallback")
        V = pred.shape[-1]
        x = pred.reshape(-1, V).contiguous()
        if x.data_ptr() % 16:
            x = x.clone()
        g = gold.reshape(-1).to(device=x.device, dtype=torch.int64).contiguous()
        rows = x.shape[0]
        if g.numel() != rows:
            raise ValueError("gold must have one entry per row of pred")
        L = _lib.lib()
        need_grad = ctx.needs_input_grad[0]
        n = ctypes.c_size_t(0)
        _lib.check(L.ctcb200_ce_workspace_bytes(rows, ctypes.byref(n)), "ctcb200_ce_workspace_bytes")
        ws = torch.empty(n.value, dtype=torch.uint8, device=x.device)
        out = torch.zeros(2, dtype=torch.float32, device=x.device)
        grad = torch.empty_like(x) if need_grad else None
        with torch.cuda.device(x.device):
            _lib.check(L.ctcb200_ce_loss_grad(x.data_ptr(), g.data_ptr(), rows, V, int(ignore_index), float(smoothing),
                                              float(weight), out.data_ptr(), grad.data_ptr() if need_grad else None,
                                              ws.data_ptr(), n.value, torch.cuda.current_stream().cuda_stream),
                       "ctcb200_ce_loss_grad")
        ctx.shape = pred.shape
        if need_grad:
            ctx.applied = torch.ones(1, dtype=torch.float32, device=x.device)
            ctx.save_for_backward(grad)
        return out[0]

    @staticmethod
    def backward(ctx, grad_out):
        (grad,) = ctx.saved_tensors
        rows, V = grad.shape
        go = grad_out.to(torch.float32).contiguous()
        with torch.cuda.device(grad.device):
            new = torch.empty_like(ctx.applied)
            _lib.check(_lib.lib().ctcb200_rescale_grad(grad.data_ptr(), go.data_ptr(), 0, ctx.applied.data_ptr(),
                                                       new.data_ptr(), 1, rows, V,
                                                       torch.cuda.current_stream().cuda_stream), "ctcb200_rescale_grad")
            ctx.applied = new
        return grad.view(ctx.shape), None, None, None, None


def attention_ce_b200(pred, gold, smoothing: float = 0.0, weight: float = 1.0, ignore_index: int = IGNORE_ID):
    """weight * cross-entropy(pred, gold) with PAD rows ignored and optional label smoothing."""
    return _CEFn.apply(pred, gold, float(smoothing), float(weight), int(ignore_index))
