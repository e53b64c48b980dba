"""The parity target: torch's own CPU CTC path.  TEST INFRASTRUCTURE ONLY.

``ref_ctc`` is exactly the expression BASELINE.json / SURVEY.md section 8b define the
drop-in op by:

    F.ctc_loss(F.log_softmax(logits, -1).transpose(0, 1),
               targets, input_lengths, target_lengths, blank, reduction, zero_infinity)

with batch-major ``[B, T, V]`` logits (the reference's convention,
Predictor/Utils/loss.py:10) and the gradient taken w.r.t. the logits.  The
reference has no call site of this op; it is the third-party op
the north star names.  It is also what ``bench.py --impl reference`` and the
``cpu_baseline`` leg time on the GPU box's host cores.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def ref_ctc(logits, targets, input_lengths, target_lengths, blank=0,
            reduction="mean", zero_infinity=False, want_grad=True, grad_output=None):
    """CPU fp32.  Returns (loss, grad_logits or None)."""
    x = logits.detach().to("cpu", torch.float32).clone().requires_grad_(want_grad)
    lp = F.log_softmax(x, dim=-1).transpose(0, 1)
    loss = F.ctc_loss(lp, targets.cpu(), input_lengths.cpu(), target_lengths.cpu(),
                      blank=blank, reduction=reduction, zero_infinity=zero_infinity)
    grad = None
    if want_grad:
        if grad_output is None:
            grad_output = torch.ones_like(loss)
        loss.backward(grad_output.to(loss))
        grad = x.grad
    return loss.detach(), grad


def ref_step(x, targets, input_lengths, target_lengths, zero_infinity=False):
    """One timed reference step: loss + backward on a leaf that already requires grad."""
    x.grad = None
    loss = F.ctc_loss(F.log_softmax(x, dim=-1).transpose(0, 1), targets, input_lengths,
                      target_lengths, blank=0, reduction="mean", zero_infinity=zero_infinity)
    loss.backward()
    return loss
