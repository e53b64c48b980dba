"""Batch-sharded CTC loss for data-parallel training.

Utterances are independent, so each rank runs the kernels on its own shard with no data-path
collective.  The only exchange is ONE all-reduce of the 2-element buffer
``[sum_b nll_b / max(U_b,1), B_local]`` (NCCL over NVLink on GPUs, gloo in the CPU tests); the global
'mean' loss is ``buf[0] / buf[1]``.  The returned tensor has the GLOBAL value and a LOCAL gradient:
d loss / d logits_local = 1 / (B_global * U_b) * (softmax - occupancy), which is what a DDP
all-reduce of parameter gradients (sum) then expects.  The division by B_global stays on the
device (no host sync), so the gradient sweep is enqueued right behind the collective.

The reference has no multi-process path at all (its only multi-GPU construct is a disabled
nn.DataParallel wrapper, Predictor/Bases/base_model.py:9-21, main.py:80).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def combine_sharded_mean(local_sum: torch.Tensor, local_count: int, group=None) -> torch.Tensor:
    """local_sum: differentiable scalar sum_b nll_b/max(U_b,1) over this rank's shard."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_sum / float(max(local_count, 1))
    buf = torch.stack([local_sum.detach().to(torch.float32),
                       torch.tensor(float(local_count), device=local_sum.device)])
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    inv = 1.0 / buf[1].clamp(min=1.0)
    local = local_sum * inv                                   # gradient flows through this term only
    return local + (buf[0] * inv - local.detach())            # value: global mean


def sharded_ctc_loss(logits, targets, input_lengths, target_lengths, blank: int = 0,
                     zero_infinity: bool = False, group=None, max_target_length=None):
    """'mean'-reduced CTC loss over the GLOBAL batch; call on every rank with its own shard."""
    from .ctc import ctc_loss_b200
    local_sum = ctc_loss_b200(logits, targets, input_lengths, target_lengths, blank=blank,
                              reduction="mean", zero_infinity=zero_infinity, inv_batch=1.0,
                              max_target_length=max_target_length)
    return combine_sharded_mean(local_sum, logits.shape[0], group)
