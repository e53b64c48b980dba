"""Batch-sharded CTC loss for data-parallel training.

Utterances are independent, so each rank runs the kernels on its own shard with no data-path
collective.  The only exchange is ONE all-reduce of the 2-element buffer
``[sum_b nll_b / max(U_b,1), B_local]`` (NCCL over NVLink on GPUs, gloo in the CPU tests); the global
'mean' loss is ``buf[0] / buf[1]``.  The returned tensor has the GLOBAL value and a LOCAL gradient whose scale
depends on how the caller combines parameter gradients across ranks (``grad_reduce``):

  'mean' (default) -- ``DistributedDataParallel`` / ``ddp.DistributedWrapper`` AVERAGE parameter gradients, so each
                      rank's gradient is that of its LOCAL mean, 1 / (B_local * U_b) * (softmax - occupancy); the
                      average over ranks is then exactly the gradient of the global mean;
  'sum'            -- a hand-rolled all-reduce(SUM) of gradients: 1 / (B_global * U_b) * (softmax - occupancy).

With equal shards (the reference's
drop_last=True) 1/B_global is known before the kernels run, so the op's upstream gradient is exactly 1
and nothing waits on the collective; nothing ever syncs with the host.  The loss value is final after the
lattice kernel, so the all-reduce is issued on a side stream behind an event recorded there and runs under
the sparse gradient patch that follows on the main stream.

The reference has no multi-process path at all (its only multi-GPU construct is a disabled
nn.DataParallel wrapper, Predictor/Bases/base_model.py:9-21, main.py:80).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def combine_sharded_mean(local_sum: torch.Tensor, local_count: int, group=None) -> torch.Tensor:
    """General form (shards may differ in size).  local_sum: differentiable scalar
    sum_b nll_b/max(U_b,1) over this rank's shard.  No host sync: the count travels as a device scalar."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_sum / float(max(local_count, 1))
    buf = torch.empty(2, dtype=torch.float32, device=local_sum.device)
    buf[0] = local_sum.detach()
    buf[1].fill_(float(local_count))
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    inv = 1.0 / buf[1].clamp(min=1.0)
    local = local_sum * inv                                   # gradient flows through this term only
    return local + (buf[0] * inv - local.detach())            # value: global mean


class _GlobalValueLocalGrad(torch.autograd.Function):
    """value: the all-reduced global mean (`total`, returned as is -- no kernel); gradient: d/d local = 1.
    Replaces `local + (total - local.detach())`, three elementwise launches (~9 us per step measured on B200) that sat
    on the critical path of every sharded step."""

    @staticmethod
    def forward(ctx, local, total):
        return total.view_as(local)

    @staticmethod
    def backward(ctx, g):
        return g, None


_SIDE = {}


def _collective_stream(dev):
    key = (dev.type, dev.index)
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(dev, priority=-1)
    return _SIDE[key]


def combine_equal_shards(local_contrib: torch.Tensor, group=None, ready_event=None, value_scale: float = 1.0) -> torch.Tensor:
    """Equal shard sizes (drop_last=True, as the reference's loader: data/data_loader/ai_shell_1.py:103):
    value_scale * sum over ranks of local_contrib is the global mean (value_scale = 1 when local_contrib is already
    this rank's share sum_b nll_b/U_b / B_global, 1/world when it is the local mean), so the upstream gradient of
    the op stays exactly 1 (no rescale sweep) and the collective is one all-reduce of a single float, enqueued
    behind the kernels with no host sync."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_contrib
    world = dist.get_world_size(group)
    # one collective, nothing else: the division by `world` (grad_reduce='mean') is NCCL's own AVG
    avg = abs(value_scale * world - 1.0) < 1e-12 and local_contrib.is_cuda
    op = dist.ReduceOp.AVG if avg else dist.ReduceOp.SUM
    if ready_event is not None and local_contrib.is_cuda:
        # The value is final at `ready_event` (after the lattice kernel) while the current stream still has the
        # sparse gradient patch queued: run the all-reduce on a side stream behind that event, so that its
        # latency (tens of microseconds for one float over NVLink) hides under the patch kernel.
        main = torch.cuda.current_stream(local_contrib.device)
        side = _collective_stream(local_contrib.device)
        side.wait_event(ready_event)
        with torch.cuda.stream(side):
            tot = local_contrib.detach().clone()
            dist.all_reduce(tot, op=op, group=group)
        tot.record_stream(main)
        main.wait_stream(side)
    else:
        tot = local_contrib.detach().clone()
        dist.all_reduce(tot, op=op, group=group)
    if not avg and value_scale != 1.0:
        tot = tot * value_scale
    return _GlobalValueLocalGrad.apply(local_contrib, tot)    # value: global mean; gradient: d/d local = 1


def sharded_ctc_loss(logits, targets, input_lengths, target_lengths, blank: int = 0,
                     zero_infinity: bool = False, group=None, max_target_length=None, equal_shards: bool = True,
                     grad_reduce: str = "mean"):
    """'mean'-reduced CTC loss over the GLOBAL batch; call on every rank with its own shard.

    grad_reduce: how the caller combines parameter gradients across ranks -- 'mean' (DistributedDataParallel,
    ``DistributedWrapper``; the default) or 'sum' (see the module doc).  The returned VALUE is the global mean
    either way."""
    from .ctc import ctc_loss_b200
    if grad_reduce not in ("mean", "sum"):
        raise ValueError("grad_reduce must be 'mean' or 'sum'")
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    B = logits.shape[0]
    if equal_shards:
        ev = torch.cuda.Event() if (world > 1 and logits.is_cuda) else None
        gw = world if grad_reduce == "sum" else 1              # gradient normaliser: B_global or B_local
        local = ctc_loss_b200(logits, targets, input_lengths, target_lengths, blank=blank, reduction="mean",
                              zero_infinity=zero_infinity, inv_batch=1.0 / max(gw * B, 1),
                              max_target_length=max_target_length, lattice_event=ev)
        return combine_equal_shards(local, group, ready_event=ev, value_scale=float(gw) / world)
    local_sum = ctc_loss_b200(logits, targets, input_lengths, target_lengths, blank=blank,
                              reduction="mean", zero_infinity=zero_infinity, inv_batch=1.0,
                              max_target_length=max_target_length)
    out = combine_sharded_mean(local_sum, B, group)
    if grad_reduce == "mean" and world > 1:                    # same value, gradient scaled for an averaging reducer
        out = out.detach() + (out - out.detach()) * world
    return out
