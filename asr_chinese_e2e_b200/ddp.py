"""One-process-per-GPU data parallelism for the reference's model API.

Replaces ``Predictor/Bases/base_model.py:9-21`` (``Wrapper`` around ``nn.DataParallel``, commented out at
``main.py:80``): same constructor idea and the same attribute pass-through to the wrapped model, but every rank
owns one GPU and its own DataLoader shard, gradients are averaged by DDP's bucketed NCCL all-reduce during
``loss.backward()``.  Because DDP AVERAGES, every rank's loss is normalised by its LOCAL batch:
``JointCTCAttention.joint_loss`` passes inv_batch = w / B_local, and ``sharded_ctc_loss(grad_reduce='mean')``
(the default) does the same while returning the global value.  ``torch.distributed`` is plumbing here, not a
kernel of this repo.
"""
import os

import torch
import torch.distributed as dist
from torch.nn.parallel import DistributedDataParallel

__all__ = ["init_from_env", "DistributedWrapper", "shard_batch"]


def init_from_env(backend=None):
    """Join the process group described by torchrun's environment (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*).
    Returns (rank, world_size, device).  NCCL when CUDA is present, gloo otherwise (CPU tests)."""
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cuda = torch.cuda.is_available()
    if cuda:
        torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend or ("nccl" if cuda else "gloo"), rank=rank, world_size=world)
    return rank, world, torch.device("cuda", local) if cuda else torch.device("cpu")


class DistributedWrapper(torch.nn.Module):
    """``Wrapper(model, device_ids)`` of the reference, one process per GPU.

    ``forward`` goes through DDP (so the backward all-reduces gradients); any other attribute -- ``iterate``,
    ``cal_metrics``, ``save``, ``load``, ``vocab`` ... -- resolves on the wrapped model exactly like the
    reference's ``__getattr__`` pass-through, which is what lets ``Trainer11`` drive it unchanged.  ``iterate``
    is re-bound so that its ``self(...)`` call runs through DDP rather than the bare module."""

    def __init__(self, model, device=None, **ddp_kwargs):
        super().__init__()
        if device is not None:
            model = model.to(device)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            ids = [device.index] if device is not None and device.type == "cuda" else None
            self.model = DistributedDataParallel(model, device_ids=ids, **ddp_kwargs)
        else:
            self.model = model

    @property
    def module(self):
        return self.model.module if isinstance(self.model, DistributedDataParallel) else self.model

    def forward(self, *inputs, **kw):
        return self.model(*inputs, **kw)

    def iterate(self, *args, **kw):
        fn = type(self.module).iterate
        return fn(self, *args, **kw)        # the model's own iterate, with self(...) routed through DDP

    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(self.module, name)


def shard_batch(batch, rank, world_size):
    """This rank's contiguous slice of every batch-major tensor in a dict / Pack (equal shards, like the
    reference's ``drop_last=True`` loader would hand each rank).  The container type is preserved (a ``Pack`` stays
    a ``Pack``); a batch size that is not a multiple of world_size is an error rather than a silent drop."""
    out = type(batch)()
    for k, v in batch.items():
        if torch.is_tensor(v) and v.dim() >= 1:
            if v.shape[0] % world_size:
                raise ValueError(f"batch dimension {v.shape[0]} of '{k}' is not a multiple of world_size {world_size} "
                                 "(the reference's loader uses drop_last=True)")
            per = v.shape[0] // world_size
            v = v[rank * per:(rank + 1) * per]
        out[k] = v
    return out
