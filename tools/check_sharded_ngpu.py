"""Developer tool (GPU box, torchrun): sharded_ctc_loss on N ranks against the un-sharded op on one GPU.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_sharded_ngpu.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from asr_chinese_e2e_b200 import ctc_loss_b200, sharded_ctc_loss
from oracle.synth import make_case

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
per = 16
c = make_case(per * world, 200, 4234, 40, 4321)
sl = slice(rank * per, (rank + 1) * per)
for it in range(3):
    x = c["logits"][sl].cuda().requires_grad_(True)
    loss = sharded_ctc_loss(x, c["targets"][sl].cuda(), c["input_lengths"][sl].cuda(), c["target_lengths"][sl].cuda(),
                            grad_reduce="sum")       # local gradient = this rank's slab of the global-mean gradient
    loss.backward()
xf = c["logits"].cuda().requires_grad_(True)
full = ctc_loss_b200(xf, c["targets"].cuda(), c["input_lengths"].cuda(), c["target_lengths"].cuda(), reduction="mean")
full.backward()
torch.cuda.synchronize()
dl = abs(loss.item() - full.item()) / abs(full.item())
dg = (x.grad - xf.grad[sl]).abs().max().item()
print(f"rank {rank}: sharded {loss.item():.6f} full {full.item():.6f} rel {dl:.2e} grad max abs diff {dg:.2e}")
assert dl < 1e-6 and dg < 1e-7
# DDP semantics: with grad_reduce='mean' every rank's gradient is world x larger (DDP averages them afterwards)
x2 = c["logits"][sl].cuda().requires_grad_(True)
l2 = sharded_ctc_loss(x2, c["targets"][sl].cuda(), c["input_lengths"][sl].cuda(), c["target_lengths"][sl].cuda())
l2.backward()
assert abs(l2.item() - full.item()) < 1e-6 * abs(full.item())
assert (x2.grad - world * xf.grad[sl]).abs().max().item() < 1e-6
dist.barrier()
dist.destroy_process_group()
