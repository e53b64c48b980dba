"""CPU, world_size 2, gloo: the one collective of the multi-GPU path.  Each rank holds
half of the batch; combine_sharded_mean must return the un-sharded 'mean' loss on every rank and a
gradient equal to that rank's slab of the un-sharded gradient."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from oracle.synth import make_case


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from asr_chinese_e2e_b200.sharded import combine_equal_shards, combine_sharded_mean
    c = make_case(6, 24, 19, 5, 99)
    h = 3
    sl = slice(rank * h, (rank + 1) * h)
    x = c["logits"][sl].clone().requires_grad_(True)
    nll = F.ctc_loss(F.log_softmax(x, -1).transpose(0, 1), c["targets"][sl], c["input_lengths"][sl],
                     c["target_lengths"][sl], reduction="none")       # oracle stands in for the CUDA op
    local_sum = (nll / c["target_lengths"][sl].clamp(min=1)).sum()
    loss = combine_sharded_mean(local_sum, h)
    loss.backward(retain_graph=True)
    g_general = x.grad.clone()
    x.grad = None
    loss2 = combine_equal_shards(local_sum / (world * h))     # equal shards: 1/B_global folded in up front
    loss2.backward()
    assert abs(loss2.item() - loss.item()) < 1e-5 * abs(loss.item())
    assert torch.allclose(x.grad, g_general, atol=1e-7)
    q.put((rank, loss.item(), x.grad.clone()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_mean_matches_unsharded():
    c = make_case(6, 24, 19, 5, 99)
    x = c["logits"].clone().requires_grad_(True)
    full = F.ctc_loss(F.log_softmax(x, -1).transpose(0, 1), c["targets"], c["input_lengths"],
                      c["target_lengths"], reduction="mean")
    full.backward()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda r: r[0])
    [p.join(60) for p in procs]
    for rank, loss, grad in res:
        assert abs(loss - full.item()) < 1e-6 * abs(full.item())
        assert torch.allclose(grad, x.grad[rank * 3:(rank + 1) * 3], atol=1e-7)


def test_single_process_is_plain_mean():
    from asr_chinese_e2e_b200.sharded import combine_sharded_mean
    s = torch.tensor(6.0, requires_grad=True)
    out = combine_sharded_mean(s, 4)
    out.backward()
    assert out.item() == 1.5 and s.grad.item() == 0.25
