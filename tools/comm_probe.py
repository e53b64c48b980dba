"""Developer tool (torchrun, N >= 2): where do the ~20 us of the sharded step's collective go?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from asr_chinese_e2e_b200 import ctc_loss_b200
from asr_chinese_e2e_b200.sharded import sharded_ctc_loss, _collective_stream
from oracle.synth import make_case

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
c = make_case(256, 400, 4234, 50, 1002 + rank)
x = c["logits"].cuda().requires_grad_(True)
tg, il, tl = c["targets"].cuda(), c["input_lengths"].cuda(), c["target_lengths"].cuda()
inv = 1.0 / 256
side = _collective_stream(x.device)
buf = torch.zeros(1, device="cuda")


def v_local():
    return ctc_loss_b200(x, tg, il, tl, inv_batch=inv)


def v_split():
    ev = torch.cuda.Event()
    return ctc_loss_b200(x, tg, il, tl, inv_batch=inv, lattice_event=ev)


def v_current():
    return sharded_ctc_loss(x, tg, il, tl)


def v_inline():
    loc = ctc_loss_b200(x, tg, il, tl, inv_batch=inv)
    tot = loc.detach().clone()
    dist.all_reduce(tot)
    return loc + (tot / world - loc.detach())


def v_split_side_nocombine():
    ev = torch.cuda.Event()
    loc = ctc_loss_b200(x, tg, il, tl, inv_batch=inv, lattice_event=ev)
    main = torch.cuda.current_stream()
    side.wait_event(ev)
    with torch.cuda.stream(side):
        tot = loc.detach().clone()
        dist.all_reduce(tot)
    main.wait_stream(side)
    return loc


def v_allreduce_only():
    dist.all_reduce(buf)
    return None


def timeit(f, n=200, bwd=True):
    def step():
        x.grad = None
        l = f()
        if bwd and l is not None:
            l.backward()
    for _ in range(10):
        step()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        step()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


for name, f in (("local (one call)", v_local), ("split call, no collective", v_split), ("current: split + side all-reduce + combine", v_current),
                ("one call + inline all-reduce + combine", v_inline), ("split + side all-reduce, no combine", v_split_side_nocombine),
                ("all-reduce of 1 float alone (back to back)", v_allreduce_only)):
    ms = timeit(f)
    if rank == 0:
        print(f"{name:50s} {ms * 1e3:8.1f} us/step", flush=True)
dist.barrier()
dist.destroy_process_group()
