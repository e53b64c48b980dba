"""Developer tool (GPU box): loss + gradient step, eager calls vs one captured CUDA graph replayed (C2 shape)."""
import sys
import torch
sys.path.insert(0, ".")
from asr_chinese_e2e_b200 import ctc_loss_b200
from oracle.synth import make_config

full = len(sys.argv) > 1 and sys.argv[1] == "full"
c = make_config("C2", full_lengths=full)
x = c["logits"].cuda().requires_grad_(True)
tg, il, tl = (c[k].cuda() for k in ("targets", "input_lengths", "target_lengths"))


def step():
    x.grad = None
    loss = ctc_loss_b200(x, tg, il, tl, reduction="mean", zero_infinity=False)
    loss.backward()
    return loss


def timeit(fn, n=200):
    for _ in range(10):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


eager = timeit(step)
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    step(); step()
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
x.grad = None
with torch.cuda.graph(g):
    loss = ctc_loss_b200(x, tg, il, tl, reduction="mean", zero_infinity=False)
    loss.backward()
graphed = timeit(g.replay)
print(f"lengths={'full' if full else 'var'}: eager {eager:.4f} ms/step, graph replay {graphed:.4f} ms/step, loss {loss.item():.5f}")
