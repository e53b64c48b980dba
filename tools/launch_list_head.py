"""Developer tool: one fused-head evaluation call and one training call at the C2 shape (for an ncu launch list)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asr_chinese_e2e_b200 import ctc_head_loss_b200
from oracle.synth import make_lengths, make_targets
B, T, K, V, U = 256, 400, 512, 4234, 50
g = torch.Generator().manual_seed(1)
tg, tl = make_targets(B, U, V, g); il = make_lengths(B, T, g)
enc = torch.randn(B, T, K, generator=g).cuda().requires_grad_(True)
W = (torch.randn(V, K, generator=g) / K ** 0.5).cuda().requires_grad_(True)
b = torch.zeros(V).cuda().requires_grad_(True)
tg, il, tl = tg.cuda(), il.cuda(), tl.cuda()
prec = sys.argv[1] if len(sys.argv) > 1 else "3xtf32"
for it in range(3):
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    with torch.no_grad():
        l = ctc_head_loss_b200(enc, W, b, tg, il, tl, zero_infinity=True, precision=prec)
    e1.record()
    l2 = ctc_head_loss_b200(enc, W, b, tg, il, tl, zero_infinity=True, precision=prec)
    l2.backward()
    e2.record()
    torch.cuda.synchronize()
    print(f"iter {it}: eval {e0.elapsed_time(e1):.3f} ms, train {e1.elapsed_time(e2):.3f} ms, loss {l.item():.4f}")
