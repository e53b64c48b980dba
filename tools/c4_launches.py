"""Developer tool: a few loss+grad steps on BASELINE config C4 (B=64, T=1500, U<=120, zero_infinity) for an ncu launch list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle.synth import make_config
from asr_chinese_e2e_b200 import ctc_loss_b200
c = make_config(sys.argv[1] if len(sys.argv) > 1 else "C4", dist="D1")
x = c["logits"].cuda().requires_grad_(True)
tg, il, tl = c["targets"].cuda(), c["input_lengths"].cuda(), c["target_lengths"].cuda()
for _ in range(4):
    x.grad = None
    loss = ctc_loss_b200(x, tg, il, tl, reduction="mean", zero_infinity=True)
    loss.backward()
torch.cuda.synchronize()
print("loss", loss.item())
