"""Developer tool (GPU box): fp32 error of this repo's gradient and of torch's CPU/CUDA ctc vs the float64 oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn.functional as F
from asr_chinese_e2e_b200 import ctc_loss_b200
from oracle.c_oracle import ctc_c_f64
from oracle.synth import make_case
from oracle.torch_ref import ref_ctc

for (B, T, V, U, seed, dist) in [(6, 284, 4234, 127, 1000, "D1"), (8, 400, 4234, 50, 5, "D1"), (8, 400, 4234, 50, 5, "D2"),
                                 (4, 1500, 4234, 120, 6, "D1")]:
    c = make_case(B, T, V, U, seed, dist=dist)
    a = [c[k].numpy() for k in ("logits", "targets", "input_lengths", "target_lengths")]
    _, n64, g64 = ctc_c_f64(*a, reduction="sum")
    for fused in (True, False):
        x = c["logits"].cuda().requires_grad_(True)
        nll = ctc_loss_b200(x, c["targets"].cuda(), c["input_lengths"].cuda(), c["target_lengths"].cuda(), reduction="none", fused=fused)
        nll.sum().backward()
        print(f"B={B} T={T} U={U} {dist} ours(fused={fused}): nll rel {np.abs(nll.detach().cpu().numpy()-n64).max()/np.abs(n64).max():.2e}  grad(sum) abs {np.abs(x.grad.cpu().numpy()-g64).max():.2e}")
    rl, rg = ref_ctc(c["logits"], c["targets"], c["input_lengths"], c["target_lengths"], reduction="sum")
    rn, _ = ref_ctc(c["logits"], c["targets"], c["input_lengths"], c["target_lengths"], reduction="none", want_grad=False)
    print(f"   torch CPU : nll rel {np.abs(rn.numpy()-n64).max()/np.abs(n64).max():.2e}  grad(sum) abs {np.abs(rg.numpy()-g64).max():.2e}")
    xc = c["logits"].cuda().requires_grad_(True)
    l = F.ctc_loss(F.log_softmax(xc, -1).transpose(0, 1), c["targets"].cuda(), c["input_lengths"].cuda(), c["target_lengths"].cuda(), reduction="none")
    l.sum().backward()
    print(f"   torch CUDA: nll rel {np.abs(l.detach().cpu().numpy()-n64).max()/np.abs(n64).max():.2e}  grad(sum) abs {np.abs(xc.grad.cpu().numpy()-g64).max():.2e}")
