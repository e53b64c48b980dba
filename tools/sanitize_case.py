"""Developer tool (GPU box): small cases through every kernel variant, for compute-sanitizer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asr_chinese_e2e_b200 import ctc_loss_b200
from oracle.synth import make_case

# odd V / T -> k1_lse_gather; even V and T -> k1p_sweep (groups of two frames, incl. an odd number of valid frames)
for (B, T, V, U, kw) in [(5, 23, 37, 6, {}), (3, 17, 4234, 9, {}), (4, 160, 131, 70, {}), (3, 290, 67, 130, {}),
                         (5, 24, 38, 6, {}), (3, 18, 4234, 9, {}), (4, 160, 130, 70, {}), (2, 40, 64, 7, {})]:
    c = make_case(B, T, V, U, 11, dist="D2", n_infeasible=1)
    for fused in (True, False):
        x = c["logits"].cuda().requires_grad_(True)
        loss = ctc_loss_b200(x, c["targets"].cuda(), c["input_lengths"].cuda(), c["target_lengths"].cuda(),
                             reduction="mean", zero_infinity=True, fused=fused, chunks=(2 if fused else 1), **kw)
        loss.backward()
        with torch.no_grad():
            l2 = ctc_loss_b200(x.detach(), c["targets"].cuda(), c["input_lengths"].cuda(), c["target_lengths"].cuda(),
                               reduction="sum", zero_infinity=True)
    x = c["logits"].cuda().requires_grad_(True)
    ctc_loss_b200(x, c["targets"].cuda(), c["input_lengths"].cuda(), c["target_lengths"].cuda()).backward()
    torch.cuda.synchronize()
    print(B, T, V, U, "ok", float(loss), float(l2))
