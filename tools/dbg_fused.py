import torch, sys
sys.path.insert(0,'.')
from asr_chinese_e2e_b200 import ctc_loss_b200
from oracle.synth import make_case
c = make_case(13, 70, 4234 // 4, 17, 555, dist="D2", n_infeasible=1, n_partial=1)
tg, il, tl = c["targets"].cuda(), c["input_lengths"].cuda(), c["target_lengths"].cuda()
def run(**kw):
    x = c["logits"].cuda().requires_grad_(True)
    loss = ctc_loss_b200(x, tg, il, tl, reduction="mean", zero_infinity=True, **kw)
    loss.backward()
    return loss.detach(), x.grad
l0,g0 = run(fused=False, chunks=1)
lf,gf = run(fused=True, chunks=1)
d=(g0-gf).abs()
print('max abs diff', d.max().item(), 'max |g|', g0.abs().max().item(), 'rel at max', (d/(g0.abs()+1e-12)).max().item())
i=d.argmax(); print(g0.flatten()[i].item(), gf.flatten()[i].item())
