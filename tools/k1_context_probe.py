"""Developer tool (GPU box): how long does the fused sweep kernel take depending on what ran before it?
 (a) back to back with itself, (b) after an idle gap, (c) inside the full step (prep | sweep | lattice + patch),
 (d) inside the step but with an L2-sized memset between the patch and the next sweep."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle.synth import make_config, make_lengths
from asr_chinese_e2e_b200 import _lib
from asr_chinese_e2e_b200.ctc import _prepare, _RED

full = len(sys.argv) > 1 and sys.argv[1] == "full"
c = make_config("C2", dist="D1")
if full:
    c["input_lengths"][:] = c["logits"].shape[1]
x = c["logits"].cuda()
x, tg, stride, il, tl, B, T, V, umax = _prepare(x, c["targets"].cuda(), c["input_lengths"].cuda(), c["target_lengths"].cuda(), 0, None)
L = _lib.lib()
wsb = _lib.workspace_bytes(B, T, V, umax)
ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
nll = torch.empty(B, device="cuda"); sums = torch.zeros(4, device="cuda"); grad = torch.empty_like(x)
junk = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream()


def call(stages):
    _lib.check(L.ctcb200_loss_grad_stages(stages, x.data_ptr(), tg.data_ptr(), stride, tg.numel(), il.data_ptr(), tl.data_ptr(),
                                          B, T, V, umax, 0, 0, 1, 1.0 / B, nll.data_ptr(), sums.data_ptr(), grad.data_ptr(),
                                          ws.data_ptr(), wsb, st.cuda_stream), "stages")


def timed(pre, post, n=30, gap=0.0):
    out = []
    for i in range(n + 5):
        pre()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); call(16); e1.record(st)
        post()
        if gap:
            torch.cuda.synchronize(); time.sleep(gap)
        e1.synchronize()
        if i >= 5:
            out.append(e0.elapsed_time(e1))
    torch.cuda.synchronize()
    out.sort()
    return out[len(out) // 2], out[0]


call(8); call(16); call(6); torch.cuda.synchronize()
nop = lambda: None
print("lengths:", "full" if full else "var")
print("(a) sweep back to back with itself        median %.4f  min %.4f ms" % timed(nop, nop))
print("(b) sweep after a 2 ms idle gap            median %.4f  min %.4f ms" % timed(nop, nop, gap=0.002))
print("(c) inside the step (prep|sweep|lat+patch) median %.4f  min %.4f ms" % timed(lambda: call(8), lambda: call(6)))
print("(c2) step, idle gap after every step       median %.4f  min %.4f ms" % timed(lambda: call(8), lambda: call(6), gap=0.002))
print("(d) step + 256 MB memset before the sweep  median %.4f  min %.4f ms" % timed(lambda: (junk.zero_(), call(8)), lambda: call(6)))
print("(e) sweep, lattice only after it           median %.4f  min %.4f ms" % timed(lambda: call(8), lambda: call(2)))
# reference point: torch device-to-device copy of the same bytes (read B*T*V*4, write the same)
y = torch.empty_like(x)
for _ in range(3): y.copy_(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): y.copy_(x)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print("torch copy of [B,T,V] (2 x %.3f GB): %.4f ms -> %.1f GB/s" % (x.numel() * 4 / 1e9, ms, 2 * x.numel() * 4 / ms / 1e6))
