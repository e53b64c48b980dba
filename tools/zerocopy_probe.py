"""Developer tool (GPU box): can the fused sweep read its logits straight from pinned HOST memory (zero-copy over PCIe)?
Times ctcb200_loss_grad on a 32-utterance chunk with the logits pointer in device memory vs in pinned host memory."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle.synth import make_config
from asr_chinese_e2e_b200 import _lib

c = make_config("C2", dist="D1")
B, T, V = c["logits"].shape
U = c["targets"].shape[1]
n = 32
h_x = c["logits"][:n].contiguous().pin_memory()
d_x = h_x.cuda()
tg, il, tl = c["targets"][:n].cuda(), c["input_lengths"][:n].cuda(), c["target_lengths"][:n].cuda()
L = _lib.lib()
wsb = _lib.workspace_bytes(n, T, V, U)
ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
nll = torch.empty(n, device="cuda")
grad = torch.empty(n, T, V, device="cuda")
st = torch.cuda.current_stream().cuda_stream
valid_bytes = int(c["input_lengths"][:n].sum()) * V * 4


def call(xptr):
    rc = L.ctcb200_loss_grad(xptr, tg.data_ptr(), U, tg.numel(), il.data_ptr(), tl.data_ptr(), n, T, V, U, 0, 0, 1, 1.0 / B,
                             nll.data_ptr(), None, grad.data_ptr(), ws.data_ptr(), wsb, st, None)
    assert rc == 0, _lib.strerror(rc)


res = {}
for name, ptr in (("device logits", d_x.data_ptr()), ("pinned host logits (zero-copy)", h_x.data_ptr())):
    for _ in range(2):
        call(ptr)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        call(ptr)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    res[name] = (nll.clone(), grad.clone())
    print(f"{name}: {ms:.3f} ms per 32-utterance call; valid logits {valid_bytes / 1e6:.1f} MB -> {valid_bytes / ms / 1e6:.1f} GB/s", flush=True)
a, b = res["device logits"], res["pinned host logits (zero-copy)"]
print("bit-identical:", torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]))
