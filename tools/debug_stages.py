"""Developer tool (GPU box): run the C-ABI stage by stage on a small case and report the error of
every intermediate (lp_lab frames, nll, occupancies, gradient) against the CPU oracle."""
import ctypes
import math
import sys
import os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from asr_chinese_e2e_b200 import _lib  # noqa: E402
from oracle.synth import make_case  # noqa: E402
from oracle.torch_ref import ref_ctc  # noqa: E402

LOG2E = 1.4426950408889634


def align(x, a=256):
    return (x + a - 1) // a * a


def layout(B, T, Umax):
    NS = 4 if Umax <= 63 else (8 if Umax <= 127 else 16)
    Lp, Sp = 4 + 16 * NS, 32 * NS
    o, off = 0, {}
    for name, n in (("hdr", 256), ("Tb", 4 * B), ("Ub", 4 * B), ("flags", 4 * B), ("toff", 8 * B),
                    ("rowstart", 4 * (B + 1)), ("lp_lab", 4 * B * T * Lp), ("gam", 4 * B * T * Lp),
                    ("ab", 4 * B * T * Sp), ("best", 4 * B * T), ("tile_off", 8 * B * ((T + 7) // 8))):
        off[name] = o
        o += align(n)
    return NS, Lp, Sp, off, o


def main(B=5, T=37, V=53, Umax=9, seed=1, zi=0, dist="D1"):
    c = make_case(B, T, V, Umax, seed, dist=dist, n_infeasible=1 if B > 2 else 0, n_partial=1 if B > 3 else 0)
    L = _lib.lib()
    x = c["logits"].cuda()
    tg, il, tl = c["targets"].cuda(), c["input_lengths"].cuda(), c["target_lengths"].cuda()
    NS, Lp, Sp, off, total = layout(B, T, Umax)
    wsb = _lib.workspace_bytes(B, T, V, Umax)
    assert wsb == total, (wsb, total)
    ws = torch.zeros(wsb, dtype=torch.uint8, device="cuda")
    nll = torch.full((B,), -1.0, device="cuda")
    sums = torch.zeros(4, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    rc = L.ctcb200_forward(x.data_ptr(), tg.data_ptr(), tg.shape[1], tg.numel(), il.data_ptr(), tl.data_ptr(),
                           B, T, V, Umax, 0, zi, nll.data_ptr(), sums.data_ptr(), ws.data_ptr(), wsb, st, None)
    print("forward rc", rc, _lib.strerror(rc)); torch.cuda.synchronize()
    w = ws.cpu().numpy()
    i32 = lambda name, n: w[off[name]: off[name] + 4 * n].view(np.int32)
    print("status", i32("hdr", 2), "Tb", i32("Tb", B), "Ub", i32("Ub", B), "flags", i32("flags", B))
    print("rowstart", i32("rowstart", B + 1), "toff", w[off["toff"]: off["toff"] + 8 * B].view(np.int64))
    lp_lab = w[off["lp_lab"]: off["lp_lab"] + 4 * B * T * Lp].view(np.float32).reshape(B, T, Lp)
    gam = w[off["gam"]: off["gam"] + 4 * B * T * Lp].view(np.float32).reshape(B, T, Lp)
    lp = torch.log_softmax(c["logits"].double(), -1).numpy() * LOG2E
    err = 0.0
    for b in range(B):
        for t in range(int(c["input_lengths"][b])):
            err = max(err, abs(lp_lab[b, t, 0] - lp[b, t, 0]))
            for j in range(int(c["target_lengths"][b])):
                err = max(err, abs(lp_lab[b, t, 4 + j] - lp[b, t, int(c["targets"][b, j])]))
    print("lp_lab max abs err (log2 units):", err)
    rn, _ = ref_ctc(c["logits"], c["targets"], c["input_lengths"], c["target_lengths"], reduction="none",
                    zero_infinity=bool(zi), want_grad=False)
    print("nll gpu", nll.cpu().numpy()); print("nll ref", rn.numpy()); print("sums", sums.cpu().numpy())
    for b in range(B):
        tb, ub = int(c["input_lengths"][b]), int(c["target_lengths"][b])
        if tb and np.isfinite(rn[b].item()) and rn[b].item() != 0:
            s = gam[b, :tb, 0] + gam[b, :tb, 4:4 + ub].sum(-1)
            print(f"  b={b} occupancy row-sum range [{s.min():.6f}, {s.max():.6f}]")
    go = torch.ones((), device="cuda")
    grad = torch.full_like(x, 7.0)
    rc = L.ctcb200_backward(x.data_ptr(), tg.data_ptr(), tg.shape[1], tg.numel(), go.data_ptr(), 0, 1, 1.0 / B,
                            B, T, V, Umax, 0, zi, grad.data_ptr(), ws.data_ptr(), wsb, st)
    print("backward rc", rc, _lib.strerror(rc)); torch.cuda.synchronize()
    rl, rg = ref_ctc(c["logits"], c["targets"], c["input_lengths"], c["target_lengths"], reduction="mean",
                     zero_infinity=bool(zi))
    g = grad.cpu()
    nanok = torch.equal(g.isnan(), rg.isnan())
    d = (g - rg).abs()
    d[rg.isnan()] = 0
    print("grad max abs err", d.max().item(), "nan pattern ok", nanok, "loss ref", rl.item(),
          "gpu", sums[0].item() / B)
    if d.max().item() > 1e-4:
        bad = (d > 1e-4).nonzero()
        print("  first bad", bad[:10].tolist())
        for (b, t, v) in bad[:5].tolist():
            print("   ", b, t, v, g[b, t, v].item(), rg[b, t, v].item())


if __name__ == "__main__":
    args = [int(a) if a.lstrip("-").isdigit() else a for a in sys.argv[1:]]
    main(*args)
