"""Developer tool (GPU box): the fused CTC head (ctc_head_loss_b200, tcgen05) against F.linear + the CPU oracle and
against the unfused path of this repo.  Every configuration runs in its own process: a faulting kernel must not take
the other checks down with it.   python tools/head_check.py [config-index]"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# (B, T, K, V, Umax, dist, lengths, precision, time_it)
CONFIGS = [
    (2, 64, 32, 40, 5, "D1", "full", "tf32", False),          # one tile, one class tile, one K chunk
    (2, 64, 32, 40, 5, "D1", "full", "3xtf32", False),
    (3, 100, 64, 300, 7, "D1", "var", "3xtf32", False),       # 3 row tiles spanning utterances, 2 class tiles
    (2, 400, 128, 700, 20, "D2", "dead", "3xtf32", False),    # a dead (fully padded) tile
    (8, 200, 512, 4234, 30, "D1", "var", "3xtf32", False),
    (8, 200, 512, 4234, 30, "D2", "var", "tf32", False),
    (256, 400, 512, 4234, 50, "D1", "var", "3xtf32", True),   # C2
    (256, 400, 512, 4234, 50, "D1", "var", "tf32", True),
]


def run(idx):
    import torch
    import torch.nn.functional as F
    from asr_chinese_e2e_b200 import ctc_head_loss_b200, ctc_loss_b200
    from oracle.synth import make_lengths, make_targets
    from oracle.torch_ref import ref_ctc
    B, T, K, V, U, dist, lengths, prec, time_it = CONFIGS[idx]
    g = torch.Generator().manual_seed(100 + idx)
    tg, tl = make_targets(B, U, V, g)
    il = make_lengths(B, T, g, full=(lengths == "full"))
    if lengths == "dead":
        il = torch.tensor([T, 50][:B] + [T] * max(0, B - 2))
    il = torch.maximum(il, torch.minimum(2 * tl + 1, torch.tensor(T)))
    enc = torch.randn(B, T, K, generator=g)
    W = torch.randn(V, K, generator=g) / K ** 0.5
    bias = torch.randn(V, generator=g) * 0.1
    if dist == "D2":                                   # sharp posteriors: push the head towards a random alignment
        W = W * 3.0
    dev = "cuda"
    enc_d, W_d, b_d = enc.to(dev).requires_grad_(True), W.to(dev).requires_grad_(True), bias.to(dev).requires_grad_(True)
    tg_d, il_d, tl_d = tg.to(dev), il.to(dev), tl.to(dev)
    out = {"config": CONFIGS[idx]}
    # reference: cuBLAS fp32 logits -> CPU oracle
    with torch.no_grad():
        logits = F.linear(enc_d, W_d, b_d)
        logits64 = (enc_d.double() @ W_d.double().t() + b_d.double())
        out["cublas_fp32_logit_err_vs_f64"] = float((logits.double() - logits64).abs().max())
    ref_nll, _ = ref_ctc(logits, tg, il, tl, reduction="none", zero_infinity=True, want_grad=False)
    _, ref_g = ref_ctc(logits, tg, il, tl, reduction="mean", zero_infinity=True)
    # fused, evaluation path
    with torch.no_grad():
        nll_eval = ctc_head_loss_b200(enc_d, W_d, b_d, tg_d, il_d, tl_d, reduction="none", zero_infinity=True, precision=prec)
    torch.cuda.synchronize()
    fin = torch.isfinite(ref_nll)
    out["nll_rel_max_eval"] = float(((nll_eval.cpu()[fin] - ref_nll[fin]).abs() / ref_nll[fin].abs().clamp(min=1)).max())
    # fused, training path
    loss = ctc_head_loss_b200(enc_d, W_d, b_d, tg_d, il_d, tl_d, reduction="mean", zero_infinity=True, precision=prec)
    loss.backward()
    torch.cuda.synchronize()
    fused = [p.grad.clone() for p in (enc_d, W_d, b_d)]
    for p in (enc_d, W_d, b_d):
        p.grad = None
    lg = F.linear(enc_d, W_d, b_d)
    lg.retain_grad()
    loss_u = ctc_loss_b200(lg, tg_d, il_d, tl_d, reduction="mean", zero_infinity=True)
    loss_u.backward()
    torch.cuda.synchronize()
    out["loss_fused"], out["loss_unfused"] = float(loss), float(loss_u)
    out["loss_rel"] = abs(float(loss) - float(loss_u)) / abs(float(loss_u))
    out["dlogits_abs_max_vs_oracle(unfused)"] = float((lg.grad.cpu() - ref_g).abs().max())
    for name, a, p in zip(("d_enc", "d_weight", "d_bias"), fused, (enc_d, W_d, b_d)):
        out[name + "_abs_max"] = float((a - p.grad).abs().max())
        out[name + "_scale"] = float(p.grad.abs().max())
    if time_it:
        def timeit(fn, n=5):
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n

        def f_eval():
            with torch.no_grad():
                ctc_head_loss_b200(enc_d, W_d, b_d, tg_d, il_d, tl_d, zero_infinity=True, precision=prec)

        def u_eval():
            with torch.no_grad():
                ctc_loss_b200(F.linear(enc_d, W_d, b_d), tg_d, il_d, tl_d, zero_infinity=True)

        def f_train():
            for p in (enc_d, W_d, b_d):
                p.grad = None
            ctc_head_loss_b200(enc_d, W_d, b_d, tg_d, il_d, tl_d, zero_infinity=True, precision=prec).backward()

        def u_train():
            for p in (enc_d, W_d, b_d):
                p.grad = None
            ctc_loss_b200(F.linear(enc_d, W_d, b_d), tg_d, il_d, tl_d, zero_infinity=True).backward()

        def lin_only():
            with torch.no_grad():
                F.linear(enc_d, W_d, b_d)
        out["ms"] = {"fused_eval": timeit(f_eval), "unfused_eval": timeit(u_eval), "fused_train": timeit(f_train),
                     "unfused_train": timeit(u_train), "cublas_fp32_linear_only": timeit(lin_only)}
        torch.backends.cuda.matmul.allow_tf32 = True
        out["ms"]["cublas_tf32_linear_only"] = timeit(lin_only)
        torch.backends.cuda.matmul.allow_tf32 = False
    print("HEADCHECK " + json.dumps(out), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run(int(sys.argv[1]))
    else:
        for i in range(len(CONFIGS)):
            t0 = time.time()
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), str(i)], capture_output=True, text=True, timeout=300)
                tail = [l for l in r.stdout.splitlines() if l.startswith("HEADCHECK")]
                print(f"[{i}] rc={r.returncode} {time.time() - t0:.1f}s", tail[0] if tail else (r.stdout[-800:] + r.stderr[-1500:]), flush=True)
            except subprocess.TimeoutExpired:
                print(f"[{i}] TIMEOUT", flush=True)
