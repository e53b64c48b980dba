"""Developer tool: how long do the two backward GEMMs of the CTC head take in cuBLAS fp32 / TF32 at the C2 shape?"""
import torch
M, V, K, pitch = 102400, 4234, 512, 4236
dl = torch.randn(M, pitch, device="cuda")[:, :V]
x = torch.randn(M, K, device="cuda"); w = torch.randn(V, K, device="cuda")
def t(f, n=5):
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for tf32 in (False, True):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    print(f"allow_tf32={tf32}: d_enc = dl @ W {t(lambda: dl @ w):.3f} ms, d_W = dl^T @ enc {t(lambda: dl.t() @ x):.3f} ms, bias sum {t(lambda: dl.sum(0)):.3f} ms")
