"""Developer tool (GPU box): the tcgen05 3xTF32 parameter-gradient GEMMs (ctcb200_head_param_grads) against float64
matmuls, each configuration in its own process.   python tools/gemm_check.py [config-index]"""
import json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
# (B, T, V, K, time_it)
CONFIGS = [(1, 128, 32, 32, False), (2, 64, 40, 32, False), (3, 100, 300, 64, False), (2, 333, 700, 128, False),
           (8, 200, 4234, 512, False), (256, 400, 4234, 512, True)]


def run(idx):
    import torch
    from asr_chinese_e2e_b200 import _lib
    B, T, V, K, time_it = CONFIGS[idx]
    L = _lib.lib()
    g = torch.Generator().manual_seed(idx)
    pitch = (V + 3) // 4 * 4
    dl = torch.zeros(B * T, pitch)
    dl[:, :V] = torch.randn(B * T, V, generator=g) * 1e-2
    enc = torch.randn(B * T, K, generator=g)
    w = torch.randn(V, K, generator=g) / K ** 0.5
    dl_d, enc_d, w_d = dl.cuda(), enc.cuda(), w.cuda()
    g_enc = torch.full((B * T, K), 7.0, device="cuda"); g_w = torch.full((V, K), 7.0, device="cuda")
    wsb = _lib.head_param_grads_workspace_bytes(V, K)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream

    def call():
        rc = L.ctcb200_head_param_grads(dl_d.data_ptr(), pitch, enc_d.data_ptr(), w_d.data_ptr(), B, T, V, K, g_enc.data_ptr(),
                                        g_w.data_ptr(), ws.data_ptr(), wsb, st)
        assert rc == 0, _lib.strerror(rc)
    call()
    torch.cuda.synchronize()
    ref_enc = (dl_d[:, :V].double() @ w_d.double())
    ref_w = (dl_d[:, :V].double().t() @ enc_d.double())
    f32_enc = dl_d[:, :V] @ w_d
    f32_w = dl_d[:, :V].t() @ enc_d
    out = {"config": CONFIGS[idx],
           "d_enc_err": float((g_enc.double() - ref_enc).abs().max()), "d_enc_scale": float(ref_enc.abs().max()),
           "d_enc_err_cublas_fp32": float((f32_enc.double() - ref_enc).abs().max()),
           "d_w_err": float((g_w.double() - ref_w).abs().max()), "d_w_scale": float(ref_w.abs().max()),
           "d_w_err_cublas_fp32": float((f32_w.double() - ref_w).abs().max())}
    if time_it:
        for _ in range(2):
            call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            call()
        e1.record(); torch.cuda.synchronize()
        out["ms_both_gemms"] = e0.elapsed_time(e1) / 5
    out["env"] = {k: v for k, v in os.environ.items() if k.startswith("CTCB200_G3")}
    out["d_enc_sample"] = [round(float(v), 5) for v in g_enc[0, :4]] + [round(float(v), 5) for v in ref_enc[0, :4]]
    out["d_w_sample"] = [round(float(v), 5) for v in g_w[0, :4]] + [round(float(v), 5) for v in ref_w[0, :4]]
    print("GEMMCHECK " + json.dumps(out), flush=True)


# descriptor variants of the MN-major operand tiles (developer bisect; the first one is the kernel's default)
VARIANTS = [{}, {"CTCB200_G3_SBO": "1024"}, {"CTCB200_G3_SWZ": "3", "CTCB200_G3_LAYOUT": "2", "CTCB200_G3_SBO": "1024"},
            {"CTCB200_G3_SWZ": "3", "CTCB200_G3_LAYOUT": "2", "CTCB200_G3_SBO": "512"},
            {"CTCB200_G3_SWZ": "3"}, {"CTCB200_G3_LAYOUT": "2"}]

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "split":
        for v in ({}, {"CTCB200_G3_RNA_SPLIT": "1"}):
            for cfg in ("3", "5"):
                r = subprocess.run([sys.executable, os.path.abspath(__file__), cfg], capture_output=True, text=True, timeout=300,
                                   env={**os.environ, **v})
                tail = [l for l in r.stdout.splitlines() if l.startswith("GEMMCHECK")]
                print(v, f"rc={r.returncode}", tail[0][:520] if tail else (r.stdout[-300:] + r.stderr[-800:]), flush=True)
    elif len(sys.argv) > 1 and sys.argv[1] == "variants":
        for v in VARIANTS:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "3"], capture_output=True, text=True, timeout=300,
                               env={**os.environ, **v})
            tail = [l for l in r.stdout.splitlines() if l.startswith("GEMMCHECK")]
            print(v, f"rc={r.returncode}", tail[0] if tail else (r.stdout[-300:] + r.stderr[-800:]), flush=True)
    elif len(sys.argv) > 1:
        run(int(sys.argv[1]))
    else:
        for i in range(len(CONFIGS)):
            t0 = time.time()
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), str(i)], capture_output=True, text=True, timeout=300)
                tail = [l for l in r.stdout.splitlines() if l.startswith("GEMMCHECK")]
                print(f"[{i}] rc={r.returncode} {time.time() - t0:.1f}s", tail[0] if tail else (r.stdout[-500:] + r.stderr[-1200:]), flush=True)
            except subprocess.TimeoutExpired:
                print(f"[{i}] TIMEOUT", flush=True)
