"""Developer tool (GPU box): where does the host-buffer pipeline's time go?  Variants of HostCTCPipeline on the C2 batch."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from concurrent.futures import ThreadPoolExecutor
from oracle.synth import make_config
from asr_chinese_e2e_b200 import host_pipeline as hp

c = make_config("C2", dist="D1")
B, T, V = c["logits"].shape
U = c["targets"].shape[1]
h_x = c["logits"].pin_memory()
h_tg, h_il, h_tl = c["targets"].pin_memory(), c["input_lengths"].pin_memory(), c["target_lengths"].pin_memory()
h_g = torch.empty(B, T, V, pin_memory=True)
h_n = torch.empty(B, pin_memory=True)


def run(label, **kw):
    workers = kw.pop("workers", None)
    nozero = kw.pop("nozero", False)
    pipe = hp.HostCTCPipeline(B, T, V, U, device="cuda", **kw)
    if workers and pipe._pool is not None:
        pipe._pool = ThreadPoolExecutor(max_workers=workers)
    if nozero and pipe._pool is not None:
        class _NoPool:
            def submit(self, *a, **k):
                class F:
                    def result(self): return None
                return F()
        pipe._pool = _NoPool()
    for _ in range(3):
        pipe(h_x, h_tg, h_il, h_tl, h_g, h_n)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(8):
        pipe(h_x, h_tg, h_il, h_tl, h_g, h_n)
    torch.cuda.synchronize()
    print(f"{label}: {1e3 * (time.perf_counter() - t0) / 8:.2f} ms/step  h2d {pipe.h2d_bytes / 1e9:.3f} GB d2h {pipe.d2h_bytes / 1e9:.3f} GB", flush=True)


run("all frames, chunk 32", valid_frames_only=False)
run("all frames, chunk 16", valid_frames_only=False, chunk=16)
for ch, ns in ((4, 3), (8, 3), (8, 6), (16, 3), (16, 4), (16, 6), (32, 3), (32, 4)):
    run(f"valid frames, chunk {ch}, {ns} slots", chunk=ch, n_slots=ns)
run("valid frames, chunk 16, no host zeroing (floor)", chunk=16, nozero=True)
run("valid frames, grad on device, chunk 16", grad_to_host=False, chunk=16)
# raw host fill bandwidth on this box
import numpy as np
a = h_g.numpy()
lens = h_il.numpy()
for w in (1, 4, 8, 16):
    pool = ThreadPoolExecutor(w)
    def z(b0, b1):
        for b in range(b0, b1):
            a[b, lens[b]:].fill(0.0)
    t = time.perf_counter()
    fs = [pool.submit(z, b, b + 8) for b in range(0, B, 8)]
    [f.result() for f in fs]
    dt = time.perf_counter() - t
    print(f"host fill, {w} threads: {dt * 1e3:.1f} ms, {(T - lens).sum() * V * 4 / dt / 1e9:.1f} GB/s", flush=True)
