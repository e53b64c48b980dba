#!/usr/bin/env python
"""bench.py -- CTC loss+grad utterances/s on synthetic AISHELL-shaped batches (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--lengths var|full]

A "step" is one pass of the hot path over one batch: prep + fused log-softmax/label-gather sweep +
alpha/beta lattice + fused gradient sweep (4 kernel launches), producing the 'mean' loss and the
gradient w.r.t. the [B,T,V] logits.  Workload at every N: BASELINE.json configs[1] per GPU
(B=256, T=400, V=4234, U<=50, variable lengths with padding, reduction=mean) -- weak scaling,
batch sharded by utterance, one NCCL all-reduce of [sum_b nll_b/U_b, B_local] per step.

One JSON line on stdout (rank 0).  `value` = whole-job utterances/s with the logits resident in
HBM; `e2e` = the same metric through the host-buffer API (pinned host logits in, gradient + loss
back to pinned host memory, copies inside the timed region); `roofline` = the dominant kernel
(k3_grad) against the measured HBM peak; `cpu_baseline` = torch's CPU ctc_loss path on a bounded
sample of the same batch, timed on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_, T_, V_, U_ = 256, 400, 4234, 50
SEED = 1002
METRIC = "ctc_loss_grad_utterances_per_s"
UNIT = "utt/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--lengths", default="var", choices=["var", "full"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=64, help="utterances in the CPU-baseline sample")
    ap.add_argument("--vocab", type=int, default=V_, help="developer experiments only (default = BASELINE V)")
    ap.add_argument("--shape", default=None, help="developer experiments only: B,T,U (default = BASELINE C2 256,400,50)")
    return ap.parse_args()


def make_batch(rank, lengths):
    from oracle.synth import make_case   # input generator only (seeded, CPU); not on the product path
    return make_case(B_, T_, V_, U_, SEED + rank, dist="D1", full_lengths=(lengths == "full"))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms.  Started before the warm-up so that it is
    already running when the (short) timed region begins; summary() keeps the samples whose arrival time
    falls inside the timed window and, if the window was shorter than one sampling period, falls back to
    the samples taken under the same load during warm-up + timed steps (and says so)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def wait_first_sample(self, timeout=5.0):
        t0 = time.time()
        while self.proc and not self.rows and time.time() - t0 < timeout:
            time.sleep(0.02)

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0, t1, load_t0):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def digest(rows):
            sm, mx, reasons = [], 0.0, set()
            for _, r in rows:
                try:
                    sm.append(float(r[0])); mx = max(mx, float(r[1]))
                except (ValueError, IndexError):
                    continue
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            return sm, mx, reasons
        window = "timed region"
        rows = [x for x in self.rows if t0 <= x[0] <= t1 + 0.06]
        if not digest(rows)[0]:
            rows = [x for x in self.rows if load_t0 <= x[0] <= t1 + 0.06]
            window = "warm-up + timed region (timed region shorter than one 50 ms sampling period)"
        sm, mx, reasons = digest(rows)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def cpu_reference_line(args, c, steps, warmup):
    """torch CPU F.log_softmax + F.ctc_loss + backward on a bounded sample (first n utterances)."""
    import torch
    from oracle.torch_ref import ref_step
    n = min(args.cpu_sample, B_)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    x = c["logits"][:n].clone().requires_grad_(True)
    tg, il, tl = c["targets"][:n], c["input_lengths"][:n], c["target_lengths"][:n]
    for _ in range(warmup):
        ref_step(x, tg, il, tl)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        ref_step(x, tg, il, tl)
        ts.append(time.perf_counter() - t0)
    ms = 1e3 * sum(ts) / len(ts)
    return {"value": n / (ms / 1e3), "unit": UNIT, "cores": cores, "kind": "reference",
            "sample": f"first {n} utterances of the C2 batch (T={T_}, V={V_}, U<={U_}, lengths={args.lengths}), "
                      f"torch {torch.__version__} CPU F.log_softmax+F.ctc_loss(mean)+backward fp32, "
                      f"{warmup} warm-up + {steps} timed, mean",
            "ms_per_step": ms, "threads": torch.get_num_threads()}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    c = make_batch(0, args.lengths)
    steps, warmup = max(1, min(args.steps, 8)), max(1, min(args.warmup, 2))
    cb = cpu_reference_line(args, c, steps, warmup)
    line = {"metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": workload_cfg(args, 1),
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_cfg(args, n):
    return {"workload": f"C2: AISHELL-1-shaped CTC loss+grad, B={B_}/GPU, T={T_}, V={V_}, U<={U_}, "
                        f"{'variable lengths ~U[T/2,T] with padding' if args.lengths == 'var' else 'full lengths'}, "
                        "reduction=mean, zero_infinity=False, randn logits (D1)",
            "global_batch": B_ * n, "lengths": args.lengths, "seed": SEED,
            "parallelism": f"batch-sharded x{n}, 1 all-reduce of 1 float/step" if n > 1 else "single GPU",
            "l2_policy": "inputs (1.73 GB logits + 1.73 GB grad per step) exceed the 126 MB L2; no flush needed"}


def main():
    args = parse()
    global V_, B_, T_, U_
    V_ = args.vocab
    if args.shape:
        B_, T_, U_ = (int(v) for v in args.shape.split(","))
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from asr_chinese_e2e_b200 import _lib, ctc_loss_b200
    from asr_chinese_e2e_b200.sharded import sharded_ctc_loss

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the CTC hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    _lib.lib()                                                  # fail loudly if the .so is missing
    dev = torch.device("cuda", local)

    c = make_batch(rank, args.lengths)
    x = c["logits"].to(dev).requires_grad_(True)
    tg, il, tl = c["targets"].to(dev), c["input_lengths"].to(dev), c["target_lengths"].to(dev)
    sum_T = int(c["input_lengths"].sum())
    bytes_step = 4 * V_ * (2 * sum_T + B_ * T_)                 # 3-sweep algorithmic bytes
    bytes_k3 = 4 * V_ * (sum_T + B_ * T_)                       # k3: re-read valid frames + write all of grad

    def fwd():
        if world == 1:
            return ctc_loss_b200(x, tg, il, tl, blank=0, reduction="mean", zero_infinity=False)   # fused, chunked
        return sharded_ctc_loss(x, tg, il, tl, blank=0, zero_infinity=False)   # 1 all-reduce of one float

    def step():
        x.grad = None
        loss = fwd()
        loss.backward()
        return loss

    ev = lambda: torch.cuda.Event(enable_timing=True)
    K = args.steps
    # ---- headline: K steps of the public op (two-sweep path, gradient produced in the forward call) ----
    with ClockSampler(local) as clocks:
        clocks.wait_first_sample()
        load_t0 = time.time()
        for _ in range(max(args.warmup, 3)):
            loss = step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t_start, t_end = ev(), ev()
        wall0 = time.time()
        t_start.record()
        for i in range(K):
            loss = step()
        t_end.record()
        torch.cuda.synchronize()
        wall1 = time.time()
        clock_summary = clocks.summary(wall0, wall1, load_t0)
    total_ms = t_start.elapsed_time(t_end)
    if world > 1:
        t = torch.tensor([total_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_step = total_ms / K
    value = world * B_ * K / (total_ms / 1e3)
    loss_val = float(loss.item())

    # ---- per-kernel timing for the roofline (live, CUDA events on the launching stream): one un-chunked
    #      two-sweep call; the library's sweep_done event splits it into [k0_prep + fused sweep kernel]
    #      and [lattice + sparse patch] ----
    from asr_chinese_e2e_b200.profiling import time_stages
    st = time_stages(x, tg, il, tl, reduction="mean", zero_infinity=False, iters=min(K, 50), warmup=3)
    sweep_ms, rest_ms = statistics.mean(st["sweep_ms"]), statistics.mean(st["rest_ms"])

    n_chunks = int(os.environ.get("CTCB200_CHUNKS", "1"))
    launches_per_step = 4 * n_chunks + 1          # per chunk: k0, k1(fused), k2, k3p; plus the (empty) rescale launch
    peak, peak_src = peaks()
    bytes_2sweep = bytes_k3                                     # read valid frames once + write all of grad once
    k1f_gbs = bytes_2sweep / (sweep_ms / 1e3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "k1f_traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        traffic = tj.get("dram_bytes_per_launch", {}).get(args.lengths)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "b200",
            "config": workload_cfg(args, world), "loss": loss_val,
            "gpu_launches": launches_per_step * K,
            "roofline": {"bound": "hbm", "kernel": "k1_lse_gather<FUSED> (log-softmax stats + label gather + dense gradient; "
                                                   "timed with k0_prep, ~4 us)",
                         "achieved": k1f_gbs, "peak": peak, "unit": "GB/s", "frac": k1f_gbs / peak, "traffic": traffic,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_2sweep,
                         "ms_per_launch": sweep_ms},
            "roofline_step": {"algorithmic_bytes_3sweep": bytes_step, "algorithmic_bytes_2sweep": bytes_2sweep,
                              "achieved_vs_3sweep": bytes_step / (ms_step / 1e3) / 1e9,
                              "achieved_vs_2sweep": bytes_2sweep / (ms_step / 1e3) / 1e9,
                              "frac": bytes_step / (ms_step / 1e3) / 1e9 / peak,
                              "frac_2sweep": bytes_2sweep / (ms_step / 1e3) / 1e9 / peak,
                              "peak": peak, "unit": "GB/s",
                              "note": "BASELINE.md's primary figure is the 3-sweep byte count; this implementation "
                                      "needs only 2 sweeps (+ a sparse correction), so frac can exceed the share of "
                                      "HBM actually used (frac_2sweep)",
                              "pipeline": f"{n_chunks} utterance chunk(s), two-sweep path, gradient produced in the "
                                          "forward call (speculative upstream gradient 1)",
                              "sweep_ms": sweep_ms, "lattice_plus_patch_ms": rest_ms},
            "clocks": clock_summary}

    if rank == 0 and world == 1:
        if not args.no_e2e:
            line["e2e"] = run_e2e(torch, c, args, dev)
            # the training-shaped variant: the gradient's consumer is the next GPU kernel, only the loss goes back
            line["e2e_grad_on_device"] = run_e2e(torch, c, args, dev, grad_to_host=False)
        line["torch_cuda_baseline"] = run_torch_cuda(torch, x, tg, il, tl)
        if not args.no_cpu:
            cb = cpu_reference_line(args, c, steps=5, warmup=2)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    elif world > 1:
        e = run_e2e(torch, c, args, dev) if not args.no_e2e else None
        if e is not None:
            t = torch.tensor([e["ms_per_step"]], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e["ms_per_step"] = float(t.item())
            e["value"] = world * B_ / (e["ms_per_step"] / 1e3)
            e["h2d_bytes_per_step"] *= world; e["d2h_bytes_per_step"] *= world
            line["e2e"] = e
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_torch_cuda(torch, x, tg, il, tl):
    """Secondary GPU baseline: torch's own CUDA log_softmax + ctc_loss + backward on the same
    B200 and inputs -- what the reference as written would execute once a CTC call were added."""
    import torch.nn.functional as F

    def step():
        x.grad = None
        F.ctc_loss(F.log_softmax(x, -1).transpose(0, 1), tg, il, tl, blank=0, reduction="mean",
                   zero_infinity=False).backward()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        step()
    e1.record()
    torch.cuda.synchronize()
    x.grad = None
    ms = e0.elapsed_time(e1) / 10
    return {"value": B_ / (ms / 1e3), "unit": UNIT, "ms_per_step": ms,
            "what": f"torch {torch.__version__} CUDA F.log_softmax + F.ctc_loss(mean) + backward, same inputs, 10 steps"}


def run_e2e(torch, c, args, dev, grad_to_host=True):
    """Host-buffer API: pinned logits -> device, kernels, gradient + nll -> pinned host, every step."""
    from asr_chinese_e2e_b200.host_pipeline import HostCTCPipeline
    pipe = HostCTCPipeline(B_, T_, V_, U_, chunk=32, device=dev, grad_to_host=grad_to_host)
    h_x = c["logits"].pin_memory()
    h_tg, h_il, h_tl = c["targets"].pin_memory(), c["input_lengths"].pin_memory(), c["target_lengths"].pin_memory()
    h_g = torch.empty(B_, T_, V_, pin_memory=True)
    h_n = torch.empty(B_, pin_memory=True)
    k = max(3, min(args.steps, 10))
    for _ in range(3):
        pipe(h_x, h_tg, h_il, h_tl, h_g, h_n)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(k):
        pipe(h_x, h_tg, h_il, h_tl, h_g, h_n)      # returns after the step's D2H has completed
    torch.cuda.synchronize()
    ms = 1e3 * (time.perf_counter() - t0) / k
    loss = float((h_n / h_tl.clamp(min=1).float()).mean())
    return {"value": B_ / (ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": pipe.h2d_bytes,
            "d2h_bytes_per_step": pipe.d2h_bytes, "ms_per_step": ms, "steps": k, "loss": loss,
            "api": "asr_chinese_e2e_b200.host_pipeline.HostCTCPipeline (pinned host logits in; "
                   + ("grad[B,T,V] + " if grad_to_host else "gradient left on the device, ")
                   + "nll[B] back to pinned host; 3-stream chunked pipeline, chunk=32 utterances)",
            "gpu_launches_per_step": pipe.launches_per_step}


if __name__ == "__main__":
    main()
