#!/usr/bin/env python
"""bench.py -- CTC loss+grad utterances/s on synthetic AISHELL-shaped batches (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--lengths var|full]

A "step" is one pass of the hot path over one batch through the public op (``loss = ctc_loss_b200(...)``,
``loss.backward()``): prep + fused log-softmax/label-gather/dense-gradient sweep + alpha/beta lattice + sparse
occupancy patch (+ one empty rescale launch), producing the 'mean' loss and the gradient w.r.t. the [B,T,V] logits.
Workload at every N: BASELINE.json configs[1] per GPU (B=256, T=400, V=4234, U<=50, variable lengths with padding,
reduction=mean) -- weak scaling, batch sharded by utterance, one all-reduce of a single float per step.

One JSON line on stdout (rank 0):
  value         whole-job utterances/s with the logits resident in HBM (CUDA events, max over ranks)
  e2e           the same metric through the host-buffer API (pinned host logits in, gradient + nll back to pinned
                host memory, copies inside the timed region)
  roofline      the dominant kernel, the fused sweep (k1p_sweep<FUSED> at this shape), timed live with CUDA events on the
                launching stream against the measured HBM peak; roofline_step = whole step
  parity        the step's nll / gradient against the CPU oracle (torch's CPU ctc_loss), checked in this very run
  cpu_baseline  torch's CPU log_softmax + ctc_loss + backward on the full batch (median; + a 1-thread figure)
  configs       driver-visible sub-records: C2 full lengths, C2 with sharp (trained-like) posteriors, C4 long
                utterances, C5 joint step with the 12-layer encoder -- each with ms/step, roofline fractions and the
                lattice kernel's path counters
  comm          (N > 1) what the single collective costs: step time with and without it
``--impl reference`` times the reference arm: torch's CPU op on this box's host cores (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import platform
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
REL_LOSS, ABS_GRAD = 1e-5, 1e-4


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--lengths", default="var", choices=["var", "full"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the C2-full / C2-D2 / C4 / C5 sub-records")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--same-seed", action="store_true", help="N>1 experiment: every rank gets rank 0's batch")
    ap.add_argument("--per-rank-lengths", action="store_true",
                    help="N>1: every rank also draws its own lengths / transcripts (per-GPU work then varies with the "
                         "rank; default: rank 0's lengths and transcripts on every rank, each rank its own logits)")
    ap.add_argument("--vocab", type=int, default=V_, help="developer experiments only (default = BASELINE V)")
    ap.add_argument("--shape", default=None, help="developer experiments only: B,T,U (default = BASELINE C2 256,400,50)")
    return ap.parse_args()


def make_batch(rank, lengths, dist="D1"):
    from oracle.synth import make_case   # input generator only (seeded, CPU); not on the product path
    return make_case(B_, T_, V_, U_, SEED + rank, dist=dist, full_lengths=(lengths == "full"))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return platform.processor() or "unknown"


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


# ------------------------------------------------------------------------------------------------------------
# CPU reference (the oracle / reference arm): torch CPU F.log_softmax + F.ctc_loss + backward
# ------------------------------------------------------------------------------------------------------------
def time_cpu(c, n, steps, warmup, threads):
    import torch
    from oracle.torch_ref import ref_step
    torch.set_num_threads(threads)
    x = c["logits"][:n].clone().requires_grad_(True)
    tg, il, tl = c["targets"][:n], c["input_lengths"][:n], c["target_lengths"][:n]
    for _ in range(warmup):
        ref_step(x, tg, il, tl)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        ref_step(x, tg, il, tl)
        ts.append(time.perf_counter() - t0)
    return ts


def cpu_reference_line(args, c, steps, warmup, one_thread=True):
    """full batch, all host threads, `warmup` warm-up + `steps` timed, MEDIAN; plus a
    1-thread figure on a bounded 16-utterance sample."""
    import torch
    cores = os.cpu_count() or 1
    ts = time_cpu(c, B_, steps, warmup, cores)
    med = statistics.median(ts)
    out = {"value": B_ / med, "unit": UNIT, "cores": cores, "kind": "reference", "cpu_model": cpu_model(),
           "sample": f"the full C2 batch ({B_} utterances, T={T_}, V={V_}, U<={U_}, lengths={args.lengths}), "
                     f"torch {torch.__version__} CPU F.log_softmax+F.ctc_loss(mean)+backward fp32, "
                     f"{warmup} warm-up + {steps} timed, median",
           "ms_per_step": 1e3 * med, "ms_per_step_mean": 1e3 * sum(ts) / len(ts), "threads": cores}
    if one_thread:
        n1 = min(16, B_)
        t1 = time_cpu(c, n1, 2, 1, 1)
        torch.set_num_threads(cores)
        out["one_thread"] = {"value": n1 / statistics.median(t1), "unit": UNIT, "threads": 1,
                             "sample": f"first {n1} utterances of the same batch, 1 warm-up + 2 timed, median"}
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    c = make_batch(0, args.lengths)
    steps, warmup = max(1, min(args.steps, 60)), max(1, min(args.warmup, 10))     # ~0.5 s per step on 16 cores
    cb = cpu_reference_line(args, c, steps, warmup, one_thread=False)
    line = {"metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": workload_cfg(args, 1),
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "cpu_model")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_cfg(args, n):
    return {"workload": f"C2: AISHELL-1-shaped CTC loss+grad, B={B_}/GPU, T={T_}, V={V_}, U<={U_}, "
                        f"{'variable lengths ~U[T/2,T] with padding' if args.lengths == 'var' else 'full lengths'}, "
                        "reduction=mean, zero_infinity=False, randn logits (D1)",
            "global_batch": B_ * n, "lengths": args.lengths, "seed": SEED,
            "per_rank_data": ("n/a" if n == 1 else "rank 0's batch on every rank" if args.same_seed else
                              "own logits, lengths and transcripts per rank" if args.per_rank_lengths else
                              "own logits per rank; rank 0's lengths and transcripts (fixed work per GPU)"),
            "parallelism": f"batch-sharded x{n}, 1 all-reduce of 1 float/step" if n > 1 else "single GPU",
            "l2_policy": "inputs (1.73 GB logits + 1.73 GB grad per step) exceed the 126 MB L2; no flush needed"}


# ------------------------------------------------------------------------------------------------------------
# helpers of the b200 arm
# ------------------------------------------------------------------------------------------------------------
def time_steps(torch, step, K, W):
    for _ in range(W):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        out = step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K, out


def parity_check(torch, c, nll_gpu, grad_gpu, n, zero_infinity=False, inv_batch=None):
    """The first n utterances of the batch against the CPU oracle: per-utterance nll (relative) and the
    'mean' gradient (absolute).  Infeasible utterances are compared by class."""
    from oracle.torch_ref import ref_ctc
    B = c["logits"].shape[0]
    sub = {k: v[:n] for k, v in c.items()}
    ref_nll, _ = ref_ctc(sub["logits"], sub["targets"], sub["input_lengths"], sub["target_lengths"], reduction="none",
                         zero_infinity=zero_infinity, want_grad=False)
    go = (inv_batch if inv_batch is not None else 1.0 / B) / sub["target_lengths"].clamp(min=1).float()
    _, ref_g = ref_ctc(sub["logits"], sub["targets"], sub["input_lengths"], sub["target_lengths"], reduction="none",
                       zero_infinity=zero_infinity, grad_output=go)
    nll = nll_gpu[:n].detach().float().cpu()
    fin = torch.isfinite(ref_nll)
    same_class = bool(torch.equal(torch.isinf(nll), torch.isinf(ref_nll)))
    rel = ((nll[fin] - ref_nll[fin]).abs() / ref_nll[fin].abs().clamp(min=1.0)).max().item() if fin.any() else 0.0
    g = grad_gpu[:n].detach().cpu()
    ok = ~torch.isnan(ref_g)
    nan_same = bool(torch.equal(torch.isnan(g), torch.isnan(ref_g)))
    gerr = (g[ok] - ref_g[ok]).abs().max().item() if ok.any() else 0.0
    return {"nll_rel_max": rel, "grad_abs_max": gerr, "n_utts": int(n), "tol": {"nll_rel": REL_LOSS, "grad_abs": ABS_GRAD},
            "oracle": "torch CPU F.log_softmax + F.ctc_loss (oracle/torch_ref.py), same inputs",
            "pass": bool(rel <= REL_LOSS and gerr <= ABS_GRAD and same_class and nan_same)}


def sub_record(torch, name, c, zero_infinity, K, peak, parity_n, note):
    """One driver-visible sub-record: step time of the public op on another configuration + roofline fractions +
    the lattice kernel's path counters + a parity gate on a bounded sample."""
    from asr_chinese_e2e_b200 import ctc_loss_b200
    from asr_chinese_e2e_b200.profiling import time_stages
    dev = torch.device("cuda", torch.cuda.current_device())
    x = c["logits"].to(dev).requires_grad_(True)
    tg, il, tl = c["targets"].to(dev), c["input_lengths"].to(dev), c["target_lengths"].to(dev)
    B, T, V = x.shape

    def step():
        x.grad = None
        loss = ctc_loss_b200(x, tg, il, tl, blank=0, reduction="mean", zero_infinity=zero_infinity)
        loss.backward()
        return loss
    ms, loss = time_steps(torch, step, K, 3)
    st = time_stages(x, tg, il, tl, reduction="mean", zero_infinity=zero_infinity, iters=min(K, 20), warmup=2)
    sum_T = int(c["input_lengths"].sum())
    b2, b3 = 4 * V * (sum_T + B * T), 4 * V * (2 * sum_T + B * T)
    rec = {"workload": note, "B": B, "T": T, "V": V, "value": B / (ms / 1e3), "unit": UNIT, "ms_per_step": ms,
           "loss": float(loss.item()), "steps": K,
           "algorithmic_bytes_2sweep": b2, "frac_2sweep": b2 / (ms / 1e3) / 1e9 / peak,
           "frac_3sweep": b3 / (ms / 1e3) / 1e9 / peak,
           "sweep_ms": statistics.mean(st["sweep_ms"]), "lattice_plus_patch_ms": statistics.mean(st["rest_ms"]),
           "sweep_frac": b2 / (statistics.mean(st["sweep_ms"]) / 1e3) / 1e9 / peak,
           "lattice_stats": {"utterances": B, "log_space": st["lattice_stats"][0],
                             "of_which_underflowed_in_linear_domain": st["lattice_stats"][1]}}
    if parity_n:
        nll = ctc_loss_b200(x.detach(), tg, il, tl, reduction="none", zero_infinity=zero_infinity)
        step()
        rec["parity"] = parity_check(torch, c, nll, x.grad, parity_n, zero_infinity)
    del x
    torch.cuda.empty_cache()
    return rec


def fused_head_record(torch, dev, peak):
    """f1: the CTC head GEMM fused with the loss (tcgen05) at the C2 shape: enc [256,400,512] x W [4234,512]^T.
    Times the CTC branch as a training step sees it (head forward + loss + backward down to d enc / d W / d bias),
    fused and unfused, and the evaluation forward; reports the tensor-core fraction of the fused forward and the
    HBM traffic of the logits tensor that the fusion removes."""
    import torch.nn.functional as F
    from asr_chinese_e2e_b200 import ctc_head_loss_b200, ctc_loss_b200
    from oracle.synth import make_lengths, make_targets
    B, T, K, V, U = B_, T_, 512, V_, U_
    g = torch.Generator().manual_seed(SEED)
    tg, tl = make_targets(B, U, V, g)
    il = make_lengths(B, T, g)
    enc = torch.randn(B, T, K, generator=g).to(dev).requires_grad_(True)
    W = (torch.randn(V, K, generator=g) / K ** 0.5).to(dev).requires_grad_(True)
    bias = (torch.randn(V, generator=g) * 0.1).to(dev).requires_grad_(True)
    tg, il_d, tl_d = tg.to(dev), il.to(dev), tl.to(dev)
    ps = (enc, W, bias)

    def clear():
        for p in ps:
            p.grad = None

    def unfused_train():
        clear()
        loss = ctc_loss_b200(F.linear(enc, W, bias), tg, il_d, tl_d, zero_infinity=True)
        loss.backward()
        return loss

    def fused_train(prec):
        def f():
            clear()
            loss = ctc_head_loss_b200(enc, W, bias, tg, il_d, tl_d, zero_infinity=True, precision=prec)
            loss.backward()
            return loss
        return f

    def unfused_eval():
        with torch.no_grad():
            return ctc_loss_b200(F.linear(enc, W, bias), tg, il_d, tl_d, zero_infinity=True)

    def fused_eval(prec):
        def f():
            with torch.no_grad():
                return ctc_head_loss_b200(enc, W, bias, tg, il_d, tl_d, zero_infinity=True, precision=prec)
        return f

    def linear_only():
        with torch.no_grad():
            return F.linear(enc, W, bias)
    ms = {}
    ms["unfused_eval"], lu = time_steps(torch, unfused_eval, 5, 2)
    ms["unfused_train"], _ = time_steps(torch, unfused_train, 5, 2)
    ms["cublas_fp32_linear_only"], _ = time_steps(torch, linear_only, 5, 2)
    rec = {"workload": f"C2 shape through the CTC head: enc [{B},{T},{K}] fp32 x W [{V},{K}]^T + bias -> CTC loss (+ d enc, d W, d bias)",
           "kernel": "k_head<NPASS, GRADPASS>: TMA (UTMALDG) operand ring -> tcgen05.mma kind::tf32 (UTCHMMA), M128 N256 K8, fp32 "
                     "accumulators in TMEM -> epilogue from TMEM (LDTM): online log-sum-exp + label gather / recomputed gradient",
           "ms": ms, "parity_vs_unfused": {}}
    sum_T = int(il.sum())
    flops = 2.0 * B * T * V * K
    for prec, npass in (("3xtf32", 3), ("tf32", 1)):
        ms[f"fused_eval_{prec}"], lf = time_steps(torch, fused_eval(prec), 5, 2)
        ms[f"fused_train_{prec}"], _ = time_steps(torch, fused_train(prec), 5, 2)
        rec["parity_vs_unfused"][prec] = {"loss_rel": abs(float(lf) - float(lu)) / abs(float(lu)),
                                          "tol": 1e-5 if prec == "3xtf32" else 2e-3}
        rec[f"tensor_tflops_eval_{prec}"] = npass * flops / (ms[f"fused_eval_{prec}"] / 1e3) / 1e12
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    bf16_sus = float(json.load(open(pk))["bf16_tflops_sustained"]) if os.path.exists(pk) else 1400.0
    rec["tensor_peak"] = {"bf16_tflops_sustained_measured": bf16_sus, "tf32_tflops_assumed": bf16_sus / 2,
                          "note": "kind::tf32 runs at half the bf16 rate; MEASURED_PEAKS.json holds no tf32 figure"}
    rec["tensor_frac_eval_3xtf32"] = rec["tensor_tflops_eval_3xtf32"] / (bf16_sus / 2)
    rec["tensor_frac_eval_tf32"] = rec["tensor_tflops_eval_tf32"] / (bf16_sus / 2)
    slab = 4.0 * B * T * V
    rec["hbm_bytes_per_step"] = {
        "unfused_logits_traffic": {"write_logits_by_cublas": slab, "read_logits_by_sweep": 4.0 * V * sum_T,
                                   "write_grad_by_sweep": slab, "read_grad_by_two_backward_gemms": 2 * slab},
        "fused_logits_traffic": {"write_dlogits_by_pass2": slab, "read_dlogits_by_two_backward_gemms": 2 * slab},
        "saved": slab + 4.0 * V * sum_T,
        "note": "the fusion removes the logits write and the logits read (the gradient buffer and the two library "
                "backward GEMMs over it stay); the operand split of 3xtf32 adds 3 x 4*B*T*K bytes"}
    rec["speedup_train"] = ms["unfused_train"] / ms["fused_train_3xtf32"]
    rec["speedup_eval"] = ms["unfused_eval"] / ms["fused_eval_3xtf32"]
    del enc, W, bias
    torch.cuda.empty_cache()
    return rec


def c5_record(torch, dist, world, rank, dev, K=5):
    """BASELINE config C5: joint CTC/attention step with the reference's 12-layer encoder architecture (PyTorch,
    fp32) feeding the CTC kernels, B=128/GPU, T=400; under DistributedDataParallel (DistributedWrapper) at N>1.
    Reports the step time and the CTC branch's share of it (step with the CTC branch minus step without)."""
    import torch.nn.functional as F
    from asr_chinese_e2e_b200 import DistributedWrapper, JointCTCAttention, Pack
    from asr_chinese_e2e_b200.speech_encoder import SpeechEncoder
    from oracle.synth import make_lengths, make_targets
    B, T, V, U, D = 128, 400, V_, 50, 512

    class Dec(torch.nn.Module):                 # a light attention-branch stand-in (the decoder body is outside this path)
        def __init__(self):
            super().__init__()
            self.emb, self.out = torch.nn.Embedding(V, D), torch.nn.Linear(D, V)

        def forward(self, tgt, enc, lens):
            n = tgt.size(0)
            ys = torch.cat([torch.full((n, 1), 2, device=tgt.device), tgt], 1)
            gold = torch.cat([tgt, torch.zeros(n, 1, dtype=torch.long, device=tgt.device)], 1)
            gold[torch.arange(n, device=tgt.device), lens] = 3
            return self.out(self.emb(ys) + enc.mean(1, keepdim=True)), gold

    class Model(JointCTCAttention, torch.nn.Module):
        def __init__(self):
            torch.nn.Module.__init__(self)
            self.encoder, self.decoder = SpeechEncoder(n_layers=12, dropout=0.1), Dec()
            self.init_ctc(D, V, ctc_weight=0.3, ctc_zero_infinity=True)
            self.use_ctc = True

        def joint_loss(self, output, input):
            if self.use_ctc:
                return JointCTCAttention.joint_loss(self, output, input)
            att = F.cross_entropy(output.pred.reshape(-1, V), output.gold.reshape(-1), ignore_index=0)
            att = att + 0.0 * output.ctc_logits.sum()          # keep the head in the graph (same GEMMs, no CTC)
            return att, att.detach(), att

    torch.manual_seed(1005)
    g = torch.Generator().manual_seed(1005 + rank)
    tg, tl = make_targets(B, U, V, g)
    il = make_lengths(B, T, g)
    batch = Pack(wave=torch.randn(B, T, 320, generator=g), wave_len=il, tgt_for_input=tg, tgt_len=tl).to(dev)
    model = DistributedWrapper(Model(), dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, betas=(0.9, 0.98), eps=1e-9)
    res = {}
    for kind in ("with_ctc", "with_ctc_fused_head", "without_ctc"):
        model.module.use_ctc = kind != "without_ctc"
        model.module.ctc_fused_head = kind == "with_ctc_fused_head"
        ms, met = time_steps(torch, lambda: model.iterate(batch, optimizer=opt, is_train=True)[0], K, 2)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        res[kind] = ms
        if kind == "with_ctc":
            loss = float(met.loss.item())
    del model, opt, batch
    torch.cuda.empty_cache()
    return {"workload": f"C5: joint CTC/attention step (ctc_weight 0.3), 12-layer encoder (d_model 512, 8 heads, FFN 1024) "
                        f"in PyTorch fp32 + CTC head + this repo's CTC / CE / CER kernels, B={B}/GPU, T={T}, V={V}; "
                        + ("DistributedDataParallel via DistributedWrapper" if world > 1 else "single GPU"),
            "ms_per_step": res["with_ctc"], "ms_per_step_fused_head": res["with_ctc_fused_head"],
            "ms_per_step_without_ctc_branch": res["without_ctc"],
            "note": "the encoder (fp32 cuBLAS / SDPA, ~200 ms) dominates and the step-to-step spread is ~1 ms, so the CTC "
                    "branch's share measured as a difference is within the noise; configs.fused_head times the branch "
                    "in isolation",
            "ctc_branch_ms": res["with_ctc"] - res["without_ctc"],
            "ctc_share": (res["with_ctc"] - res["without_ctc"]) / res["with_ctc"],
            "value": world * B / (res["with_ctc"] / 1e3), "unit": UNIT, "steps": K, "loss": loss}


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

    c = make_batch(0 if args.same_seed else rank, args.lengths)
    if world > 1 and rank > 0 and not args.same_seed and not args.per_rank_lengths:
        # weak scaling = FIXED work per GPU as N grows: every rank sweeps its own logits, but with rank 0's lengths and
        # transcripts, so that sum(T_b) -- the bytes a rank moves -- does not depend on the rank (with per-rank lengths
        # the step is as slow as the rank that happened to draw the longest utterances: +14 us at N = 8)
        c0 = make_batch(0, args.lengths)
        for key in ("targets", "input_lengths", "target_lengths"):
            c[key] = c0[key]
        del c0
    x = c["logits"].to(dev).requires_grad_(True)
    tg, il, tl = c["targets"].to(dev), c["input_lengths"].to(dev), c["target_lengths"].to(dev)
    sum_T = int(c["input_lengths"].sum())
    bytes_3sweep = 4 * V_ * (2 * sum_T + B_ * T_)               # BASELINE.md section 3, primary figure
    bytes_2sweep = 4 * V_ * (sum_T + B_ * T_)                   # what this implementation moves: read once, write once

    def fwd():
        if world == 1:
            return ctc_loss_b200(x, tg, il, tl, blank=0, reduction="mean", zero_infinity=False)
        return sharded_ctc_loss(x, tg, il, tl, blank=0, zero_infinity=False)   # 1 all-reduce of one float

    def step():
        x.grad = None
        loss = fwd()
        loss.backward()
        return loss

    ev = lambda: torch.cuda.Event(enable_timing=True)
    K = args.steps
    W = max(args.warmup, 3)
    # ---- headline: K steps of the public op ----
    with ClockSampler(local) as clocks:
        clocks.wait_first_sample()
        load_t0 = time.time()
        for _ in range(W):
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

    # ---- per-kernel timing for the roofline (live, CUDA events on the launching stream): one un-chunked call; the
    #      library's sweep_done event splits it into [k0_prep + fused sweep kernel] and [lattice + sparse patch] ----
    from asr_chinese_e2e_b200.profiling import time_stages, time_sweep_kernel
    st = time_stages(x, tg, il, tl, reduction="mean", zero_infinity=False, iters=min(K, 50), warmup=3)
    sweep_ms, rest_ms = statistics.mean(st["sweep_ms"]), statistics.mean(st["rest_ms"])
    k1_ms = time_sweep_kernel(x, tg, il, tl, reduction="mean", zero_infinity=False, iters=min(K, 50), warmup=3)

    launches_per_step = 5                          # k0_prep, k1p_sweep<FUSED>, k2_lattice, k3p_patch, k4_rescale (early exit)
    peak, peak_src = peaks()
    k1f_gbs = bytes_2sweep / (k1_ms / 1e3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "k1f_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_launch", {}).get(args.lengths)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "b200",
            "config": workload_cfg(args, world), "loss": loss_val,
            "gpu_launches": launches_per_step * K,
            "roofline": {"bound": "hbm", "kernel": "k1p_sweep<128,17,FUSED,BULKST> (the fused sweep: log-softmax stats + label gather + dense "
                                                   "gradient; aligned two-frame groups through a bulk-TMA ring, in and out), "
                                                   "timed by itself: CUDA events around its launch on the launching stream "
                                                   "(stage-split ABI call: k0_prep | event | sweep | event)",
                         "achieved": k1f_gbs, "peak": peak, "unit": "GB/s", "frac": k1f_gbs / peak, "traffic": traffic,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_2sweep,
                         "ms_per_launch": k1_ms, "ms_prep_plus_sweep_in_one_call": sweep_ms},
            "roofline_step": {"algorithmic_bytes_3sweep": bytes_3sweep, "algorithmic_bytes_2sweep": bytes_2sweep,
                              "achieved_vs_3sweep": bytes_3sweep / (ms_step / 1e3) / 1e9,
                              "achieved_vs_2sweep": bytes_2sweep / (ms_step / 1e3) / 1e9,
                              "frac": bytes_3sweep / (ms_step / 1e3) / 1e9 / peak,
                              "frac_2sweep": bytes_2sweep / (ms_step / 1e3) / 1e9 / peak,
                              "peak": peak, "unit": "GB/s",
                              "note": "BASELINE.md's primary figure is the 3-sweep byte count; this implementation "
                                      "needs only 2 sweeps (+ a sparse correction): frac_2sweep is the share of HBM "
                                      "bandwidth the step really uses",
                              "sweep_ms": sweep_ms, "lattice_plus_patch_ms": rest_ms},
            "lattice_stats": {"utterances": B_, "log_space": st["lattice_stats"][0],
                              "of_which_underflowed_in_linear_domain": st["lattice_stats"][1]},
            "clocks": clock_summary}

    # ---- parity gate on the very tensors that were timed ----
    if not args.no_parity:
        nll = ctc_loss_b200(x.detach(), tg, il, tl, reduction="none", zero_infinity=False)
        step()
        n_par = B_ if world == 1 else 32             # N>1: a bounded per-rank sample (every rank shares the host cores)
        par = parity_check(torch, c, nll, x.grad, n_par, inv_batch=1.0 / B_)
        if world > 1:
            # global loss against a host-side float64 sum over every rank's per-utterance nll
            allnll = [torch.empty_like(nll) for _ in range(world)]
            alltl = [torch.empty_like(tl) for _ in range(world)]
            dist.all_gather(allnll, nll); dist.all_gather(alltl, tl)
            host = (torch.cat(allnll).double().cpu() / torch.cat(alltl).clamp(min=1).double().cpu()).mean().item()
            t = torch.tensor([par["nll_rel_max"], par["grad_abs_max"], 0.0 if par["pass"] else 1.0], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            par.update(nll_rel_max=float(t[0]), grad_abs_max=float(t[1]), n_utts=n_par * world,
                       global_loss=loss_val, global_loss_host_sum=host,
                       global_loss_rel=abs(loss_val - host) / abs(host))
            par["pass"] = bool(t[2].item() == 0.0 and par["global_loss_rel"] <= REL_LOSS)
        line["parity"] = par

    # ---- N>1: what the collective costs (same shard, same kernels, no all-reduce) ----
    if world > 1:
        def local_step():
            x.grad = None
            loss = ctc_loss_b200(x, tg, il, tl, blank=0, reduction="mean", zero_infinity=False, inv_batch=1.0 / (world * B_))
            loss.backward()
            return loss
        kk = min(K, 100)
        dist.barrier()
        ms_local, _ = time_steps(torch, local_step, kk, 3)
        dist.barrier()
        ms_coll, _ = time_steps(torch, step, kk, 3)
        t = torch.tensor([ms_local, ms_coll], device=dev)
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tmin = t.clone(); dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        line["comm"] = {"ms_per_step_without_collective_max_rank": float(tmax[0]),
                        "ms_per_step_without_collective_min_rank": float(tmin[0]),
                        "ms_per_step_with_collective_max_rank": float(tmax[1]),
                        "comm_us": 1e3 * float(tmax[1] - tmax[0]),
                        "slowest_rank_us": 1e3 * float(tmax[0] - tmin[0]),
                        "note": "comm_us = step with the all-reduce minus the same step without it (max over ranks); "
                                "slowest_rank_us = spread of the collective-free step across ranks (each rank draws its "
                                "own lengths only with --per-rank-lengths)", "same_seed": bool(args.same_seed),
                        "per_rank_lengths": bool(args.per_rank_lengths)}

    if not args.no_e2e:
        e = run_e2e(torch, c, args, dev)
        if world > 1:
            t = torch.tensor([e["ms_per_step"]], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e["ms_per_step"] = float(t.item())
            e["value"] = world * B_ / (e["ms_per_step"] / 1e3)
            nb = torch.tensor([e["h2d_bytes_per_step"], e["d2h_bytes_per_step"]], device=dev, dtype=torch.float64)
            dist.all_reduce(nb, op=dist.ReduceOp.SUM)              # each rank copies its own valid frames
            e["h2d_bytes_per_step"], e["d2h_bytes_per_step"] = int(nb[0].item()), int(nb[1].item())
        line["e2e"] = e
        if world == 1:
            # the training-shaped variant: the gradient's consumer is the next GPU kernel, only the loss goes back
            line["e2e_grad_on_device"] = run_e2e(torch, c, args, dev, grad_to_host=False)
            line["e2e_all_frames"] = run_e2e(torch, c, args, dev, valid_frames_only=False)

    if not args.no_configs:
        cfgs = {}
        if world == 1:
            from oracle.synth import make_config
            Kc = max(10, min(K, 50))
            cfgs["C2_full"] = sub_record(torch, "C2_full", make_batch(0, "full"), False, Kc, peak, 0,
                                         "C2 with all lengths = T (the headline roofline point of BASELINE.md section 3)")
            cfgs["C2_D2_sharp"] = sub_record(torch, "C2_D2", make_batch(0, args.lengths, dist="D2"), False, Kc, peak, 64,
                                             "C2 with trained-like sharp posteriors (D2: randn + 8*onehot(alignment))")
            x = None
            torch.cuda.empty_cache()
            cfgs["C4"] = sub_record(torch, "C4", make_config("C4"), True, Kc, peak, 64,
                                    "C4: B=64, T=1500, U<=120, zero_infinity=True, 8 infeasible + 8 partial-lattice utterances")
        x = None
        torch.cuda.empty_cache()
        if world == 1:
            cfgs["fused_head"] = fused_head_record(torch, dev, peak)
        cfgs["C5"] = c5_record(torch, dist, world, rank, dev)
        line["configs"] = cfgs

    if rank == 0 and world == 1:
        line["torch_cuda_baseline"] = run_torch_cuda(torch, c, dev)
        if not args.no_cpu:
            cb = cpu_reference_line(args, c, steps=5, warmup=2)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "cpu_model",
                                                        "ms_per_step", "one_thread")}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_torch_cuda(torch, c, dev):
    """Secondary GPU baseline: torch's own CUDA log_softmax + ctc_loss + backward on the same
    B200 and inputs -- what the reference as written would execute once a CTC call were added."""
    import torch.nn.functional as F
    x = c["logits"].to(dev).requires_grad_(True)
    tg, il, tl = c["targets"].to(dev), c["input_lengths"].to(dev), c["target_lengths"].to(dev)

    def step():
        x.grad = None
        F.ctc_loss(F.log_softmax(x, -1).transpose(0, 1), tg, il, tl, blank=0, reduction="mean",
                   zero_infinity=False).backward()
    ms, _ = time_steps(torch, step, 10, 3)
    return {"value": B_ / (ms / 1e3), "unit": UNIT, "ms_per_step": ms,
            "what": f"torch {torch.__version__} CUDA F.log_softmax + F.ctc_loss(mean) + backward, same inputs, 10 steps"}


def run_e2e(torch, c, args, dev, grad_to_host=True, valid_frames_only=True):
    """Host-buffer API: pinned logits -> device, kernels, gradient + nll -> pinned host, every step."""
    from asr_chinese_e2e_b200.host_pipeline import HostCTCPipeline
    pipe = HostCTCPipeline(B_, T_, V_, U_, chunk=16, device=dev, grad_to_host=grad_to_host,
                           valid_frames_only=valid_frames_only)
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
    h2d, d2h, launches = pipe.h2d_bytes, pipe.d2h_bytes, pipe.launches_per_step
    pipe.close()
    del pipe
    torch.cuda.empty_cache()
    return {"value": B_ / (ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": d2h, "ms_per_step": ms, "steps": k, "loss": loss,
            "api": "asr_chinese_e2e_b200.host_pipeline.HostCTCPipeline (pinned host logits in; "
                   + ("grad[B,T,V] + " if grad_to_host else "gradient left on the device, ")
                   + "nll[B] back to pinned host; 3-stream chunked pipeline, chunk=16 utterances; "
                   + ("only the valid frames (t < input_lengths[b]) cross PCIe, the padded gradient rows are zeroed "
                      "on the host inside the timed call" if valid_frames_only else "whole [B,T,V] tensors cross PCIe")
                   + ")",
            "valid_frames_only": bool(valid_frames_only), "gpu_launches_per_step": launches}


if __name__ == "__main__":
    main()
