"""Python boundary of the B200 CTC op: ``ctc_loss_b200`` / ``CTCLossB200``.

Mirrors ``torch.nn.functional.ctc_loss`` (blank / reduction / zero_infinity semantics, padded
2-D or concatenated 1-D targets, int32/int64 lengths) but takes the reference's batch-major
``[B, T, V]`` LOGITS (Predictor/Utils/loss.py:10 "pred: N x T x C", the convention of
``cal_performance``), i.e. it is

    F.ctc_loss(F.log_softmax(logits, -1).transpose(0, 1), targets, input_lengths,
               target_lengths, blank, reduction, zero_infinity)

differentiable w.r.t. ``logits`` only.  It plugs in next to ``cal_performance`` inside
``TransformerOffical.cal_metrics`` (Predictor/Models/transformer_official.py:83-94); see
``asr_chinese_e2e_b200.joint``.  All arithmetic runs in libctcb200.so (sm_100a CUDA); torch
provides device memory, the current stream and autograd plumbing only.
"""
from __future__ import annotations

import math
import os

import torch

from . import _lib

_RED = {"none": 0, "mean": 1, "sum": 2}

# Host-side developer knobs: read from the environment ONCE, at import; `configure()` changes them afterwards.
_CFG = {
    "debug": bool(int(os.environ.get("CTCB200_DEBUG", "0"))),        # read the device status word after every call (syncs)
    "fused": bool(int(os.environ.get("CTCB200_FUSED", "1"))),        # gradient produced inside the forward call
    "two_sweep": bool(int(os.environ.get("CTCB200_TWO_SWEEP", "1"))),
    "split": float(os.environ.get("CTCB200_SPLIT", "1.0")),          # two-stream utterance split (lost: DESIGN.md 5)
    "chunks": int(os.environ["CTCB200_CHUNKS"]) if "CTCB200_CHUNKS" in os.environ else None,
    "lattice_log": False,                                            # CTCB200_FLAG_LATTICE_LOG on every call
}


def configure(**kw):
    """Change a host-side developer knob (debug, fused, two_sweep, split, chunks, lattice_log).
    Returns the previous values."""
    old = {k: _CFG[k] for k in kw}
    for k, v in kw.items():
        if k not in _CFG:
            raise KeyError(k)
        _CFG[k] = v
    return old


def _validate_host(name, x, lo, hi):
    """Lengths that are still on the host are checked for free, with F.ctc_loss's error behaviour (it raises);
    device-resident ones are checked by the kernels, which poison the result with NaN (no host sync)."""
    if torch.is_tensor(x):
        if x.is_cuda or x.numel() == 0:
            return
        mn, mx = int(x.min()), int(x.max())
    else:
        seq = list(x)
        if not seq:
            return
        mn, mx = int(min(seq)), int(max(seq))
    if mn < lo or mx > hi:
        raise ValueError(f"{name} must lie in [{lo}, {hi}] (got min {mn}, max {mx})")


def _as_i64_cuda(x, device):
    if not torch.is_tensor(x):
        x = torch.as_tensor(x)
    return x.to(device=device, dtype=torch.int64, non_blocking=True).contiguous()


def _prepare(logits, targets, input_lengths, target_lengths, blank, max_target_length):
    if not (torch.is_tensor(logits) and logits.is_cuda):
        raise _lib.CtcB200Error("ctc_loss_b200 needs CUDA logits: the hot path has no CPU fallback")
    if logits.dtype != torch.float32:
        raise _lib.CtcB200Error(f"logits must be float32 (got {logits.dtype}); the path computes in f32")
    if logits.dim() != 3:
        raise ValueError("logits must be [B, T, V] (batch-major, as in the reference's loss API)")
    x = logits.contiguous()
    if x.data_ptr() % 16:
        x = x.clone()
    B, T, V = x.shape
    dev = x.device
    _validate_host("input_lengths", input_lengths, 0, T)
    if torch.is_tensor(targets) and targets.dim() == 2:
        _validate_host("target_lengths", target_lengths, 0, targets.shape[1])
    if torch.is_tensor(targets) and not targets.is_cuda and targets.numel():
        if int(targets.min()) < 0 or int(targets.max()) >= V:
            raise ValueError(f"targets must lie in [0, {V})")
    il = _as_i64_cuda(input_lengths, dev)
    tl = _as_i64_cuda(target_lengths, dev)
    tg = _as_i64_cuda(targets, dev)
    if il.numel() != B or tl.numel() != B:
        raise ValueError("input_lengths and target_lengths must have B elements")
    if tg.dim() == 2:
        if tg.shape[0] != B:
            raise ValueError("2-D targets must be [B, Umax]")
        umax, stride = tg.shape[1], tg.shape[1]
        if umax > 255:   # wider padding than the lattice kernel supports: use the true maximum
            umax = int(max_target_length if max_target_length is not None else tl.max().item())
        if stride == 0:
            tg, stride = tg.new_zeros(B, 1), 1
    elif tg.dim() == 1:
        stride = 0
        if max_target_length is not None:
            umax = int(max_target_length)
        elif torch.is_tensor(target_lengths) and not target_lengths.is_cuda:
            umax = int(target_lengths.max()) if B else 0
        else:
            umax = int(tl.max().item()) if B else 0    # one host sync; pass max_target_length to avoid it
        if tg.numel() == 0:
            tg = tg.new_zeros(1)
    else:
        raise ValueError("targets must be [B, Umax] or 1-D concatenated")
    return x, tg, stride, il, tl, B, T, V, int(umax)


def _check_status(ws, stream):
    import ctypes
    st = ctypes.c_int(0)
    _lib.check(_lib.lib().ctcb200_read_status(ws.data_ptr(), ctypes.byref(st), stream), "ctcb200_read_status")
    if st.value:
        raise _lib.CtcB200Error(f"invalid CTC inputs, device status word = {st.value} "
                                "(1: input length, 2: target length, 4: label)")


# Side streams for the chunk pipeline (torch supplies streams/events; one pair per device).
_SIDE = {}


def _side_streams(dev):
    key = (dev.type, dev.index)
    if key not in _SIDE:
        # high priority: the few lattice CTAs must be placed ahead of the thousands of queued sweep/patch CTAs
        _SIDE[key] = (torch.cuda.Stream(dev, priority=-1), torch.cuda.Stream(dev, priority=-1))
    return _SIDE[key]


def _event_handle(ev):
    """Raw cudaEvent_t of a torch event (created lazily by torch on first record: force it)."""
    if ev is None:
        return None
    if not ev.cuda_event:
        ev.record()                      # materialises the handle; re-recorded by the library on its stream
    return ev.cuda_event


_ONES = {}


def _ones(dev, n):
    """Read-only ones[n] per device (the 'gradient already applied for upstream 1' marker): no fill kernel per call."""
    key = (dev.type, dev.index, n)
    if key not in _ONES:
        _ONES[key] = torch.ones(n, dtype=torch.float32, device=dev)
    return _ONES[key]


def _n_chunks(B, requested):
    n = requested if requested is not None else (_CFG["chunks"] or 1)
    return max(1, min(int(n), B)) if B else 1


def _decode_chunks(L, decode, tg, stride, chunks, T, V, umax, blank, stream, dev, B):
    """Greedy CTC decode + edit distance from the argmax the sweep left in each chunk's workspace.
    chunks: list of (lo, n, ws_ptr, ws_bytes).  Fills the caller's `decode` dict with device tensors."""
    edit = torch.empty(B, dtype=torch.int32, device=dev)
    hlen = torch.empty(B, dtype=torch.int32, device=dev)
    hyp = torch.empty(B, T, dtype=torch.int64, device=dev) if decode.get("want_hyp", True) else None
    for lo, n, wsp, wsb in chunks:
        _lib.check(L.ctcb200_greedy_decode(tg.data_ptr() + lo * stride * 8, stride, tg.numel() - lo * stride, n, T, V, umax,
                                           blank, wsp, wsb, edit.data_ptr() + lo * 4, hlen.data_ptr() + lo * 4,
                                           hyp.data_ptr() + lo * T * 8 if hyp is not None else None, stream.cuda_stream),
                   "ctcb200_greedy_decode")
    decode.update(edit_distance=edit, hyp_len=hlen, hyp=hyp)


def _two_sweep_pipeline(ctx, L, x, tg, stride, il, tl, B, T, V, umax, blank, zi, red, inv_b, nll, grad, reduction,
                        decode=None, lattice_event=None):
    """Default training path: ONE call (prep, fused sweep, lattice, sparse patch back to back on the caller's
    stream; with ``lattice_event`` split after the lattice so that the caller's collective can start there).

    ``CTCB200_SPLIT=s`` (0 < s < 1, experiment) cuts the batch into two utterance chunks (s / 1-s):

        main stream:  sweep(a)  sweep(b)            patch(a)   patch(b)
        side stream:            lattice(a)  ......  lattice(b)

    so that the lattice of one chunk runs under the sweep of the other.  Measured on B200 this LOSES (the lattice
    doubles its time when it shares the memory system with a sweep), hence the default 1.0."""
    dev = x.device
    m = 4 // math.gcd(T * V, 4)                      # chunk starts must stay 16-byte aligned
    split = _CFG["split"]                            # measured on B200: one chunk wins (profiles/), see DESIGN.md 5
    cut = int(round(B * split / m)) * m
    bounds = [0, B] if (B < 16 or cut <= 0 or cut >= B or stride == 0) else [0, cut, B]
    n_ch = len(bounds) - 1
    ws_bytes = [_lib.workspace_bytes(bounds[c + 1] - bounds[c], T, V, umax) for c in range(n_ch)]
    ws_off = [sum(ws_bytes[:c]) for c in range(n_ch)]
    ws = torch.empty(sum(ws_bytes), dtype=torch.uint8, device=dev)
    sums = torch.empty(n_ch, 4, dtype=torch.float32, device=dev) if B else torch.zeros(n_ch, 4, device=dev)
    xs = 4 * T * V

    def call(c, stages, stream):
        lo, n = bounds[c], bounds[c + 1] - bounds[c]
        _lib.check(L.ctcb200_loss_grad_stages(
            stages, x.data_ptr() + lo * xs, tg.data_ptr() + lo * stride * 8, stride, tg.numel() - lo * stride,
            il.data_ptr() + lo * 8, tl.data_ptr() + lo * 8, n, T, V, umax, blank, zi, red, inv_b,
            nll.data_ptr() + lo * 4, sums.data_ptr() + c * 16, grad.data_ptr() + lo * xs,
            ws.data_ptr() + ws_off[c], ws_bytes[c], stream.cuda_stream), "ctcb200_loss_grad_stages")

    with torch.cuda.device(dev):
        main = torch.cuda.current_stream()
        if n_ch == 1 and lattice_event is not None:
            # the loss value is final after the lattice: let the caller start its collective on it while the
            # sparse patch of the gradient still runs (sharded.py)
            call(0, 3, main)
            lattice_event[0].record(main)
            lattice_event[1] = True
            call(0, 4, main)
        elif n_ch == 1:
            call(0, 7, main)
        else:
            side = _side_streams(dev)[0]
            ev_s = [torch.cuda.Event() for _ in range(n_ch)]
            ev_l = [torch.cuda.Event() for _ in range(n_ch)]
            for c in range(n_ch):
                call(c, 1, main)
                ev_s[c].record(main)
                side.wait_event(ev_s[c])
                call(c, 2, side)
                ev_l[c].record(side)
            for c in range(n_ch):
                main.wait_event(ev_l[c])
                call(c, 4, main)
        if _CFG["debug"]:
            for c in range(n_ch):
                _check_status(ws[ws_off[c]:], main.cuda_stream)
        if decode is not None:
            _decode_chunks(L, decode, tg, stride, [(bounds[c], bounds[c + 1] - bounds[c], ws.data_ptr() + ws_off[c],
                                                    ws_bytes[c]) for c in range(n_ch)], T, V, umax, blank, main, dev, B)
    ctx.cfg = (stride, B, T, V, umax, blank, zi, red, 0, 0, n_ch, inv_b, True)
    ctx.chunk_map = [(bounds[c], bounds[c + 1] - bounds[c], ws_off[c], ws_bytes[c]) for c in range(n_ch)]
    ctx.speculative_used = False
    ctx.save_for_backward(grad, x, tg, ws)
    if reduction == "none":
        return nll
    if n_ch == 1:                                     # the lattice kernel already reduced (and scaled) the batch
        return sums[0, 1] if reduction == "sum" else sums[0, 3]
    col = sums[:, 1] if reduction == "sum" else sums[:, 0]
    return col.sum() if reduction == "sum" else col.sum() * inv_b


class _CTCLossB200Fn(torch.autograd.Function):
    """Forward runs prep + lse/gather sweep + lattice (+ with ``fused``: the gradient sweep, computed
    speculatively for an upstream gradient of 1).  The batch is cut into utterance chunks that
    alternate between two side streams, so the latency-bound lattice of one chunk runs under the
    HBM-bound sweeps of the next; utterances are independent, so chunking changes no result bit."""

    @staticmethod
    def forward(ctx, logits, targets, input_lengths, target_lengths, blank, reduction, zero_infinity,
                inv_batch, max_target_length, fused, chunks, decode, lattice_event=None, grad_mode=True):
        x, tg, stride, il, tl, B, T, V, umax = _prepare(logits, targets, input_lengths, target_lengths,
                                                        blank, max_target_length)
        L = _lib.lib()
        # (needs_input_grad ignores torch.no_grad(): an evaluation pass over tensors that require grad must still take
        #  the loss-only path -- grad_mode is torch.is_grad_enabled() at the call site)
        need_grad = bool(ctx.needs_input_grad[0] and grad_mode)
        fused = bool(need_grad and fused)
        two_sweep = _CFG["two_sweep"]
        red = _RED[reduction]
        zi = (int(bool(zero_infinity)) | (2 if decode is not None else 0)   # bit 1: record per-frame argmax
              | (4 if _CFG["lattice_log"] else 0))                           # bit 2: log-space recursion everywhere
        inv_b = float(inv_batch) if inv_batch is not None else (1.0 / max(B, 1))
        dev = x.device
        nll = torch.empty(B, dtype=torch.float32, device=dev)
        grad = torch.empty_like(x) if fused else None
        if fused and two_sweep and chunks is None and _CFG["chunks"] is None:
            out = _two_sweep_pipeline(ctx, L, x, tg, stride, il, tl, B, T, V, umax, int(blank), zi, red, inv_b,
                                      nll, grad, reduction, decode, lattice_event)
            return out
        n_ch = _n_chunks(B, chunks)
        per = (B + n_ch - 1) // n_ch if B else 0
        # every chunk's slab must start 16-byte aligned: chunk size multiple of 4/gcd(T*V, 4) utterances
        m = 4 // math.gcd(T * V, 4)
        per = (per + m - 1) // m * m
        n_ch = (B + per - 1) // per if B else 1
        ws_bytes = _lib.workspace_bytes(max(per, 1), T, V, umax)
        ws = torch.empty(n_ch * ws_bytes, dtype=torch.uint8, device=dev)
        one = torch.ones((), dtype=torch.float32, device=dev) if fused else None
        sums = torch.zeros(n_ch, 4, dtype=torch.float32, device=dev)   # per chunk: [sum nll/U, sum nll, n, mean]
        fwd = L.ctcb200_forward if need_grad else L.ctcb200_loss_only
        with torch.cuda.device(dev):
            main = torch.cuda.current_stream()
            side = _side_streams(dev) if n_ch > 1 else (main, main)
            if n_ch > 1:
                ev0 = torch.cuda.Event()
                ev0.record(main)
                side[0].wait_event(ev0)
                side[1].wait_event(ev0)
            xs, ts, es = x.element_size() * T * V, 8, 4
            prev_sweep = None
            for c in range(n_ch):
                lo, hi = c * per, min((c + 1) * per, B)
                n = hi - lo
                if n <= 0:
                    break
                st = side[c & 1].cuda_stream
                # stagger: chunk c's sweep starts when chunk c-1's sweep is done, i.e. under c-1's lattice
                sweep_ev = None
                if n_ch > 1:
                    if prev_sweep is not None:
                        side[c & 1].wait_event(prev_sweep)
                    sweep_ev = torch.cuda.Event()
                    prev_sweep = sweep_ev
                tgp = tg.data_ptr() + (lo * stride * ts if stride else 0)
                wsp = ws.data_ptr() + c * ws_bytes
                # 1-D targets: every chunk sees the whole concatenation and needs its own offset base;
                # keep it simple and exact by running 1-D targets as a single chunk (see _prepare)
                args = (x.data_ptr() + lo * xs, tgp, stride, tg.numel() - lo * stride, il.data_ptr() + lo * 8,
                        tl.data_ptr() + lo * 8, n, T, V, umax, int(blank), zi)
                if fused and two_sweep:
                    _lib.check(L.ctcb200_loss_grad(*args, red, inv_b, nll.data_ptr() + lo * es, sums.data_ptr() + c * 16,
                                                   grad.data_ptr() + lo * xs, wsp, ws_bytes, st,
                                                   _event_handle(sweep_ev)), "ctcb200_loss_grad")
                    continue
                _lib.check(fwd(*args, nll.data_ptr() + lo * es, sums.data_ptr() + c * 16, wsp, ws_bytes, st,
                               _event_handle(sweep_ev)),
                           "ctcb200_forward" if need_grad else "ctcb200_loss_only")
                if fused:
                    _lib.check(L.ctcb200_backward(x.data_ptr() + lo * xs, tgp, stride, tg.numel() - lo * stride,
                                                  one.data_ptr(), 0, red, inv_b, n, T, V, umax, int(blank), zi,
                                                  grad.data_ptr() + lo * xs, wsp, ws_bytes, st), "ctcb200_backward")
            if n_ch > 1:
                for s_ in side:
                    e_ = torch.cuda.Event()
                    e_.record(s_)
                    main.wait_event(e_)
            if _CFG["debug"]:
                for c in range(n_ch):
                    _check_status(ws[c * ws_bytes:], main.cuda_stream)
            if decode is not None:
                _decode_chunks(L, decode, tg, stride,
                               [(c * per, min((c + 1) * per, B) - c * per, ws.data_ptr() + c * ws_bytes, ws_bytes)
                                for c in range(n_ch) if c * per < B], T, V, umax, int(blank), main, dev, B)
        ctx.cfg = (stride, B, T, V, umax, int(blank), zi, red, ws_bytes, per, n_ch, inv_b, fused)
        ctx.chunk_map = [(c * per, min((c + 1) * per, B) - c * per, c * ws_bytes, ws_bytes)
                         for c in range(n_ch) if c * per < B]
        ctx.speculative_used = False
        if need_grad:
            if fused:
                ctx.save_for_backward(grad, x, tg, ws)
            else:
                ctx.save_for_backward(x, tg, ws)
        if reduction == "none":
            return nll
        # the lattice kernel's last CTA of each chunk reduced its chunk in a fixed order (deterministic)
        col = sums[:, 1] if reduction == "sum" else sums[:, 0]
        tot = col[0] if n_ch == 1 else col.sum()
        return tot if reduction == "sum" else tot * inv_b

    @staticmethod
    def backward(ctx, grad_out):
        stride, B, T, V, umax, blank, zi, red, ws_bytes, per, n_ch, inv_b, fused = ctx.cfg
        go = grad_out.to(dtype=torch.float32).contiguous()
        L = _lib.lib()
        if fused and not ctx.speculative_used:
            # First backward of this graph: the forward call left the gradient for an upstream gradient of 1 in
            # `grad`.  k4 compares the real upstream gradient with 1 on the device and exits without touching memory
            # when they agree (loss.backward()), or scales the slab by it otherwise.  The buffer is handed to
            # autograd here and never written again: a later backward of a retained graph recomputes (below).
            grad = ctx.saved_tensors[0]
            ctx.speculative_used = True
            with torch.cuda.device(grad.device):
                stream = torch.cuda.current_stream().cuda_stream
                ones = _ones(grad.device, B)
                scratch = torch.empty(B, dtype=torch.float32, device=grad.device)
                _lib.check(L.ctcb200_rescale_grad(grad.data_ptr(), go.data_ptr(), 1 if red == 0 else 0,
                                                  ones.data_ptr(), scratch.data_ptr(), B, T, V, stream),
                           "ctcb200_rescale_grad")
            return (grad,) + (None,) * 13
        x, tg, ws = ctx.saved_tensors[-3:]
        grad = torch.empty_like(x)
        xs = x.element_size() * T * V
        gs = 4 if red == 0 else 0
        with torch.cuda.device(x.device):
            stream = torch.cuda.current_stream().cuda_stream
            for lo, n, wo, wb in ctx.chunk_map:
                tgp = tg.data_ptr() + (lo * stride * 8 if stride else 0)
                _lib.check(L.ctcb200_backward(x.data_ptr() + lo * xs, tgp, stride, tg.numel() - lo * stride,
                                              go.data_ptr() + lo * gs, 1 if red == 0 else 0, red, inv_b, n, T, V,
                                              umax, blank, zi, grad.data_ptr() + lo * xs,
                                              ws.data_ptr() + wo, wb, stream), "ctcb200_backward")
        return (grad,) + (None,) * 13


def ctc_loss_b200(logits, targets, input_lengths, target_lengths, blank: int = 0,
                  reduction: str = "mean", zero_infinity: bool = False, *, inv_batch=None,
                  max_target_length=None, fused=None, chunks=None, decode=None, lattice_event=None):
    """CTC loss on batch-major logits; same flags and results as ``F.ctc_loss`` (see module doc).

    inv_batch: 1/(global batch) for 'mean' when the batch is sharded over ranks (default 1/B).
    max_target_length: upper bound on target_lengths for 1-D targets (avoids one host sync).
    fused: compute the gradient inside the forward call, speculatively for an upstream gradient of
        1, and only rescale it in backward if autograd hands over something else (one empty launch
        otherwise).  This lets the gradient sweep of one chunk overlap the lattice of the next and
        is the default when the logits require grad (env CTCB200_FUSED=0 to disable); costs one
        extra [B,T,V] buffer held until backward, like autograd's own saved log-probs would.
    chunks: utterance chunks of the two-stream pipeline (default env CTCB200_CHUNKS or 1).
    lattice_event: optional ``torch.cuda.Event``; recorded on the current stream as soon as the loss VALUE is final
        (after the lattice kernel, before the sparse gradient patch), so a caller can overlap work that needs only
        the value -- the all-reduce of ``sharded_ctc_loss`` -- with the rest of the call.
    decode: optional dict; if given it is filled with the on-device greedy (best-path) decode of the same
        forward pass -- ``edit_distance`` int32[B] (token-level Levenshtein distance to the targets),
        ``hyp_len`` int32[B], ``hyp`` int64[B,T] (blank-padded; skip with ``decode={"want_hyp": False}``).
    """
    if reduction not in _RED:
        raise ValueError(f"reduction must be one of {list(_RED)}")
    if fused is None:
        fused = _CFG["fused"]
    if torch.is_tensor(targets) and targets.dim() == 1:
        chunks = 1
    ev = [lattice_event, False] if lattice_event is not None else None
    out = _CTCLossB200Fn.apply(logits, targets, input_lengths, target_lengths, blank, reduction,
                               zero_infinity, inv_batch, max_target_length, fused, chunks, decode, ev,
                               torch.is_grad_enabled())
    if ev is not None and not ev[1]:
        lattice_event.record(torch.cuda.current_stream(out.device))   # paths that do not split the call
    return out


def ctc_greedy_cer_b200(logits, targets, input_lengths, target_lengths, blank: int = 0):
    """Best-path CTC decode and character error rate, all on the device (one read sweep of the logits).
    Returns (cer_percent 0-dim tensor = 100 * mean_b(edit_b / max(U_b,1)), info dict as in ``decode``)."""
    info = {}
    with torch.no_grad():
        ctc_loss_b200(logits, targets, input_lengths, target_lengths, blank=blank, reduction="none",
                      zero_infinity=True, decode=info)
    tl = torch.as_tensor(target_lengths).to(info["edit_distance"].device)
    cer = (info["edit_distance"].float() / tl.clamp(min=1).float()).mean() * 100.0
    return cer, info


class CTCLossB200(torch.nn.Module):
    """``nn.CTCLoss``-shaped module over ``ctc_loss_b200`` (batch-major logits in, not log-probs)."""

    def __init__(self, blank: int = 0, reduction: str = "mean", zero_infinity: bool = False):
        super().__init__()
        if reduction not in _RED:
            raise ValueError(f"reduction must be one of {list(_RED)}")
        self.blank, self.reduction, self.zero_infinity = blank, reduction, zero_infinity

    def forward(self, logits, targets, input_lengths, target_lengths):
        return ctc_loss_b200(logits, targets, input_lengths, target_lengths, self.blank, self.reduction,
                             self.zero_infinity)
