"""End-to-end CTC loss+grad for HOST-resident batches (pinned memory in, pinned memory out).

This is the call a host-side user of the op makes when the logits are not already on the device:
the batch is cut into utterance chunks and pushed through a pipeline

    logits of chunk i+1 over PCIe  ||  prep + fused sweep + lattice + sparse patch kernels (chunk i)  ||  D2H(grad chunk i-1)

so PCIe traffic in both directions overlaps the kernels.  All arithmetic is the same C-ABI calls
(ctcb200_loss_grad, the two-sweep path) as the autograd op; torch supplies pinned/device memory,
streams and events.  Utterances are independent, so chunking does not change any result bit.

``valid_frames_only`` (default): the step is PCIe-bound (two 1.73 GB transfers at C2 against 0.7 ms of kernels), and
the kernels never read a padded frame (t >= input_lengths[b]) nor write anything but zeros there -- so only the valid
frames of every utterance cross the bus, in both directions, and the padded rows of the host gradient buffer are
zeroed by host threads (inside the call) while the transfers are in flight.  With AISHELL-like lengths ~U[T/2, T]
that is 25 % fewer bytes each way.  The lengths are read from the caller's HOST tensor; nothing syncs with the
device for it.

``zero_copy_logits`` (opt-in experiment): no H2D copy of the logits at all -- the fused sweep kernel is handed the
address of the caller's PINNED host buffer and its bulk-TMA row ring pulls the valid frames straight over PCIe
(unified addressing: a pinned allocation is addressable from the device under the same pointer).  Bit-identical
results; measured on B200 (tools/zerocopy_probe.py, tools/e2e_probe.py): 49.7 GB/s through the kernel against 52 GB/s
for the copy engine when nothing else is on the bus (gradient left on the device: 26.3 vs 25.0 ms per C2 step), but
with the gradient's D2H copies running at the same time the SM-issued reads starve (44 ms against 37 ms per step with
both directions on the copy engines) -- so the staged copy stays the default.  Needs page-locked logits
(``tensor.pin_memory()`` / ``cudaHostRegister``).
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor

import torch

from . import _lib


class HostCTCPipeline:
    def __init__(self, B, T, V, Umax, chunk=32, device="cuda", blank=0, zero_infinity=False, n_slots=3,
                 grad_to_host=True, valid_frames_only=True, zero_copy_logits=False):
        self.B, self.T, self.V, self.Umax = B, T, V, Umax
        self.chunk = min(chunk, B)
        self.blank, self.zi = int(blank), int(bool(zero_infinity))
        self.dev = torch.device(device)
        self.n_slots = n_slots
        self.grad_to_host = grad_to_host          # False: the gradient stays on the device (self.g_full)
        self.valid_frames_only = bool(valid_frames_only)
        self.zero_copy_logits = bool(zero_copy_logits)
        # host threads that zero the padded rows of the caller's gradient buffer (numpy fills release the GIL) while
        # the calling thread keeps enqueueing copies and kernels
        self._pool = (ThreadPoolExecutor(max_workers=max(2, min(16, (os.cpu_count() or 4) // 2)))
                      if (self.valid_frames_only and grad_to_host) else None)
        c = self.chunk
        self.ws_bytes = _lib.workspace_bytes(c, T, V, Umax)
        self.x = None if self.zero_copy_logits else [torch.empty(c, T, V, device=self.dev) for _ in range(n_slots)]
        self.g_full = None if grad_to_host else torch.empty(B, T, V, device=self.dev)
        self.g = [torch.empty(c, T, V, device=self.dev) for _ in range(n_slots)] if grad_to_host else None
        self.ws = [torch.empty(self.ws_bytes, dtype=torch.uint8, device=self.dev) for _ in range(n_slots)]
        self.tg = torch.empty(B, max(Umax, 1), dtype=torch.int64, device=self.dev)
        self.il = torch.empty(B, dtype=torch.int64, device=self.dev)
        self.tl = torch.empty(B, dtype=torch.int64, device=self.dev)
        self.nll = torch.empty(B, device=self.dev)
        self.one = torch.ones((), device=self.dev)
        self.s_in, self.s_cmp, self.s_out = (torch.cuda.Stream(self.dev) for _ in range(3))
        self.ev_in = [torch.cuda.Event() for _ in range(n_slots)]
        self.ev_cmp = [torch.cuda.Event() for _ in range(n_slots)]
        self.ev_out = [torch.cuda.Event() for _ in range(n_slots)]
        # bytes that crossed PCIe in the LAST call (they depend on the lengths); before the first call: the full batch
        self.h2d_bytes = B * T * V * 4 + self.tg.numel() * 8 + 2 * B * 8
        self.d2h_bytes = (B * T * V * 4 if grad_to_host else 0) + B * 4
        self.launches_per_step = 4 * ((B + self.chunk - 1) // self.chunk)

    def close(self):
        """Stops the host threads that zero the padded gradient rows (idempotent)."""
        if self._pool is not None:
            self._pool.shutdown(wait=True)
            self._pool = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _runs(lens, lo, hi, T):
        """Utterances [lo, hi) as (first, end, frames) copy operations: a run of full-length utterances is contiguous
        in [B,T,V] and crosses the bus as ONE copy; a shorter utterance is a copy of its first ``frames`` frames."""
        b = lo
        while b < hi:
            if lens[b] == T:
                e = b + 1
                while e < hi and lens[e] == T:
                    e += 1
                yield b, e, T
                b = e
            else:
                yield b, b + 1, lens[b]
                b += 1

    def __call__(self, h_logits, h_targets, h_il, h_tl, h_grad, h_nll, reduction="mean"):
        """h_* are pinned CPU tensors; h_grad [B,T,V] and h_nll [B] are written. Returns after sync."""
        L = _lib.lib()
        B, T, V, c = self.B, self.T, self.V, self.chunk
        red = {"none": 0, "mean": 1, "sum": 2}[reduction]
        vfo, zc, to_host = self.valid_frames_only, self.zero_copy_logits, self.grad_to_host
        if zc and not (h_logits.is_pinned() and h_logits.is_contiguous() and h_logits.dtype == torch.float32):
            raise _lib.CtcB200Error("zero_copy_logits needs a contiguous float32 PINNED host tensor "
                                    "(tensor.pin_memory()); use zero_copy_logits=False for pageable memory")
        lens = [min(max(int(v), 0), T) for v in h_il.tolist()]          # host tensor: no device sync
        cur = torch.cuda.current_stream(self.dev)
        for s in (self.s_in, self.s_cmp, self.s_out):
            s.wait_stream(cur)
        with torch.cuda.stream(self.s_in):
            self.tg.copy_(h_targets, non_blocking=True)
            self.il.copy_(h_il, non_blocking=True)
            self.tl.copy_(h_tl, non_blocking=True)
        if zc:
            self.s_cmp.wait_stream(self.s_in)                            # targets / lengths
        g_np = h_grad.numpy() if (vfo and to_host) else None             # zero-copy view of the pinned buffer
        pending = []

        def zero_padded(b0, b1):
            for b in range(b0, b1):
                if lens[b] < T:
                    g_np[b, lens[b]:].fill(0.0)

        row_bytes = V * 4
        h2d = self.tg.numel() * 8 + 2 * B * 8
        d2h = B * 4
        n_chunks = (B + c - 1) // c
        for i in range(n_chunks):
            k = i % self.n_slots
            lo, hi = i * c, min((i + 1) * c, B)
            n = hi - lo
            # ---- logits of the chunk: read in place by the sweep kernel (zero copy), or staged by the copy engine ----
            if zc:
                h2d += sum(lens[lo:hi]) * row_bytes                     # the sweep only ever reads valid frames
            else:
                with torch.cuda.stream(self.s_in):
                    if i >= self.n_slots:
                        self.s_in.wait_event(self.ev_cmp[k])            # slot's previous kernels are done
                    if vfo:
                        for b0, b1, nt in self._runs(lens, lo, hi, T):
                            if nt == T:
                                self.x[k][b0 - lo:b1 - lo].copy_(h_logits[b0:b1], non_blocking=True)
                            elif nt > 0:
                                self.x[k][b0 - lo, :nt].copy_(h_logits[b0, :nt], non_blocking=True)
                            h2d += (b1 - b0) * nt * row_bytes
                    else:
                        self.x[k][:n].copy_(h_logits[lo:hi], non_blocking=True)
                        h2d += n * T * row_bytes
                    self.ev_in[k].record(self.s_in)
            # ---- kernels ----
            with torch.cuda.stream(self.s_cmp):
                if not zc:
                    self.s_cmp.wait_event(self.ev_in[k])
                if i >= self.n_slots and to_host:
                    self.s_cmp.wait_event(self.ev_out[k])               # slot's previous grad has left
                gdst = self.g[k] if to_host else self.g_full[lo:hi]
                xptr = h_logits[lo:hi].data_ptr() if zc else self.x[k].data_ptr()
                st = self.s_cmp.cuda_stream
                _lib.check(L.ctcb200_loss_grad(xptr, self.tg[lo:hi].data_ptr(), self.tg.shape[1],
                                               n * self.tg.shape[1], self.il[lo:hi].data_ptr(),
                                               self.tl[lo:hi].data_ptr(), n, T, V, self.Umax, self.blank, self.zi,
                                               red, 1.0 / B, self.nll[lo:hi].data_ptr(), None, gdst.data_ptr(),
                                               self.ws[k].data_ptr(), self.ws_bytes, st, None), "ctcb200_loss_grad")
                self.ev_cmp[k].record(self.s_cmp)
            # ---- gradient of the chunk back to the host ----
            if to_host:
                with torch.cuda.stream(self.s_out):
                    self.s_out.wait_event(self.ev_cmp[k])
                    if vfo:
                        for b0, b1, nt in self._runs(lens, lo, hi, T):
                            if nt == T:
                                h_grad[b0:b1].copy_(self.g[k][b0 - lo:b1 - lo], non_blocking=True)
                            elif nt > 0:
                                h_grad[b0, :nt].copy_(self.g[k][b0 - lo, :nt], non_blocking=True)
                            d2h += (b1 - b0) * nt * row_bytes
                    else:
                        h_grad[lo:hi].copy_(self.g[k][:n], non_blocking=True)
                        d2h += n * T * row_bytes
                    self.ev_out[k].record(self.s_out)
                if vfo:
                    # the padded rows of the host buffer: zeros written by host threads (disjoint from the rows the DMA
                    # writes) while the copies are in flight
                    step = max(1, (n + 3) // 4)
                    if self._pool is None:                           # (closed pipeline reused: zero inline)
                        zero_padded(lo, hi)
                    else:
                        for b0 in range(lo, hi, step):
                            pending.append(self._pool.submit(zero_padded, b0, min(b0 + step, hi)))
        self.h2d_bytes, self.d2h_bytes = h2d, d2h
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_stream(self.s_cmp)
            h_nll.copy_(self.nll, non_blocking=True)
        cur.wait_stream(self.s_out)
        cur.wait_stream(self.s_in)
        for f in pending:
            f.result()
        cur.synchronize()
        return h_nll, h_grad


def ctc_loss_grad_host(logits, targets, input_lengths, target_lengths, blank=0, reduction="mean",
                       zero_infinity=False, chunk=32, pipeline=None):
    """Convenience wrapper: CPU tensors in (pinned on the fly if needed), (loss, nll, grad) CPU tensors out."""
    B, T, V = logits.shape
    Umax = targets.shape[1] if targets.dim() == 2 else int(target_lengths.max())
    p = pipeline or HostCTCPipeline(B, T, V, Umax, chunk=chunk, blank=blank, zero_infinity=zero_infinity)
    pin = lambda t: t if t.is_pinned() else t.contiguous().pin_memory()
    h_grad = torch.empty(B, T, V, pin_memory=True)
    h_nll = torch.empty(B, pin_memory=True)
    tg = targets if targets.dim() == 2 else _pad_1d(targets, target_lengths, Umax)
    p(pin(logits.to(torch.float32)), pin(tg.to(torch.int64)), pin(input_lengths.to(torch.int64)),
      pin(target_lengths.to(torch.int64)), h_grad, h_nll, reduction)
    tl = target_lengths.clamp(min=1).to(torch.float32)
    loss = {"none": h_nll, "sum": h_nll.sum(), "mean": (h_nll / tl).mean()}[reduction]
    return loss, h_nll, h_grad


def _pad_1d(cat, tl, Umax):
    out = torch.zeros(len(tl), max(Umax, 1), dtype=torch.int64)
    off = 0
    for b, u in enumerate(tl.tolist()):
        out[b, :u] = cat[off:off + u]
        off += u
    return out
