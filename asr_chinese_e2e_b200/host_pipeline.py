"""End-to-end CTC loss+grad for HOST-resident batches (pinned memory in, pinned memory out).

This is the call a host-side user of the op makes when the logits are not already on the device:
the batch is cut into utterance chunks and pushed through a 3-stream pipeline

    H2D(logits chunk i+1)  ||  prep + fused sweep + lattice + sparse patch kernels (chunk i)  ||  D2H(grad chunk i-1)

so PCIe traffic in both directions overlaps the kernels.  All arithmetic is the same C-ABI calls
(ctcb200_loss_grad, the two-sweep path) as the autograd op; torch supplies pinned/device memory,
streams and events.  Utterances are independent, so chunking does not change any result bit.
"""
from __future__ import annotations

import torch

from . import _lib


class HostCTCPipeline:
    def __init__(self, B, T, V, Umax, chunk=32, device="cuda", blank=0, zero_infinity=False, n_slots=3,
                 grad_to_host=True):
        self.B, self.T, self.V, self.Umax = B, T, V, Umax
        self.chunk = min(chunk, B)
        self.blank, self.zi = int(blank), int(bool(zero_infinity))
        self.dev = torch.device(device)
        self.n_slots = n_slots
        self.grad_to_host = grad_to_host          # False: the gradient stays on the device (self.g_full)
        c = self.chunk
        self.ws_bytes = _lib.workspace_bytes(c, T, V, Umax)
        self.x = [torch.empty(c, T, V, device=self.dev) for _ in range(n_slots)]
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
        self.h2d_bytes = B * T * V * 4 + self.tg.numel() * 8 + 2 * B * 8
        self.d2h_bytes = (B * T * V * 4 if grad_to_host else 0) + B * 4
        self.launches_per_step = 4 * ((B + self.chunk - 1) // self.chunk)

    def __call__(self, h_logits, h_targets, h_il, h_tl, h_grad, h_nll, reduction="mean"):
        """h_* are pinned CPU tensors; h_grad [B,T,V] and h_nll [B] are written. Returns after sync."""
        L = _lib.lib()
        B, T, V, c = self.B, self.T, self.V, self.chunk
        red = {"none": 0, "mean": 1, "sum": 2}[reduction]
        cur = torch.cuda.current_stream(self.dev)
        for s in (self.s_in, self.s_cmp, self.s_out):
            s.wait_stream(cur)
        with torch.cuda.stream(self.s_in):
            self.tg.copy_(h_targets, non_blocking=True)
            self.il.copy_(h_il, non_blocking=True)
            self.tl.copy_(h_tl, non_blocking=True)
        n_chunks = (B + c - 1) // c
        for i in range(n_chunks):
            k = i % self.n_slots
            lo, hi = i * c, min((i + 1) * c, B)
            n = hi - lo
            with torch.cuda.stream(self.s_in):
                if i >= self.n_slots:
                    self.s_in.wait_event(self.ev_cmp[k])          # slot's previous kernels are done
                self.x[k][:n].copy_(h_logits[lo:hi], non_blocking=True)
                self.ev_in[k].record(self.s_in)
            with torch.cuda.stream(self.s_cmp):
                self.s_cmp.wait_event(self.ev_in[k])
                if i >= self.n_slots and self.grad_to_host:
                    self.s_cmp.wait_event(self.ev_out[k])         # slot's previous grad has left
                gdst = self.g[k] if self.grad_to_host else self.g_full[lo:hi]
                st = self.s_cmp.cuda_stream
                _lib.check(L.ctcb200_loss_grad(self.x[k].data_ptr(), self.tg[lo:hi].data_ptr(), self.tg.shape[1],
                                               n * self.tg.shape[1], self.il[lo:hi].data_ptr(),
                                               self.tl[lo:hi].data_ptr(), n, T, V, self.Umax, self.blank, self.zi,
                                               red, 1.0 / B, self.nll[lo:hi].data_ptr(), None, gdst.data_ptr(),
                                               self.ws[k].data_ptr(), self.ws_bytes, st, None), "ctcb200_loss_grad")
                self.ev_cmp[k].record(self.s_cmp)
            if self.grad_to_host:
                with torch.cuda.stream(self.s_out):
                    self.s_out.wait_event(self.ev_cmp[k])
                    h_grad[lo:hi].copy_(self.g[k][:n], non_blocking=True)
                    self.ev_out[k].record(self.s_out)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_stream(self.s_cmp)
            h_nll.copy_(self.nll, non_blocking=True)
        cur.wait_stream(self.s_out)
        cur.wait_stream(self.s_in)
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
    p(pin(logits), pin(tg.to(torch.int64)), pin(input_lengths.to(torch.int64)),
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
