"""f1: the CTC head fused with the CTC loss -- ``ctc_head_loss_b200(enc, weight, bias, targets, ...)``.

Equivalent to

    ctc_loss_b200(F.linear(enc, weight, bias), targets, input_lengths, target_lengths, ...)

i.e. the ``ctc_head = nn.Linear(d_model, vocab)`` that ``JointCTCAttention`` attaches to the encoder output
(tap point Predictor/Models/transformer_official.py:76) followed by the CTC loss, but the
``[B, T, V]`` logits never exist in HBM: libctcb200's ``ctcb200_head_loss*`` computes each 128 x 256 logits tile on
the tensor cores (tcgen05.mma kind::tf32, fp32 accumulators in TMEM, TMA-staged operands) and consumes it in the GEMM
epilogue (online log-sum-exp + label gather; in training a second GEMM pass recomputes the tile and writes
``d loss / d logits``).  ``d weight``, ``d enc`` and ``d bias`` are then plain library GEMMs / a column sum over that
one gradient buffer (torch.matmul: plumbing, like the ``nn.Linear`` backward they replace).

precision='3xtf32' (default) splits every fp32 operand into two tf32 halves and accumulates three products in fp32:
fp32-GEMM grade, meets the path's 1e-5 relative bar on the per-utterance loss.  precision='tf32' is a single pass
(10-bit mantissa operands, what ``torch.backends.cuda.matmul.allow_tf32 = True`` would give): 3x fewer tensor-core
passes, looser stated tolerance (tests/test_gpu_head.py).
"""
from __future__ import annotations

import torch

from . import _lib
from .ctc import _CFG, _RED, _as_i64_cuda, _validate_host

_PREC = {"3xtf32": 0, "tf32": 1}
# how d enc / d W are formed from the gradient buffer: "tcgen05" = this repo's 3xTF32 tensor-core GEMM (k_gemm3),
# "torch" = torch.matmul (cuBLAS fp32 SIMT GEMMs, what the nn.Linear backward would run)
_CFG_HEAD = {"param_grads": "tcgen05"}


class _CTCHeadLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, enc, weight, bias, targets, input_lengths, target_lengths, blank, reduction, zero_infinity,
                inv_batch, precision, max_target_length, grad_mode=True):
        for name, x in (("enc", enc), ("weight", weight)):
            if not (torch.is_tensor(x) and x.is_cuda):
                raise _lib.CtcB200Error(f"ctc_head_loss_b200 needs CUDA tensors ({name}): the hot path has no CPU fallback")
            if x.dtype != torch.float32:
                raise _lib.CtcB200Error(f"{name} must be float32 (got {x.dtype})")
        if enc.dim() != 3 or weight.dim() != 2 or enc.shape[2] != weight.shape[1]:
            raise ValueError("enc must be [B, T, K] and weight [V, K]")
        B, T, K = enc.shape
        V = weight.shape[0]
        if K % 32:
            raise ValueError("the fused head needs d_model to be a multiple of 32")
        dev = enc.device
        x = enc.contiguous()
        w = weight.contiguous()
        bs = bias.contiguous().float() if bias is not None else None
        _validate_host("input_lengths", input_lengths, 0, T)
        il, tl, tg = _as_i64_cuda(input_lengths, dev), _as_i64_cuda(target_lengths, dev), _as_i64_cuda(targets, dev)
        if tg.dim() == 2:
            umax, stride = tg.shape[1], tg.shape[1]
            if umax > 255:
                umax = int(max_target_length if max_target_length is not None else tl.max().item())
            if stride == 0:
                tg, stride = tg.new_zeros(B, 1), 1
        else:
            stride = 0
            umax = int(max_target_length) if max_target_length is not None else (int(tl.max().item()) if B else 0)
            if tg.numel() == 0:
                tg = tg.new_zeros(1)
        L = _lib.lib()
        prec = _PREC[precision]
        need_grad = bool(any(ctx.needs_input_grad[:3]) and grad_mode)     # (needs_input_grad ignores torch.no_grad())
        red = _RED[reduction]
        flags = int(bool(zero_infinity)) | (4 if _CFG["lattice_log"] else 0)
        inv_b = float(inv_batch) if inv_batch is not None else 1.0 / max(B, 1)
        wsb = _lib.head_workspace_bytes(B, T, V, K, umax, prec)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        nll = torch.empty(B, dtype=torch.float32, device=dev)
        sums = torch.empty(4, dtype=torch.float32, device=dev) if B else torch.zeros(4, device=dev)
        common = (x.data_ptr(), w.data_ptr(), bs.data_ptr() if bs is not None else None, tg.data_ptr(), stride, tg.numel(),
                  il.data_ptr(), tl.data_ptr(), B, T, V, K, umax, int(blank), flags, prec)
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream().cuda_stream
            if need_grad:
                pitch = (V + 3) // 4 * 4
                dl = torch.empty(B * T, pitch, dtype=torch.float32, device=dev)
                _lib.check(L.ctcb200_head_loss_grad(*common, red, inv_b, nll.data_ptr(), sums.data_ptr(), dl.data_ptr(), pitch,
                                                    ws.data_ptr(), wsb, st), "ctcb200_head_loss_grad")
                ctx.save_for_backward(dl, x, w)
                ctx.meta = (B, T, K, V, red, bias is not None)
            else:
                _lib.check(L.ctcb200_head_loss(*common, nll.data_ptr(), sums.data_ptr(), ws.data_ptr(), wsb, st),
                           "ctcb200_head_loss")
        if reduction == "none":
            return nll
        return sums[1] if reduction == "sum" else (sums[3] if need_grad else sums[0] * inv_b)

    @staticmethod
    def backward(ctx, grad_out):
        dl, x, w = ctx.saved_tensors
        B, T, K, V, red, has_bias = ctx.meta
        go = grad_out.to(torch.float32)
        if red == 0:                                         # per-utterance upstream gradient: a scaled copy, same pitch
            dl = (dl.view(B, T, -1) * go.view(B, 1, 1)).view(B * T, -1)
            go = None
        d = dl[:, :V]                                        # [B*T, V] view of the pitched buffer
        g_enc = g_w = g_b = None
        if _CFG_HEAD["param_grads"] == "tcgen05" and (ctx.needs_input_grad[0] or ctx.needs_input_grad[1]):
            # the two parameter-gradient GEMMs on the tensor cores with fp32-grade accuracy (k_gemm3: 3xTF32, operands
            # split inside the shared-memory ring, MN-major operand tiles -- no transposed copy of the gradient)
            L = _lib.lib()
            dev = dl.device
            if ctx.needs_input_grad[0]:
                g_enc = torch.empty(B, T, K, dtype=torch.float32, device=dev)
            if ctx.needs_input_grad[1]:
                g_w = torch.empty(V, K, dtype=torch.float32, device=dev)
            wsb = _lib.head_param_grads_workspace_bytes(V, K)
            ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
            with torch.cuda.device(dev):
                st = torch.cuda.current_stream().cuda_stream
                _lib.check(L.ctcb200_head_param_grads(dl.data_ptr(), dl.shape[1], x.data_ptr(), w.data_ptr(), B, T, V, K,
                                                      g_enc.data_ptr() if g_enc is not None else None,
                                                      g_w.data_ptr() if g_w is not None else None, ws.data_ptr(), wsb, st),
                           "ctcb200_head_param_grads")
            if go is not None:
                g_enc = g_enc * go if g_enc is not None else None
                g_w = g_w * go if g_w is not None else None
        else:
            if ctx.needs_input_grad[0]:
                g_enc = torch.matmul(d, w).view(B, T, K)
                if go is not None:
                    g_enc = g_enc * go
            if ctx.needs_input_grad[1]:
                g_w = torch.matmul(d.t(), x.view(B * T, K))
                if go is not None:
                    g_w = g_w * go
        if has_bias and ctx.needs_input_grad[2]:
            g_b = d.sum(0)
            if go is not None:
                g_b = g_b * go
        return (g_enc, g_w, g_b) + (None,) * 10


def ctc_head_loss_b200(enc, weight, bias, targets, input_lengths, target_lengths, blank: int = 0,
                       reduction: str = "mean", zero_infinity: bool = False, *, inv_batch=None,
                       precision: str = "3xtf32", max_target_length=None):
    """CTC loss of ``F.linear(enc, weight, bias)`` without materialising the logits (see the module doc).
    enc [B, T, K] float32 CUDA, weight [V, K], bias [V] or None; other arguments as ``ctc_loss_b200``."""
    if reduction not in _RED:
        raise ValueError(f"reduction must be one of {list(_RED)}")
    if precision not in _PREC:
        raise ValueError(f"precision must be one of {list(_PREC)}")
    return _CTCHeadLossFn.apply(enc, weight, bias, targets, input_lengths, target_lengths, blank, reduction,
                                zero_infinity, inv_batch, precision, max_target_length, torch.is_grad_enabled())
