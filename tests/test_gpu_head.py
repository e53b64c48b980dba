"""GPU parity of f1, the tcgen05 fused CTC head (ctc_head_loss_b200 / ctcb200_head_loss*), against
``F.linear`` (cuBLAS fp32) + the CPU oracle and against the unfused path of this repo.

Tolerances.  precision='3xtf32' (default): the path's own bars -- per-utterance nll 1e-5 relative, 'mean' gradient
w.r.t. the logits 1e-4 absolute; parameter gradients within 1e-4 of their scale of the unfused autograd result.
precision='tf32' (single pass, 10-bit mantissa operands -- what allow_tf32 gives): a logit carries ~1e-3 absolute
error, so the separately stated bar is nll within 2e-3 relative and parameter gradients within 2e-2 of their scale."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

from oracle.synth import make_lengths, make_targets
from oracle.torch_ref import ref_ctc

pytestmark = pytest.mark.gpu


def _case(B, T, K, V, U, seed, lengths="var", sharp=False):
    g = torch.Generator().manual_seed(seed)
    tg, tl = make_targets(B, U, V, g)
    il = make_lengths(B, T, g, full=(lengths == "full"))
    if lengths == "dead":
        il = torch.tensor([T, 50] + [T] * (B - 2))
    il = torch.maximum(il, torch.minimum(2 * tl + 1, torch.tensor(T)))
    enc = torch.randn(B, T, K, generator=g)
    W = torch.randn(V, K, generator=g) / K ** 0.5 * (3.0 if sharp else 1.0)
    bias = torch.randn(V, generator=g) * 0.1
    return enc, W, bias, tg, il, tl


def _run(case, precision, reduction="mean", zero_infinity=True, bias=True, go=None):
    from asr_chinese_e2e_b200 import ctc_head_loss_b200, ctc_loss_b200
    enc, W, b, tg, il, tl = case
    params = [enc.cuda().requires_grad_(True), W.cuda().requires_grad_(True)] + ([b.cuda().requires_grad_(True)] if bias else [])
    bb = params[2] if bias else None
    tg_d, il_d, tl_d = tg.cuda(), il.cuda(), tl.cuda()
    loss = ctc_head_loss_b200(params[0], params[1], bb, tg_d, il_d, tl_d, reduction=reduction,
                              zero_infinity=zero_infinity, precision=precision)
    loss.backward(torch.ones_like(loss) if go is None else go.cuda())
    fused = [p.grad.clone() for p in params]
    for p in params:
        p.grad = None
    logits = F.linear(params[0], params[1], bb)
    logits.retain_grad()
    lu = ctc_loss_b200(logits, tg_d, il_d, tl_d, reduction=reduction, zero_infinity=zero_infinity)
    lu.backward(torch.ones_like(lu) if go is None else go.cuda())
    unfused = [p.grad.clone() for p in params]
    return loss.detach().cpu(), lu.detach().cpu(), fused, unfused, logits.detach()


@pytest.mark.parametrize("shape", [(2, 64, 32, 40, 5, "full", False), (3, 100, 64, 300, 7, "var", False),
                                   (2, 400, 128, 700, 20, "dead", True), (5, 70, 96, 1000, 9, "var", True),
                                   (8, 200, 512, 4234, 30, "var", False)])
def test_3xtf32_meets_the_paths_parity_bar(shape):
    from asr_chinese_e2e_b200 import ctc_head_loss_b200
    B, T, K, V, U, lengths, sharp = shape
    case = _case(B, T, K, V, U, 7 + B, lengths, sharp)
    enc, W, b, tg, il, tl = case
    loss, lu, fused, unfused, logits = _run(case, "3xtf32")
    # the parity target: cuBLAS fp32 logits -> torch CPU log_softmax + ctc_loss
    ref_nll, _ = ref_ctc(logits, tg, il, tl, reduction="none", zero_infinity=True, want_grad=False)
    rl, _ = ref_ctc(logits, tg, il, tl, reduction="mean", zero_infinity=True, want_grad=False)
    with torch.no_grad():
        nll = ctc_head_loss_b200(enc.cuda(), W.cuda(), b.cuda(), tg.cuda(), il.cuda(), tl.cuda(), reduction="none",
                                 zero_infinity=True).cpu()
    assert ((nll - ref_nll).abs() / ref_nll.abs().clamp(min=1)).max().item() <= 1e-5
    assert abs(loss.item() - rl.item()) <= 1e-5 * abs(rl.item())
    for f, u, name in zip(fused, unfused, ("d_enc", "d_weight", "d_bias")):
        assert (f - u).abs().max().item() <= 1e-4 * u.abs().max().item() + 1e-9, name


def test_gradient_buffer_equals_the_oracle_gradient():
    """The second GEMM pass writes d loss / d logits = g (softmax - occupancy), zeros for padded frames and for the
    pitch padding columns: checked element-wise through the C ABI against the oracle's gradient w.r.t. the logits."""
    from asr_chinese_e2e_b200 import _lib
    L = _lib.lib()
    enc, W, b, tg, il, tl = _case(4, 150, 64, 301, 11, 3, "var", True)
    tl[3] = 0                                        # an empty target
    tl[2] = tg.shape[1]                              # infeasible: 5 frames for 11 labels
    tg[2] = torch.arange(7, 7 + tg.shape[1])
    il[2] = 5
    B, T, K = enc.shape
    V = W.shape[0]
    U = tg.shape[1]
    x, w, bb = enc.cuda(), W.cuda(), b.cuda()
    tg_d, il_d, tl_d = tg.cuda(), il.cuda(), tl.cuda()
    pitch = (V + 3) // 4 * 4
    wsb = _lib.head_workspace_bytes(B, T, V, K, U, 0)
    ws = torch.zeros(wsb, dtype=torch.uint8, device="cuda")
    nll = torch.empty(B, device="cuda"); sums = torch.zeros(4, device="cuda")
    dl = torch.full((B * T, pitch), 7.0, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    rc = L.ctcb200_head_loss_grad(x.data_ptr(), w.data_ptr(), bb.data_ptr(), tg_d.data_ptr(), U, tg_d.numel(), il_d.data_ptr(),
                                  tl_d.data_ptr(), B, T, V, K, U, 0, 1, 0, 1, 1.0 / B, nll.data_ptr(), sums.data_ptr(),
                                  dl.data_ptr(), pitch, ws.data_ptr(), wsb, st)
    assert rc == 0, _lib.strerror(rc)
    torch.cuda.synchronize()
    logits = F.linear(x, w, bb)
    rn, _ = ref_ctc(logits, tg, il, tl, reduction="none", zero_infinity=True, want_grad=False)
    _, rg = ref_ctc(logits, tg, il, tl, reduction="mean", zero_infinity=True)
    assert ((nll.cpu() - rn).abs() / rn.abs().clamp(min=1)).max().item() <= 1e-5
    got = dl.view(B, T, pitch).cpu()
    assert torch.all(got[:, :, V:] == 0)                                     # pitch padding
    assert (got[:, :, :V] - rg).abs().max().item() <= 1e-4
    assert torch.all(got[2] == 0) and nll[2].item() == 0.0                   # zero_infinity
    tmask = torch.arange(T)[None, :] < il[:, None]
    assert torch.all(got[~tmask] == 0)                                       # padded frames
    # argument validation of the new entry points
    assert L.ctcb200_head_loss_grad(x.data_ptr(), w.data_ptr(), None, tg_d.data_ptr(), U, tg_d.numel(), il_d.data_ptr(),
                                    tl_d.data_ptr(), B, T, V, K, U, 0, 1, 0, 1, 1.0 / B, nll.data_ptr(), None, dl.data_ptr(),
                                    pitch + 1, ws.data_ptr(), wsb, st) == -2   # pitch not a multiple of 4
    out = ctypes.c_size_t(0)
    assert L.ctcb200_head_workspace_bytes(B, T, V, 48, U, 0, ctypes.byref(out)) == -2      # K % 32 != 0


def test_tf32_single_pass_has_its_own_stated_tolerance():
    case = _case(8, 200, 512, 4234, 30, 11, "var", True)
    loss, lu, fused, unfused, _ = _run(case, "tf32")
    assert abs(loss.item() - lu.item()) <= 2e-3 * abs(lu.item())
    for f, u, name in zip(fused, unfused, ("d_enc", "d_weight", "d_bias")):
        assert (f - u).abs().max().item() <= 2e-2 * u.abs().max().item(), name


def test_reductions_no_bias_and_upstream_gradients():
    case = _case(4, 90, 64, 120, 8, 21, "var", False)
    for red in ("sum", "none"):
        go = torch.tensor([0.5, -1.0, 2.0, 0.0]) if red == "none" else torch.tensor(0.3)
        loss, lu, fused, unfused, _ = _run(case, "3xtf32", reduction=red, go=go)
        assert torch.allclose(loss, lu, rtol=1e-5)
        for f, u in zip(fused, unfused):
            assert (f - u).abs().max().item() <= 1e-4 * u.abs().max().item() + 1e-9, red
    loss, lu, fused, unfused, _ = _run(case, "3xtf32", bias=False)
    assert torch.allclose(loss, lu, rtol=1e-5) and len(fused) == 2
    for f, u in zip(fused, unfused):
        assert (f - u).abs().max().item() <= 1e-4 * u.abs().max().item() + 1e-9


def test_joint_mixin_with_the_fused_head():
    """JointCTCAttention(fused_head=True): same loss and parameter gradients as the unfused mix-in."""
    from tiny_model import TinyJoint, _batch
    res = {}
    for fused in (False, True):
        torch.manual_seed(3)
        m = TinyJoint(V=13, d=32).cuda()
        m.ctc_fused_head = fused
        batch = _batch().to("cuda")
        met, _ = m.iterate(batch, optimizer=None, is_train=False)
        out = m.forward(batch)
        assert ("ctc_logits" in out) != fused
        loss = m.cal_metrics(out, batch).loss
        loss.backward()
        res[fused] = (loss.detach().clone(), {k: p.grad.clone() for k, p in m.named_parameters()})
    assert torch.allclose(res[True][0], res[False][0], rtol=1e-5)
    for k in res[True][1]:
        a, b = res[True][1][k], res[False][1][k]
        assert (a - b).abs().max().item() <= 1e-4 * b.abs().max().item() + 1e-8, k


@pytest.mark.parametrize("shape", [(1, 128, 32, 32), (2, 64, 40, 32), (3, 100, 300, 64), (2, 333, 700, 128),
                                   (8, 200, 4234, 512), (64, 400, 1030, 96)])
def test_param_grad_gemms_are_fp32_grade(shape):
    """ctcb200_head_param_grads (k_gemm3: 3xTF32 on tcgen05, MN-major operand tiles, accumulators drained every 256
    reduction elements) against float64 matmuls: the error must stay within 4x of cuBLAS's fp32 SIMT GEMM (plus a floor of
    2e-6 of the result's scale) -- ragged tile edges in every dimension, and a 25 600-long reduction for d W."""
    from asr_chinese_e2e_b200 import _lib
    B, T, V, K = shape
    L = _lib.lib()
    g = torch.Generator().manual_seed(B * 1000 + V)
    pitch = (V + 3) // 4 * 4
    dl = torch.zeros(B * T, pitch)
    dl[:, :V] = torch.randn(B * T, V, generator=g) * 1e-2
    enc = torch.randn(B * T, K, generator=g)
    w = torch.randn(V, K, generator=g) / K ** 0.5
    dl_d, enc_d, w_d = dl.cuda(), enc.cuda(), w.cuda()
    g_enc = torch.full((B * T, K), 7.0, device="cuda")
    g_w = torch.full((V, K), 7.0, device="cuda")
    wsb = _lib.head_param_grads_workspace_bytes(V, K)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for want_enc, want_w in ((True, True), (True, False), (False, True)):
        g_enc.fill_(7.0)
        g_w.fill_(7.0)
        rc = L.ctcb200_head_param_grads(dl_d.data_ptr(), pitch, enc_d.data_ptr(), w_d.data_ptr(), B, T, V, K,
                                        g_enc.data_ptr() if want_enc else None, g_w.data_ptr() if want_w else None,
                                        ws.data_ptr(), wsb, st)
        assert rc == 0, _lib.strerror(rc)
        torch.cuda.synchronize()
        ref_enc = dl_d[:, :V].double() @ w_d.double()
        ref_w = dl_d[:, :V].double().t() @ enc_d.double()
        if want_enc:
            err = (g_enc.double() - ref_enc).abs().max().item()
            lib = ((dl_d[:, :V] @ w_d).double() - ref_enc).abs().max().item()
            assert err <= 4 * lib + 2e-6 * ref_enc.abs().max().item(), ("d_enc", err, lib)
        else:
            assert (g_enc == 7.0).all()
        if want_w:
            err = (g_w.double() - ref_w).abs().max().item()
            lib = ((dl_d[:, :V].t() @ enc_d).double() - ref_w).abs().max().item()
            assert err <= 4 * lib + 2e-6 * ref_w.abs().max().item(), ("d_weight", err, lib)
        else:
            assert (g_w == 7.0).all()


def test_param_grad_paths_agree():
    """param_grads='tcgen05' (default) and 'torch' (cuBLAS fp32) give the same parameter gradients."""
    from asr_chinese_e2e_b200 import head
    case = _case(4, 120, 64, 500, 9, 3)
    res = {}
    for mode in ("tcgen05", "torch"):
        head._CFG_HEAD["param_grads"] = mode
        try:
            res[mode] = _run(case, "3xtf32")[2]
        finally:
            head._CFG_HEAD["param_grads"] = "tcgen05"
    for a, b in zip(res["tcgen05"], res["torch"]):
        assert (a - b).abs().max().item() <= 2e-5 * b.abs().max().item() + 1e-9   # cuBLAS fp32 itself is ~3e-6 off


def test_in_ring_split_option_meets_the_same_bar():
    """Option head_inring: k_head takes the raw fp32 tiles as the hi halves and splits in the shared-memory ring
    (truncation split, as k_gemm3 does) instead of reading operands pre-split in HBM -- same parity bar."""
    from asr_chinese_e2e_b200 import _lib
    case = _case(3, 100, 64, 300, 7, 11, "var", False)
    enc, W, b, tg, il, tl = case
    try:
        _lib.set_option("head_inring", 1)
        loss, lu, fused, unfused, logits = _run(case, "3xtf32")
    finally:
        _lib.set_option("head_inring", 0)
    rl, _ = ref_ctc(logits, tg, il, tl, reduction="mean", zero_infinity=True, want_grad=False)
    assert abs(loss.item() - rl.item()) <= 1e-5 * abs(rl.item())
    for f, u, name in zip(fused, unfused, ("d_enc", "d_weight", "d_bias")):
        assert (f - u).abs().max().item() <= 1e-4 * u.abs().max().item() + 1e-9, name
