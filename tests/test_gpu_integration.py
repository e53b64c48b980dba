"""GPU: the callers either side of the kernels -- joint CTC/attention mix-in with the real op, the
host-buffer pipeline, the stage-split C ABI, the sharded wrapper on one rank."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle.synth import make_case
from oracle.torch_ref import ref_ctc

pytestmark = pytest.mark.gpu


def test_joint_mixin_on_gpu_matches_torch_autograd():
    from tiny_model import TinyJoint, _batch
    torch.manual_seed(3)
    m = TinyJoint().cuda()
    batch = _batch().to("cuda")
    out = m.forward(batch)
    met = m.cal_metrics(out, batch)
    met.loss.backward()
    got = {k: p.grad.clone() for k, p in m.named_parameters()}
    m.zero_grad()
    # reference: identical graph with torch's own CTC on the same device
    out = m.forward(batch)
    ctc = F.ctc_loss(F.log_softmax(out.ctc_logits, -1).transpose(0, 1), batch.tgt_for_input, batch.wave_len,
                     batch.tgt_len, blank=0, reduction="mean", zero_infinity=True)
    att = F.cross_entropy(out.pred.reshape(-1, out.pred.size(-1)), out.gold.reshape(-1), ignore_index=0)
    ref = 0.3 * ctc + 0.7 * att
    ref.backward()
    assert abs(met.loss.item() - ref.item()) <= 1e-5 * abs(ref.item())
    assert abs(met.ctc_loss.item() - ctc.item()) <= 1e-5 * abs(ctc.item())
    for k, p in m.named_parameters():
        assert torch.allclose(got[k], p.grad, rtol=1e-3, atol=1e-5), k


def test_host_pipeline_matches_oracle():
    from asr_chinese_e2e_b200.host_pipeline import ctc_loss_grad_host
    c = make_case(11, 48, 4234 // 8, 9, 31, dist="D2", n_infeasible=1, n_partial=1)
    loss, nll, grad = ctc_loss_grad_host(c["logits"], c["targets"], c["input_lengths"], c["target_lengths"],
                                         reduction="mean", zero_infinity=True, chunk=4)
    rl, rg = ref_ctc(c["logits"], c["targets"], c["input_lengths"], c["target_lengths"], reduction="mean",
                     zero_infinity=True)
    assert abs(loss.item() - rl.item()) <= 1e-5 * abs(rl.item())
    assert (grad - rg).abs().max().item() <= 1e-4


def test_host_pipeline_valid_frames_only_is_bit_identical():
    """valid_frames_only (default) moves only the frames t < input_lengths[b] over PCIe and zeroes the padded rows of
    the host gradient on the host; zero_copy_logits (default) lets the sweep kernel read the pinned host logits in
    place.  Every combination gives the same bits as whole tensors through the copy engines, on a dirty host buffer,
    with full-length runs, ragged and zero-length utterances; the byte counters reflect what crossed the bus."""
    from asr_chinese_e2e_b200.host_pipeline import HostCTCPipeline
    B, T, V, U = 10, 40, 133, 7
    outs = {}
    for seed, lens in ((1, [40, 40, 40, 17, 0, 40, 25, 40, 40, 31]), (2, [12, 40, 40, 40, 40, 9, 40, 33, 40, 40])):
        c = make_case(B, T, V, U, seed, dist="D1")
        il = torch.tensor(lens)
        il = torch.where(il > 0, torch.maximum(il, 2 * c["target_lengths"] + 1), il)
        pin = lambda t: t.contiguous().pin_memory()
        for vfo, zc in ((True, True), (True, False), (False, True), (False, False)):
            pipe = HostCTCPipeline(B, T, V, U, chunk=4, zero_infinity=True, valid_frames_only=vfo, zero_copy_logits=zc)
            h_g = torch.full((B, T, V), float("nan")).pin_memory()
            h_n = torch.empty(B).pin_memory()
            pipe(pin(c["logits"]), pin(c["targets"]), pin(il), pin(c["target_lengths"]), h_g, h_n)
            outs[(seed, vfo, zc)] = (h_g.clone(), h_n.clone())
            valid = int(il.sum())
            small = B * U * 8 + 2 * B * 8                       # targets + the two length vectors
            assert pipe.d2h_bytes == (valid if vfo else B * T) * V * 4 + B * 4
            assert pipe.h2d_bytes == (valid if (vfo or zc) else B * T) * V * 4 + small   # the sweep reads valid frames only
        ref = outs[(seed, False, False)]                        # whole tensors through the copy engines
        for key, got in outs.items():
            if key[0] == seed:
                assert torch.equal(got[0], ref[0]) and torch.equal(got[1], ref[1]), key
        tmask = torch.arange(T)[None, :] < il[:, None]
        assert (ref[0][~tmask] == 0).all()


def test_stage_split_abi_equals_single_call():
    from asr_chinese_e2e_b200 import _lib
    L = _lib.lib()
    c = make_case(9, 40, 133, 7, 5, dist="D1", n_infeasible=1)
    x = c["logits"].cuda()
    tg, il, tl = c["targets"].cuda(), c["input_lengths"].cuda(), c["target_lengths"].cuda()
    B, T, V = x.shape
    U = tg.shape[1]
    wsb = _lib.workspace_bytes(B, T, V, U)
    st = torch.cuda.current_stream().cuda_stream

    def run(stage_list):
        ws = torch.zeros(wsb, dtype=torch.uint8, device="cuda")
        nll = torch.empty(B, device="cuda"); sums = torch.zeros(4, device="cuda"); grad = torch.full_like(x, 7.0)
        for stg in stage_list:
            rc = L.ctcb200_loss_grad_stages(stg, x.data_ptr(), tg.data_ptr(), U, tg.numel(), il.data_ptr(),
                                            tl.data_ptr(), B, T, V, U, 0, 1, 1, 1.0 / B, nll.data_ptr(),
                                            sums.data_ptr(), grad.data_ptr(), ws.data_ptr(), wsb, st)
            assert rc == 0, _lib.strerror(rc)
        torch.cuda.synchronize()
        return nll, sums, grad
    a = run([7])
    b = run([1, 2, 4])
    for u, v in zip(a, b):
        assert torch.equal(u, v)
    rl, rg = ref_ctc(c["logits"], c["targets"], c["input_lengths"], c["target_lengths"], reduction="mean",
                     zero_infinity=True)
    assert (a[2].cpu() - rg).abs().max().item() <= 1e-4
    assert abs(a[1][0].item() / B - rl.item()) <= 1e-5 * abs(rl.item())
    assert L.ctcb200_loss_grad_stages(0, *([None] * 2), 0, 0, None, None, 1, 1, 5, 1, 0, 0, 1, 1.0, None, None,
                                      ctypes.c_void_p(16), None, 0, None) != 0


def test_sharded_wrapper_single_rank_and_debug_status():
    from asr_chinese_e2e_b200 import _lib, sharded_ctc_loss
    c = make_case(6, 30, 41, 6, 17)
    x = c["logits"].cuda().requires_grad_(True)
    loss = sharded_ctc_loss(x, c["targets"].cuda(), c["input_lengths"].cuda(), c["target_lengths"].cuda())
    loss.backward()
    rl, rg = ref_ctc(c["logits"], c["targets"], c["input_lengths"], c["target_lengths"], reduction="mean")
    assert abs(loss.item() - rl.item()) <= 1e-5 * abs(rl.item())
    assert (x.grad.cpu() - rg).abs().max().item() <= 1e-4
    # invalid inputs are clamped and reported through the device status word (debug API)
    L = _lib.lib()
    B, T, V = x.shape
    tg = c["targets"].clone(); tg[0, 0] = V + 5                      # label out of range
    il = c["input_lengths"].clone(); il[1] = T + 3                   # input length out of range
    tgc, ilc, tlc = tg.cuda(), il.cuda(), c["target_lengths"].cuda()
    wsb = _lib.workspace_bytes(B, T, V, tg.shape[1])
    ws = torch.zeros(wsb, dtype=torch.uint8, device="cuda"); nll = torch.empty(B, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    assert L.ctcb200_forward(x.data_ptr(), tgc.data_ptr(), tg.shape[1], tg.numel(), ilc.data_ptr(), tlc.data_ptr(),
                             B, T, V, tg.shape[1], 0, 1, nll.data_ptr(), None, ws.data_ptr(), wsb, st, None) == 0
    status = ctypes.c_int(0)
    assert L.ctcb200_read_status(ws.data_ptr(), ctypes.byref(status), st) == 0
    assert status.value & 1 and status.value & 4
    assert torch.isfinite(nll[2:]).all()
    assert torch.isnan(nll[:2]).all()            # the two invalid utterances are poisoned, not silently clamped


def test_no_out_of_bounds_writes_guard_zones():
    """compute-sanitizer is closed on the GPU pool, so the write side is checked with canaries: every
    output buffer (grad, nll, loss_sums, workspace) is carved out of a larger poisoned allocation and the
    guard zones on both sides must be bit-identical after the two-sweep and the three-sweep calls.
    V=37 and an odd B*T make every row misaligned and the tensor end not 16-byte aligned."""
    from asr_chinese_e2e_b200 import _lib
    L = _lib.lib()
    c = make_case(5, 23, 37, 6, 11, dist="D2", n_infeasible=1)
    x = c["logits"].cuda()
    tg, il, tl = c["targets"].cuda(), c["input_lengths"].cuda(), c["target_lengths"].cuda()
    B, T, V = x.shape
    U = tg.shape[1]
    wsb = _lib.workspace_bytes(B, T, V, U)
    G = 4096                                                   # guard bytes (multiple of 256)
    st = torch.cuda.current_stream().cuda_stream

    def guarded(nbytes):
        raw = torch.full((G + nbytes + G,), 0xA5, dtype=torch.uint8, device="cuda")
        return raw, raw[G:G + nbytes]

    for mode in ("two_sweep", "three_sweep", "loss_only"):
        graw, gbuf = guarded(B * T * V * 4)
        nraw, nbuf = guarded(B * 4)
        sraw, sbuf = guarded(4 * 4)
        wraw, wbuf = guarded(wsb)
        assert gbuf.data_ptr() % 16 == 0 and wbuf.data_ptr() % 256 == 0
        if mode == "two_sweep":
            rc = L.ctcb200_loss_grad(x.data_ptr(), tg.data_ptr(), U, tg.numel(), il.data_ptr(), tl.data_ptr(), B, T, V, U,
                                     0, 1, 1, 1.0 / B, nbuf.data_ptr(), sbuf.data_ptr(), gbuf.data_ptr(), wbuf.data_ptr(),
                                     wsb, st, None)
        else:
            f = L.ctcb200_forward if mode == "three_sweep" else L.ctcb200_loss_only
            rc = f(x.data_ptr(), tg.data_ptr(), U, tg.numel(), il.data_ptr(), tl.data_ptr(), B, T, V, U, 0, 1,
                   nbuf.data_ptr(), sbuf.data_ptr(), wbuf.data_ptr(), wsb, st, None)
            if rc == 0 and mode == "three_sweep":
                one = torch.ones((), device="cuda")
                rc = L.ctcb200_backward(x.data_ptr(), tg.data_ptr(), U, tg.numel(), one.data_ptr(), 0, 1, 1.0 / B, B, T, V,
                                        U, 0, 1, gbuf.data_ptr(), wbuf.data_ptr(), wsb, st)
        assert rc == 0, _lib.strerror(rc)
        torch.cuda.synchronize()
        for name, raw, n in (("grad", graw, B * T * V * 4), ("nll", nraw, B * 4), ("sums", sraw, 16), ("ws", wraw, wsb)):
            assert bool((raw[:G] == 0xA5).all()) and bool((raw[G + n:] == 0xA5).all()), f"{mode}: {name} guard zone written"
        if mode != "loss_only":
            grad = gbuf.view(torch.float32).view(B, T, V)
            _, rg = ref_ctc(c["logits"], c["targets"], c["input_lengths"], c["target_lengths"], reduction="mean",
                            zero_infinity=True)
            assert (grad.cpu() - rg).abs().max().item() <= 1e-4, mode


@pytest.mark.parametrize("eps", [0.0, 0.1])
def test_attention_ce_matches_reference_formulation(eps):
    """the attention-branch loss on the sweep kernels vs the reference's cal_loss
    (Predictor/Utils/loss.py:26-51) restated with torch ops on the CPU."""
    from asr_chinese_e2e_b200 import attention_ce_b200
    g = torch.Generator().manual_seed(7)
    for (N, Ln, C) in ((3, 7, 37), (5, 11, 4234), (2, 5, 130)):
        pred = torch.randn(N, Ln, C, generator=g) * 2.0
        gold = torch.randint(1, C, (N, Ln), generator=g)
        gold[0, -2:] = 0; gold[-1, 1] = 0                          # PAD rows are ignored
        w = 0.7
        x = pred.clone().requires_grad_(True)
        p2, g2 = x.view(-1, C), gold.view(-1)
        if eps > 0:
            one_hot = torch.zeros_like(p2).scatter(1, g2.view(-1, 1), 1)
            one_hot = one_hot * (1 - eps) + (1 - one_hot) * eps / C
            ref = -(one_hot * F.log_softmax(p2, 1)).sum(1).masked_select(g2.ne(0)).sum() / g2.ne(0).sum()
        else:
            ref = F.cross_entropy(p2, g2, ignore_index=0, reduction="mean")
        (w * ref * 1.5).backward()
        xc = pred.cuda().requires_grad_(True)
        got = attention_ce_b200(xc, gold.cuda(), smoothing=eps, weight=w)
        (got * 1.5).backward()                                     # upstream gradient != 1: rescale path
        assert abs(got.item() - w * ref.item()) <= 1e-5 * abs(ref.item()), (N, Ln, C)
        assert (xc.grad.cpu() - x.grad).abs().max().item() <= 1e-6, (N, Ln, C)
        assert torch.all(xc.grad.view(-1, C)[g2.eq(0).cuda()] == 0)
        with torch.no_grad():
            assert abs(attention_ce_b200(pred.cuda(), gold.cuda(), eps).item() - ref.item()) <= 1e-5 * abs(ref.item())


def test_greedy_decode_and_edit_distance_on_device():
    """best-path decode + Levenshtein on the device vs the host implementation."""
    from asr_chinese_e2e_b200 import ctc_greedy_cer_b200, ctc_loss_b200
    from asr_chinese_e2e_b200.joint import edit_distance, greedy_ctc_ids
    for (B, T, V, U, dist) in ((7, 61, 53, 13, "D2"), (5, 200, 4234, 30, "D2"), (4, 300, 97, 100, "D1"), (3, 90, 31, 150, "D1")):
        c = make_case(B, T, V, U, 77, dist=dist)
        c["target_lengths"][0] = 0
        x = c["logits"].cuda()
        cer, info = ctc_greedy_cer_b200(x, c["targets"].cuda(), c["input_lengths"].cuda(), c["target_lengths"].cuda())
        want_h = greedy_ctc_ids(c["logits"], c["input_lengths"])
        tot = 0.0
        for b in range(B):
            n = int(info["hyp_len"][b])
            assert info["hyp"][b, :n].tolist() == want_h[b], (B, T, V, U, b)
            assert torch.all(info["hyp"][b, n:] == 0)
            ref = c["targets"][b, : int(c["target_lengths"][b])].tolist()
            d = edit_distance(want_h[b], ref)
            assert int(info["edit_distance"][b]) == d, (B, T, V, U, b)
            tot += d / max(len(ref), 1)
        assert abs(cer.item() - 100.0 * tot / B) < 1e-3
    # the training forward fills the same dict as a by-product (no extra sweep)
    info2 = {}
    xg = c["logits"].cuda().requires_grad_(True)
    ctc_loss_b200(xg, c["targets"].cuda(), c["input_lengths"].cuda(), c["target_lengths"].cuda(), zero_infinity=True,
                  decode=info2).backward()
    assert torch.equal(info2["edit_distance"], info["edit_distance"]) and torch.equal(info2["hyp"], info["hyp"])


def test_masks_on_the_device_match_the_reference_loop():
    """Section 8(f) row 4: length masks built on the GPU from GPU-resident lengths (one broadcast compare, no
    host loop, nothing read back) equal the reference's per-utterance slice assignments."""
    from asr_chinese_e2e_b200 import get_attn_pad_mask, get_non_pad_mask
    g = torch.Generator().manual_seed(11)
    x = torch.randn(64, 400, 8, generator=g).cuda()
    lens = torch.randint(200, 401, (64,), generator=g)
    want = x.new_ones(64, 400)
    for i in range(64):                                              # reference utils.py:106-108
        want[i, int(lens[i]):] = 0
    got = get_non_pad_mask(x, input_lengths=lens.cuda())
    assert got.is_cuda and torch.equal(got.squeeze(-1), want)
    att = get_attn_pad_mask(x, lens.cuda(), 5)
    assert att.shape == (64, 5, 400) and torch.equal(att[:, 0], want.lt(1))


def test_lattice_event_splits_the_call_without_changing_results():
    """`lattice_event` (used by sharded_ctc_loss to start its all-reduce under the gradient patch) is recorded
    once the loss value is final; loss and gradient are bit-identical to the single-call path."""
    from asr_chinese_e2e_b200 import ctc_loss_b200
    c = make_case(20, 60, 97, 9, 2024, n_infeasible=1)
    args = [c[k].cuda() for k in ("targets", "input_lengths", "target_lengths")]
    outs = []
    for ev in (None, torch.cuda.Event()):
        x = c["logits"].cuda().requires_grad_(True)
        loss = ctc_loss_b200(x, *args, reduction="mean", zero_infinity=True, lattice_event=ev)
        if ev is not None:
            ev.synchronize()                                   # recorded (would raise / hang otherwise)
            seen_early = loss.item()                           # the value is final at the event
        loss.backward()
        outs.append((loss.detach().clone(), x.grad.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert seen_early == outs[1][0].item()
    with torch.no_grad():                                      # paths that do not split still record the event
        ev = torch.cuda.Event()
        ctc_loss_b200(c["logits"].cuda(), *args, reduction="mean", zero_infinity=True, lattice_event=ev)
        ev.synchronize()


def test_forward_backward_is_cuda_graph_capturable():
    """No host synchronisation, host-side branching on device data or allocation outside torch's allocator on
    the default path: loss + gradient can be captured once in a CUDA graph and replayed on new data."""
    from asr_chinese_e2e_b200 import ctc_loss_b200
    c = make_case(24, 80, 211, 12, 77)
    tg, il, tl = (c[k].cuda() for k in ("targets", "input_lengths", "target_lengths"))
    x = c["logits"].cuda().requires_grad_(True)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):                               # warm-up outside capture (library load, attributes)
        for _ in range(2):
            x.grad = None
            ctc_loss_b200(x, tg, il, tl, reduction="mean", zero_infinity=True).backward()
    torch.cuda.current_stream().wait_stream(side)
    x.grad = None
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        loss = ctc_loss_b200(x, tg, il, tl, reduction="mean", zero_infinity=True)
        loss.backward()
    for seed in (78, 79):
        c2 = make_case(24, 80, 211, 12, seed)
        with torch.no_grad():
            x.copy_(c2["logits"]); tg.copy_(c2["targets"]); il.copy_(c2["input_lengths"]); tl.copy_(c2["target_lengths"])
        g.replay()
        torch.cuda.synchronize()
        rl, rg = ref_ctc(c2["logits"], c2["targets"], c2["input_lengths"], c2["target_lengths"], reduction="mean",
                         zero_infinity=True)
        assert abs(loss.item() - rl.item()) <= 1e-5 * abs(rl.item())
        assert (x.grad.cpu() - rg).abs().max().item() <= 1e-4


def test_no_grad_takes_the_loss_only_path_even_if_the_logits_require_grad():
    """Evaluation (Trainer11.evaluate runs model.iterate under torch.no_grad()) must not pay for a gradient: autograd's
    needs_input_grad ignores the grad mode, so the op checks it itself -- visible as no [B,T,V] buffer being allocated."""
    from asr_chinese_e2e_b200 import ctc_loss_b200
    c = make_case(32, 200, 4234, 20, 9)
    x = c["logits"].cuda().requires_grad_(True)
    args = [c[k].cuda() for k in ("targets", "input_lengths", "target_lengths")]
    with torch.no_grad():
        ctc_loss_b200(x, *args)                                    # warm-up (library load, workspace sizes)
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats()
        base = torch.cuda.memory_allocated()
        loss = ctc_loss_b200(x, *args)
        torch.cuda.synchronize()
        assert torch.cuda.max_memory_allocated() - base < x.numel() * 4      # no gradient slab
    assert not loss.requires_grad
    loss2 = ctc_loss_b200(x, *args)
    assert loss2.requires_grad and torch.allclose(loss, loss2.detach(), rtol=1e-6)
