"""GPU parity: the sm_100a CUDA path (through the C-ABI library) against the CPU oracle.

Tolerances are BASELINE.json's: per-utterance fp32 loss within 1e-5 relative, gradient w.r.t. the
logits within 1e-4 absolute ('mean' reduction), infeasible utterances compared by class.
The oracle is torch's CPU ctc_loss (oracle.torch_ref) and, as the adjudicator, the float64
restatement (oracle.c_oracle).  /root/reference is never read here.
"""
import numpy as np
import pytest
import torch

from oracle.c_oracle import ctc_c_f64
from oracle.synth import make_case, make_config
from oracle.torch_ref import ref_ctc

pytestmark = pytest.mark.gpu

REL_LOSS = 1e-5
ABS_GRAD = 1e-4


def _op():
    from asr_chinese_e2e_b200 import ctc_loss_b200
    return ctc_loss_b200


def run_gpu(c, reduction="mean", zero_infinity=False, grad=True, grad_output=None):
    x = c["logits"].cuda().requires_grad_(grad)
    loss = _op()(x, c["targets"].cuda(), c["input_lengths"].cuda(), c["target_lengths"].cuda(),
                 blank=0, reduction=reduction, zero_infinity=zero_infinity)
    g = None
    if grad:
        go = torch.ones_like(loss) if grad_output is None else grad_output.to(loss)
        loss.backward(go)
        g = x.grad.detach().cpu()
    torch.cuda.synchronize()
    return loss.detach().cpu(), g


def assert_loss_close(got, want, what=""):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert np.array_equal(np.isinf(got), np.isinf(want)), f"{what}: inf pattern differs"
    assert not np.isnan(got).any(), f"{what}: NaN loss"
    fin = np.isfinite(want)
    err = np.abs(got[fin] - want[fin]) / np.maximum(np.abs(want[fin]), 1.0)
    assert err.size == 0 or err.max() <= REL_LOSS, f"{what}: loss rel err {err.max():.3e}"


def assert_grad_close(got, want, what="", tol=ABS_GRAD):
    got, want = got.numpy(), want.numpy() if torch.is_tensor(want) else want
    assert np.array_equal(np.isnan(got), np.isnan(want)), f"{what}: NaN pattern differs"
    ok = ~np.isnan(want)
    err = np.abs(got[ok] - want[ok]).max() if ok.any() else 0.0
    assert err <= tol, f"{what}: grad abs err {err:.3e}"


def check_case(c, zero_infinity=False, what=""):
    ref_nll, _ = ref_ctc(c["logits"], c["targets"], c["input_lengths"], c["target_lengths"],
                         reduction="none", zero_infinity=zero_infinity, want_grad=False)
    nll, _ = run_gpu(c, "none", zero_infinity, grad=False)
    assert_loss_close(nll, ref_nll, what + " nll(loss_only)")
    for red in ("mean", "sum"):
        rl, rg = ref_ctc(c["logits"], c["targets"], c["input_lengths"], c["target_lengths"],
                         reduction=red, zero_infinity=zero_infinity)
        gl, gg = run_gpu(c, red, zero_infinity)
        assert_loss_close(gl, rl, f"{what} {red}")
        if red == "mean":
            assert_grad_close(gg, rg, f"{what} {red}", tol=ABS_GRAD)          # the north-star bar
        else:
            # un-normalised gradients are B*U times larger and fp32 log-space rounding (|alpha| ~ T ln V)
            # shows: adjudicate with the float64 oracle -- we must be no worse than 3x torch's own error
            _, _, g64 = ctc_c_f64(*[c[k].numpy() for k in ("logits", "targets", "input_lengths", "target_lengths")],
                                  reduction="sum", zero_infinity=zero_infinity)
            ok = ~np.isnan(g64)
            assert np.array_equal(np.isnan(gg.numpy()), ~ok), f"{what}: NaN pattern differs"
            torch_err = np.abs(rg.numpy()[ok] - g64[ok]).max()
            our_err = np.abs(gg.numpy()[ok] - g64[ok]).max()
            assert our_err <= max(1e-3, 3 * torch_err), f"{what} sum: ours {our_err:.3e} vs torch {torch_err:.3e}"
    # forward-with-grad path must give the same nll as the loss-only path
    x = c["logits"].cuda().requires_grad_(True)
    nll2 = _op()(x, c["targets"].cuda(), c["input_lengths"].cuda(), c["target_lengths"].cuda(),
                 reduction="none", zero_infinity=zero_infinity).detach().cpu()
    assert_loss_close(nll2, ref_nll, what + " nll(forward)")


@pytest.mark.parametrize("zi", [False, True])
def test_golden_vectors(golden, zi):
    z, names = golden
    for name in names:
        c = {k: torch.from_numpy(z[f"{name}/{k}"]) for k in ("logits", "targets", "input_lengths", "target_lengths")}
        nll, _ = run_gpu(c, "none", zi, grad=False)
        assert_loss_close(nll, z[f"{name}/zi{int(zi)}/nll"], name)
        for red in ("mean", "sum"):
            gl, gg = run_gpu(c, red, zi)
            assert_loss_close(gl, z[f"{name}/zi{int(zi)}/{red}/loss"], f"{name} {red}")
            assert_grad_close(gg, torch.from_numpy(z[f"{name}/zi{int(zi)}/{red}/grad"]), f"{name} {red}",
                              tol=ABS_GRAD if red == "mean" else 1e-3)


@pytest.mark.parametrize("dist", ["D1", "D2"])
@pytest.mark.parametrize("zi", [False, True])
def test_small_random(dist, zi):
    c = make_case(7, 61, 53, 13, 21, dist=dist, n_infeasible=1, n_partial=1)
    c["target_lengths"][6] = 0                      # empty target
    check_case(c, zi, f"small {dist}")


def test_vocab_alignment_classes():
    # V*4 mod 16 in {0,4,8,12}: every misalignment of the row hull, incl. the tensor-end tail path
    for V in (32, 33, 34, 35, 4234 // 8 + 1):
        c = make_case(3, 17, V, 5, 100 + V)
        check_case(c, False, f"V={V}")


def test_c1_config():
    c = make_config("C1")                           # B=16, T=200, V=4234, U<=30
    check_case(c, False, "C1")
    c = make_config("C1", dist="D2")
    check_case(c, False, "C1/D2")


def test_wide_lattices():
    # NS=8 (U<=127) and NS=16 (U<=255) variants of the lattice kernel
    for umax, T, seed in ((100, 260, 31), (200, 450, 32)):
        c = make_case(3, T, 301, umax, seed, dist="D2", n_partial=1)
        check_case(c, True, f"Umax={umax}")


def test_c4_like_long_with_infeasible():
    c = make_case(8, 1500, 257, 120, 1004, dist="D1", n_infeasible=2, n_partial=2)
    check_case(c, True, "C4-like")
    nll, g = run_gpu(c, "mean", True)
    il = c["input_lengths"]
    bad = [1, 2]
    for b in bad:
        assert torch.all(g[b] == 0)


def test_c4_full_size_properties():
    """BASELINE config 4 at full size (B=64, T=1500, V=4234, U<=120, zero_infinity=True, 8 infeasible + 8
    partial-lattice utterances): size-independent properties + a float64 spot check."""
    c = make_config("C4", dist="D1")
    B, T, V = c["logits"].shape
    op = _op()
    x = c["logits"].cuda().requires_grad_(True)
    tg, il, tl = c["targets"].cuda(), c["input_lengths"].cuda(), c["target_lengths"].cuda()
    loss = op(x, tg, il, tl, reduction="mean", zero_infinity=True)
    loss.backward()
    g = x.grad
    nll = op(x.detach(), tg, il, tl, reduction="none", zero_infinity=True)
    infeasible = list(range(1, 9))                      # make_config("C4"): utterances 1..8 infeasible, 9..16 partial
    assert torch.all(nll[infeasible] == 0) and torch.all(g[infeasible] == 0)
    assert torch.all(nll[9:17] > 0) and torch.isfinite(nll).all() and torch.isfinite(g).all()
    assert abs(loss.item() - (nll.double() / tl.clamp(min=1).double()).mean().item()) <= 1e-6 * abs(loss.item())
    tmask = torch.arange(T, device="cuda")[None, :] < il[:, None]
    assert torch.all(g[~tmask] == 0)
    assert g.sum(-1)[tmask].abs().max().item() < 1e-5
    idx = [0, 9, 40]
    sub = {k: v[idx] for k, v in c.items()}
    _, n64, g64 = ctc_c_f64(*[sub[k].numpy() for k in ("logits", "targets", "input_lengths", "target_lengths")],
                            reduction="sum", zero_infinity=True)
    assert np.abs(nll[idx].cpu().numpy() - n64).max() / np.abs(n64).max() < REL_LOSS
    scale = 1.0 / (B * tl[idx].clamp(min=1).double().cpu().numpy())
    assert np.abs(g[idx].cpu().numpy() - g64 * scale[:, None, None]).max() < ABS_GRAD


def test_closer_to_float64_than_the_fp32_reference():
    """The lattice keeps alpha/beta relative to a running offset (double), so its fp32 state stays small; the
    un-normalised gradient and the loss must be at least as close to the float64 oracle as torch's fp32 path
    (measured: 30-60x closer, profiles/r01_accuracy_vs_float64.txt)."""
    c = make_case(6, 400, 4234, 50, 5, dist="D1")
    a = [c[k].numpy() for k in ("logits", "targets", "input_lengths", "target_lengths")]
    _, n64, g64 = ctc_c_f64(*a, reduction="sum")
    rn, _ = ref_ctc(c["logits"], c["targets"], c["input_lengths"], c["target_lengths"], reduction="none", want_grad=False)
    _, rg = ref_ctc(c["logits"], c["targets"], c["input_lengths"], c["target_lengths"], reduction="sum")
    for fused in (True, False):
        x = c["logits"].cuda().requires_grad_(True)
        nll = _op()(x, c["targets"].cuda(), c["input_lengths"].cuda(), c["target_lengths"].cuda(), reduction="none",
                    fused=fused)
        nll.sum().backward()
        ours_n = np.abs(nll.detach().cpu().numpy() - n64).max() / np.abs(n64).max()
        ours_g = np.abs(x.grad.cpu().numpy() - g64).max()
        ref_n = np.abs(rn.numpy() - n64).max() / np.abs(n64).max()
        ref_g = np.abs(rg.numpy() - g64).max()
        assert ours_n <= max(ref_n, 1.5e-7), (ours_n, ref_n)
        assert ours_g <= 0.25 * ref_g, (ours_g, ref_g)


def test_reduction_none_with_upstream_gradient():
    c = make_case(5, 40, 29, 8, 77)
    go = torch.tensor([0.5, -1.0, 2.0, 0.0, 1.5])
    rl, rg = ref_ctc(c["logits"], c["targets"], c["input_lengths"], c["target_lengths"],
                     reduction="none", grad_output=go)
    gl, gg = run_gpu(c, "none", grad_output=go)
    assert_loss_close(gl, rl, "none")
    assert_grad_close(gg, rg, "none", tol=1e-3)


def test_target_encodings_and_length_dtypes():
    c = make_case(5, 30, 21, 6, 3)
    tl = c["target_lengths"]
    cat = torch.cat([c["targets"][b, : tl[b]] for b in range(5)])
    op = _op()
    x = c["logits"].cuda()
    a = op(x, c["targets"].cuda(), c["input_lengths"].cuda(), tl.cuda(), reduction="none")
    b1 = op(x, cat.cuda(), c["input_lengths"].cuda(), tl.cuda(), reduction="none")
    b2 = op(x, cat, c["input_lengths"].int(), tl.int(), reduction="none")          # CPU int32 lengths
    b3 = op(x, cat.cuda(), tuple(c["input_lengths"].tolist()), tuple(tl.tolist()), reduction="none")
    assert torch.equal(a, b1) and torch.equal(a, b2) and torch.equal(a, b3)


def test_properties_and_determinism_c2_shape():
    # full BASELINE size: size-independent properties instead of an element-wise oracle
    c = make_config("C2")
    B, T, V = c["logits"].shape
    x = c["logits"].cuda().requires_grad_(True)
    tg, il, tl = c["targets"].cuda(), c["input_lengths"].cuda(), c["target_lengths"].cuda()
    op = _op()
    loss = op(x, tg, il, tl, reduction="mean")
    loss.backward()
    g1 = x.grad.clone()
    nll = op(x.detach(), tg, il, tl, reduction="none")
    # (1) 'mean' == mean_b(nll_b / max(U_b,1))
    want = (nll.double() / tl.clamp(min=1).double()).mean()
    assert abs(loss.item() - want.item()) <= 1e-6 * abs(want.item())
    # (2) padded frames exactly zero; valid frames sum to ~0 over the vocabulary
    tmask = torch.arange(T, device="cuda")[None, :] < il[:, None]
    assert torch.all(g1[~tmask] == 0)
    rs = g1.sum(-1)[tmask].abs().max().item()
    assert rs < 1e-5, rs      # |g_b| ~ 1e-4 times the fp32 log-space normalisation error (~3e-3)
    # (3) bit-identical on repeat (no float atomics)
    x.grad = None
    op(x, tg, il, tl, reduction="mean").backward()
    assert torch.equal(g1, x.grad)
    # (4) sharding: running two halves with the global normaliser reproduces the slab bit-for-bit
    h = B // 2
    for sl in (slice(0, h), slice(h, B)):
        xs = c["logits"][sl].cuda().requires_grad_(True)
        ls = op(xs, tg[sl], il[sl], tl[sl], reduction="mean", inv_batch=1.0 / B)
        ls.backward()
        assert torch.equal(xs.grad, g1[sl])
    # (5) spot-check a few utterances against the float64 oracle
    idx = [0, 17, 255]
    sub = {k: v[idx] for k, v in c.items()}
    _, n64, g64 = ctc_c_f64(*[sub[k].numpy() for k in ("logits", "targets", "input_lengths", "target_lengths")],
                            reduction="sum")
    assert np.abs(nll[idx].cpu().numpy() - n64).max() / np.abs(n64).max() < REL_LOSS
    scale = 1.0 / (B * tl[idx].clamp(min=1).double().cpu().numpy())
    assert np.abs(g1[idx].cpu().numpy() - g64 * scale[:, None, None]).max() < ABS_GRAD


def test_fused_pipeline_equals_unfused_and_rescales():
    """The chunked two-stream, speculative-gradient path must equal the plain forward/backward path bit for
    bit, for any chunk count, and apply a non-unit upstream gradient exactly once."""
    c = make_case(13, 70, 4234 // 4, 17, 555, dist="D2", n_infeasible=1, n_partial=1)
    op = _op()
    tg, il, tl = c["targets"].cuda(), c["input_lengths"].cuda(), c["target_lengths"].cuda()

    def run(w=1.0, **kw):
        x = c["logits"].cuda().requires_grad_(True)
        loss = op(x, tg, il, tl, reduction="mean", zero_infinity=True, **kw)
        (loss * w).backward()
        return loss.detach(), x.grad

    l0, g0 = run(fused=False, chunks=1)                       # textbook split: k1, k2 | k3 in backward
    l1, g1 = run(fused=False, chunks=3)
    assert torch.allclose(l0, l1, rtol=1e-6) and torch.equal(g0, g1)
    lf, gf = run(fused=True, chunks=1)                        # two-sweep path: dense part in the sweep + sparse patch
    for kw in (dict(fused=True, chunks=4), dict(fused=True, chunks=13)):
        l2, g2 = run(**kw)
        assert torch.allclose(lf, l2, rtol=1e-6)              # per-chunk partial sums: order of the final add differs
        assert torch.equal(gf, g2), kw
    # the two formulations round differently (g*(p-occ) vs fl(g*p) - g*occ) but agree to fp32 rounding
    assert torch.allclose(l0, lf, rtol=1e-6)
    assert (g0 - gf).abs().max().item() <= 1e-5 * g0.abs().max().item()
    # ctc_weight-style upstream gradient (JointCTCAttention: loss = 0.3*ctc + 0.7*att)
    _, gw = run(w=0.3, fused=True, chunks=4)
    _, gu = run(w=0.3, fused=False, chunks=1)
    assert (gw - gu).abs().max().item() <= 1e-5 * gu.abs().max().item()
    # reduction='none' with a per-utterance upstream gradient, and a second backward (retain_graph)
    x = c["logits"].cuda().requires_grad_(True)
    nll = op(x, tg, il, tl, reduction="none", zero_infinity=True, fused=True, chunks=2)
    go = torch.linspace(-1, 2, 13, device="cuda")
    nll.backward(go, retain_graph=True)
    ga = x.grad.clone()
    x.grad = None
    nll.backward(go)
    # a second backward of a retained graph recomputes the gradient out of place with the three-sweep kernel (the
    # buffer handed out by the first backward is never written again): same values up to fp32 rounding
    assert (ga - x.grad).abs().max().item() <= 1e-5 * ga.abs().max().item()
    _, rg = ref_ctc(c["logits"], c["targets"], c["input_lengths"], c["target_lengths"], reduction="none",
                    zero_infinity=True, grad_output=go.cpu())
    assert_grad_close(ga.cpu(), rg, "none fused", tol=1e-3)


def test_error_codes_and_no_cpu_fallback():
    from asr_chinese_e2e_b200 import _lib
    op = _op()
    with pytest.raises(_lib.CtcB200Error):
        op(torch.zeros(1, 4, 5), torch.ones(1, 1, dtype=torch.long), [4], [1])      # CPU logits
    with pytest.raises(_lib.CtcB200Error):
        op(torch.zeros(1, 4, 5, device="cuda", dtype=torch.float64), torch.ones(1, 1, dtype=torch.long), [4], [1])
    with pytest.raises(_lib.CtcB200Error):
        op(torch.zeros(1, 4, 5, device="cuda"), torch.ones(1, 1, dtype=torch.long), [4], [1], blank=7)


def _raw_loss_grad(c, zero_infinity=True):
    """One ctcb200_loss_grad call through the C ABI ('sum' reduction); returns nll, grad and the lattice
    kernel's path counters [utterances in log space, of which: underflowed in the linear domain]."""
    import ctypes
    from asr_chinese_e2e_b200 import _lib
    L = _lib.lib()
    x = c["logits"].cuda()
    tg, il, tl = c["targets"].cuda(), c["input_lengths"].cuda(), c["target_lengths"].cuda()
    B, T, V = x.shape
    U = tg.shape[1]
    wsb = _lib.workspace_bytes(B, T, V, U)
    ws = torch.zeros(wsb, dtype=torch.uint8, device="cuda")
    nll = torch.empty(B, device="cuda"); sums = torch.zeros(4, device="cuda"); grad = torch.empty_like(x)
    st = torch.cuda.current_stream().cuda_stream
    rc = L.ctcb200_loss_grad(x.data_ptr(), tg.data_ptr(), U, tg.numel(), il.data_ptr(), tl.data_ptr(), B, T, V, U, 0,
                             int(zero_infinity), 2, 1.0, nll.data_ptr(), sums.data_ptr(), grad.data_ptr(),
                             ws.data_ptr(), wsb, st, None)
    assert rc == 0, _lib.strerror(rc)
    stats = (ctypes.c_int * 2)()
    assert L.ctcb200_read_lattice_stats(ws.data_ptr(), stats, st) == 0
    return nll.cpu().numpy(), grad.cpu().numpy(), list(stats)


def _f64(c, zero_infinity=True):
    _, n64, g64 = ctc_c_f64(*[c[k].numpy() for k in ("logits", "targets", "input_lengths", "target_lengths")],
                            reduction="sum", zero_infinity=zero_infinity)
    return n64, g64


@pytest.mark.gpu
@pytest.mark.parametrize("umax", [12, 100, 200])
def test_linear_domain_lattice_is_the_default_and_matches_float64(umax):
    """Ordinary inputs run the linear-domain (float64, per-lane exponent) recursion for every utterance, for
    all three lattice widths (4 / 8 / 16 states per lane), and agree with the float64 oracle far inside the
    parity tolerances."""
    c = make_case(5, 2 * umax + 37, 53, umax, 900 + umax, dist="D2")
    nll, g, stats = _raw_loss_grad(c)
    assert stats == [0, 0]
    n64, g64 = _f64(c)
    # (the fp32 rounding of the gathered log-probabilities themselves is ~5e-7 per frame)
    assert (np.abs(nll - n64) <= 1e-6 * np.abs(n64) + 2e-5).all()
    assert np.abs(g - g64).max() < 2e-5


@pytest.mark.gpu
def test_blank_dominated_posteriors_stay_on_the_linear_path():
    """Early-training regime: the model predicts blank with probability ~1 and every label with ~2^-35, so
    alpha spreads over ~2^-35 per label ACROSS the states of a frame (2^-1400 over 40 labels) -- beyond any
    single warp-wide scale in float64; the per-lane exponents hold it."""
    c = make_case(4, 120, 61, 40, 4242, dist="D1", full_targets=True)
    c["logits"][:, :, 0] += 24.0
    nll, g, stats = _raw_loss_grad(c)
    assert stats == [0, 0]
    n64, g64 = _f64(c)
    assert np.isfinite(nll).all()
    assert (np.abs(nll - n64) <= 1e-6 * np.abs(n64) + 2e-5).all()
    assert np.abs(g - g64).max() < 2e-5


@pytest.mark.gpu
def test_out_of_range_and_infeasible_utterances_fall_back_to_log_space():
    """Utterances with a gathered log-probability below the linear path's range (flagged by the sweep) run the
    log-space recursion; a structurally infeasible one (T_b < U_b + repeats) is settled without any recursion
    (round 2: it used to run the linear recursion to its underflow and then the log-space one); the rest of the
    batch stays linear."""
    c = make_case(6, 60, 47, 9, 31337, dist="D1", n_infeasible=1)      # utterance 1 is infeasible
    c["logits"][3] *= 60.0                                              # gaps of hundreds of nats
    c["logits"][4, 7, int(c["targets"][4, 0])] = -400.0                # one gathered value far out of range
    nll, g, stats = _raw_loss_grad(c)
    assert stats == [2, 0], stats                                      # [log-space, underflowed in the linear domain]
    n64, g64 = _f64(c)
    assert nll[1] == 0.0 and not g[1].any()                            # zero_infinity
    fin = np.isfinite(n64)
    assert np.abs(nll[fin] - n64[fin]).max() / np.abs(n64[fin]).max() < 1e-5
    assert np.abs(g - g64).max() < 1e-3
    # the flags are per call: the same workspace shape on ordinary inputs is all-linear again
    _, _, stats2 = _raw_loss_grad(make_case(6, 60, 47, 9, 31338, dist="D1"))
    assert stats2 == [0, 0]


@pytest.mark.gpu
def test_skipping_negligible_occupancies_changes_nothing_measurable():
    """The patch kernel does not apply occupancies <= 2^-40 (each would cost a DRAM read-modify-write): against
    applying everything but exact zeros the gradient moves by less than 2^-39 of the gradient scale."""
    c = make_case(6, 200, 4234, 30, 606, dist="D1")
    outs = []
    from asr_chinese_e2e_b200 import _lib
    try:
        for bits in (0, 40):
            _lib.set_option("occ_skip_bits", bits)            # (environment variables are read once, at load)
            assert _lib.get_option("occ_skip_bits") == bits
            _, g = run_gpu(c, "sum")
            outs.append(g)
    finally:
        _lib.set_option("occ_skip_bits", 40)
    diff = (outs[0] - outs[1]).abs().max().item()
    assert diff <= 2.0 ** -39                                   # occupancy + at most half an ulp of the result


@pytest.mark.parametrize("variant", [1, 2, 12, 3])
def test_direct_sweep_option_is_parity_green(variant):
    """Every sweep kernel behind the `sweep_direct` option (1: k1d_sweep, aligned frame groups, direct loads, one barrier
    per group; 2: k1p_sweep, the same groups fed by a bulk-TMA ring, two CTAs per SM -- what 0 = auto picks for even V
    and T; 3: k1_lse_gather, one frame per ring slot): same parity suite, including an odd last frame, infeasible
    utterances, the non-fused paths and V % 4 == 0; shapes the group kernels do not cover (odd V, odd T) silently use
    k1_lse_gather."""
    from asr_chinese_e2e_b200 import _lib
    try:
        _lib.set_option("sweep_direct", variant % 10)
        _lib.set_option("k1p_bulkst", 0 if variant == 12 else 1)     # 12: k1p_sweep with per-thread STG.128 stores
        for (B, T, V, U, seed) in ((7, 62, 54, 13, 21), (5, 40, 64, 7, 22), (3, 200, 4234, 30, 23), (4, 31, 53, 6, 24)):
            c = make_case(B, T, V, U, seed, dist="D2", n_infeasible=1, n_partial=1)
            c["input_lengths"][0] = T - 1 if T % 2 == 0 else T            # an odd number of valid frames
            check_case(c, True, f"direct sweep V={V} T={T}")
        c = make_config("C1")
        check_case(c, False, "direct sweep C1")
        from asr_chinese_e2e_b200 import ctc_greedy_cer_b200
        from asr_chinese_e2e_b200.joint import greedy_ctc_ids
        c = make_case(5, 200, 4234, 30, 77, dist="D2")
        _, info = ctc_greedy_cer_b200(c["logits"].cuda(), c["targets"].cuda(), c["input_lengths"].cuda(), c["target_lengths"].cuda())
        want = greedy_ctc_ids(c["logits"], c["input_lengths"])
        for b in range(5):
            assert info["hyp"][b, : int(info["hyp_len"][b])].tolist() == want[b]
    finally:
        _lib.set_option("sweep_direct", 0)
        _lib.set_option("k1p_bulkst", 1)
