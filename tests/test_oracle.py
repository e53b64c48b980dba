"""CPU: the float64 restatements (numpy and plain C) against the golden vectors made from
torch's CPU ctc_loss, plus known-answer properties."""
import numpy as np
import pytest
import torch

from oracle.c_oracle import ctc_c_f64
from oracle.ctc_f64 import ctc_f64
from oracle.synth import make_case
from oracle.torch_ref import ref_ctc

IMPLS = {"numpy": ctc_f64, "c": ctc_c_f64}


def _inputs(z, name):
    return (z[f"{name}/logits"], z[f"{name}/targets"], z[f"{name}/input_lengths"],
            z[f"{name}/target_lengths"])


@pytest.mark.parametrize("impl", ["numpy", "c"])
@pytest.mark.parametrize("zi", [False, True])
def test_oracle_matches_golden(golden, impl, zi):
    z, names = golden
    f = IMPLS[impl]
    for name in names:
        x, tg, il, tl = _inputs(z, name)
        want_nll = z[f"{name}/zi{int(zi)}/nll"]
        _, nll, _ = f(x, tg, il, tl, reduction="none", zero_infinity=zi, want_grad=False)
        assert np.array_equal(np.isinf(nll), np.isinf(want_nll)), name
        fin = np.isfinite(want_nll)
        # fp32 torch vs float64 truth: 1e-5 relative is the north-star tolerance
        assert np.allclose(nll[fin], want_nll[fin], rtol=1e-5, atol=1e-5), name
        for red in ("mean", "sum"):
            loss, _, grad = f(x, tg, il, tl, reduction=red, zero_infinity=zi)
            want = z[f"{name}/zi{int(zi)}/{red}/grad"]
            wl = z[f"{name}/zi{int(zi)}/{red}/loss"]
            if np.isfinite(wl):
                assert abs(loss - wl) <= 1e-5 * max(1.0, abs(wl)), (name, red)
            else:
                assert np.isinf(loss)
            assert np.array_equal(np.isnan(grad), np.isnan(want)), (name, red)
            ok = ~np.isnan(want)
            assert np.abs(grad[ok] - want[ok]).max() <= 1e-4, (name, red)


def test_numpy_and_c_agree_tightly():
    c = make_case(5, 50, 41, 12, 11, dist="D2", n_infeasible=1, n_partial=1)
    a = [c[k].numpy() for k in ("logits", "targets", "input_lengths", "target_lengths")]
    for zi in (False, True):
        l1, n1, g1 = ctc_f64(*a, reduction="sum", zero_infinity=zi)
        l2, n2, g2 = ctc_c_f64(*a, reduction="sum", zero_infinity=zi)
        fin = np.isfinite(n1)
        assert np.array_equal(fin, np.isfinite(n2))
        assert np.abs(n1[fin] - n2[fin]).max() < 1e-10
        assert np.array_equal(np.isnan(g1), np.isnan(g2))
        assert np.nanmax(np.abs(g1 - g2)) < 1e-12


def test_known_answers():
    V = 5
    x = np.zeros((1, 3, V), dtype=np.float32)                 # uniform: every path has prob V^-T
    # empty target: only the all-blank path
    _, nll, g = ctc_f64(x, np.zeros((1, 1), np.int64), [3], [0], reduction="none")
    assert np.isclose(nll[0], 3 * np.log(V))
    # single label over 3 frames: paths a--, -a-, --a, aa-, -aa, aaa  -> 6 paths
    _, nll, _ = ctc_f64(x, np.array([[2]]), [3], [1], reduction="none")
    assert np.isclose(nll[0], 3 * np.log(V) - np.log(6))
    # repeated label needs a separating blank: [2,2] infeasible at T=2, exactly one path at T=3
    _, nll, _ = ctc_f64(x[:, :2], np.array([[2, 2]]), [2], [2], reduction="none")
    assert np.isinf(nll[0])
    _, nll, g = ctc_f64(x, np.array([[2, 2]]), [3], [2], reduction="none")
    assert np.isclose(nll[0], 3 * np.log(V))
    # single path => occupancy is one-hot: grad = softmax - onehot(path 2,0,2)
    want = np.full((3, V), 1.0 / V)
    want[0, 2] -= 1; want[1, 0] -= 1; want[2, 2] -= 1
    assert np.allclose(g[0], want)


def test_gradient_properties_vs_torch():
    c = make_case(4, 30, 23, 7, 5, dist="D1")
    a = [c[k].numpy() for k in ("logits", "targets", "input_lengths", "target_lengths")]
    loss, nll, g = ctc_c_f64(*a, reduction="mean")
    lt, gt = ref_ctc(c["logits"], c["targets"], c["input_lengths"], c["target_lengths"])
    assert abs(loss - lt.item()) < 1e-5 * abs(loss)
    assert np.abs(g - gt.numpy()).max() < 1e-6
    il = a[2]
    for b in range(4):
        assert np.all(g[b, il[b]:] == 0)                       # padded frames: exact zeros
        assert np.abs(g[b, : il[b]].sum(-1)).max() < 1e-12     # softmax minus a distribution
    # 'mean' == mean_b(nll_b / max(U_b,1))
    assert np.isclose(loss, np.mean(nll / np.maximum(a[3], 1)))


def test_finite_difference():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((1, 7, 6)).astype(np.float32)
    tg, il, tl = np.array([[1, 1, 3]]), [7], [3]
    _, _, g = ctc_f64(x, tg, il, tl, reduction="sum")
    eps = 1e-3
    for (t, v) in [(0, 1), (3, 0), (6, 3), (2, 5)]:
        xp, xm = x.astype(np.float64).copy(), x.astype(np.float64).copy()
        xp[0, t, v] += eps; xm[0, t, v] -= eps
        fp = ctc_f64(xp, tg, il, tl, reduction="sum", want_grad=False)[0]
        fm = ctc_f64(xm, tg, il, tl, reduction="sum", want_grad=False)[0]
        assert abs((fp - fm) / (2 * eps) - g[0, t, v]) < 1e-6
