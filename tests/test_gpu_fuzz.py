"""GPU: seeded random-shape fuzz of the CUDA path against the CPU oracle (layer 3).
Shapes cover every lattice width (U up to 255), vocabularies with every row misalignment, T from 1,
empty targets, infeasible and partial-lattice utterances, both zero_infinity values, all reductions,
both target encodings.  Kept small so the oracle (torch CPU + float64 C) finishes in seconds."""
import numpy as np
import pytest
import torch

from oracle.c_oracle import ctc_c_f64
from oracle.synth import make_case
from oracle.torch_ref import ref_ctc

pytestmark = pytest.mark.gpu


def _cases():
    rng = np.random.default_rng(20261018)
    out = []
    for i in range(36):
        umax = int(rng.choice([0, 1, 2, 5, 17, 40, 63, 64, 100, 127, 128, 200, 255]))
        B = int(rng.integers(1, 7))
        T = int(rng.choice([1, 2, 3, 9, 31, 64, 130, 257])) if umax < 64 else int(rng.integers(umax + 5, 2 * umax + 60))
        V = int(rng.integers(max(3, min(umax, 40) + 2), 300)) if i % 5 else int(rng.choice([4233, 4234, 4235, 4236]))
        out.append((B, T, V, umax, 1000 + i, "D2" if i % 2 else "D1", bool(i % 3 == 0), ["mean", "sum", "none"][i % 3]))
    return out


@pytest.mark.parametrize("B,T,V,umax,seed,dist,zi,red", _cases())
def test_fuzz_case(B, T, V, umax, seed, dist, zi, red):
    from asr_chinese_e2e_b200 import ctc_loss_b200
    c = make_case(B, T, V, umax, seed, dist=dist, n_infeasible=1 if B > 2 and umax > 0 else 0,
                  n_partial=1 if B > 3 and umax > 1 else 0)
    if B > 1 and umax > 0:
        c["target_lengths"][-1] = 0
        c["targets"][-1] = 0
    c["input_lengths"].clamp_(max=T)
    tl = c["target_lengths"]
    tg = c["targets"]
    if seed % 4 == 0 and umax > 0:                                   # 1-D concatenated encoding
        tg = torch.cat([c["targets"][b, : tl[b]] for b in range(B)]) if int(tl.sum()) else torch.zeros(0, dtype=torch.int64)
    x = c["logits"].cuda().requires_grad_(True)
    loss = ctc_loss_b200(x, tg.cuda(), c["input_lengths"].cuda(), tl.cuda(), reduction=red, zero_infinity=zi,
                         fused=bool(seed % 2))
    go = torch.ones_like(loss) if red != "none" else torch.linspace(0.5, 1.5, B, device="cuda")
    loss.backward(go)
    rl, rg = ref_ctc(c["logits"], tg, c["input_lengths"], tl, reduction=red, zero_infinity=zi, grad_output=go.cpu())
    got, want = loss.detach().cpu().double().reshape(-1).numpy(), rl.double().reshape(-1).numpy()
    assert np.array_equal(np.isinf(got), np.isinf(want)) and not np.isnan(got).any()
    fin = np.isfinite(want)
    assert np.all(np.abs(got[fin] - want[fin]) <= 1e-5 * np.maximum(np.abs(want[fin]), 1.0))
    g, r = x.grad.cpu().numpy(), rg.numpy()
    assert np.array_equal(np.isnan(g), np.isnan(r))
    ok = ~np.isnan(r)
    # adjudicate with the float64 oracle: with few utterances the 'mean' scale 1/(B*U_b) is weak and BOTH fp32
    # implementations carry the log-space rounding of |alpha| ~ T*log V (tools/dbg_accuracy.py: ours is as close to
    # float64 as torch or closer), so "ours vs torch" alone would test the sum of two errors.
    _, _, g64 = ctc_c_f64(c["logits"].numpy(), c["targets"].numpy(), c["input_lengths"].numpy(), tl.numpy(),
                          reduction=("sum" if red == "none" else red), zero_infinity=zi)
    g64 = g64 * (go.cpu().numpy().reshape(-1, 1, 1) if red == "none" else 1.0)
    terr = np.abs(r[ok] - g64[ok]).max(initial=0.0)
    oerr = np.abs(g[ok] - g64[ok]).max(initial=0.0)
    floor = 1e-4 if red == "mean" else 1e-3
    assert oerr <= max(floor, 3 * terr), (oerr, terr)
    if red == "mean":
        assert np.abs(g[ok] - r[ok]).max(initial=0.0) <= max(1e-4, 2 * (terr + oerr))
