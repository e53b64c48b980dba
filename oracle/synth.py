"""Seeded synthetic AISHELL-shaped inputs.  TEST INFRASTRUCTURE.

Everything is generated on the CPU with a fixed-seed ``torch.Generator`` so the CPU
oracle and the GPU path see identical bits.  Tensor contract = the reference's
collate function (data/data_loader/ai_shell_1.py:75-88): ``targets`` int64
``[B, Umax]`` padded with 0 (no BOS/EOS), int64 length vectors, labels >= 1
(ids 0..3 are specials in Predictor/data_handler/vocab.py:10,17; only 0 = PAD =
blank is excluded here).
"""
from __future__ import annotations

import math
import torch

CONFIGS = {
    # name: (B, T, V, Umax, seed)
    "C1": (16, 200, 4234, 30, 1001),
    "C2": (256, 400, 4234, 50, 1002),
    "C4": (64, 1500, 4234, 120, 1004),
}


def make_targets(B, Umax, V, gen, min_frac=0.2, repeat_p=0.1, full=False):
    lo = max(1, math.ceil(Umax * min_frac)) if Umax > 0 else 0
    if Umax == 0:
        return torch.zeros(B, 1, dtype=torch.int64), torch.zeros(B, dtype=torch.int64)
    tl = torch.randint(lo, Umax + 1, (B,), generator=gen, dtype=torch.int64)
    if full:
        tl[:] = Umax
    lab = torch.randint(1, V, (B, Umax), generator=gen, dtype=torch.int64)
    rep = torch.rand(B, Umax, generator=gen) < repeat_p
    for i in range(1, Umax):
        lab[:, i] = torch.where(rep[:, i], lab[:, i - 1], lab[:, i])
    mask = torch.arange(Umax)[None, :] < tl[:, None]
    return lab * mask, tl


def n_repeats(targets, tl):
    r = torch.zeros_like(tl)
    for b in range(targets.shape[0]):
        u = int(tl[b])
        if u > 1:
            r[b] = int((targets[b, 1:u] == targets[b, : u - 1]).sum())
    return r


def make_lengths(B, T, gen, full=False):
    if full:
        return torch.full((B,), T, dtype=torch.int64)
    il = torch.randint(math.ceil(T / 2), T + 1, (B,), generator=gen, dtype=torch.int64)
    il[0] = T
    return il


def random_alignment(tg_row, u, t_len, gen):
    """A random valid CTC alignment (frame -> class id) of labels tg_row[:u] over t_len frames."""
    ext = [0]
    for k in range(u):
        ext += [int(tg_row[k]), 0]
    need = u + sum(1 for k in range(1, u) if tg_row[k] == tg_row[k - 1])
    if t_len < need:
        return [0] * t_len
    # choose, for each label, a frame; keep strictly increasing with a gap where labels repeat
    pos, prev, out = [], -1, [0] * t_len
    slack = t_len - need
    cuts = sorted(torch.randint(0, slack + 1, (u,), generator=gen).tolist())
    base = 0
    for k in range(u):
        gap = 1 if (k > 0 and tg_row[k] == tg_row[k - 1]) else 0
        base += gap
        p = base + cuts[k]
        out[min(p, t_len - 1)] = int(tg_row[k])
        base += 1
    return out


def make_logits(B, T, V, gen, dist="D1", targets=None, tl=None, il=None, peak=8.0):
    x = torch.randn(B, T, V, generator=gen, dtype=torch.float32)
    if dist == "D2":
        for b in range(B):
            a = random_alignment(targets[b], int(tl[b]), int(il[b]), gen)
            idx = torch.tensor(a, dtype=torch.int64)
            x[b, torch.arange(len(a)), idx] += peak
    return x


def make_case(B, T, V, Umax, seed, dist="D1", full_lengths=False, full_targets=False,
              n_infeasible=0, n_partial=0):
    """Returns dict(logits, targets, input_lengths, target_lengths)."""
    gen = torch.Generator().manual_seed(seed)
    targets, tl = make_targets(B, Umax, V, gen, full=full_targets)
    il = make_lengths(B, T, gen, full=full_lengths)
    rep = n_repeats(targets, tl)
    # guarantee feasibility: T_b >= U_b + rep_b (configs other than C4 have T >= 2 Umax anyway)
    il = torch.maximum(il, torch.minimum(tl + rep, torch.tensor(T)))
    k = 1
    for _ in range(n_infeasible):            # truly infeasible: in_len = U + rep - 1
        il[k] = min(max(int(tl[k] + rep[k]) - 1, 1), T)
        k += 1
    for _ in range(n_partial):               # partial lattice band: U+rep <= in_len < 2U+1
        lo, hi = int(tl[k] + rep[k]), int(2 * tl[k])
        il[k] = min(max(lo, (lo + hi) // 2), T)
        k += 1
    il = il.clamp(max=T)
    logits = make_logits(B, T, V, gen, dist, targets, tl, il)
    return dict(logits=logits, targets=targets, input_lengths=il, target_lengths=tl)


def make_config(name, dist="D1", **kw):
    B, T, V, Umax, seed = CONFIGS[name]
    if name == "C4":
        kw.setdefault("n_infeasible", 8)
        kw.setdefault("n_partial", 8)
    return make_case(B, T, V, Umax, seed, dist=dist, **kw)
