"""Independent float64 restatement of CTC loss + gradient w.r.t. logits.

TEST INFRASTRUCTURE (see oracle/__init__.py).  numpy only, vectorised over the
batch and the lattice states; the only Python loop is over time.

Follows SURVEY.md Appendix B (Graves et al. 2006; torch convention in which both
alpha and beta include the emission term lp[t, l'_s]).  The reference repo has
no CTC code of its own (grep -rni ctc /root/reference is empty); the tensor
contract mirrored here is the reference's:
  * logits batch-major [B, T, V]        (Predictor/Utils/loss.py:10  "pred: N x T x C")
  * targets int64 [B, Umax] padded with 0 (data/data_loader/ai_shell_1.py:75-88,
    Predictor/data_handler/padder.py:6-27), or 1-D concatenated
  * lengths int64 [B]                    (ai_shell_1.py:80-84)
  * blank = 0 = PAD id                   (Predictor/data_handler/vocab.py:10,17)
"""
from __future__ import annotations

import numpy as np

NEG_INF = -np.inf


def _lse(*xs):
    """log(sum(exp(x_i))) element-wise over equally shaped arrays, -inf safe."""
    m = xs[0]
    for x in xs[1:]:
        m = np.maximum(m, x)
    ms = np.where(np.isfinite(m), m, 0.0)
    acc = np.zeros_like(ms)
    for x in xs:
        acc = acc + np.exp(x - ms)
    with np.errstate(divide="ignore"):
        return ms + np.log(acc)


def log_softmax_f64(logits):
    x = np.asarray(logits, dtype=np.float64)
    m = x.max(axis=-1, keepdims=True)
    lse = m + np.log(np.exp(x - m).sum(axis=-1, keepdims=True))
    return x - lse


def _padded_targets(targets, target_lengths, B):
    targets = np.asarray(targets)
    tl = np.asarray(target_lengths, dtype=np.int64)
    if targets.ndim == 2:
        return targets.astype(np.int64)
    umax = int(tl.max()) if B > 0 else 0
    out = np.zeros((B, max(umax, 1)), dtype=np.int64)
    off = 0
    for b in range(B):
        out[b, : tl[b]] = targets[off : off + tl[b]]
        off += int(tl[b])
    return out


def ctc_f64(logits, targets, input_lengths, target_lengths, blank=0,
            reduction="mean", zero_infinity=False, want_grad=True,
            grad_output=None):
    """Returns (loss, nll[B], grad_logits[B,T,V] or None) in float64.

    Semantics of F.ctc_loss(F.log_softmax(logits,-1).transpose(0,1), ...):
      'none' -> nll[B]; 'sum' -> sum_b nll_b; 'mean' -> mean_b(nll_b / max(U_b,1)).
      zero_infinity: inf -> 0 and that utterance's gradient slab -> 0.
      Without zero_infinity an infeasible utterance has nll=+inf and NaN gradient
      (torch behaviour).
    """
    x = np.asarray(logits, dtype=np.float64)
    B, T, V = x.shape
    il = np.asarray(input_lengths, dtype=np.int64)
    tl = np.asarray(target_lengths, dtype=np.int64)
    tg = _padded_targets(targets, tl, B)
    Umax = max(int(tl.max()) if B else 0, 0)
    S = 2 * Umax + 1
    lp = log_softmax_f64(x)                                  # [B,T,V]

    # extended label sequence l' (blank, y1, blank, ..., yU, blank)
    ext = np.full((B, S), blank, dtype=np.int64)
    if Umax > 0:
        ext[:, 1::2] = tg[:, :Umax]
    s_idx = np.arange(S)[None, :]
    valid = s_idx < (2 * tl[:, None] + 1)                    # [B,S]
    is_lab = (s_idx % 2 == 1) & valid
    # skip transition s-2 -> s allowed for label states whose label differs from l'_{s-2}
    skip_f = np.zeros((B, S), dtype=bool)
    if S > 2:
        skip_f[:, 2:] = is_lab[:, 2:] & (ext[:, 2:] != ext[:, :-2])
    skip_b = np.zeros((B, S), dtype=bool)                    # s -> s+2 seen from s
    if S > 2:
        skip_b[:, :-2] = is_lab[:, :-2] & valid[:, 2:] & (ext[:, 2:] != ext[:, :-2])

    bi = np.arange(B)[:, None]
    lpe = np.take_along_axis(lp, np.broadcast_to(ext[:, None, :], (B, T, S)), axis=2)  # [B,T,S]
    lpe = np.where(valid[:, None, :], lpe, NEG_INF)

    alpha = np.full((B, T, S), NEG_INF)
    beta = np.full((B, T, S), NEG_INF)
    has_t = il > 0
    # forward
    a0 = np.full((B, S), NEG_INF)
    a0[:, 0] = lpe[:, 0, 0]
    if S > 1:
        a0[:, 1] = np.where(tl > 0, lpe[:, 0, 1], NEG_INF)
    a0[~has_t] = NEG_INF
    alpha[:, 0] = a0
    for t in range(1, T):
        p = alpha[:, t - 1]
        p1 = np.concatenate([np.full((B, 1), NEG_INF), p[:, :-1]], axis=1)
        p2 = np.concatenate([np.full((B, 2), NEG_INF), p[:, :-2]], axis=1) if S > 2 else np.full((B, S), NEG_INF)
        p2 = np.where(skip_f, p2, NEG_INF)
        a = lpe[:, t] + _lse(p, p1, p2)
        alpha[:, t] = np.where((t < il)[:, None], a, NEG_INF)
    # log-likelihood from alpha at the last valid frame
    last = np.clip(il - 1, 0, T - 1)
    aT = alpha[np.arange(B), last]                           # [B,S]
    sl = 2 * tl                                              # index of final blank
    end1 = aT[np.arange(B), sl]
    end2 = np.where(tl > 0, aT[np.arange(B), np.maximum(sl - 1, 0)], NEG_INF)
    ll = _lse(end1, end2)
    ll = np.where(has_t, ll, np.where(tl == 0, 0.0, NEG_INF))
    nll = -ll
    # backward
    for b in range(B):
        if il[b] > 0:
            tb = il[b] - 1
            beta[b, tb, sl[b]] = lpe[b, tb, sl[b]]
            if tl[b] > 0:
                beta[b, tb, sl[b] - 1] = lpe[b, tb, sl[b] - 1]
    for t in range(T - 2, -1, -1):
        n = beta[:, t + 1]
        n1 = np.concatenate([n[:, 1:], np.full((B, 1), NEG_INF)], axis=1)
        n2 = np.concatenate([n[:, 2:], np.full((B, 2), NEG_INF)], axis=1) if S > 2 else np.full((B, S), NEG_INF)
        n2 = np.where(skip_b, n2, NEG_INF)
        bt = lpe[:, t] + _lse(n, n1, n2)
        upd = (t < il - 1)[:, None]
        beta[:, t] = np.where(upd, bt, beta[:, t])

    infeasible = ~np.isfinite(ll)
    nll_out = nll.copy()
    if zero_infinity:
        nll_out[infeasible] = 0.0
    denom = np.maximum(tl, 1).astype(np.float64)
    if reduction == "none":
        loss = nll_out
    elif reduction == "sum":
        loss = nll_out.sum()
    elif reduction == "mean":
        loss = (nll_out / denom).mean() if B else np.float64(0.0)
    else:
        raise ValueError(reduction)

    grad = None
    if want_grad:
        # posterior state occupancy gamma_t(s) = alpha*beta / (y_t(l'_s) * P)
        with np.errstate(invalid="ignore"):
            lg = alpha + beta - lpe - ll[:, None, None]
        lg = np.where(np.isfinite(alpha) & np.isfinite(beta), lg, NEG_INF)
        gam = np.exp(lg)                                      # [B,T,S]
        occ = np.zeros((B, T, V))
        for s in range(S):
            np.add.at(occ, (np.arange(B), slice(None), ext[:, s]), gam[:, :, s])
        grad = np.exp(lp) - occ
        tmask = np.arange(T)[None, :] < il[:, None]
        grad = grad * tmask[:, :, None]
        if grad_output is None:
            go = np.ones(B) if reduction == "none" else np.float64(1.0)
        else:
            go = np.asarray(grad_output, dtype=np.float64)
        if reduction == "none":
            scale = np.broadcast_to(go, (B,)).astype(np.float64)
        elif reduction == "sum":
            scale = np.full(B, float(go))
        else:
            scale = float(go) / (B * denom)
        grad = grad * scale[:, None, None]
        if zero_infinity:
            grad[infeasible] = 0.0
        else:
            # torch: every valid frame of an infeasible utterance is NaN, padded frames stay 0
            grad = np.where((infeasible[:, None] & tmask)[:, :, None], np.nan, grad)
    return loss, nll_out, grad
