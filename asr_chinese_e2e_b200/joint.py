"""Model-side glue: the joint CTC/attention objective under the reference's model API.

The reference's model contract (Predictor/Bases/base_model.py:24-71) is
``forward(Pack) -> Pack``, ``cal_metrics(output, input) -> Pack(loss, cer)``,
``iterate(input, optimizer, is_train) -> (Pack, None)``; ``Trainer11.train_epoch``
(Trainer/trainer11.py:51-80) and ``BaseTrainer.train_epoch`` (Trainer/base_trainer.py:52-71) only
ever see ``model.iterate`` and read ``metrics.loss`` / ``metrics.cer`` (every Pack value must be a
tensor: trainer11.py:108-112, metric_manager.py:24-26).  ``JointCTCAttention`` is a mix-in that adds
the CTC branch at the two insertion points SURVEY.md section 3.1 names, leaving the trainers
untouched:

  forward():      ctc_logits = ctc_head(encoder_out)           (tap: transformer_official.py:76)
  cal_metrics():  loss = w * ctc_loss_b200(ctc_logits, tgt_for_input, wave_len, tgt_len)
                         + (1 - w) * attention cross-entropy    (sibling of cal_performance, :86)

``tgt_for_input`` / ``tgt_len`` are exactly the CTC targets/lengths (no BOS/EOS at collate time,
data/data_loader/ai_shell_1.py:52-53,75-88) and the encoder does no time subsampling, so
``wave_len`` are the CTC input lengths.  New config keys ``ctc_weight`` (0.3) and
``ctc_zero_infinity`` (True) ride on the existing ``get_default_config`` mechanism.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .ce import attention_ce_b200
from .ctc import ctc_loss_b200
from .head import ctc_head_loss_b200
from .metrics import seq_cer_b200

IGNORE_ID = 0   # PAD id of the reference vocab == CTC blank (Predictor/Utils/loss.py:5, vocab.py:10)


class Pack(dict):
    """Attribute-access dict used as batch / output / metrics container (mirrors the behaviour of the
    reference's Predictor/Utils/pack.py:3-27: missing keys read as None, ``add(**kw)``, ``cuda()``)."""

    def __getattr__(self, name):
        return self.get(name)

    def __setattr__(self, name, value):
        self[name] = value

    def add(self, **kwargs):
        self.update(kwargs)
        return self

    def to(self, device):
        move = lambda v: tuple(move(i) for i in v) if isinstance(v, tuple) else (
            v.to(device, non_blocking=True) if torch.is_tensor(v) else v)
        return Pack({k: move(v) for k, v in self.items()})

    def cuda(self):
        return self.to("cuda")


def edit_distance(a, b) -> int:
    """Levenshtein distance between two sequences (host metric; the reference calls
    python-Levenshtein in Predictor/Utils/score.py:4-13)."""
    if len(a) < len(b):
        a, b = b, a
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


def reference_cer(hyp_ids, gold_ids, pad: int = IGNORE_ID) -> float:
    """Host restatement of the reference's ``cer`` (transformer_official.py:87-91): per utterance the Levenshtein
    distance between the SPACE-JOINED strings of the non-PAD ids (vocab.py:74-78, score.py:4-13; every token is one
    character, so the joined string is the symbol sequence t1 SP t2 ... tn), divided by the number of gold words
    (``len(''.split(' ')) == 1``), times 100, averaged over the batch.  Used on CPU tensors and as the checker of the
    device kernel (metrics.seq_cer_b200)."""
    def joined(ids):
        toks = [int(t) for t in ids if int(t) != pad]
        out = []
        for k, t in enumerate(toks):
            if k:
                out.append(-7)
            out.append(t)
        return out, max(len(toks), 1)
    tot, n = 0.0, 0
    for h, g in zip(hyp_ids, gold_ids):
        hs, _ = joined(h)
        gs, words = joined(g)
        tot += edit_distance(hs, gs) / words
        n += 1
    return tot * 100.0 / max(n, 1)


def attention_ce(pred, gold, smoothing: float = 0.0):
    """Attention-branch loss with the reference's convention (Predictor/Utils/loss.py:7-51):
    pred [N,T,C] logits, gold [N,T] with PAD=0 ignored, scalar mean over non-pad tokens."""
    pred = pred.reshape(-1, pred.size(-1))
    gold = gold.reshape(-1)
    if smoothing > 0.0:
        n_class = pred.size(1)
        logp = F.log_softmax(pred, dim=1)
        nonpad = gold.ne(IGNORE_ID)
        nll = -logp.gather(1, gold.unsqueeze(1)).squeeze(1)
        smooth = -logp.sum(dim=1)
        # target distribution: (1-eps) on the label + eps/C elsewhere  (loss.py:38-39)
        loss = (1.0 - smoothing - smoothing / n_class) * nll + (smoothing / n_class) * smooth
        return loss.masked_select(nonpad).sum() / nonpad.sum().clamp(min=1)
    return F.cross_entropy(pred, gold, ignore_index=IGNORE_ID, reduction="mean")


def greedy_ctc_ids(ctc_logits, lengths, blank: int = 0):
    """Best-path decode: argmax per frame, collapse repeats, drop blanks (host lists)."""
    best = ctc_logits.argmax(-1).cpu()
    out = []
    for row, n in zip(best, lengths.tolist()):
        ids, prev = [], blank
        for v in row[:n].tolist():
            if v != prev and v != blank:
                ids.append(v)
            prev = v
        out.append(ids)
    return out


class JointCTCAttention:
    """Mix-in for a reference-style encoder/decoder model (class M(JointCTCAttention, BaseModel)).

    The host class provides ``self.encoder(wave, wave_len) -> (enc, ...)`` and
    ``self.decoder(tgt, enc, lens) -> (pred, gold, ...)`` like ``TransformerOffical``
    (transformer_official.py:68-81).  Call ``init_ctc`` at the end of ``__init__``.
    """

    ctc_weight: float = 0.3
    ctc_zero_infinity: bool = True
    ctc_fused_head: bool = False
    ctc_head_precision: str = "3xtf32"

    def init_ctc(self, d_model: int, vocab_size: int, ctc_weight: float = 0.3, ctc_zero_infinity: bool = True,
                 smoothing: float = 0.0, ctc_cer_on_device: bool = True, fused_head: bool = False,
                 head_precision: str = "3xtf32"):
        self.ctc_head = torch.nn.Linear(d_model, vocab_size)
        self.ctc_weight = float(ctc_weight)
        self.ctc_zero_infinity = bool(ctc_zero_infinity)
        self.att_smoothing = float(smoothing)
        # on CUDA batches: adds a `ctc_cer` metric (greedy CTC decode + edit distance of the same forward pass,
        # device-side, no sync); the attention-branch `cer` is computed on the device either way
        self.ctc_cer_on_device = bool(ctc_cer_on_device)
        # f1: compute ctc_head and the CTC loss in one tcgen05 kernel pair (head.py); the [B,T,V] CTC logits then never
        # exist, forward() hands the encoder output on instead (and there is no `ctc_cer` by-product)
        self.ctc_fused_head = bool(fused_head)
        self.ctc_head_precision = head_precision

    # -- forward: encoder tap + CTC head, then the decoder exactly as the reference calls it --------
    def forward(self, input):
        enc, *_ = self.encoder(input.wave, input.wave_len)
        pred, gold, *_ = self.decoder(input.tgt_for_input, enc, input.tgt_len)
        if self.ctc_fused_head and enc.is_cuda and enc.shape[-1] % 32 == 0:
            return Pack(pred=pred, gold=gold, ctc_enc=enc)
        return Pack(pred=pred, gold=gold, ctc_logits=self.ctc_head(enc))

    def joint_loss(self, output, input):
        if output.ctc_logits is None and output.ctc_enc is None:   # Pack returns None for a missing key (pack.py:7-8)
            raise KeyError("output Pack has no 'ctc_logits': forward() must add the CTC head output")
        w = self.ctc_weight
        eps = getattr(self, "att_smoothing", 0.0)
        if output.pred.is_cuda and w < 1.0:
            # attention branch on the same sweep kernels, (1-w) folded in (no rescaling sweep in backward)
            att = attention_ce_b200(output.pred.float(), output.gold, eps, weight=1.0 - w) / (1.0 - w)
        else:
            att = attention_ce(output.pred, output.gold, eps)
        B = (output.ctc_logits if output.ctc_logits is not None else output.ctc_enc).shape[0]
        if output.ctc_logits is None:          # fused head: logits = ctc_head(enc) never materialised
            self._last_decode = None
            wctc = ctc_head_loss_b200(output.ctc_enc.float(), self.ctc_head.weight, self.ctc_head.bias,
                                      input.tgt_for_input, input.wave_len, input.tgt_len, blank=IGNORE_ID,
                                      reduction="mean", zero_infinity=self.ctc_zero_infinity,
                                      inv_batch=(w if w > 0 else 1.0) / max(B, 1), precision=self.ctc_head_precision)
            if w > 0:
                return wctc + (1.0 - w) * att, wctc / w, att
            return att + 0.0 * wctc, wctc, att
        # ctc_weight is folded into the op's normaliser (inv_batch = w/B): the op returns w*ctc and its
        # speculative gradient is already the final one, so backward() costs one empty launch instead of a
        # rescaling sweep over [B,T,V]
        dec = {"want_hyp": False} if (getattr(self, "ctc_cer_on_device", False) and output.ctc_logits.is_cuda) else None
        wctc = ctc_loss_b200(output.ctc_logits.float(), input.tgt_for_input, input.wave_len, input.tgt_len,
                             blank=IGNORE_ID, reduction="mean", zero_infinity=self.ctc_zero_infinity,
                             inv_batch=(w if w > 0 else 1.0) / max(B, 1), decode=dec)
        self._last_decode = dec
        if w > 0:
            return wctc + (1.0 - w) * att, wctc / w, att
        return att + 0.0 * wctc, wctc, att

    def cal_metrics(self, output, input):
        loss, ctc, att = self.joint_loss(output, input)
        assert not torch.isinf(loss)           # the reference's only numerical guard (transformer_official.py:88)
        # `cer` is the reference's metric (arg-max ids -> space-joined strings -> Levenshtein / #gold words,
        # transformer_official.py:87-91).  On CUDA it is one kernel launch and the value stays on the device: the train
        # step has no .tolist() / Python Levenshtein loop; Trainer11 reads it with .item() like every other metric.
        hyp = output.pred.topk(1)[1].squeeze(-1)   # the reference's own op (ties in all-zero padded rows resolve alike)
        if hyp.is_cuda:
            cer = seq_cer_b200(hyp, output.gold, pad=IGNORE_ID, mode="string")
        else:
            cer = torch.tensor([reference_cer(hyp.tolist(), output.gold.tolist())])
        pack = Pack(loss=loss, cer=cer, ctc_loss=ctc.detach(), att_loss=att.detach())
        dec = getattr(self, "_last_decode", None)
        if dec:   # CTC-branch CER of the same forward pass, computed entirely on the device (no sync here)
            tl = input.tgt_len.to(dec["edit_distance"].device).clamp(min=1).float()
            pack.add(ctc_cer=((dec["edit_distance"].float() / tl).mean() * 100.0).reshape(1))
        return pack

    def iterate(self, input, optimizer=None, is_train=True):
        output = self.forward(input)
        metrics = self.cal_metrics(output, input)
        if optimizer is not None and is_train:
            optimizer.zero_grad()
            metrics.loss.backward()
            torch.nn.utils.clip_grad_norm_(self.parameters(), 5.0)   # transformer_official.py:102
            optimizer.step()
        return metrics, None
