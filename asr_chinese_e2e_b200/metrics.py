"""The reference's ``cer`` metric on the device.

``TransformerOffical.cal_metrics`` (Predictor/Models/transformer_official.py:87-91) takes the arg-max ids of the
decoder output, turns hypothesis and gold into space-joined strings with ``Vocab.convert_id2str``
(Predictor/data_handler/vocab.py:74-78: PAD ids dropped) and averages ``calculate_cer``
(Predictor/Utils/score.py:4-13): ``Lev.distance(hyp_str, gold_str) / len(gold_str.split(' '))`` -- a Levenshtein
distance over the CHARACTERS of the joined strings, separators included.  That costs a device->host copy of the ids and
a Python loop per utterance every training step.  ``seq_cer_b200`` computes the same number with one kernel launch
(``ctcb200_edit_distance``, one warp per utterance) and returns a device tensor: nothing synchronises.

The string-level distance equals the distance between the symbol sequences ``t1 SP t2 SP ... tn`` because every token
of the reference's vocabulary is one character (``tokenize_fn`` in vocab.py:4-5 splits by character; the four special
tokens ``$ % ^ &`` are single characters too).
"""
from __future__ import annotations

import torch

from . import _lib

_MODES = {"token": 0, "string": 1}


def seq_edit_distance_b200(hyp: torch.Tensor, gold: torch.Tensor, pad: int = 0, mode: str = "string"):
    """hyp, gold: integer id matrices [B, L] on the GPU.  Returns (edit int32[B], words int32[B]) device tensors."""
    if mode not in _MODES:
        raise ValueError("mode must be 'string' (the reference's metric) or 'token'")
    if not (hyp.is_cuda and gold.is_cuda):
        raise _lib.CtcB200Error("seq_edit_distance_b200 needs CUDA tensors: no CPU fallback")
    if hyp.dim() != 2 or gold.dim() != 2 or hyp.shape[0] != gold.shape[0]:
        raise ValueError("hyp and gold must be [B, Lh] and [B, Lg]")
    B = hyp.shape[0]
    L = max(hyp.shape[1], gold.shape[1], 1)

    def prep(x):
        x = x.to(torch.int64)
        if x.shape[1] < L:
            x = torch.nn.functional.pad(x, (0, L - x.shape[1]), value=pad)
        return x.contiguous()
    h, g = prep(hyp), prep(gold)
    edit = torch.empty(B, dtype=torch.int32, device=h.device)
    words = torch.empty(B, dtype=torch.int32, device=h.device)
    with torch.cuda.device(h.device):
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().ctcb200_edit_distance(h.data_ptr(), L, g.data_ptr(), L, B, L, int(pad), _MODES[mode],
                                                    edit.data_ptr(), words.data_ptr(), st), "ctcb200_edit_distance")
    return edit, words


def seq_cer_b200(hyp: torch.Tensor, gold: torch.Tensor, pad: int = 0, mode: str = "string") -> torch.Tensor:
    """100 * mean_b(edit_b / words_b) as a 1-element float tensor on the device (the reference's ``cer`` value when
    mode='string').  No host synchronisation."""
    edit, words = seq_edit_distance_b200(hyp, gold, pad, mode)
    return ((edit.float() / words.float()).mean() * 100.0).reshape(1)
