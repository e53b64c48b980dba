"""Test helper: import the reference package in THIS container (it never exists on the GPU box).

The reference imports six third-party modules that are not installed here; they are irrelevant to the
CTC path, so they are replaced by in-memory stubs.  ``Levenshtein.distance`` -- the one stubbed function the metric
path really calls (Predictor/Utils/score.py:13) -- is given a plain Levenshtein DP.
"""
import os
import sys
import types

REF_ROOT = "/root/reference"
HAVE_REFERENCE = os.path.isdir(os.path.join(REF_ROOT, "Predictor"))


def lev(a, b):
    if len(a) < len(b):
        a, b = b, a
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return lambda *a, **k: None


def mount_reference(monkeypatch):
    """Make ``import Predictor`` / ``import Trainer`` resolve to /root/reference for the duration of one test."""
    for name in ("fire", "Levenshtein", "python_speech_features", "librosa", "librosa.core", "seaborn", "pyaudio"):
        monkeypatch.setitem(sys.modules, name, _Stub(name))
    sys.modules["Levenshtein"].distance = lev
    for k in [k for k in sys.modules if k.split(".")[0] in ("Predictor", "Trainer", "data")]:
        monkeypatch.delitem(sys.modules, k)
    monkeypatch.syspath_prepend(REF_ROOT)


class CharVocab:
    """The id <-> character contract of Predictor/data_handler/vocab.py (PAD '$' = 0, UNK '%', BOS '^', EOS '&',
    then one id per character; convert_id2str drops PAD and joins with spaces, :74-78)."""

    def __init__(self, n=40):
        self._id2token = ["$", "%", "^", "&"] + [chr(0x4E00 + i) for i in range(n - 4)]
        self.vocab_size = n

    def convert_id2str(self, ids):
        return " ".join(self._id2token[int(i)] for i in ids if int(i) != 0)
