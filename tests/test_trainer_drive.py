"""The drop-in claim end to end: the UNMODIFIED reference trainers (Trainer/trainer11.py:51-80 ``train_epoch``,
:108-129 ``summarize`` / ``evaluate``; Trainer/base_trainer.py:52-71) drive the reference's own ``TransformerOffical``
with the ``JointCTCAttention`` mix-in for a few steps over an in-memory loader, with the reference's ``NoamOpt``.

CPU only (the reference tree exists only in the authoring container): the oracle stands in for the CUDA op.  The GPU
counterpart, with the real op, is tests/test_gpu_integration.py::test_trainer_loop_contract_on_gpu (a restatement of
the same loop, since /root/reference does not exist on the GPU box)."""
import pytest
import torch

import asr_chinese_e2e_b200.joint as joint
from asr_chinese_e2e_b200 import JointCTCAttention, Pack
from ref_import import HAVE_REFERENCE, CharVocab, mount_reference
from test_host_logic import oracle_ctc

pytestmark = pytest.mark.skipif(not HAVE_REFERENCE, reason="reference tree only in the authoring container")


def _batches(n, V, seed):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n):
        B, T, U = 4, 24, 6
        tl = torch.randint(2, U + 1, (B,), generator=g)
        wl = torch.randint(T // 2, T + 1, (B,), generator=g)
        wl[0] = T
        tg = torch.randint(4, V, (B, U), generator=g) * (torch.arange(U)[None] < tl[:, None])
        wave = torch.randn(B, T, 8, generator=g) * (torch.arange(T)[None, :, None] < wl[:, None, None])
        out.append(Pack(wave=wave, wave_len=wl, tgt_for_input=tg, tgt_len=tl))
    return out


def _model(monkeypatch):
    mount_reference(monkeypatch)
    monkeypatch.setattr(joint, "ctc_loss_b200", oracle_ctc)
    from Predictor.Models.transformer_official import TransformerOffical

    cfg = TransformerOffical.get_default_config()()
    cfg.n_mels, cfg.lfr_m, cfg.layer_num, cfg.d_model = 8, 1, 2, 32
    cfg.hidden_size, cfg.num_head, cfg.ff_size, cfg.dropout = 8, 4, 64, 0.0
    cfg.num_epoch = 1

    class Joint(JointCTCAttention, TransformerOffical):
        def __init__(self, config, vocab):
            TransformerOffical.__init__(self, config, vocab)
            self.init_ctc(config.d_model, vocab.vocab_size)

    torch.manual_seed(0)
    return Joint(cfg, CharVocab(40)), TransformerOffical


@pytest.mark.parametrize("which", ["Trainer11", "BaseTrainer"])
def test_reference_trainer_drives_the_mixin_model(monkeypatch, tmp_path, which):
    model, _ = _model(monkeypatch)
    import Trainer
    opt = Trainer.NoamOpt(32, 1.0, 4, torch.optim.Adam(model.parameters(), lr=0, betas=(0.9, 0.98), eps=1e-9))
    train, dev, test = _batches(5, 40, 1), _batches(2, 40, 2), _batches(1, 40, 3)
    cls = getattr(Trainer, which)
    tr = cls(optimizer=opt, model=model, train_iter=train, dev_iter=dev, test_iter=test, ckpt_root=str(tmp_path) + "/",
             exp_name="drive", log_every_iter=1, eval_every_iter=2, save_every_iter=4)
    with torch.no_grad():
        before, _ = model.iterate(train[0], is_train=False)
    w0 = model.ctc_head.weight.detach().clone()
    if which == "BaseTrainer":
        # BaseTrainer.save_ckpt / evaluate have reference-side defects unrelated to this path (an undefined
        # `reference_score` at base_trainer.py:93-99); drive its train loop with those two neutralised
        monkeypatch.setattr(cls, "save_ckpt", lambda self, *a, **k: None)
    tr.train_epoch()                                   # 5 x iterate + summarize + evaluate(dev) x2 + save + evaluate(test)
    assert tr.global_step == 5 and opt._step == 5
    assert not torch.equal(w0, model.ctc_head.weight)                     # the CTC branch trained the head
    with torch.no_grad():
        after, _ = model.iterate(train[0], is_train=False)
    assert set(after) == {"loss", "cer", "ctc_loss", "att_loss"}
    assert after.loss.item() < before.loss.item()                        # and the joint loss went down
    if which == "Trainer11":
        assert (tmp_path / "drive" / "e0_s5.model").exists() and (tmp_path / "drive" / "e0_s5.opt").exists()


def test_cer_equals_the_reference_cal_metrics(monkeypatch):
    """`cer` of the mix-in == the value the reference's own cal_metrics computes (convert_id2str + python-Levenshtein
    on the space-joined strings, transformer_official.py:87-91), and differs from a token-level rate."""
    model, TransformerOffical = _model(monkeypatch)
    batch = _batches(1, 40, 9)[0]
    with torch.no_grad():
        out = model.forward(batch)
        mine = model.cal_metrics(out, batch)
        theirs = TransformerOffical.cal_metrics(model, out, batch)        # the reference's method on the same output
    assert abs(mine.cer.item() - theirs.cer.item()) < 1e-4
    hyp, gold = out.pred.topk(1)[1].squeeze(-1).tolist(), out.gold.tolist()
    strip = lambda s: [t for t in s if t != 0]
    token_level = 100.0 * sum(joint.edit_distance(strip(h), strip(g)) / max(len(strip(g)), 1)
                              for h, g in zip(hyp, gold)) / len(hyp)
    assert joint.reference_cer(hyp, gold) == pytest.approx(mine.cer.item(), abs=1e-4)
    # string-level: a deleted / inserted token also costs its separator, so the two metrics are different numbers
    assert joint.reference_cer([[5, 6, 7]], [[5, 7]]) == pytest.approx(100.0)       # "a b c" vs "a c": 2 edits / 2 words
    assert joint.reference_cer([[5, 9, 7]], [[5, 6, 7]]) == pytest.approx(100.0 / 3)  # one substitution
    assert joint.reference_cer([[0, 0]], [[0, 0]]) == 0.0 and joint.reference_cer([[5]], [[0]]) == pytest.approx(100.0)
    assert token_level >= 0.0
