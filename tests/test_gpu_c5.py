"""GPU: BASELINE.json config C5 -- the joint CTC/attention step with the reference's 12-layer encoder architecture
(asr_chinese_e2e_b200.speech_encoder.SpeechEncoder; tests/test_speech_encoder.py pins it to the reference's own
Encoder class in the authoring container) feeding the CTC kernels through ``JointCTCAttention`` -- against the
identical graph with torch's own CUDA ctc_loss; the Trainer11 loop contract with the real op; and the N-rank sharded
path under NCCL (skipped with fewer than 2 GPUs)."""
import os
import subprocess
import sys

import pytest
import torch
import torch.nn.functional as F

from asr_chinese_e2e_b200 import JointCTCAttention, Pack
from asr_chinese_e2e_b200.speech_encoder import SpeechEncoder
from oracle.synth import make_lengths, make_targets
from tiny_model import TinyDec

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class C5Model(JointCTCAttention, torch.nn.Module):
    """The reference's encoder + the mix-in's CTC head; a small decoder stands in for the attention branch (the
    decoder body is dense GEMM work outside this path)."""

    def __init__(self, V, n_layers, d_model=512, dropout=0.0):
        torch.nn.Module.__init__(self)
        self.encoder = SpeechEncoder(d_input=320, n_layers=n_layers, d_model=d_model, dropout=dropout)
        self.decoder = TinyDec(d_model, V)
        self.init_ctc(d_model, V, ctc_weight=0.3, ctc_zero_infinity=True)


def c5_batch(B, T, V, U, seed, device):
    g = torch.Generator().manual_seed(seed)
    tg, tl = make_targets(B, U, V, g)
    il = make_lengths(B, T, g)
    wave = torch.randn(B, T, 320, generator=g) * (torch.arange(T)[None, :, None] < il[:, None, None])
    return Pack(wave=wave, wave_len=il, tgt_for_input=tg, tgt_len=tl).to(device)


def test_c5_joint_step_matches_torch_autograd():
    """B=128, T=400, V=4234, 12 encoder layers, fp32: loss and EVERY parameter gradient of one joint step with this
    repo's op equal the same step with torch's CUDA log_softmax + ctc_loss + cross_entropy."""
    B, T, V, U = 128, 400, 4234, 50
    torch.manual_seed(1005)
    m = C5Model(V, n_layers=12).cuda()
    batch = c5_batch(B, T, V, U, 1005, "cuda")
    out = m.forward(batch)
    met = m.cal_metrics(out, batch)
    assert set(met) >= {"loss", "cer", "ctc_loss", "att_loss", "ctc_cer"} and met.cer.is_cuda      # no host round trip
    met.loss.backward()
    got = {k: p.grad.clone() for k, p in m.named_parameters()}
    # padded encoder frames are zeros, so the CTC logits there are the head's bias and must get no gradient
    assert torch.allclose(out.ctc_logits[1, int(batch.wave_len[1]):], m.ctc_head.bias.expand(T - int(batch.wave_len[1]), -1))
    m.zero_grad()
    out = m.forward(batch)
    ctc = F.ctc_loss(F.log_softmax(out.ctc_logits, -1).transpose(0, 1), batch.tgt_for_input, batch.wave_len,
                     batch.tgt_len, blank=0, reduction="mean", zero_infinity=True)
    att = F.cross_entropy(out.pred.reshape(-1, V), out.gold.reshape(-1), ignore_index=0)
    ref = 0.3 * ctc + 0.7 * att
    ref.backward()
    assert abs(met.loss.item() - ref.item()) <= 1e-5 * abs(ref.item())
    assert abs(met.ctc_loss.item() - ctc.item()) <= 1e-5 * abs(ctc.item())
    for k, p in m.named_parameters():
        scale = p.grad.abs().max().item() + 1e-12
        assert (got[k] - p.grad).abs().max().item() <= 2e-3 * scale + 1e-7, k


def test_trainer_loop_contract_on_gpu():
    """What Trainer11.train_epoch / evaluate do with the model (trainer11.py:51-80,108-129), restated because the
    reference tree does not exist on the GPU box (tests/test_trainer_drive.py runs the unmodified trainer on CPU):
    iterate(data, optimizer, is_train) -> (Pack, None); every Pack value a tensor with .item(); summarize reads
    `.detach().cpu().numpy()` of every key; evaluate under no_grad with model.eval()."""
    V = 97
    torch.manual_seed(0)
    m = C5Model(V, n_layers=2, d_model=64).cuda()
    opt = torch.optim.Adam(m.parameters(), lr=2e-3, betas=(0.9, 0.98), eps=1e-9)
    data = [c5_batch(8, 60, V, 9, s, "cuda") for s in range(4)]
    m.train()
    hist = []
    for step in range(12):
        metrics, none = m.iterate(data[step % 4], optimizer=opt, is_train=True)
        assert none is None
        hist.append(metrics.loss.item())                                  # trainer11.py:73
        assert isinstance(metrics.cer.item(), float)                      # trainer11.py:74
        for k in metrics:                                                 # summarize, trainer11.py:108-112
            assert metrics[k].detach().cpu().numpy().size == 1, k
    assert sum(hist[-4:]) < sum(hist[:4])                                 # the joint objective trains
    m.eval()
    with torch.no_grad():                                                 # evaluate, trainer11.py:114-129
        for d in data:
            metrics, _ = m.iterate(d, is_train=False)
            for key, val in metrics.items():
                assert torch.is_tensor(val) and val.numel() == 1          # MetricsManager.update -> val.item()
                float(val.item())


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs (gpurun --gpus 2)")
def test_sharded_loss_under_nccl_matches_the_unsharded_op():
    """C3-style: N ranks x one shard each, sharded_ctc_loss + its single all-reduce against the un-sharded op."""
    n = min(torch.cuda.device_count(), 8)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29731",
                        os.path.join(ROOT, "tools", "check_sharded_ngpu.py")], cwd=ROOT, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("grad max abs diff") == n
