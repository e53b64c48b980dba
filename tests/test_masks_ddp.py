"""CPU: the vectorised mask utilities against a literal restatement of the reference's loops (and against the
reference's own functions when /root/reference is present), and the one-process-per-GPU wrapper on gloo,
world_size 2."""
import importlib.util
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from asr_chinese_e2e_b200 import masks


# ---- literal restatement of Predictor/Models/utils.py:99-145 (host loops and all) ----
def ref_non_pad(padded_input, input_lengths=None, pad_idx=None):
    if input_lengths is not None:
        m = padded_input.new_ones(padded_input.size()[:-1])
        for i in range(padded_input.size(0)):
            m[i, input_lengths[i]:] = 0
    if pad_idx is not None:
        m = padded_input.ne(pad_idx).float()
    return m.unsqueeze(-1)


def ref_subsequent(seq):
    b, l = seq.size()
    return torch.triu(torch.ones((l, l), dtype=torch.uint8), diagonal=1).unsqueeze(0).expand(b, -1, -1)


def ref_key_pad(seq_k, seq_q, pad_idx):
    return seq_k.eq(pad_idx).unsqueeze(1).expand(-1, seq_q.size(1), -1)


def ref_attn_pad(padded_input, input_lengths, expand_length):
    return ref_non_pad(padded_input, input_lengths=input_lengths).squeeze(-1).lt(1).unsqueeze(1).expand(-1, expand_length, -1)


def _reference_module():
    path = "/root/reference/Predictor/Models/utils.py"
    if not os.path.exists(path):
        return None
    src = open(path).read()
    start = src.index("# -- Transformer Related --")           # the mask helpers only (the file's top imports kaldi_io)
    mod = type(masks)("ref_masks")
    exec(compile(src[start:], path, "exec"), mod.__dict__)
    return mod


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_masks_equal_the_reference_loops(dtype):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(5, 9, 4, generator=g).to(dtype)
    for lens in ([9, 1, 0, 5, 9], torch.tensor([3, 9, 9, 2, 7]), torch.tensor([9, 0, -2, 20, 4], dtype=torch.int32)):
        a, b = masks.get_non_pad_mask(x, input_lengths=lens), ref_non_pad(x, input_lengths=lens)
        assert a.dtype == b.dtype and a.shape == b.shape and torch.equal(a, b)
        a, b = masks.get_attn_pad_mask(x, lens, 6), ref_attn_pad(x, lens, 6)
        assert a.dtype == b.dtype and a.shape == b.shape and torch.equal(a, b)
    ids = torch.randint(0, 4, (5, 7), generator=g)
    a, b = masks.get_non_pad_mask(ids, pad_idx=0), ref_non_pad(ids, pad_idx=0)
    assert a.dtype == b.dtype and torch.equal(a, b)
    a, b = masks.get_subsequent_mask(ids), ref_subsequent(ids)
    assert a.dtype == b.dtype and a.shape == b.shape and torch.equal(a, b)
    a, b = masks.get_attn_key_pad_mask(ids, ids[:, :3], 0), ref_key_pad(ids, ids[:, :3], 0)
    assert a.dtype == b.dtype and a.shape == b.shape and torch.equal(a, b)
    ref = _reference_module()
    if ref is not None:                                          # the real functions, when the reference is mounted
        lens = torch.tensor([3, 9, 9, 2, 7])
        assert torch.equal(masks.get_non_pad_mask(x, input_lengths=lens), ref.get_non_pad_mask(x, input_lengths=lens))
        assert torch.equal(masks.get_attn_pad_mask(x, lens, 4), ref.get_attn_pad_mask(x, lens, 4))
        assert torch.equal(masks.get_subsequent_mask(ids), ref.get_subsequent_mask(ids))
        assert torch.equal(masks.get_attn_key_pad_mask(ids, ids, 0), ref.get_attn_key_pad_mask(ids, ids, 0))
        assert torch.equal(masks.get_non_pad_mask(ids, pad_idx=0), ref.get_non_pad_mask(ids, pad_idx=0))


class _Toy(torch.nn.Module):
    """Reference-shaped model: forward + iterate(input, optimizer, is_train) and an extra attribute."""
    tag = "toy"

    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.lin = torch.nn.Linear(4, 3)

    def forward(self, x):
        return self.lin(x)

    def iterate(self, x, y):
        out = self(x)                                            # through DDP when wrapped
        loss = ((out - y) ** 2).sum() / x.shape[0]              # LOCAL mean: DDP averages the gradients over ranks
        loss.backward()
        return loss.detach()


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from asr_chinese_e2e_b200.ddp import DistributedWrapper, init_from_env, shard_batch
    r, w, dev = init_from_env("gloo")
    assert (r, w) == (rank, world)
    model = DistributedWrapper(_Toy(), dev)
    assert model.tag == "toy" and isinstance(model.module, _Toy)          # attribute pass-through
    g = torch.Generator().manual_seed(1)
    from asr_chinese_e2e_b200 import Pack
    batch = Pack(x=torch.randn(6, 4, generator=g), y=torch.randn(6, 3, generator=g), note="kept")
    mine = shard_batch(batch, rank, world)
    assert isinstance(mine, Pack) and mine.x.shape[0] == 3 and mine.note == "kept"      # container type preserved
    try:
        shard_batch({"x": torch.zeros(5, 1)}, rank, world)
        raise AssertionError("a ragged batch must not be dropped silently")
    except ValueError:
        pass
    model.iterate(mine.x, mine.y)
    # DDP averages gradients over ranks: with every rank's loss normalised by its LOCAL batch the averaged gradient
    # IS the full-batch-mean gradient -- no compensation factor
    q.put((rank, "toy", model.module.lin.weight.grad.tolist()))     # plain lists: no fd passing after exit

    # the documented recipe end to end: DistributedWrapper + sharded_ctc_loss(grad_reduce='mean') must give the
    # parameter gradient of the un-sharded global-mean CTC loss (the oracle stands in for the CUDA op on CPU)
    import asr_chinese_e2e_b200.ctc as ctc_mod
    from asr_chinese_e2e_b200.sharded import sharded_ctc_loss
    from test_host_logic import oracle_ctc
    ctc_mod.ctc_loss_b200 = oracle_ctc
    head = DistributedWrapper(_Head(), dev)
    c = _ctc_case()
    sl = slice(rank * 3, rank * 3 + 3)
    for mode in ("mean", "sum"):
        head.module.zero_grad()
        loss = sharded_ctc_loss(head(c["enc"][sl]), c["targets"][sl], c["input_lengths"][sl], c["target_lengths"][sl],
                                grad_reduce=mode)
        loss.backward()
        q.put((rank, mode, (loss.item(), head.module.lin.weight.grad.tolist())))
    dist.barrier()
    dist.destroy_process_group()


class _Head(torch.nn.Module):
    def __init__(self):
        super().__init__()
        torch.manual_seed(3)
        self.lin = torch.nn.Linear(5, 11)

    def forward(self, enc):
        return self.lin(enc)


def _ctc_case():
    g = torch.Generator().manual_seed(7)
    tl = torch.tensor([3, 2, 4, 1, 3, 2])
    return dict(enc=torch.randn(6, 9, 5, generator=g), input_lengths=torch.tensor([9, 8, 9, 5, 7, 9]), target_lengths=tl,
                targets=torch.randint(1, 11, (6, 4), generator=g) * (torch.arange(4)[None] < tl[:, None]))


def test_distributed_wrapper_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=180) for _ in range(6)]
    [p.join(60) for p in procs]
    ref = _Toy()
    g = torch.Generator().manual_seed(1)
    x, y = torch.randn(6, 4, generator=g), torch.randn(6, 3, generator=g)
    (((ref(x) - y) ** 2).sum() / 6).backward()
    for _, _, grad in [r for r in res if r[1] == "toy"]:
        assert torch.allclose(torch.tensor(grad), ref.lin.weight.grad, atol=1e-6)
    import torch.nn.functional as F
    head, c = _Head(), _ctc_case()
    full = F.ctc_loss(F.log_softmax(head(c["enc"]), -1).transpose(0, 1), c["targets"], c["input_lengths"],
                      c["target_lengths"], reduction="mean")
    full.backward()
    for _, mode, (loss, grad) in [r for r in res if r[1] in ("mean", "sum")]:
        assert abs(loss - full.item()) < 1e-5 * abs(full.item()), mode               # global value on every rank
        # DDP averaged the gradients: 'mean' reproduces the global-mean gradient, 'sum' is world times too small
        scale = 1.0 if mode == "mean" else 2.0
        assert torch.allclose(torch.tensor(grad) * scale, head.lin.weight.grad, atol=1e-6), mode


def test_wrapper_without_process_group_is_transparent():
    from asr_chinese_e2e_b200.ddp import DistributedWrapper
    m = DistributedWrapper(_Toy())
    assert m.module is m.model and m.tag == "toy"
    out = m(torch.zeros(2, 4))
    assert out.shape == (2, 3)
