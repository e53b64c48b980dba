"""Generates tests/golden/ctc_golden.npz from torch's own CPU CTC path.

Run in the authoring container (torch 2.11.0+cu128, CPU):  python tests/golden/make_golden.py
The reference repo has no CTC code or fixtures, so the golden
vectors pin the oracle and the CUDA path against the op BASELINE.json names:
F.ctc_loss(F.log_softmax(logits,-1).transpose(0,1), ...), gradient w.r.t. logits.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.synth import make_case  # noqa: E402
from oracle.torch_ref import ref_ctc  # noqa: E402


def cases():
    g = torch.Generator().manual_seed(20261018)
    V = 11
    out = {}

    def add(name, logits, targets, il, tl):
        out[name] = dict(logits=logits.float(), targets=torch.as_tensor(targets, dtype=torch.int64),
                         input_lengths=torch.as_tensor(il, dtype=torch.int64),
                         target_lengths=torch.as_tensor(tl, dtype=torch.int64))

    # hand-built known-answer style cases
    add("kat_edge", torch.randn(8, 6, V, generator=g),
        [[0, 0, 0], [3, 0, 0], [2, 2, 0], [2, 2, 0], [4, 5, 6], [7, 7, 7], [1, 2, 1], [9, 0, 0]],
        [6, 6, 2, 3, 3, 5, 4, 1],      # [2,2]@T=2 infeasible, @T=3 single path, [4,5,6]@T=3 single path
        [0, 1, 2, 2, 3, 3, 3, 1])
    add("kat_zero_len", torch.randn(3, 4, V, generator=g), [[1, 2], [0, 0], [3, 0]], [0, 0, 4], [2, 0, 1])
    c = make_case(6, 40, 37, 9, 7, dist="D1", n_infeasible=1, n_partial=1)
    add("rand_d1", c["logits"], c["targets"], c["input_lengths"], c["target_lengths"])
    c = make_case(6, 40, 37, 9, 8, dist="D2", n_infeasible=1, n_partial=1)
    add("rand_d2", c["logits"], c["targets"], c["input_lengths"], c["target_lengths"])
    c = make_case(4, 33, 4234 // 16, 14, 9, dist="D2")     # odd T, wider vocab, more states than a warp quarter
    add("rand_wide", c["logits"], c["targets"], c["input_lengths"], c["target_lengths"])
    c = make_case(5, 30, 21, 6, 3)
    tl = c["target_lengths"]
    cat = torch.cat([c["targets"][b, : tl[b]] for b in range(5)])
    add("concat_1d", c["logits"], cat, c["input_lengths"], tl)
    return out


def main():
    blob = {}
    for name, c in cases().items():
        for k, v in c.items():
            blob[f"{name}/{k}"] = v.numpy()
        for zi in (False, True):
            nll, _ = ref_ctc(c["logits"], c["targets"], c["input_lengths"], c["target_lengths"],
                             reduction="none", zero_infinity=zi, want_grad=False)
            blob[f"{name}/zi{int(zi)}/nll"] = nll.numpy()
            for red in ("mean", "sum"):
                loss, grad = ref_ctc(c["logits"], c["targets"], c["input_lengths"], c["target_lengths"],
                                     reduction=red, zero_infinity=zi)
                blob[f"{name}/zi{int(zi)}/{red}/loss"] = loss.numpy()
                blob[f"{name}/zi{int(zi)}/{red}/grad"] = grad.numpy()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ctc_golden.npz")
    np.savez_compressed(path, torch_version=np.array(torch.__version__), **blob)
    print(path, os.path.getsize(path) // 1024, "KiB", len(blob), "arrays")


if __name__ == "__main__":
    main()
