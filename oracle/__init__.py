"""CPU oracle for the CTC loss+grad hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is on the product path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import it, and there only as the checker (or the timed CPU baseline),
never as the thing shipped.  The product (``asr_chinese_e2e_b200``) raises if its
CUDA library is missing; it never routes through this package.

Parity status: the reference repository (zqs01/ASR_chinese_e2e) ships no CTC code
and no tests, so the reference itself pins
nothing: **parity unpinned by the reference**.  The operative oracle is
``torch.nn.functional.ctc_loss`` (torch 2.11.0, the third-party op BASELINE.json
names).  ``oracle.ctc_f64`` is an independent float64 restatement of the
published algorithm (Graves et al. 2006) and is pinned
against outputs of torch's op generated in the authoring container
(``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``).
"""
