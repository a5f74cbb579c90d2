"""CPU oracle for the PASSION training hot path.  TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a plain PyTorch-CPU (fp32) / numpy restatement of
the reference's algorithm (Jun-Jie-Shi/PASSION, ``code/models/rfnet.py``,
``code/models/blocks.py``, ``code/utils/criterions.py``, ``code/train.py``).
It exists so that the CUDA product path in ``passion_b200/`` can be checked
against something that does not need ``/root/reference`` at run time.

Rules (enforced by tests/test_layout.py):
  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
    ``--impl reference`` legs may import this package;
  * nothing under ``passion_b200/`` imports it; the product path has no CPU
    fallback and raises when the CUDA library is missing.

Parity pinning: the restatement was checked here against the *real* reference
modules imported from ``/root/reference/code`` (``oracle/gen_golden.py``), and the
reference's outputs on seeded inputs are committed under ``tests/golden/``.
The reference ships no numeric golden vectors of its own for logits/losses; the
only shipped known-answer data are the missing-modality CSV tables, which
``oracle/masks.py`` reproduces byte-for-byte (tests/test_masks.py).
"""
