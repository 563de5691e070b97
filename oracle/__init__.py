"""ORACLE — test infrastructure only.

A plain-PyTorch fp32 restatement of the reference's algorithm for the pretraining hot path
(backbones + contrastive objectives), written from the reference's behaviour, each function citing
the reference file:line it follows (paths under /root/reference).

Who may import this package: ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs — as the checker or the timed CPU baseline, never as
the product. ``dualvar_b200`` itself must not import it; the product path fails loudly when the
CUDA extension is missing.

Parity pinning: the reference ships no tests, golden vectors or fixtures for this path
(SURVEY.md §4, §8c — "parity unpinned" by the reference itself). The oracle is therefore pinned
against outputs of the reference run in the build container: ``tests/golden/make_golden.py``
imports the unmodified reference modules from /root/reference (with the 4-line shim of SURVEY.md
Appendix B), runs them on seeded inputs and writes ``tests/golden/*.npz``; ``tests/test_oracle_*``
check this restatement against those files on CPU.
"""
