"""CPU oracle for the SMG grasp-affordance hot path.  TEST INFRASTRUCTURE ONLY.

This package is a plain torch/numpy CPU restatement of what the reference
(fukangl/SMG-multimodal-grasping, `code/models.py`, `code/trainer.py`,
`code/utils.py`, `code/NMS.py`, `code/main.py:137-233`) computes on the hot
path.  Every function cites the reference file:line it follows.

Rules (enforced by tests/test_layout.py):
  * only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
    `--impl reference` legs may import anything from here;
  * the product package (`smg-multimodal-grasping_b200/`, alias `smg_b200`)
    never imports it, and has no CPU fallback: it raises if the CUDA library
    is missing.

Parity pinning: the reference ships no tests, golden vectors or fixtures
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference
itself, produced in the build container by `tests/golden/make_golden.py`
(which imports the unmodified reference modules through `oracle/refshim.py`)
and committed as small fixtures under `tests/golden/`.
`tests/test_oracle_golden.py` replays them.
"""
