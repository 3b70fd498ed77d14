"""oracle/ — TEST INFRASTRUCTURE ONLY.

CPU restatement (torch fp32 / numpy) of the SOME inference hot path
(/root/reference: modules/rmvpe/spec.py, modules/conform/Gconform.py,
modules/attention/base_attention.py, modules/conv/base_conv.py,
modules/model/Gmidi_conform.py, utils/infer_utils.py, inference/me_infer.py,
inference/me_quant_infer.py), plus a shim (refshim.py) that imports the
UNMODIFIED reference in the build container to pin the restatement.

Parity status: PINNED against the reference itself, run in the build container
through refshim.py (the reference ships no tests / golden vectors of its own —
SURVEY.md §4).  `tests/golden/make_golden.py` executes the real reference and
commits its outputs under tests/golden/*.npz; tests/test_oracle_golden.py checks
this restatement against those files on every CPU test run.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  The product (some_b200/, inference/)
never does: it fails loudly when the CUDA library is missing.
"""
