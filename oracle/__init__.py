"""CPU oracle for the acoustic-model hot path -- TEST INFRASTRUCTURE ONLY.

This package is a from-scratch CPU (torch fp32 / numpy fp64) restatement of the
reference algorithm of boji123/pytorch-kaldi-asr for the path named in
BASELINE.json `north_star`.  It is the *checker*:

  * only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
    `--impl reference` legs may import it;
  * the product package (`pytorch-kaldi-asr_b200/`) never imports it and has no
    CPU fallback -- it fails loudly when the CUDA library is missing.

Pinning: the restatement is checked (tests/test_oracle_vs_golden.py) against
golden vectors produced by importing the *real* reference from /root/reference
(tests/golden/make_golden.py, committed together with the fixtures), against the
one known-answer vector the reference ships (Lattice demo, T/Lattice.py:109-130),
and -- when /root/reference is present -- live against the reference modules
(tests/test_oracle_vs_reference_live.py).  The CMVN stage is the exception: its
arithmetic lives in the external Kaldi binary `apply-cmvn`, no fixture of the
reference pins it, so `oracle.cmvn` is "parity unpinned" (see its header).

Path abbreviations used in citations (relative to /root/reference/):
  T/ = project/attention-transformer-timit/local/pytorch/transformer/
  L/ = project/attention-transformer-timit/local/
  U/ = pytorch/utils/
"""
