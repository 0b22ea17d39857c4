"""CPU oracle for the openzl_b200 hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (Python big-int + plain C under ``oracle/c``) of the
arkworks 0.3.0 algorithms that ``plugins/arkworks`` forwards to
(``/root/reference/plugins/arkworks/src/groth16.rs:445-457`` ->
``ark_groth16::Groth16::prove`` -> ``ark_ec::msm::VariableBaseMSM`` /
``ark_poly::Radix2EvaluationDomain``).  Those crates are crates.io dependencies pinned at
``^0.3.0`` (``/root/reference/plugins/arkworks/Cargo.toml:113-146``) and are NOT vendored in
``/root/reference``; no Rust toolchain exists in the build image, so the reference binary
cannot be built (``oracle/_ref`` therefore does not exist).

PARITY STATUS
-------------
* Field arithmetic (BLS12-381 Fr add/mul/pow/inverse) is PINNED against the reference's own
  golden vectors: the Poseidon width-3 KAT (``openzl-tutorials/src/poseidon.rs:388-401``),
  the Cauchy MDS matrices widths 2..12
  (``plugins/arkworks/src/poseidon/mds_hardcoded_tests``) and the 189 Grain-LFSR round
  constants (``plugins/arkworks/src/poseidon/parameters_hardcoded_test/lfsr_values``) --
  see ``tests/golden/`` and ``tests/test_oracle_golden.py``.
* MSM and NTT results: **parity unpinned** by the reference (no test in the reference ever
  calls ``multi_scalar_mul``, an FFT, or ``Groth16::prove``).  They are pinned by mathematics
  only: MSM == naive sum of double-and-add, MSM over bases [d_i]G == [sum s_i d_i]G,
  NTT == O(n^2) DFT, inverse(forward) == identity.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference legs
may import this package.  The product (``openzl_b200``) never does.
"""
