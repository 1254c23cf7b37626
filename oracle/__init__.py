"""CPU oracle for the phase-1 / phase-2 hot path — TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference (nimiq/snark-setup-operator) holds no golden
vectors, known-answer tests or unit tests for this path and
its arithmetic lives in un-vendored crates (`nimiq/snark-setup` rev
bd530da9804b628107e29ccd31098ed061d1cd66 over arkworks 0.4.2, Cargo.lock:150-368,
2603-2694, 3477-3500) that cannot be built here (no cargo/rustc, no network).
This package restates the *published* algorithms of those crates with Python
big integers.  What IS pinned against files inside /root/reference: the Fr
byte encoding (e2e/circuit_* fixtures, tests/golden/), the 64-byte hash-file
convention (src/utils.rs:264-276, 618-623) and the curve constants (checked
for primality / on-curve / order in tests/test_oracle_constants.py).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this package.  The product
(`snark-setup-operator_b200/`) never does.

Layout
  fields.py    Fp / Fq2 / Fq3 arithmetic on Python ints          (ark-ff 0.4.2)
  curves.py    the four pairing curves, affine SW group law      (ark-ec 0.4.2)
  serialize.py canonical point / field byte formats              (ark-serialize 0.4.2)
  chacha.py    ChaCha20 word stream, Fr::rand / G::rand          (rand_chacha 0.3.1)
  params.py    Phase1Parameters size arithmetic                  (phase1 crate)
  phase1.py    key generation, chunk contribution, chunk checks  (phase1 / setup-utils)
  phase2.py    delta^-1 scaling of the H / L queries             (phase2 crate)
  synth.py     deterministic synthetic accumulators
  c/           multi-threaded C++ restatement (64-bit limbs) used for the
               timed CPU baseline and for parity at sizes Python cannot reach
"""
