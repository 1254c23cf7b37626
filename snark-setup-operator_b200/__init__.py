"""B200-native compute core for the phase-1 / phase-2 hot path of nimiq/snark-setup-operator.

The product is `libsso_b200.so` (CUDA kernels for sm_100a behind the C ABI of
include/sso_b200.h).  This Python package is the thin host-side binding used by the tests
and the benchmark: it mirrors the reference's interface for the path — `Phase1Parameters`
(reference src/utils.rs:326-352), `contribute`, `transform_pok_and_correctness`
(src/bin/contribute.rs:809-824, 966-987) — and fails loudly when the CUDA library is
missing.  There is no CPU fallback.
"""
from ._lib import SsoError, lib, library_path  # noqa: F401
from .phase1 import (  # noqa: F401
    CHECK_FULL, CHECK_NO, CHECK_NONZERO, CURVES, Phase1Parameters, batch_exp, batch_mul, contribute_buf, contribute_many_buf, contribute_seeded_many_buf, verify_chunk_many_buf,
    contribute, contribute_dev, contribute_seeded_buf, keygen, transform_pok_and_correctness, verify_chunk_buf, imad_peak, merge_pairs, new_challenge_dev, power_pairs, profile_enable, profile_read, profile_reset, reencode, same_ratio,
    new_challenge, set_generators, combine, transform_ratios, dist_init_from_torch, dist_barrier, dist_finalize, dist_stats,
)
