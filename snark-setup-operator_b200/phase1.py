"""Host-side mirror of the reference's phase-1 interface over the C ABI.

`Phase1Parameters` follows `phase1::Phase1Parameters` as constructed at reference
src/utils.rs:326-352 (`new_chunk`, `new_full`); the functions follow the call shapes of
`phase1_cli::contribute` (src/bin/contribute.rs:809-824) with the RNG-dependent part (key
generation) factored out: scalars are passed explicitly, as SURVEY.md §8b proposes for the
`_buf` / `_dev` entry points.  Device buffers are torch uint8 CUDA tensors — torch is only the
allocator here; every computation happens inside libsso_b200.so.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass

from . import _lib
from ._lib import P1Params, SsoError, call

CURVES = {"bls12_377": 0, "bw6_761": 1, "bw6": 1, "mnt4_753": 2, "mnt6_753": 3}
CHECK_NO, CHECK_NONZERO, CHECK_FULL = 0, 1, 2
G1, G2 = 0, 1


def curve_id(curve) -> int:
    return CURVES[curve] if isinstance(curve, str) else int(curve)


def curve_sizes(curve) -> dict:
    out = (ctypes.c_uint64 * 5)()
    rc = _lib.lib().sso_curve_sizes(curve_id(curve), out)
    if rc:
        raise SsoError(rc, "unknown curve")
    return {"g1_c": out[0], "g1_u": out[1], "g2_c": out[2], "g2_u": out[3], "fr": out[4]}


def scalar_bytes(curve, k: int) -> bytes:
    return int(k).to_bytes(curve_sizes(curve)["fr"], "little")


@dataclass(frozen=True)
class Phase1Parameters:
    curve: str
    power: int
    chunk_index: int = 0
    chunk_size: int = 0
    batch_size: int = 0
    contribution_mode: int = 0       # 0 chunked, 1 full
    proving_system: int = 0          # Groth16

    @staticmethod
    def new_chunk(curve: str, chunk_index: int, chunk_size: int, power: int, batch_size: int) -> "Phase1Parameters":
        return Phase1Parameters(curve, power, chunk_index, chunk_size, batch_size, 0)

    @staticmethod
    def new_full(curve: str, power: int, batch_size: int) -> "Phase1Parameters":
        return Phase1Parameters(curve, power, 0, 0, batch_size, 1)

    def c_struct(self) -> P1Params:
        return P1Params(curve_id(self.curve), self.proving_system, self.contribution_mode, self.power, self.chunk_index,
                        self.chunk_size, self.batch_size)

    def sizes(self) -> dict:
        out = (ctypes.c_uint64 * 8)()
        call("sso_p1_sizes", ctypes.byref(self.c_struct()), out)
        keys = ("powers_length", "powers_g1_length", "g1_count", "other_count", "accumulator_size", "contribution_size",
                "public_key_size", "num_chunks")
        return dict(zip(keys, out))

    @property
    def accumulator_size(self) -> int:
        return self.sizes()["accumulator_size"]

    @property
    def contribution_size(self) -> int:
        return self.sizes()["contribution_size"]


def _dptr(t) -> int:
    if not t.is_cuda or not t.is_contiguous():
        raise ValueError("expected a contiguous CUDA tensor")
    return t.data_ptr()


def _elem_size(curve, group, compressed) -> int:
    s = curve_sizes(curve)
    return s[("g1" if group == G1 else "g2") + ("_c" if compressed else "_u")]


def batch_exp(curve, group: int, d_in, n: int, first_index: int, tau: int, coeff, d_out, in_compressed=False,
              out_compressed=True, check=CHECK_NO, device=0):
    """out[j] = (coeff * tau^(first_index + j)) * in[j] on device buffers (setup_utils::batch_exp)."""
    assert d_in.numel() == n * _elem_size(curve, group, in_compressed)
    assert d_out.numel() == n * _elem_size(curve, group, out_compressed)
    call("sso_batch_exp_dev", curve_id(curve), group, _dptr(d_in), int(in_compressed), n, first_index,
         scalar_bytes(curve, tau), None if coeff is None else scalar_bytes(curve, coeff), _dptr(d_out),
         int(out_compressed), check, device)


def batch_mul(curve, group: int, d_in, n: int, scalar: int, d_out, in_compressed=False, out_compressed=True,
              check=CHECK_NO, device=0):
    """out[j] = scalar * in[j] (phase-2 batch_mul of the H / L queries)."""
    assert d_in.numel() == n * _elem_size(curve, group, in_compressed)
    assert d_out.numel() == n * _elem_size(curve, group, out_compressed)
    call("sso_batch_mul_dev", curve_id(curve), group, _dptr(d_in), int(in_compressed), n, scalar_bytes(curve, scalar),
         _dptr(d_out), int(out_compressed), check, device)


def reencode(curve, group: int, d_in, n: int, d_out, in_compressed=True, out_compressed=False, check=CHECK_FULL,
             subgroup_check=True, device=0):
    """Decompress / recompress with the correctness checks of transform_pok_and_correctness."""
    assert d_in.numel() == n * _elem_size(curve, group, in_compressed)
    assert d_out.numel() == n * _elem_size(curve, group, out_compressed)
    call("sso_reencode_dev", curve_id(curve), group, _dptr(d_in), int(in_compressed), n, _dptr(d_out), int(out_compressed),
         check, int(subgroup_check), device)


def contribute_dev(params: Phase1Parameters, d_challenge, d_response, tau: int, alpha: int, beta: int, check=CHECK_NONZERO,
                   device=0):
    """Phase1::computation on a device-resident chunk (hash slot and public key untouched)."""
    assert d_challenge.numel() == params.accumulator_size and d_response.numel() == params.contribution_size
    c = params.curve
    call("sso_p1_contribute_dev", ctypes.byref(params.c_struct()), _dptr(d_challenge), _dptr(d_response),
         scalar_bytes(c, tau), scalar_bytes(c, alpha), scalar_bytes(c, beta), check, device)


def _host_ptr(buf):
    """bytes / bytearray / numpy / pinned CPU torch tensor -> (address, length, keepalive)"""
    if hasattr(buf, "data_ptr"):
        return buf.data_ptr(), buf.numel() * buf.element_size(), buf
    if isinstance(buf, (bytes, bytearray)):
        arr = (ctypes.c_char * len(buf)).from_buffer(buf) if isinstance(buf, bytearray) else ctypes.c_char_p(buf)
        return ctypes.cast(arr, ctypes.c_void_p).value, len(buf), arr
    import numpy as np
    a = np.ascontiguousarray(buf)
    return a.ctypes.data, a.nbytes, a


def contribute_buf(params: Phase1Parameters, challenge, response, tau: int, alpha: int, beta: int, pubkey: bytes | None = None,
                   check=CHECK_NONZERO, device=0):
    """The RNG-free core of phase1_cli::contribute on HOST buffers (H2D, compute, D2H inside).
    `response` must be a writable buffer of contribution_size bytes."""
    c = params.curve
    ch_ptr, ch_len, k1 = _host_ptr(challenge)
    rs_ptr, rs_len, k2 = _host_ptr(response)
    call("sso_p1_contribute_buf", ctypes.byref(params.c_struct()), ch_ptr, ch_len, rs_ptr, rs_len, scalar_bytes(c, tau),
         scalar_bytes(c, alpha), scalar_bytes(c, beta), pubkey, 0 if pubkey is None else len(pubkey), check, device)
    del k1, k2
    return response


def _ptr_arrays(bufs):
    """list of host buffers -> (void*[n], size_t[n], keepalives)"""
    n = len(bufs)
    keep, ptrs, lens = [], [], []
    for b in bufs:
        p, ln, k = _host_ptr(b)
        keep.append(k)
        ptrs.append(p if isinstance(p, int) else ctypes.cast(p, ctypes.c_void_p).value)
        lens.append(ln)
    return (ctypes.c_void_p * n)(*ptrs), (ctypes.c_size_t * n)(*lens), keep


def contribute_many_buf(params_list, challenges, responses, tau: int, alpha: int, beta: int, pubkey: bytes | None = None,
                        check=CHECK_NONZERO, host_threads=0, device=0):
    """contribute_buf over several chunks in flight (the reference's Process lane, src/bin/contribute.rs:64-71,158-163):
    `host_threads` workers (0 = library default, 3 per device) each run one chunk at a time, so one chunk's Blake2b
    overlaps the copies and kernels of the others.  params_list[i] / challenges[i] / responses[i] describe chunk i;
    device < 0 spreads the workers over all visible GPUs."""
    n = len(params_list)
    if not (len(challenges) == len(responses) == n):
        raise SsoError(-1, "params, challenges and responses must have the same length")
    if n == 0:
        return responses
    c = params_list[0].curve
    P = (_lib.P1Params * n)(*[p.c_struct() for p in params_list])
    ch_p, ch_l, k1 = _ptr_arrays(challenges)
    rs_p, rs_l, k2 = _ptr_arrays(responses)
    call("sso_p1_contribute_many_buf", P, n, ch_p, ch_l, rs_p, rs_l, scalar_bytes(c, tau), scalar_bytes(c, alpha),
         scalar_bytes(c, beta), pubkey, 0 if pubkey is None else len(pubkey), check, host_threads, device)
    del k1, k2
    return responses


def contribute_seeded_many_buf(params_list, challenges, responses, seed32: bytes, check=CHECK_NONZERO, host_threads=0, device=0):
    """contribute_seeded_buf (key generation, proofs of knowledge, computation) over several chunks in flight with the same
    seed-derived key, as the contributor does for the chunks it holds (src/bin/contribute.rs:789, 809-823)."""
    n = len(params_list)
    if not (len(challenges) == len(responses) == n):
        raise SsoError(-1, "params, challenges and responses must have the same length")
    if n == 0:
        return responses
    P = (_lib.P1Params * n)(*[p.c_struct() for p in params_list])
    ch_p, ch_l, k1 = _ptr_arrays(challenges)
    rs_p, rs_l, k2 = _ptr_arrays(responses)
    call("sso_p1_contribute_seeded_many_buf", P, n, ch_p, ch_l, rs_p, rs_l, seed32, check, host_threads, device)
    del k1, k2
    return responses


def verify_chunk_many_buf(params_list, challenges, responses, new_challenges, check_input=CHECK_NO, check_output=CHECK_FULL,
                          subgroup_check_mode=0, ratio_check=True, rlc_seed32=None, host_threads=0, device=0):
    """verify_chunk_buf over several chunks in flight: the chunk loop of verify_transcript (src/bin/verify_transcript.rs:293-569)
    as a work queue.  Raises SsoError for the first rejected chunk ("chunk i of the batch: ...")."""
    n = len(params_list)
    if not (len(challenges) == len(responses) == len(new_challenges) == n):
        raise SsoError(-1, "params, challenges, responses and new challenges must have the same length")
    if n == 0:
        return new_challenges
    P = (_lib.P1Params * n)(*[p.c_struct() for p in params_list])
    ch_p, ch_l, k1 = _ptr_arrays(challenges)
    rs_p, rs_l, k2 = _ptr_arrays(responses)
    nc_p, nc_l, k3 = _ptr_arrays(new_challenges)
    call("sso_p1_verify_chunk_many_buf", P, n, ch_p, ch_l, rs_p, rs_l, nc_p, nc_l, check_input, check_output, subgroup_check_mode,
         int(ratio_check), rlc_seed32, host_threads, device)
    del k1, k2, k3
    return new_challenges


def new_challenge_dev(params: Phase1Parameters, d_challenge, device=0):
    """phase1_cli::new_challenge on a device buffer (all-generator accumulator, blank hash)."""
    assert d_challenge.numel() == params.accumulator_size
    call("sso_p1_new_challenge_dev", ctypes.byref(params.c_struct()), _dptr(d_challenge), device)


PROFILE_KINDS = ("tau_tables", "batch_exp_g1", "batch_exp_g2", "normalize_g1", "normalize_g2", "reencode_g1", "reencode_g2",
                 "fill", "msm", "other", "batch_exp_chunk", "normalize_chunk")


def profile_enable(on=True):
    """True/1: time every kernel with events; 2: also serialise each call on one stream (isolated kernel times)."""
    _lib.lib().sso_profile_enable(int(on))


def profile_reset():
    _lib.lib().sso_profile_reset()


def profile_read() -> dict:
    """{kind: {"launches": n, "ms": total device milliseconds (0 unless profiling), "elems": n}}"""
    out = (ctypes.c_uint64 * (3 * 16))()
    n = _lib.lib().sso_profile_read(out, len(out))
    if n < 0:
        raise SsoError(n, "profile buffer too small")
    return {PROFILE_KINDS[i]: {"launches": out[3 * i], "ms": out[3 * i + 1] / 1e6, "elems": out[3 * i + 2]} for i in range(n)}


def imad_peak(variant: int = 0, device=0) -> float:
    """Measured multiply-accumulate peak (MAC/s) of the probe kernel; variant 0 = mad.wide.u32."""
    out = ctypes.c_double(0)
    call("sso_imad_peak", device, variant, ctypes.byref(out))
    return out.value


def power_pairs(curve, group: int, d_in, n: int, in_compressed=False, check=CHECK_NO, subgroup_check=False, seed32=None,
                device=0) -> bytes:
    """setup_utils::power_pairs on a device vector -> the two result points, uncompressed."""
    usz = _elem_size(curve, group, False)
    out = ctypes.create_string_buffer(2 * usz)
    call("sso_power_pairs_dev", curve_id(curve), group, _dptr(d_in), int(in_compressed), n, check, int(subgroup_check), seed32,
         out, len(out), device)
    return out.raw


def merge_pairs(curve, group: int, d_a, d_b, n: int, in_compressed=False, check=CHECK_NO, subgroup_check=False, seed32=None,
                device=0) -> bytes:
    """setup_utils::merge_pairs on two device vectors -> the two result points, uncompressed."""
    usz = _elem_size(curve, group, False)
    out = ctypes.create_string_buffer(2 * usz)
    call("sso_merge_pairs_dev", curve_id(curve), group, _dptr(d_a), _dptr(d_b), int(in_compressed), n, check,
         int(subgroup_check), seed32, out, len(out), device)
    return out.raw


def same_ratio(curve, checks, device=0):
    """Batch of setup_utils::same_ratio checks.  `checks`: list of (a, b, c, d) uncompressed point byte
    strings (a, b in G1; c, d in G2).  Returns a list of booleans e(a, d) == e(b, c)."""
    buf = b"".join(a + b + c + d for a, b, c, d in checks)
    verdicts = (ctypes.c_uint32 * max(1, len(checks)))()
    call("sso_same_ratio", curve_id(curve), buf, len(checks), verdicts, device)
    return [bool(v) for v in verdicts[:len(checks)]]


SUBGROUP_AUTO, SUBGROUP_DIRECT, SUBGROUP_BATCHED, SUBGROUP_NO = 0, 1, 2, 3


def keygen(curve, seed32: bytes, digest64: bytes, device=0):
    """Phase1::key_generation -> ((tau, alpha, beta), serialized public key)."""
    s = curve_sizes(curve)
    sc = ctypes.create_string_buffer(3 * s["fr"])
    pk = ctypes.create_string_buffer(6 * s["g1_u"] + 3 * s["g2_u"])
    call("sso_p1_keygen", curve_id(curve), seed32, digest64, sc, len(sc), pk, len(pk), device)
    fr = s["fr"]
    return tuple(int.from_bytes(sc.raw[i * fr:(i + 1) * fr], "little") for i in range(3)), pk.raw


def contribute_seeded_buf(params: Phase1Parameters, challenge, response, seed32: bytes, check=CHECK_NONZERO, device=0):
    """phase1_cli::contribute on host buffers with the RNG given by its 32-byte seed."""
    ch_ptr, ch_len, k1 = _host_ptr(challenge)
    rs_ptr, rs_len, k2 = _host_ptr(response)
    call("sso_p1_contribute_seeded_buf", ctypes.byref(params.c_struct()), ch_ptr, ch_len, rs_ptr, rs_len, seed32, check, device)
    del k1, k2
    return response


def contribute(challenge_filename, challenge_hash_filename, response_filename, response_hash_filename, check_input, batch_exp_mode,
               parameters: Phase1Parameters, seed32: bytes, device=0):
    """phase1_cli::contribute, argument for argument (reference src/bin/contribute.rs:811-823); rng -> its seed."""
    call("sso_p1_contribute_file", ctypes.byref(parameters.c_struct()), challenge_filename.encode(), challenge_hash_filename.encode(),
         response_filename.encode(), response_hash_filename.encode(), check_input, batch_exp_mode, seed32, device)


def verify_chunk_buf(params: Phase1Parameters, challenge, response, new_challenge, check_input=CHECK_NO, check_output=CHECK_FULL,
                     subgroup_check_mode=SUBGROUP_AUTO, ratio_check=True, rlc_seed32=None, device=0):
    """Phase1::verification of one chunk on host buffers; raises SsoError(code -4) on a rejected contribution."""
    ch_ptr, ch_len, k1 = _host_ptr(challenge)
    rs_ptr, rs_len, k2 = _host_ptr(response)
    nc_ptr, nc_len, k3 = _host_ptr(new_challenge)
    call("sso_p1_verify_chunk_buf", ctypes.byref(params.c_struct()), ch_ptr, ch_len, rs_ptr, rs_len, nc_ptr, nc_len, check_input,
         check_output, subgroup_check_mode, int(ratio_check), rlc_seed32, device)
    del k1, k2, k3
    return new_challenge


def transform_pok_and_correctness(challenge_filename, challenge_hash_filename, check_input, response_filename, response_hash_filename,
                                  check_output, new_challenge_filename, new_challenge_hash_filename, subgroup_check_mode, ratio_check,
                                  parameters: Phase1Parameters, device=0):
    """phase1_cli::transform_pok_and_correctness, argument for argument (reference src/bin/contribute.rs:968-986)."""
    call("sso_p1_verify_chunk_file", ctypes.byref(parameters.c_struct()), challenge_filename.encode(), challenge_hash_filename.encode(),
         check_input, response_filename.encode(), response_hash_filename.encode(), check_output, new_challenge_filename.encode(),
         new_challenge_hash_filename.encode(), subgroup_check_mode, int(ratio_check), device)


# ---- file-level calls on whole accumulators (verify_transcript / control / new_setup) ---------------------------------

def _devices_arg(devices):
    if not devices:
        return None, 0
    arr = (ctypes.c_int * len(devices))(*devices)
    return arr, len(devices)


def new_challenge(challenge_filename, challenge_hash_filename, parameters: Phase1Parameters, device=0):
    """phase1_cli::new_challenge, argument for argument (reference src/bin/new_setup.rs:105-109)."""
    call("sso_p1_new_challenge_file", challenge_filename.encode(), challenge_hash_filename.encode(), ctypes.byref(parameters.c_struct()),
         device)


def set_generators(curve, g1_uncompressed: bytes | None, g2_uncompressed: bytes | None, device=0):
    """Hand the reference's G1 / G2 generators over (needed for MNT4-753 / MNT6-753 G2, see include/sso_b200.h); None, None resets."""
    call("sso_p1_set_generators", curve_id(curve), g1_uncompressed, 0 if g1_uncompressed is None else len(g1_uncompressed),
         g2_uncompressed, 0 if g2_uncompressed is None else len(g2_uncompressed), device)


def combine(response_list_filename, combined_filename, parameters: Phase1Parameters, devices=None, device=0):
    """phase1_cli::combine, argument for argument (reference src/bin/verify_transcript.rs:603-607); `parameters` = the chunk-0
    parameters of the ceremony.  devices: GPUs of this process to decode on (default: `device`)."""
    arr, n = _devices_arg(devices)
    call("sso_p1_combine_file", response_list_filename.encode(), combined_filename.encode(), ctypes.byref(parameters.c_struct()), arr, n,
         device)


def transform_ratios(response_filename, check_input, parameters: Phase1Parameters, devices=None, device=0, rlc_seed32=None):
    """phase1_cli::transform_ratios, argument for argument (reference src/bin/verify_transcript.rs:646-653, 811-822); raises
    SsoError(code -4) on rejection.  The partial MSM results of the devices (or of the ranks of the process group set up with
    dist_init) meet in one NCCL all-gather inside the library."""
    arr, n = _devices_arg(devices)
    call("sso_p1_verify_ratios_file", ctypes.byref(parameters.c_struct()), response_filename.encode(), check_input, arr, n, device,
         rlc_seed32)


def dist_init_from_torch(device: int):
    """One process per GPU under torchrun: build the library's NCCL communicator for the cooperative calls.  Rank 0 draws
    the NCCL id, torch.distributed (any backend) broadcasts it."""
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    box = [None]
    if rank == 0:
        buf = ctypes.create_string_buffer(128)
        call("sso_dist_unique_id", buf)
        box[0] = buf.raw
    dist.broadcast_object_list(box, src=0)
    call("sso_dist_init", rank, world, box[0], device)


def dist_barrier():
    call("sso_dist_barrier")


def dist_finalize():
    _lib.lib().sso_dist_finalize()


def dist_stats() -> dict:
    out = (ctypes.c_uint64 * 5)()
    _lib.lib().sso_dist_stats(out)
    return {"initialised": bool(out[0]), "rank": out[1], "world": out[2], "all_gathers": out[3], "nccl_version": out[4]}
