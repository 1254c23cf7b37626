"""The exchange step on real hardware (needs two GPUs; skipped on a one-GPU box): (a) one process driving two devices —
sso_p1_combine_file / sso_p1_verify_ratios_file with devices = [0, 1], ncclCommInitAll inside; (b) one process per GPU with
a process group (sso_dist_init over an id broadcast by torch.distributed) making the cooperative Full-mode calls on shared
files.  Outputs are compared with the oracle exactly as in tests/test_gpu_transcript.py; the all-gather count is asserted."""
import os

import pytest

import snark_setup_operator_b200 as sso

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _two_gpus():
    return torch.cuda.is_available() and torch.cuda.device_count() >= 2


@pytest.mark.skipif(not _two_gpus(), reason="needs two GPUs")
def test_transcript_over_two_devices_one_process(tmp_path):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_gpu_transcript import _run_transcript
    _run_transcript(tmp_path, "bls12_377", 10, 8, 64, devices=[0, 1])


def _rank_main(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import faulthandler
    import hashlib
    import sys
    import torch.distributed as dist
    faulthandler.dump_traceback_later(150, exit=True, file=sys.stderr)       # a hang must not hold the two GPUs
    from oracle import cport, phase1, synth
    from oracle.chacha import ChaChaRng
    from oracle.params import Phase1Params
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    sso.dist_init_from_torch(rank)
    name, power, cs, batch = "bls12_377", 8, 64, 32
    f = lambda n: os.path.join(tmp, n)
    p0 = sso.Phase1Parameters.new_chunk(name, 0, cs, power, batch)
    o0 = Phase1Params.new_chunk(name, 0, cs, power, batch)
    resps = []
    # chunk-level calls stay local: rank r contributes to chunks r, r + world, ...
    for k in range(o0.num_chunks):
        if k % world == rank:
            p = sso.Phase1Parameters.new_chunk(name, k, cs, power, batch)
            sso.new_challenge(f("ch%d" % k), f("ch%d.hash" % k), p, device=rank)
            sso.contribute(f("ch%d" % k), f("ch%d.h2" % k), f("resp%d" % k), f("resp%d.hash" % k), sso.CHECK_NONZERO, 0, p, synth.SEED_CONTRIB, device=rank)
    dist.barrier()
    if rank == 0:
        open(f("list"), "w").write("\n".join(f("resp%d" % k) for k in range(o0.num_chunks)))
    dist.barrier()
    pf = sso.Phase1Parameters.new_full(name, power, batch)
    of = Phase1Params.new_full(name, power, batch)
    beacon = hashlib.blake2s(b"beacon").digest()
    # cooperative calls: every rank makes the same call on the same files
    sso.combine(f("list"), f("combined"), p0)
    sso.contribute(f("combined"), f("combined.hash"), f("beacon"), f("beacon.hash"), sso.CHECK_NONZERO, 0, pf, beacon)
    sso.transform_pok_and_correctness(f("combined"), f("c.vhash"), sso.CHECK_NO, f("beacon"), f("b.vhash"), sso.CHECK_NONZERO, f("final"),
                                      f("final.hash"), 0, True, pf)
    before = sso.dist_stats()["all_gathers"]
    sso.transform_ratios(f("final"), sso.CHECK_NO, pf)
    stats = sso.dist_stats()
    assert stats["all_gathers"] == before + 1 and stats["world"] == world      # ONE all-gather per transform_ratios
    if rank == 0:
        resps = [open(f("resp%d" % k), "rb").read() for k in range(o0.num_chunks)]
        combined = open(f("combined"), "rb").read()
        assert combined == cport.combine(o0, resps)
        digest = phase1.calculate_hash(combined)
        pub, key = phase1.key_generation(of.curve, ChaChaRng(beacon), digest)
        want = cport.contribute_with_key(of, combined, key, pub.to_bytes(of.curve))
        got = open(f("beacon"), "rb").read()
        assert got == want
        assert open(f("final"), "rb").read() == cport.decompress_response(of, got, check=0, subgroup=False)
        # a tampered accumulator is rejected on every rank (the verdict is agreed through the group)
        import numpy as np
        mm = np.memmap(f("final"), dtype=np.uint8, mode="r+")
        mm[64 + 200 * 96 + 3] ^= 1
        mm.flush(); del mm
    dist.barrier()
    try:
        sso.transform_ratios(f("final"), sso.CHECK_NO, pf)
        raise AssertionError("tampered accumulator accepted on rank %d" % rank)
    except sso.SsoError as e:
        assert e.code in (-3, -4)
    sso.dist_finalize()
    dist.destroy_process_group()
    faulthandler.cancel_dump_traceback_later()


@pytest.mark.skipif(not _two_gpus(), reason="needs two GPUs")
def test_cooperative_calls_one_process_per_gpu(tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_rank_main, args=(2, 29631, str(tmp_path)), nprocs=2, join=True)
