"""bench.py contract checks that need no GPU: the reference arm (CPU restatement on the host cores) prints one JSON
line with the keys the driver reads, and the declared work counts are consistent."""
import json
import subprocess
import sys

from conftest import ROOT

import bench


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0", "--curve", "bls12_377",
                          "--chunk-log", "9"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["value"] > 0 and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


def test_declared_work_counts():
    # the kernels do fewer field multiplications than the reference's double-and-add: b*D + (b/2)*11
    ref = {"bls12_377": 253 * 7 + 126.5 * 11, "bw6_761": 377 * 7 + 188.5 * 11, "mnt4_753": 753 * 10 + 376.5 * 11}
    for curve, ref_fm in ref.items():
        ours = bench.declared_fq_muls_per_point(curve, 0)
        assert 0.4 * ref_fm < ours < ref_fm
    assert bench.macs_per_fq_mul(12) == 300 and bench.macs_per_fq_mul(24) == 1176 and bench.macs_per_fq_mul(8) == 136
