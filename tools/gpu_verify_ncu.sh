#!/bin/bash
# ncu launch list (gpu__time_duration) of one chunk verification per curve, aggregated per kernel name
#   tools/gpu_verify_ncu.sh TAG curve...
TAG=$1; shift
mkdir -p gpurun_out
for c in "$@"; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_verify_$c.csv python tools/gpu_verify_profile.py $c > gpurun_out/${TAG}_verify_$c.out 2>&1
  python - <<PY
import csv, re, collections
rows = list(csv.reader(l for l in open("gpurun_out/${TAG}_verify_$c.csv") if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
launches = [(r[ki], float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6}.get(r[ui], 1e-6)) for r in rows[1:]]
# the profiled verification is the second one: take the second half of the launches after the setup kernels
names = [n for n, _ in launches]
def short(n): return re.sub(r"<.*", "", n.split("(")[0]).replace("sso::", "").replace("void ", "")
idx = [i for i, n in enumerate(names) if "k_same_ratio" in n]
start = idx[-2] + 1 if len(idx) >= 2 else 0
agg = collections.OrderedDict()
for n, ms in launches[start:]:
    k = short(n); a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += ms
tot = sum(a[1] for a in agg.values())
print("$c: one chunk verification, %d launches, %.1f ms of kernel time" % (sum(a[0] for a in agg.values()), tot))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("  %-40s x%-3d %9.3f ms  %5.1f %%" % (k[:40], a[0], a[1], 100 * a[1] / tot))
PY
done
