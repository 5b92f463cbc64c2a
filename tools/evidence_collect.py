#!/usr/bin/env python
"""Copy what tools/evidence_run.sh left under gpurun_out/final/ into profiles/ (r02_* names, ncu_traffic.json) and
summarise the launch list of the bench run.   python tools/evidence_collect.py [round_prefix]"""
import csv
import os
import re
import shutil
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out", "final")
DST = os.path.join(ROOT, "profiles")
FILES = """bench_all.json bench_bc7u.json bench_reference.json convert_bench.log ncu_bc1_1e8_bench.txt ncu_bc1_bench.txt
ncu_bc6h_bench.txt ncu_bc6h_signed.txt ncu_bc7_bench.txt ncu_bc7_tiled.txt ncu_etc2_tiled.txt ncu_mixed_bench.txt pytest_gpu.log
smoke.log sweep_all_formats.log sweep_bc7_16m.log sweep_bc7_single_modes.log sweep_etc2_33m.log version.txt sanitizer.txt""".split()


def main():
    prefix = sys.argv[1] if len(sys.argv) > 1 else "r02"
    for f in FILES:
        shutil.copy(os.path.join(SRC, f), os.path.join(DST, "%s_%s" % (prefix, f)))
    shutil.copy(os.path.join(SRC, "ncu_traffic.json"), os.path.join(DST, "ncu_traffic.json"))
    rows = list(csv.reader(l for l in open(os.path.join(SRC, "launches_bench.csv")) if l.startswith('"')))
    ix = {h: i for i, h in enumerate(rows[0])}
    agg = defaultdict(lambda: [0, 0.0])
    for x in rows[1:]:
        if x[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v, u = float(x[ix["Metric Value"]].replace(",", "")), x[ix["Metric Unit"]]
        us = v / 1000 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1000)
        name = re.sub(r"\(.*", "", x[ix["Kernel Name"]])[:70]
        agg[name][0] += 1
        agg[name][1] += us
    tot = sum(v[1] for v in agg.values())
    ver = open(os.path.join(SRC, "version.txt")).read().strip()
    with open(os.path.join(DST, "%s_launches_bench.txt" % prefix), "w") as f:
        f.write("# launch list of `python bench.py --steps 3 --warmup 3 --no-cpu` under ncu --metrics gpu__time_duration.sum --clock-control none\n")
        f.write("# (per-launch times are cold-cache and serialised: shares, not absolutes; evidence run, build %s)\n" % ver)
        f.write("%-72s %6s %12s %7s\n" % ("kernel", "count", "total us", "share"))
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-72s %6d %12.1f %6.1f%%\n" % (k, v[0], v[1], 100 * v[1] / tot))
    print("profiles/%s_* refreshed from %s (%s)" % (prefix, SRC, ver))


if __name__ == "__main__":
    main()
