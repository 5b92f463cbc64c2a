#!/usr/bin/env python
"""Record the DRAM traffic of one kernel launch from an .ncu-rep into profiles/ncu_traffic.json, together with the
build id of the library that was profiled, so that bench.py's `roofline.traffic` can tell a current figure from a
stale one.  Run on the GPU box right after the capture (the same snapshot = the same build):

  python tools/ncu_traffic.py gpurun_out/x/ncu_bc7_bench.ncu-rep "decode_sorted_kernel<BPTC,linear>" [json path]
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    rep, key = sys.argv[1], sys.argv[2]
    path = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "profiles", "ncu_traffic.json")
    import detex_b200 as D
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, row = rows[0], rows[1], rows[2]
    d = dict(zip(hdr, row))
    u = dict(zip(hdr, units))

    def val(name):
        return float(d[name].replace(",", "")) * UNIT[u[name]]
    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    try:
        table = json.load(open(path))
    except Exception:
        table = {}
    table[key] = {"bytes": int(rd + wr), "read": int(rd), "write": int(wr), "build": D.build_id(),
                  "duration_us": float(d["gpu__time_duration.sum"].replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(u["gpu__time_duration.sum"], 1),
                  "source": os.path.basename(rep), "kernel": d.get("Kernel Name", "?")[:120]}
    table["_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the bench.py workload (ncu --set full "
                      "--clock-control none), each with the build id of the profiled library; bench.py reports a figure only "
                      "when that id equals the loaded library's (detexb200Version).")
    json.dump(table, open(path, "w"), indent=1)
    print(key, table[key])


if __name__ == "__main__":
    main()
