#!/usr/bin/env python
"""Read the SASS-level source page of an ncu capture (`ncu -i x.ncu-rep --page source --csv`) and print where the
samples are: per opcode, and the hottest instruction windows with their stall reasons.
  python tools/ncu_source_hot.py gpurun_out/x/ncu_foo_source.csv [top_n]"""
import csv
import re
import sys
from collections import Counter


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    recs = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        recs.append(dict(src=r[ix["Source"]].strip(), samples=int(r[ix["# Samples"]] or 0), inst=int(r[ix["Instructions Executed"]] or 0),
                         exc=int(r[ix["L1 Wavefronts Shared Excessive"]] or 0),
                         stalls={c[6:]: int(r[ix[c]] or 0) for c in stall_cols}))
    tot_s = sum(x["samples"] for x in recs)
    tot_i = sum(x["inst"] for x in recs)
    print("instructions (SASS lines): %d, warp instructions executed: %d, samples: %d" % (len(recs), tot_i, tot_s))
    by_op_s, by_op_i = Counter(), Counter()
    for x in recs:
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", x["src"])
        op = m.group(2).split(".")[0] if m else "?"
        by_op_s[op] += x["samples"]
        by_op_i[op] += x["inst"]
    print("\nper opcode: %-10s %8s %7s %8s %7s" % ("", "inst", "%", "samples", "%"))
    for op, n in by_op_i.most_common(30):
        print("  %-18s %10d %6.1f%% %8d %6.1f%%" % (op, n, 100.0 * n / tot_i, by_op_s[op], 100.0 * by_op_s[op] / max(1, tot_s)))
    tot_st = Counter()
    for x in recs:
        tot_st.update(x["stalls"])
    print("\nstall samples:", ", ".join("%s %.1f%%" % (k, 100.0 * v / max(1, tot_s)) for k, v in tot_st.most_common(10)))
    print("\nhottest instructions:")
    for i in sorted(range(len(recs)), key=lambda i: -recs[i]["samples"])[:top]:
        x = recs[i]
        st = ", ".join("%s %d" % (k, v) for k, v in sorted(x["stalls"].items(), key=lambda kv: -kv[1])[:3] if v)
        print("  #%4d %5.2f%% inst %9d exc_wavefronts %8d  %-60s %s" % (i, 100.0 * x["samples"] / max(1, tot_s), x["inst"], x["exc"], x["src"][:60], st))
    # regions: split the listing at big jumps of executed-instruction counts
    print("\nregions (runs of lines with similar execution counts):")
    start = 0
    for i in range(1, len(recs) + 1):
        if i == len(recs) or (recs[i]["inst"] and recs[start]["inst"] and not (0.5 < recs[i]["inst"] / recs[start]["inst"] < 2.0)) or \
                (bool(recs[i]["inst"]) != bool(recs[start]["inst"])):
            seg = recs[start:i]
            s, n = sum(x["samples"] for x in seg), sum(x["inst"] for x in seg)
            if n and (s > 0.01 * tot_s or n > 0.01 * tot_i):
                print("  lines %4d-%4d (%4d)  exec/line %9d  inst %5.1f%%  samples %5.1f%%" %
                      (start, i - 1, i - start, seg[0]["inst"], 100.0 * n / tot_i, 100.0 * s / max(1, tot_s)))
            start = i


if __name__ == "__main__":
    main()
