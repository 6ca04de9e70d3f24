#!/usr/bin/env python3
"""Summarise an `ncu --page source --csv` export: instructions executed and stall samples per
barrier-delimited SASS segment, plus runs of equal execution count inside the hottest segment.
    ncu -i prof.ncu-rep --page source --csv > src.csv ; python tools/ncu_segments.py src.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
ia, isamp, isrc = col["Instructions Executed"], col["# Samples"], col["Source"]
stalls = [c for c in ("stall_barrier", "stall_long_sb", "stall_math", "stall_wait", "stall_short_sb", "stall_not_selected",
                      "stall_lg", "stall_mio", "stall_branch_resolving", "stall_no_inst") if c in col]
tot = sum(int(r[ia]) for r in data)
ts = sum(int(r[isamp]) for r in data)
print("kernel:", rows[0][1][:90])
print("total warp instructions %d, samples %d, SASS lines %d" % (tot, ts, len(data)))
seg, cur = [], None
for k, r in enumerate(data):
    if cur is None:
        cur = {"start": k, "inst": 0, "samp": 0, "n": 0, **{s: 0 for s in stalls}}
    cur["inst"] += int(r[ia]); cur["samp"] += int(r[isamp]); cur["n"] += 1
    for s in stalls:
        cur[s] += int(r[col[s]])
    if "BAR.SYNC" in r[isrc] or "EXIT" in r[isrc]:
        cur["end"] = k
        seg.append(cur)
        cur = None
for s in seg:
    if s["inst"] > tot * 0.005 or s["samp"] > ts * 0.005:
        print("sass[%d..%d] n=%d inst=%.1f%% samples=%.1f%% | " % (s["start"], s["end"], s["n"], 100 * s["inst"] / tot, 100 * s["samp"] / ts)
              + " ".join("%s=%.1f" % (k.replace("stall_", ""), 100 * s[k] / ts) for k in stalls if s[k] > ts * 0.003))
if len(sys.argv) > 3:
    lo, hi = int(sys.argv[2]), int(sys.argv[3])
    prev, start, acc, sm = None, lo, 0, 0
    for k in range(lo, hi + 1):
        e = data[k][ia] if k < hi else None
        if e != prev:
            if prev is not None:
                print("  sass[%d..%d] n=%d exec=%s total=%.1fM samples=%d : %s" % (start, k - 1, k - start, prev, acc / 1e6, sm, data[start][isrc].strip()[:70]))
            prev, start, acc, sm = e, k, 0, 0
        if k < hi:
            acc += int(data[k][ia]); sm += int(data[k][isamp])
