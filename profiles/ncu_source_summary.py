"""Summarise an `ncu --page source --csv` dump: executed warp-instructions by opcode and
warp-stall samples by reason, per kernel.   usage: ncu -i X.ncu-rep --page source --csv | python ncu_source_summary.py [n_perm]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(sys.stdin))
n_perm = int(sys.argv[1]) if len(sys.argv) > 1 else 0
blocks = []
cur = None
for r in rows:
    if len(r) >= 2 and r[0] == "Kernel Name":
        cur = dict(name=r[1], hdr=None, rows=[])
        blocks.append(cur)
    elif cur is not None and len(r) > 5 and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None and len(r) >= len(cur["hdr"]):
        cur["rows"].append(r)
for b in blocks:
    h = b["hdr"]
    ix = {n: i for i, n in enumerate(h)}
    ops, stalls = collections.Counter(), collections.Counter()
    tot = 0
    for r in b["rows"]:
        src = r[ix["Source"]].strip()
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
        op = m.group(2).split(".")[0] if m else "?"
        n = int(r[ix["Instructions Executed"]])
        ops[op] += n
        tot += n
        for c in h:
            if c.startswith("stall_"):
                stalls[c] += int(r[ix[c]] or 0)
    print("kernel:", b["name"][:90])
    print("  warp instructions executed:", tot, (f"= {tot * 32 / n_perm:.1f} thread-instr per Keccak-f" if n_perm else ""))
    for k, v in ops.most_common(12):
        print(f"    {k:10s} {v:12d} {100 * v / tot:6.2f}%")
    s = sum(stalls.values())
    print("  warp stall samples:", s)
    for k, v in stalls.most_common(8):
        print(f"    {k:24s} {v:8d} {100 * v / max(1, s):6.1f}%")
