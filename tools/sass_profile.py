"""Instruction-mix / stall summary of one kernel from `ncu --page source --csv --print-source sass` output.
usage: sass_profile.py file.csv [top_n]"""
import csv
import collections
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
# the export holds one block per kernel ("Kernel Name" row, header row, instruction rows): pick block sys.argv[3]
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
lo = starts[which]
hi = starts[which + 1] if which + 1 < len(starts) else len(rows)
print("kernel:", rows[lo][1][:110], f"(block {which} of {len(starts)})")
rows = rows[lo:hi]
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
tot_inst = 0
by_op = collections.Counter()
samples_by_op = collections.Counter()
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
stall_tot = collections.Counter()
recs = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    src = r[ix["Source"]].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = m.group(2) if m else src.split()[0]
    base = op.split(".")[0]
    n = int(r[ix["Instructions Executed"]] or 0)
    s = int(r[ix["# Samples"]] or 0)
    tot_inst += n
    by_op[base] += n
    samples_by_op[base] += s
    for c in stall_cols:
        stall_tot[c] += int(r[ix[c]] or 0)
    recs.append((n, s, src))
print("total warp instructions", tot_inst)
for op, n in by_op.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 40):
    print(f"  {op:12s} {n:12d} {100.0 * n / tot_inst:6.2f}%   samples {samples_by_op[op]}")
tot_s = sum(stall_tot.values())
print("stall samples:", {k: round(100.0 * v / tot_s, 1) for k, v in stall_tot.most_common(12)})
