"""Top stall lines of one kernel from `ncu --page source --csv` output (stdin): SASS, samples, dominant stall reason."""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
which = int(sys.argv[1]) if len(sys.argv) > 1 else 0
i0 = hdr_idx[which]
i1 = hdr_idx[which + 1] if which + 1 < len(hdr_idx) else len(rows)
hdr = rows[i0]
body = [r for r in rows[i0 + 1:i1] if len(r) == len(hdr)]
S, I, src = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
ci = {c: hdr.index(c) for c in cols}
print(rows[i0 - 1][:2] if i0 > 0 else "")
tot = {c: sum(int(r[ci[c]] or 0) for r in body) for c in cols}
T = sum(tot.values()) or 1
print("samples", T, {k: round(100 * v / T, 1) for k, v in sorted(tot.items(), key=lambda x: -x[1])[:9]})
print("instructions executed", sum(int(r[I]) for r in body))
top = sorted(range(len(body)), key=lambda k: -int(body[k][S]))[:40]
for k in sorted(top):
    r = body[k]
    dom = max(cols, key=lambda c: int(r[ci[c]] or 0))
    print(k, r[S], r[I], dom, "|", body[k - 1][src].strip()[:60], "||", r[src].strip()[:80])
