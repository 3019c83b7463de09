"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares."""
import collections
import csv
import gzip
import json
import sys


def main(path, command):
    op = gzip.open if path.endswith(".gz") else open
    tot, cnt = collections.Counter(), collections.Counter()
    with op(path, "rt") as f:
        rows = [r for r in csv.reader(f) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for r in rows[1:]:
        if r[hdr.index("Metric Name")] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
        name = r[ki].split("(")[0].replace("stpde::", "")
        tot[name] += v * scale
        cnt[name] += 1
    total = sum(tot.values())
    out = {"command": command, "note": "cold-cache serialized per-launch times: compare SHARES with bench.py's kernel_ms, "
           "not absolutes", "total_ms": total,
           "kernels": [{"kernel": k, "launches": cnt[k], "ms": round(v, 3), "share": round(v / total, 4)}
                       for k, v in tot.most_common()]}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
