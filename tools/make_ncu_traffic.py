"""profiles/ncu_dominant_kernel.json: DRAM bytes per launch of the dominant kernel (hidden layer 1) from an
`ncu --set full` capture of `tools/profile_step.py <precision> <npts>` with <npts> = points per chunk (rows = 8 * npts per launch)."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, precision, npts = sys.argv[1], sys.argv[2], int(sys.argv[3])
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, first = rows[0], rows[1], rows[2]


def val(key):
    i = hdr.index(key)
    v = float(first[i].replace(",", ""))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1, "%": 1}[units[i]]
    return v * scale


path = os.path.join(ROOT, "profiles", "ncu_dominant_kernel.json")
data = json.load(open(path)) if os.path.exists(path) else {}
data[precision] = {
    "kernel": first[hdr.index("Kernel Name")].split("(")[0],
    "rows_per_launch": 8 * npts,
    "dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
    "dram_read_bytes": val("dram__bytes_read.sum"), "dram_write_bytes": val("dram__bytes_write.sum"),
    "duration_s_under_ncu": val("gpu__time_duration.sum"),
    "tensor_pipe_active_pct": val("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    "source": os.path.basename(rep),
}
json.dump(data, open(path, "w"), indent=1)
print(data[precision])
