"""Static SASS opcode counts per kernel family of libstpde.so (cuobjdump -sass; all template instantiations of a family
summed).  usage: python tools/sass_histogram.py [path/to/libstpde.so] > profiles/rNN_sass_opcode_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "space_time_pde_b200", "libstpde.so")
OPS = ["UTCHMMA", "UTCQMMA", "LDTM", "UTMALDG", "UTMASTG", "UTMACMDFLUSH", "UTCBAR", "SYNCS", "STS", "LDS", "STG", "LDG",
       "LDGSTS", "RED", "REDG", "ATOMG", "MUFU", "F2FP", "LDL", "STL"]
FAMILIES = ["blend_backward_kernel", "final_blend_kernel", "layer0_jets_tc_kernel", "layer_gemm_kernel", "prep_points_kernel",
            "residual", "tc_layer_pair_kernel", "tc_layer_kernel", "tc_wgrad_pair_kernel", "vertex_bias_kernel",
            "vertex_backward"]
proc = subprocess.Popen(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True)
counts = collections.defaultdict(collections.Counter)
nfn = collections.Counter()
fam = None
op_re = re.compile(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)")
for line in proc.stdout:
    if line.lstrip().startswith("Function :"):
        name = line.split(":", 1)[1].strip()
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
        fam = next((f for f in FAMILIES if f in dem), "other")
        nfn[fam] += 1
        continue
    m = op_re.match(line)
    if m and fam is not None:
        op = m.group(1)
        if op in OPS:
            counts[fam][op] += 1
print(f"Static SASS opcode counts per kernel family of {os.path.relpath(lib, ROOT)} (cuobjdump -sass, all template instantiations")
print("of a family summed; tools/sass_histogram.py)")
print()
print(f"{'family':28s} {'#fn':>4s} " + " ".join(f"{o:>8s}" for o in OPS))
tot = collections.Counter()
for f in sorted(nfn):
    print(f"{f:28s} {nfn[f]:4d} " + " ".join(f"{counts[f][o]:8d}" for o in OPS))
    tot.update(counts[f])
print(f"{'TOTAL':28s} {sum(nfn.values()):4d} " + " ".join(f"{tot[o]:8d}" for o in OPS))
