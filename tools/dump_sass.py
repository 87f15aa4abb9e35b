"""SASS listings + opcode histograms of the hot kernels of libvrdd.so, and the ptxas -v lines, for profiles/.
    python tools/dump_sass.py r2"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
LIB = os.path.join(ROOT, "volume-rendering-based-on-distribution-data_b200", "libvrdd.so")
BUILD = os.path.join(ROOT, "volume-rendering-based-on-distribution-data_b200", "csrc", "build")
KERNELS = [("decode_hist_tma", r"decode_hist_tma_kernel"), ("decode_fractal_moments2", r"decode_fractal_moments2_kernelILb1ELb0E"),
           ("raycast", r"raycast_kernelILi0ELi1ELb0ELi4E"), ("raycast_gather", r"raycast_gather_kernelILi1ELi1ELb1ELb0ELi2E"),
           ("raycast_mode7", r"raycast_mode7_kernelILb0ELi4ELb1E"), ("raycast_brick", r"raycast_brick_kernelILi2ELb0ELb1E")]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
funcs = re.split(r"(?=\t*Function : )", sass)
for name, pat in KERNELS:
    body = next((f for f in funcs if re.search(r"Function : \S*" + pat, f)), None)
    if body is None:
        print("not found:", name); continue
    ops = collections.Counter()
    listing = []
    for line in body.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?)\s*;\s*/\*", line)
        if m:
            ins = m.group(2)
            listing.append(f"/*{m.group(1)}*/ {ins}")
            op = re.sub(r"^@!?U?P\d+\s+", "", ins).split()[0]
            ops[op.split(".")[0]] += 1
        elif "Function :" in line:
            listing.append(line.strip())
    marks = {k: sum(v for o, v in ops.items() if o.startswith(k)) for k in ("UBLKCP", "SYNCS", "TEX", "TLD4", "LDS", "STS", "SULD", "SUST", "LDG", "STG", "MUFU", "RED", "ATOM", "BAR")}
    with open(os.path.join(ROOT, "profiles", f"sass_{name}_{tag}.txt"), "w") as f:
        f.write(f"# cuobjdump -sass libvrdd.so, kernel {name} ({len(listing) - 1} instructions, static)\n")
        f.write("# instructions of interest: " + ", ".join(f"{k} {v}" for k, v in marks.items() if v) + "\n")
        f.write("# opcode histogram: " + ", ".join(f"{o} {n}" for o, n in ops.most_common()) + "\n\n")
        f.write("\n".join(listing) + "\n")
    print(name, len(listing) - 1, "instructions;", ", ".join(f"{k} {v}" for k, v in marks.items() if v))
with open(os.path.join(ROOT, "profiles", f"ptxas_{tag}.txt"), "w") as f:
    f.write("# nvcc -Xptxas -v of every kernel of libvrdd.so (csrc/build/*.ptxas.log): registers, spills, shared memory\n")
    for log in sorted(os.listdir(BUILD)):
        if not log.endswith(".ptxas.log"):
            continue
        txt = open(os.path.join(BUILD, log)).read()
        for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\nptxas info\s+: Function properties for \S+\n\s+(.*)\nptxas info\s+: (Used .*)", txt):
            dem = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            dem = re.sub(r"vrdd::\(anonymous namespace\)::", "", dem)
            f.write(f"{log[:-10]:16s} {dem[:110]:110s} {m.group(3)}; {m.group(2)}\n")
print("ptxas log written")
