"""Condenses gpurun_out/<tag>/ (tools/gpu_measure_r2.sh <tag>) into the tracked files under profiles/:
bench lines, launch list + shares, per-view DRAM traffic of the orbit (one GPU, and rank 0's share of the N = 2/4/8
frames, weak and strong), key metrics of the full captures, traffic.json (what bench.py reports as roofline.traffic).
    python tools/summarize_profiles_r2.py r2"""
import ast
import collections
import csv
import json
import os
import re
import subprocess
import sys

tag = sys.argv[1]
G, P = os.path.join("gpurun_out", tag), "profiles"
os.makedirs(P, exist_ok=True)
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tex.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__m_xbar2l1tex_read_bytes.sum"]
STALLS = re.compile(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active.ratio")


def metric_rows(path):
    """rows of an `ncu --csv --metrics ...` log: launch id -> {metric: value in base units}, kernel name."""
    out = collections.OrderedDict()
    if not os.path.exists(path):
        return out
    for r in csv.reader(l for l in open(path) if l.startswith('"')):
        if not r or not r[0].isdigit():
            continue
        d = out.setdefault(int(r[0]), {"kernel": r[4]})
        d[r[12]] = float(r[14].replace(",", "")) * UNIT.get(r[13], 1.0)
    return out


def prof_info(log):
    for l in open(log):
        if l.startswith("PROF_ORBIT"):
            return ast.literal_eval(l[len("PROF_ORBIT"):].strip())
    return None


def orbit(name):
    """second half of the launches of tools/prof_orbit.py (the first half built the copies and counted samples)"""
    rows = list(metric_rows(os.path.join(G, name + ".csv")).values())
    info = prof_info(os.path.join(G, name + ".log")) if os.path.exists(os.path.join(G, name + ".log")) else None
    if not rows or not info:
        return None
    n = len(info["views"])
    rows = rows[-n:]
    per = []
    for v, s, r in zip(info["views"], info["samples"], rows):
        per.append({"view": v, "samples": s, "dram_read": r.get("dram__bytes_read.sum", 0.0), "dram_write": r.get("dram__bytes_write.sum", 0.0),
                    "us": r.get("gpu__time_duration.sum", 0.0) * 1e6, "kernel": "gather" if "gather" in r["kernel"] else "array"})
    return info, per


lines = [f"# ncu summaries, tag {tag}", ""]
traffic = {}
try:
    traffic = json.load(open(os.path.join(P, "traffic.json")))
except Exception:
    pass

for f in ("bench.json", "bench_64.json", "bench_ref.json"):
    src = os.path.join(G, f)
    if os.path.exists(src) and os.path.getsize(src):
        subprocess.run(["cp", src, os.path.join(P, f.replace(".json", f"_{tag}.json"))])

# ---- launch list ---------------------------------------------------------------------------
ll = os.path.join(G, "launches.csv")
if os.path.exists(ll):
    rows = list(csv.DictReader(l for l in open(ll) if l.startswith('"')))
    agg = collections.OrderedDict()
    for r in rows:
        k = re.sub(r"\(.*", "", r["Kernel Name"]).replace("vrdd::<unnamed>::", "").replace("void ", "")
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1e-9) * 1e3
    tot = sum(a[1] for a in agg.values())
    lines += ["## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, cold-cache and serialised: compare shares)", "",
              "Command: `python bench.py --steps 16 --warmup 3 --decode-reps 1 --no-cpu --e2e-decode-z 0 --config1 0 --mode7 0 --flex 0 --l1tex 0 --matched 0`", "",
              "| kernel | launches | total ms | avg ms | share |", "|---|---:|---:|---:|---:|"]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{k[:80]}` | {n} | {t:.3f} | {t / n:.4f} | {t / tot * 100:.1f}% |")
    lines += ["", f"{len(rows)} launches, {tot:.1f} ms of kernel time. The `synth_*` kernels generate the inputs (untimed by bench.py); "
              "`at::native::*` are torch fills/copies of the harness.", ""]
    subprocess.run(["cp", ll, os.path.join(P, f"launches_{tag}.csv")])

# ---- orbit traffic ---------------------------------------------------------------------------
lines += ["## DRAM traffic of the ray-cast launches over the orbit (`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum`)", "",
          "`tools/prof_orbit.py`: one GPU renders rank 0's share of the N-rank frame (64x64 tiles round-robin); bytes and times are per launch, "
          "cold L2 before each (ncu). B/sample = DRAM bytes / transfer-function lookups of that launch.", "",
          "| shape | frame | views | mean DRAM MB / launch | min..max MB | mean B/sample | array / gather launches | mean us under ncu |", "|---|---|---:|---:|---:|---:|---:|---:|"]
for name, key in [("orbit_n1", "raycast_kernel"), ("orbit_n1_array", "raycast_kernel_array_only")] + [(f"orbit_weak_n{n}", f"raycast_tiles_n{n}") for n in (2, 4, 8)] + \
        [(f"orbit_strong_n{n}", f"raycast_strong_n{n}") for n in (1, 2, 4, 8)]:
    res = orbit(name)
    if not res:
        continue
    info, per = res
    by = [p["dram_read"] + p["dram_write"] for p in per]
    s = sum(p["samples"] for p in per)
    traffic[key] = {"launches": len(per), "dram_bytes_per_launch": sum(by) / len(by), "dram_read": sum(p["dram_read"] for p in per) / len(per),
                    "dram_bytes_min": min(by), "dram_bytes_max": max(by), "samples_per_launch": s / len(per),
                    "bytes_per_sample": sum(by) / s, "duration_us_under_ncu": sum(p["us"] for p in per) / len(per),
                    "frame": info["frame"], "layout": info["layout"], "source": f"profiles/{name}_{tag}.csv (tools/prof_orbit.py)"}
    lines.append(f"| {key} | {info['frame'][0]}x{info['frame'][1]} | {len(per)} | {sum(by) / len(by) / 1e6:.1f} | {min(by) / 1e6:.0f}..{max(by) / 1e6:.0f} | "
                 f"{sum(by) / s:.1f} | {sum(p['kernel'] == 'array' for p in per)} / {sum(p['kernel'] == 'gather' for p in per)} | {sum(p['us'] for p in per) / len(per):.1f} |")
    subprocess.run(["cp", os.path.join(G, name + ".csv"), os.path.join(P, f"{name}_{tag}.csv")])
    if name in ("orbit_n1", "orbit_n1_array"):
        lines += ["", f"per view ({name}): " + "  ".join(f"v{p['view']}:{(p['dram_read'] + p['dram_write']) / 1e6:.0f}MB/{(p['dram_read'] + p['dram_write']) / p['samples']:.0f}B/{p['kernel'][0]}" for p in per), ""]
lines.append("")


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return [dict(zip(rows[0], r)) for r in rows[2:]], dict(zip(rows[0], rows[1]))


for name, what in (("prof_decode_hist", "decode_hist_tma_kernel, one 1024x1024x256 slab"), ("prof_decode_fractal", "decode_fractal_moments2_kernel, one slab"),
                   ("prof_raycast_v0", "ray caster, view 0 (frontal: 3-D array, texture unit)"), ("prof_raycast_v6", "ray caster, view 6 (34 deg off z: 3-D array, oblique)"),
                   ("prof_raycast_v16", "ray caster, view 16 (along x: layered copy, tld4 + integer weights)"),
                   ("prof_raycast_v6_dense", "ray caster, view 6 at density 0.3 (most rays terminate early)")):
    rep = os.path.join(G, name + ".ncu-rep")
    if not os.path.exists(rep):
        continue
    recs, units = raw(rep)
    if not recs:
        continue
    d = recs[0]
    with open(os.path.join(P, f"ncu_raw_{name[5:]}_{tag}.csv"), "w") as f:
        f.write(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)
    lines += [f"## {what} (`ncu --set full --clock-control none`)", "", f"`{d.get('Kernel Name', '')[:120]}`", "", "| metric | value | unit |", "|---|---:|---|"]
    for k in KEYS:
        if k in d and d[k] not in ("", "n/a"):
            lines.append(f"| {k} | {d[k]} | {units.get(k, '')} |")
    st = sorted(((float(v.replace(",", "")), STALLS.match(k).group(1)) for k, v in d.items() if STALLS.match(k) and v not in ("", "n/a")), reverse=True)[:6]
    lines.append("| top stalls (warps per issue) | " + ", ".join(f"{n} {v:.2f}" for v, n in st) + " | |")
    lines.append("")

    def num(k):
        return float(d[k].replace(",", "")) * UNIT.get(units.get(k, ""), 1.0)
    if name == "prof_decode_hist":
        traffic["decode_hist_tma_kernel"] = {"algorithmic_bytes_per_launch": 1024 * 1024 * 256 * 140, "dram_bytes_per_launch": num("dram__bytes_read.sum") + num("dram__bytes_write.sum"),
                                             "dram_read": num("dram__bytes_read.sum"), "dram_write": num("dram__bytes_write.sum"), "source": f"profiles/ncu_raw_decode_hist_{tag}.csv"}
    if name == "prof_decode_fractal":
        traffic["decode_fractal_moments2_kernel"] = {"dram_bytes_per_launch": num("dram__bytes_read.sum") + num("dram__bytes_write.sum"), "dram_read": num("dram__bytes_read.sum"),
                                                     "dram_write": num("dram__bytes_write.sum"), "duration_us_under_ncu": num("gpu__time_duration.sum") * 1e6,
                                                     "source": f"profiles/ncu_raw_decode_fractal_{tag}.csv"}

json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
open(os.path.join(P, f"summary_{tag}.md"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines)[:6000])
