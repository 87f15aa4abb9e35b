"""Turns the ncu captures in gpurun_out/ (tools/gpu_measure.sh <tag>) into the tracked
summaries under profiles/: launch-list shares, key metrics per kernel, traffic.json.
    python tools/summarize_profiles.py r1c"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

tag = sys.argv[1]
G, P = "gpurun_out", "profiles"
os.makedirs(P, exist_ok=True)

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_tex.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__t_sectors_pipe_tex_mem_texture.sum", "l1tex__t_requests_pipe_tex_mem_texture.sum",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sectors.sum",
]
STALLS = re.compile(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active.ratio")


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, r)) for r in rows[2:]], dict(zip(hdr, units))


lines = [f"# ncu summaries, tag {tag}", ""]
traffic = {}

# ---- launch list ---------------------------------------------------------------------------
ll = os.path.join(G, f"launches_{tag}.csv")
if os.path.exists(ll):
    rows = list(csv.DictReader(l for l in open(ll) if l.startswith('"')))
    agg = collections.OrderedDict()
    for r in rows:
        k = re.sub(r"\(.*", "", r["Kernel Name"]).replace("vrdd::<unnamed>::", "").replace("void ", "")
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(r["Metric Value"]) / 1e6
    tot = sum(a[1] for a in agg.values())
    lines += ["## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, cold-cache and serialised: compare shares)",
              "", "Command: `python bench.py --steps 16 --warmup 3 --decode-reps 1 --no-cpu --e2e-decode-z 0 --config1 0 --mode7 0`", "",
              "| kernel | launches | total ms | avg ms | share |", "|---|---:|---:|---:|---:|"]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{k[:70]}` | {n} | {t:.3f} | {t / n:.4f} | {t / tot * 100:.1f}% |")
    lines += ["", f"{len(rows)} launches, {tot:.1f} ms of kernel time. The `synth_*` kernels generate the inputs (untimed by "
              "bench.py); `at::native::*` are torch fills/copies of the harness.", ""]
    subprocess.run(["cp", ll, os.path.join(P, f"launches_{tag}.csv")])

# ---- full captures ---------------------------------------------------------------------------
ALGO = {"decode_hist": lambda d: None}
for name in ("decode_hist", "raycast", "decode_fractal_moments"):
    rep = os.path.join(G, f"prof_{name}_{tag}.ncu-rep")
    if not os.path.exists(rep):
        continue
    recs, units = raw(rep)
    lines += [f"## `{name}` (`ncu --set full --clock-control none --import-source on`)", ""]
    for i, d in enumerate(recs):
        lines += [f"launch {i}: `{d.get('Kernel Name', '')[:110]}`", "", "| metric | value | unit |", "|---|---:|---|"]
        for k in KEYS:
            if k in d and d[k] not in ("", "n/a"):
                lines.append(f"| {k} | {d[k]} | {units.get(k, '')} |")
        st = sorted(((float(v.replace(",", "")), STALLS.match(k).group(1)) for k, v in d.items()
                     if STALLS.match(k) and v not in ("", "n/a")), reverse=True)[:6]
        lines.append("| top stalls (warps per issue) | " + ", ".join(f"{n} {v:.2f}" for v, n in st) + " | |")
        lines.append("")
        if i == 0:
            def num(k):
                v, u = float(d[k].replace(",", "")), units.get(k, "")
                return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12}.get(u, 1)
            traffic[name] = {"dram_bytes_per_launch": num("dram__bytes_read.sum") + num("dram__bytes_write.sum"),
                             "dram_read": num("dram__bytes_read.sum"), "dram_write": num("dram__bytes_write.sum"),
                             "duration_us_under_ncu": float(d["gpu__time_duration.sum"].replace(",", "")) *
                             {"us": 1, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(units.get("gpu__time_duration.sum"), 1)}
    # raw csv page for the record (small)
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    open(os.path.join(P, f"ncu_raw_{name}_{tag}.csv"), "w").write(out)

# ---- DRAM bytes of each of the 64 timed orbit launches of the ray caster ----------------------
orbit = None
oc = os.path.join(G, f"raycast_orbit_{tag}.csv")
if os.path.exists(oc):
    per = collections.defaultdict(dict)
    for r in csv.DictReader(l for l in open(oc) if l.startswith('"')):
        v = float(r["Metric Value"].replace(",", ""))
        v *= {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12, "us": 1, "ms": 1e3, "ns": 1e-3, "s": 1e6,
              "usecond": 1, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}.get(r["Metric Unit"], 1)
        per[r["ID"]][r["Metric Name"]] = v
    rd = [d["dram__bytes_read.sum"] for d in per.values()]
    wr = [d["dram__bytes_write.sum"] for d in per.values()]
    us = [d["gpu__time_duration.sum"] for d in per.values()]
    n = len(rd)
    orbit = {"launches": n, "dram_bytes_per_launch": (sum(rd) + sum(wr)) / n, "dram_read": sum(rd) / n, "dram_write": sum(wr) / n,
             "dram_bytes_min": min(a + b for a, b in zip(rd, wr)), "dram_bytes_max": max(a + b for a, b in zip(rd, wr)),
             "duration_us_under_ncu": sum(us) / n}
    lines += ["## Ray caster, DRAM bytes of each timed launch of the 64-view orbit", "",
              "`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:raycast_kernel -s 69 -c 64` on "
              "`python bench.py --steps 64 --warmup 5 ...` (cold L2 before every launch)", "",
              f"{n} launches: mean {orbit['dram_bytes_per_launch'] / 1e6:.1f} MB per launch (read {orbit['dram_read'] / 1e6:.1f}, "
              f"write {orbit['dram_write'] / 1e6:.1f}), min {orbit['dram_bytes_min'] / 1e6:.1f}, max {orbit['dram_bytes_max'] / 1e6:.1f}; "
              f"mean duration under ncu {orbit['duration_us_under_ncu']:.1f} us", ""]
    subprocess.run(["cp", oc, os.path.join(P, f"raycast_orbit_{tag}.csv")])

open(os.path.join(P, f"summary_{tag}.md"), "w").write("\n".join(lines) + "\n")

# traffic.json: keyed by the kernel names bench.py reports, with the launch's algorithmic bytes
bench = os.path.join(G, f"bench_{tag}.json")
tj = {}
if os.path.exists(bench) and "decode_hist" in traffic:
    b = json.loads(open(bench).read().strip().splitlines()[-1])
    t = traffic["decode_hist"]
    rd = b.get("roofline_decode", b["roofline"])
    tj[rd["kernel"]] = {"algorithmic_bytes_per_launch": rd["bytes_per_launch"],
                                   "dram_bytes_per_launch": t["dram_bytes_per_launch"], "dram_read": t["dram_read"],
                                   "dram_write": t["dram_write"], "source": f"profiles/ncu_raw_decode_hist_{tag}.csv"}
    if "raycast" in traffic:
        tj["raycast_kernel"] = dict(traffic["raycast"], source=f"profiles/ncu_raw_raycast_{tag}.csv")
        if orbit:                      # the per-launch figure bench.py quotes is the orbit mean, like its launch time
            tj["raycast_kernel"] = dict(orbit, source=f"profiles/raycast_orbit_{tag}.csv",
                                        full_capture_side_view=dict(traffic["raycast"], source=f"profiles/ncu_raw_raycast_{tag}.csv"))
    if "decode_fractal_moments" in traffic:
        tj["decode_fractal_moments2_kernel"] = dict(traffic["decode_fractal_moments"],
                                                   source=f"profiles/ncu_raw_decode_fractal_moments_{tag}.csv")
    json.dump(tj, open(os.path.join(P, "traffic.json"), "w"), indent=1)
    subprocess.run(["cp", bench, os.path.join(P, f"bench_{tag}.json")])
    ref = os.path.join(G, f"bench_ref_{tag}.json")
    if os.path.exists(ref):
        subprocess.run(["cp", ref, os.path.join(P, f"bench_ref_{tag}.json")])
print(open(os.path.join(P, f"summary_{tag}.md")).read()[:6000])
print(json.dumps(tj, indent=1))
