#!/bin/bash
# Fractal-decode variant experiment: timings of every variant, then a short ncu metric list per variant.
out=gpurun_out; mkdir -p $out
variants="${@:-moments moments2 moments2r moments2b moments2br}"
python tools/bench_fractal.py 1024 256 dense $variants > $out/frac_variants.log 2>&1; cat $out/frac_variants.log
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio
timeout 300 ncu --metrics $M --clock-control none -k regex:decode_fractal_moments -c 12 --csv --log-file $out/frac_ncu.csv python tools/bench_fractal.py 1024 256 $variants > $out/frac_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/frac_ncu.csv')) if len(r)>10]
h=rows[0]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value'); ii=h.index('ID')
seen={}
for r in rows[1:]:
    seen.setdefault((r[ii],r[ki][:60]),{})[r[mi]]=r[vi]
done=set()
for (i,k),m in seen.items():
    if k in done: continue
    done.add(k); print(k)
    for a,b in m.items(): print('   ',a,b)
PY
