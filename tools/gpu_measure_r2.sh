#!/bin/bash
# Round-2 measurement pass on ONE B200: tests, smoke, both bench arms, the ncu launch list of the bench command, the DRAM
# traffic of every ray-cast launch of the orbit (single GPU and rank 0's share of the N = 2/4/8 frames, weak and strong),
# full captures of the hot kernels, and the early-exit capture.  Everything lands in gpurun_out/<tag>/; tools/
# summarize_profiles_r2.py condenses it into profiles/.
set -u
tag=${1:-r2}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 > $out/tests.log; cat $out/tests.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench (reference arm)"; ( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $out/bench_ref.json 2> $out/bench_ref.err ) 2>&1 | grep real
echo "== bench (ours)"; ( time timeout 900 python bench.py --steps 20 --warmup 5 > $out/bench.json 2> $out/bench.err ) 2>&1 | grep real
tail -c 300 $out/bench.err
( time timeout 900 python bench.py > $out/bench_64.json 2> $out/bench_64.err ) 2>&1 | grep real
CMD="python bench.py --steps 16 --warmup 3 --decode-reps 1 --no-cpu --e2e-decode-z 0 --config1 0 --mode7 0 --flex 0 --l1tex 0 --matched 0"
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $out/launches.csv $CMD > $out/ncu_launches.log 2>&1; echo "rc=$?"
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
K='regex:raycast_(gather_)?kernel'
echo "== ncu DRAM traffic of the orbit (64 views, one GPU)"
timeout 900 ncu --metrics $M --clock-control none -k "$K" --csv --log-file $out/orbit_n1.csv python tools/prof_orbit.py --views 64 > $out/orbit_n1.log 2>&1; echo "rc=$?"
timeout 900 ncu --metrics $M --clock-control none -k "$K" --csv --log-file $out/orbit_n1_array.csv python tools/prof_orbit.py --views 32 --layout array > $out/orbit_n1_array.log 2>&1; echo "rc=$?"
for N in 2 4 8; do
  timeout 900 ncu --metrics $M --clock-control none -k "$K" --csv --log-file $out/orbit_weak_n$N.csv python tools/prof_orbit.py --views 16 --world $N > $out/orbit_weak_n$N.log 2>&1; echo "weak $N rc=$?"
done
for N in 1 2 4 8; do
  timeout 900 ncu --metrics $M --clock-control none -k "$K" --csv --log-file $out/orbit_strong_n$N.csv python tools/prof_orbit.py --views 16 --world $N --strong 1 > $out/orbit_strong_n$N.log 2>&1; echo "strong $N rc=$?"
done
echo "== full captures"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_hist_tma_kernel -s 1 -c 1 -f -o $out/prof_decode_hist $CMD > $out/ncu_decode_hist.log 2>&1; echo "rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_fractal_moments2_kernel -s 1 -c 1 -f -o $out/prof_decode_fractal $CMD > $out/ncu_decode_fractal.log 2>&1; echo "rc=$?"
# view 10 (56 deg off z: oblique, 3-D array... auto picks the copy), view 16 (straight along x: layered copy), view 0 (frontal: 3-D array), view 6 (34 deg: 3-D array, oblique)
for v in 0 6 16; do
  timeout 900 ncu --set full --clock-control none --import-source on -k "$K" -s 3 -c 1 -f -o $out/prof_raycast_v$v python tools/prof_orbit.py --views 3 --only-view $v > $out/ncu_raycast_v$v.log 2>&1; echo "view $v rc=$?"
done
echo "== early exit: density 0.3 (most rays terminate early), oblique view 6"
timeout 900 ncu --set full --clock-control none -k "$K" -s 3 -c 1 -f -o $out/prof_raycast_v6_dense python tools/prof_orbit.py --views 3 --only-view 6 --density 0.3 > $out/ncu_raycast_v6_dense.log 2>&1; echo "rc=$?"
ls -la $out
