#!/bin/bash
# One measurement pass on the GPU box: tests, bench (both arms), ncu launch list and full
# captures of the three hot kernels.  Everything lands in gpurun_out/.
set -u
tag=${1:-r1}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -5 > $out/tests_$tag.log
cat $out/tests_$tag.log
echo "== bench (ours)"; ( time timeout 900 python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err ) 2>&1 | grep real
tail -c 400 $out/bench_$tag.err
echo "== bench (reference arm)"; ( time timeout 600 python bench.py --impl reference --steps 8 --warmup 2 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err ) 2>&1 | grep real
SMALL="python bench.py --volume 512 --slab-z 128 --steps 4 --warmup 3 --decode-reps 1 --no-cpu --e2e-decode-z 0"
echo "== ncu launch list"
$SMALL > $out/plain_$tag.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_$tag.csv $SMALL > $out/ncu_launches_$tag.log 2>&1
echo "rc=$?"
echo "== ncu full: decode_hist"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_hist -s 2 -c 2 -f -o $out/prof_decode_hist_$tag $SMALL > $out/ncu_dh_$tag.log 2>&1; echo "rc=$?"
echo "== ncu full: raycast"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:raycast_kernel -s 12 -c 2 -f -o $out/prof_raycast_$tag $SMALL > $out/ncu_rc_$tag.log 2>&1; echo "rc=$?"
echo "== ncu full: decode_fractal"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_fractal -s 1 -c 1 -f -o $out/prof_decode_fractal_$tag $SMALL > $out/ncu_df_$tag.log 2>&1; echo "rc=$?"
ls -la $out | tail -20
