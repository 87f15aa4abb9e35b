#!/bin/bash
# One measurement pass on the GPU box: tests, bench (both arms), the ncu launch list of the
# bench command and full captures of the hot kernels at the bench's own launch shapes.
# Everything lands in gpurun_out/ (copy what should be judged into profiles/).
set -u
tag=${1:-r1}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -5 > $out/tests_$tag.log
cat $out/tests_$tag.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench (ours)"; ( time timeout 900 python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err ) 2>&1 | grep real
tail -c 400 $out/bench_$tag.err
echo "== bench (reference arm)"; ( time timeout 600 python bench.py --impl reference --steps 8 --warmup 2 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err ) 2>&1 | grep real
CMD="python bench.py --steps 16 --warmup 3 --decode-reps 1 --no-cpu --e2e-decode-z 0 --config1 0 --mode7 0"
echo "== ncu launch list"
$CMD > $out/plain_$tag.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $out/launches_$tag.csv $CMD > $out/ncu_launches_$tag.log 2>&1
echo "rc=$?"
echo "== ncu DRAM traffic of every view of the orbit (the 64 timed ray-cast launches)"
ORB="python bench.py --steps 64 --warmup 5 --decode-reps 1 --no-cpu --e2e-decode-z 0 --config1 0 --mode7 0"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:raycast_kernel -s 69 -c 64 --csv --log-file $out/raycast_orbit_$tag.csv $ORB > $out/ncu_orbit_$tag.log 2>&1
echo "rc=$?"
for spec in "decode_hist_tma_kernel 1 1 decode_hist" "raycast_kernel 30 2 raycast" "decode_fractal_moments2_kernel 1 1 decode_fractal_moments"; do
  set -- $spec
  echo "== ncu full: $1"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c $3 -f -o $out/prof_$4_$tag $CMD > $out/ncu_$4_$tag.log 2>&1; echo "rc=$?"
done
ls -la $out | grep $tag
