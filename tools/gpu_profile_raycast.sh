#!/bin/bash
# ncu capture of the ray caster on the full-size (1024^3) volume.
tag=${1:-r1}; shift
extra="$@"
out=gpurun_out
CMD="python bench.py --volume 1024 --steps 4 --warmup 3 --decode-reps 1 --no-cpu --e2e-decode-z 0 --fractal 0 $extra"
$CMD > $out/plain_rc_$tag.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:raycast_kernel -s 12 -c 2 -f -o $out/prof_raycast_$tag $CMD > $out/ncu_rc_$tag.log 2>&1
echo "rc=$?"; tail -c 600 $out/plain_rc_$tag.log
