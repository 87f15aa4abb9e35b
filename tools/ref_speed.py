"""The reference's own kernels and ours on the same GPU at the REFERENCE'S configuration (50x50x10 blocks x 32 bins,
512x512 frames, its self-test view): oracle/_ref/ref_driver ... time, then libvrdd.so through the C ABI.  No torch."""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import ref_pin as R

d = os.path.abspath(sys.argv[1])
R.gen(d)
exe = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
out = subprocess.run([exe, d, "64", "64", str(len(R.VIEWS)), "time"], capture_output=True, text=True, timeout=300)
print("\n".join(l for l in (out.stdout + out.stderr).splitlines() if l.startswith("ref_")))

import vrdd_b200 as V
from oracle.vrdd_oracle import Oracle
o = Oracle()
hist, cb, tmpl, err, views = R.inputs(o)
r = V.Renderer(0)
r.enable_interpolated_mean(True)
r.set_volume(*R.DIMS)
r.set_histograms_host(hist); r.set_fractal_host(cb, err, tmpl)
r.decode(V.SRC_ORIGINAL); r.decode(V.SRC_FRACTAL); r.synchronize()
t0 = time.perf_counter()
for _ in range(50):
    r.decode(V.SRC_ORIGINAL); r.decode(V.SRC_FRACTAL)
r.synchronize()
print(f"ours decode (25000 blocks, both sources): {(time.perf_counter() - t0) / 50 * 1e3:.4f} ms per pair of launches (wall clock)")
W = H = 512
frame = r.frame_alloc(W * H * 4)
r.set_view(views[0])
for qm, iters in ((1, 500), (4, 500), (7, 200)):
    p = V.default_render_params(query_method=qm)
    r.render(frame, W, H, p, clear_misses=True); r.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        r.render(frame, W, H, p, clear_misses=True)
    r.synchronize()
    print(f"ours render queryMethod {qm} {W}x{H}: {(time.perf_counter() - t0) / iters * 1e3:.4f} ms per frame (wall clock, launch-bound)")
r.close()
