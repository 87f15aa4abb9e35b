"""The reference's OWN d_render at the headline size, on the same B200 and the same data as ours.

    python tools/ref_speed_headline.py <dir> [views] [reps]

1. libvrdd decodes the synthetic 1024^3 volume of bench.py (seed 1234) and writes its mean plane to <dir>/mean.f32;
2. oracle/_ref/ref_driver ... headline binds it, as the .x lane of a 1024^3 float4 array (the reference's layout), to
   originalQueryTex and runs the reference's render_kernel (volumeRender_kernel.cu:2387-2401; d_render as written, its
   16x16 blocks, default parameters, queryMethod 1) at 1024x1024 over views spread over bench.py's 64-view orbit;
3. our raycast kernels render the same views: time per frame, and every byte of every frame compared.
Prints one JSON record (-> profiles/traffic.json "reference_gpu_kernel", which bench.py reports next to cpu_baseline).
Keep <dir> outside gpurun_out/ (4.3 GB plane + frames)."""
import json
import os
import re
import subprocess
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import vrdd_b200 as V
import ref_pin as R
import bench

d = os.path.abspath(sys.argv[1])
nviews = int(sys.argv[2]) if len(sys.argv) > 2 else 16
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
vol = img = 1024
R.gen(d)                                                            # the small inputs initCuda wants
views_k = bench.timed_views(nviews)
views = np.array([V.view_matrix(0.0, k * 360.0 / 64) for k in views_k], np.float32)
views.tofile(os.path.join(d, "in", "views.f32"))

r = V.Renderer(0); r.set_stream(torch.cuda.current_stream().cuda_stream)
r.keep_linear_planes(True)
r.set_volume(vol, vol, vol)
slab = 128
buf = torch.empty(slab * vol * vol * 32, dtype=torch.float32, device="cuda")
for z0 in range(0, vol, slab):
    r.synth_histograms_device(1234, z0, slab, buf); r.set_histograms_device(buf, z0, slab); r.decode(V.SRC_ORIGINAL, z0, slab)
r.synchronize(); del buf; torch.cuda.empty_cache()
plane = os.path.join(d, "mean.f32")
V.as_torch(r.get_decoded_planes_device(V.SRC_ORIGINAL)[0], (vol ** 3,)).cpu().numpy().tofile(plane)

# ours, first (the reference run needs 17 GB of device memory on top)
out = torch.zeros(img, img, dtype=torch.int32, device="cuda")
p = V.default_render_params(query_method=1)
ours_ms, frames, samples = [], [], []
for k in views_k:
    r.set_view(V.view_matrix(0.0, k * 360.0 / 64))
    r.count_samples(True); r.render(out, img, img, p, clear_misses=False); samples.append(r.get_sample_count()); r.count_samples(False)
    out.zero_(); r.render(out, img, img, p, clear_misses=False); torch.cuda.synchronize()
    frames.append(out.cpu().numpy().view(np.uint32).copy())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r.render(out, img, img, p, clear_misses=False)
    e1.record(); torch.cuda.synchronize()
    ours_ms.append(e0.elapsed_time(e1) / reps)
r.close(); torch.cuda.empty_cache()

exe = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
res = subprocess.run([exe, d, str(img), str(img), str(nviews), "headline", plane, str(vol), str(reps)], capture_output=True, text=True, timeout=3000)
txt = res.stdout + res.stderr
if res.returncode != 0:
    print(txt[-3000:]); raise SystemExit("ref_driver failed")
ref_ms = [float(m) for m in re.findall(r"ref_headline view \d+: ([0-9.]+) ms", txt)]
assert len(ref_ms) == nviews, txt[-2000:]
worst, differing = 0, 0
for i in range(nviews):
    ref = np.fromfile(os.path.join(d, "out", f"headline_v{i}.u32"), np.uint32).reshape(img, img)
    dd = np.abs(ref.view(np.uint8).astype(np.int16) - frames[i].view(np.uint8).astype(np.int16))
    worst = max(worst, int(dd.max())); differing = max(differing, int((dd != 0).sum()))
rec = {"what": "the reference's own d_render (volumeRender_kernel.cu:272-717, compiled where it lies behind oracle/ref_shim, launched by its "
               "render_kernel with 16x16 blocks) on the SAME B200, the same decoded 1024^3 volume as a float4 array (its layout, 17 GB) and the "
               "same 1024x1024 views; ours = vrdd_render on the fp32 plane(s)",
       "views": views_k, "reference_ms_per_frame": sum(ref_ms) / nviews, "ours_ms_per_frame": sum(ours_ms) / nviews,
       "reference_gsamples_per_s": sum(samples) / sum(ref_ms) / 1e6, "ours_gsamples_per_s": sum(samples) / sum(ours_ms) / 1e6,
       "speedup": sum(ref_ms) / sum(ours_ms), "per_view_ms": {"reference": ref_ms, "ours": ours_ms},
       "frames": {"max_lsb_diff": worst, "max_differing_bytes_per_frame": differing, "bytes_per_frame": img * img * 4},
       "source": "tools/ref_speed_headline.py"}
print(json.dumps(rec))
