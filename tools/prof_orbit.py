"""The ray-cast launches of bench.py's orbit on ONE GPU, for ncu: rank 0's share of an N-rank frame (image-space tiles,
64x64 round-robin — every rank's share is statistically the same), weak (frame grows with N) or strong (fixed 2048^2).

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        -k regex:"raycast_(gather_)?kernel" --csv --log-file out.csv python tools/prof_orbit.py --world 4 --views 16

One untimed launch per view comes first (it builds the layered copies: kernel build_gather_copy_kernel, not matched by
the regex), then exactly one launch per view: the launches ncu sees are 2 x views, the SECOND half is the set to use."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vrdd_b200 as V
import vrdd_b200.dist as D
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--world", type=int, default=1)
ap.add_argument("--strong", type=int, default=0)
ap.add_argument("--views", type=int, default=64)
ap.add_argument("--density", type=float, default=0.05)
ap.add_argument("--layout", default="auto")
ap.add_argument("--vol", type=int, default=1024)
ap.add_argument("--image", type=int, default=1024)
ap.add_argument("--only-view", type=int, default=-1)
a = ap.parse_args()
vol = a.vol
r = V.Renderer(0); r.set_stream(torch.cuda.current_stream().cuda_stream); r.set_volume(vol, vol, vol)
slab = 128
buf = torch.empty(slab * vol * vol * 32, dtype=torch.float32, device="cuda")
for z0 in range(0, vol, slab):
    r.synth_histograms_device(1234, z0, slab, buf); r.set_histograms_device(buf, z0, slab); r.decode(V.SRC_ORIGINAL, z0, slab)
r.synchronize(); del buf; torch.cuda.empty_cache()
fw, fh = (bench.STRONG_IMAGE, bench.STRONG_IMAGE) if a.strong else D.frame_size(a.image, a.world)
part = V.TilePartition(D.TILE, D.TILE, 0, a.world) if a.world > 1 else None
out = torch.zeros(fh, fw, dtype=torch.int32, device="cuda")
p = V.default_render_params(query_method=1, density=a.density)
r.set_variant("raycast_layout", a.layout)
views = bench.timed_views(a.views) if a.only_view < 0 else [a.only_view] * a.views
counts = []
for rep in range(2):
    r.count_samples(rep == 0)
    for k in views:
        r.set_view(bench.orbit_view(V, k))
        r.render(out, fw, fh, p, part=part, clear_misses=True)
        if rep == 0:
            counts.append(r.get_sample_count())
torch.cuda.synchronize()
print("PROF_ORBIT", {"world": a.world, "strong": a.strong, "frame": [fw, fh], "views": views, "samples": counts, "density": a.density,
                     "layout": a.layout})
