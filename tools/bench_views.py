"""Per-view time of the ray caster over the 64-view orbit (1024^3 volume, 1024^2 frames)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vrdd_b200 as V
vol, img = 1024, 1024
r = V.Renderer(0); r.set_stream(torch.cuda.current_stream().cuda_stream); r.set_volume(vol, vol, vol)
slab = 128
buf = torch.empty(slab * vol * vol * 32, dtype=torch.float32, device="cuda")
for z0 in range(0, vol, slab):
    r.synth_histograms_device(1234, z0, slab, buf); r.set_histograms_device(buf, z0, slab); r.decode(V.SRC_ORIGINAL, z0, slab)
r.synchronize(); del buf; torch.cuda.empty_cache()
out = torch.zeros(img, img, dtype=torch.int32, device="cuda")
p = V.default_render_params(query_method=1)
for U in ("4", "8"):
    r.set_variant("raycast_unroll", U)
    rows = []
    for k in range(0, 64, 4):
        r.set_view(V.view_matrix(0.0, k * 360.0 / 64))
        r.count_samples(True); r.render(out, img, img, p, clear_misses=True); S = r.get_sample_count(); r.count_samples(False)
        for _ in range(3): r.render(out, img, img, p, clear_misses=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): r.render(out, img, img, p, clear_misses=True)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        rows.append((k, ms, S))
    print("U=" + U + ": " + "  ".join(f"v{k}:{ms:.3f}ms/{S/1e6:.0f}M/{S/ms/1e6:.0f}G" for k, ms, S in rows), flush=True)
