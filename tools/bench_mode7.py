"""queryMethod 7 (interpolated mean, point-sampled cells) over the orbit, per fetch path: time per view, samples,
Gsamples/s, and whether the paths produce the same frame.  The array behind a path is chosen at decode time, so
every variant gets its own decode.
    python tools/bench_mode7.py [edge] [image]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vrdd_b200 as V
vol = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
img = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
out = torch.zeros(img, img, dtype=torch.int32, device="cuda")
frames = {}
for qm, variant in ((7, "gather"), (7, "texture"), (7, "linear"), (1, "texture")):
    r = V.Renderer(0); r.set_stream(torch.cuda.current_stream().cuda_stream); r.set_volume(vol, vol, vol)
    r.enable_interpolated_mean(qm == 7)
    r.set_variant("raycast_mode7", variant)
    slab = min(128, vol)
    buf = torch.empty(slab * vol * vol * 32, dtype=torch.float32, device="cuda")
    for z0 in range(0, vol, slab):
        r.synth_histograms_device(1234, z0, slab, buf); r.set_histograms_device(buf, z0, slab); r.decode(V.SRC_ORIGINAL, z0, slab)
    r.synchronize(); del buf; torch.cuda.empty_cache()
    p = V.default_render_params(query_method=qm)
    rows = []
    for k in range(0, 64, 8):
        r.set_view(V.view_matrix(0.0, k * 360.0 / 64))
        r.count_samples(True); r.render(out, img, img, p, clear_misses=True); S = r.get_sample_count(); r.count_samples(False)
        for _ in range(2): r.render(out, img, img, p, clear_misses=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): r.render(out, img, img, p, clear_misses=True)
        e1.record(); torch.cuda.synchronize()
        rows.append((k, e0.elapsed_time(e1) / 5, S))
    if qm == 7:
        frames[variant] = out.clone()
    ms = sum(m for _, m, _ in rows); S = sum(s for _, _, s in rows)
    print(f"qm {qm} ({variant}): {S / ms / 1e6:.1f} Gsamples/s over {len(rows)} views;  " +
          "  ".join(f"v{k}:{m:.3f}ms/{s/1e6:.0f}M/{s/m/1e6:.0f}G" for k, m, s in rows), flush=True)
    r.close()
print("gather frame == texture frame == linear frame:",
      bool(torch.equal(frames["gather"], frames["texture"]) and torch.equal(frames["linear"], frames["texture"])))
