"""Per-view time of the ray caster over the 64-view orbit (1024^3 volume, 1024^2 frames) for each sector layout:
x = the 3-D array filtered by the texture unit, y / z = layered copies read with tld4 (raycast_gather_kernel), auto.
    python tools/bench_layouts.py [--tstep T] [--every K] [--unroll U]"""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vrdd_b200 as V
if os.environ.get("VRDD_L2_GRAN"):
    from cuda import cudart
    torch.cuda.init(); torch.zeros(1, device="cuda")
    print("cudaLimitMaxL2FetchGranularity before:", cudart.cudaDeviceGetLimit(cudart.cudaLimit.cudaLimitMaxL2FetchGranularity))
    print("set:", cudart.cudaDeviceSetLimit(cudart.cudaLimit.cudaLimitMaxL2FetchGranularity, int(os.environ["VRDD_L2_GRAN"])))
    print("after:", cudart.cudaDeviceGetLimit(cudart.cudaLimit.cudaLimitMaxL2FetchGranularity))
ap = argparse.ArgumentParser()
ap.add_argument("--tstep", type=float, default=0.01)
ap.add_argument("--every", type=int, default=4)
ap.add_argument("--unroll", default="4")
ap.add_argument("--layouts", default="array,layers_x,auto")
ap.add_argument("--rot-x", type=float, default=0.0)
ap.add_argument("--vol", type=int, default=1024)
ap.add_argument("--img", type=int, default=1024)
ap.add_argument("--world", type=int, default=1, help="render rank 0's tiles of the WORLD-rank weak frame")
ap.add_argument("--persist", default="100")
ap.add_argument("--gather-tf", default="follow")
ap.add_argument("--gather-unroll", default="")
ap.add_argument("--tf", default="")
ap.add_argument("--cos", default="", help="auto: cosine of the largest angle between view direction and stacking axis for which the copy is used")
ap.add_argument("--array-blocks", default="", help="resident blocks per SM of the 3-D array kernel (0 = all that fit)")
ap.add_argument("--tile", type=int, default=0, help="enumerate the 16x16-pixel items tile by tile (TILE x TILE pixels) instead of row by row")
a = ap.parse_args()
vol, img = a.vol, a.img
r = V.Renderer(0); r.set_stream(torch.cuda.current_stream().cuda_stream); r.set_volume(vol, vol, vol)
slab = 128
buf = torch.empty(slab * vol * vol * 32, dtype=torch.float32, device="cuda")
for z0 in range(0, vol, slab):
    r.synth_histograms_device(1234, z0, slab, buf); r.set_histograms_device(buf, z0, slab); r.decode(V.SRC_ORIGINAL, z0, slab)
r.synchronize(); del buf; torch.cuda.empty_cache()
import vrdd_b200.dist as D
fw, fh = D.frame_size(img, a.world)
part = V.TilePartition(D.TILE, D.TILE, 0, a.world) if a.world > 1 else (V.TilePartition(a.tile, a.tile, 0, 1) if a.tile else None)
out = torch.zeros(fh, fw, dtype=torch.int32, device="cuda")
r.set_variant("raycast_persist_pct", a.persist)
r.set_variant("raycast_gather_tf", a.gather_tf)
if a.tf:
    r.set_variant("raycast_tf", a.tf)
import math
ms_ = int(math.ceil(2 * math.sqrt(3.0) / a.tstep)) + 1
p = V.default_render_params(query_method=1, tstep=a.tstep, max_steps=max(500, ms_))
r.set_variant("raycast_unroll", a.unroll if a.unroll in ("1", "2", "4", "8") else "4")
if a.array_blocks:
    r.set_variant("raycast_array_blocks_per_sm", a.array_blocks)
if a.cos:
    r.set_variant("raycast_layout_cos", a.cos)
if a.gather_unroll:
    r.set_variant("raycast_gather_unroll", a.gather_unroll)
res = {}
ref_frames = {}
for lay in a.layouts.split(","):
    r.set_variant("raycast_layout", lay)
    rows = []
    for k in range(0, 64, a.every):
        r.set_view(V.view_matrix(a.rot_x, k * 360.0 / 64))
        r.count_samples(True); r.render(out, fw, fh, p, part=part, clear_misses=True); S = r.get_sample_count(); r.count_samples(False)
        for _ in range(3): r.render(out, fw, fh, p, part=part, clear_misses=True)
        torch.cuda.synchronize()
        if lay == "array":
            ref_frames[k] = out.clone()
        elif k in ref_frames:
            d = (out.view(torch.uint8).to(torch.int16) - ref_frames[k].view(torch.uint8).to(torch.int16)).abs()
            assert int(d.max()) <= 1, (lay, k, int(d.max()))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): r.render(out, fw, fh, p, part=part, clear_misses=True)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        rows.append((k, ms, S))
    tot_ms = sum(m for _, m, _ in rows); tot_s = sum(s for _, _, s in rows)
    res[lay] = {"views": rows, "gsamples_per_s": tot_s / tot_ms / 1e6, "ms_per_view": tot_ms / len(rows)}
    print(f"layout {lay:4s} ({tot_s / tot_ms / 1e6:6.1f} Gsamples/s, {tot_ms / len(rows):.3f} ms/view): " +
          "  ".join(f"v{k}:{ms:.3f}/{S/ms/1e6:.0f}G" for k, ms, S in rows), flush=True)
print(json.dumps({"tstep": a.tstep, "unroll": a.unroll, "rot_x": a.rot_x, "vol": vol, "img": img,
                  "layouts": {k: {"gsamples_per_s": v["gsamples_per_s"], "ms_per_view": v["ms_per_view"]} for k, v in res.items()}}))
