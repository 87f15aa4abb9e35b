"""One view of the 1024^3 volume rendered a few times with a given sector layout — the target of an ncu capture.
    python tools/prof_layout.py <layout x|y|z> <orbit view k> [frames] [tstep]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vrdd_b200 as V
if os.environ.get("VRDD_L2_GRAN"):
    from cuda import cudart
    torch.cuda.init(); torch.zeros(1, device="cuda")
    print("cudaLimitMaxL2FetchGranularity before:", cudart.cudaDeviceGetLimit(cudart.cudaLimit.cudaLimitMaxL2FetchGranularity))
    print("set:", cudart.cudaDeviceSetLimit(cudart.cudaLimit.cudaLimitMaxL2FetchGranularity, int(os.environ["VRDD_L2_GRAN"])))
    print("after:", cudart.cudaDeviceGetLimit(cudart.cudaLimit.cudaLimitMaxL2FetchGranularity))
lay, k = sys.argv[1], int(sys.argv[2])
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 3
tstep = float(sys.argv[4]) if len(sys.argv) > 4 else 0.01
vol = img = 1024
r = V.Renderer(0); r.set_stream(torch.cuda.current_stream().cuda_stream); r.set_volume(vol, vol, vol)
slab = 128
buf = torch.empty(slab * vol * vol * 32, dtype=torch.float32, device="cuda")
for z0 in range(0, vol, slab):
    r.synth_histograms_device(1234, z0, slab, buf); r.set_histograms_device(buf, z0, slab); r.decode(V.SRC_ORIGINAL, z0, slab)
r.synchronize(); del buf; torch.cuda.empty_cache()
out = torch.zeros(img, img, dtype=torch.int32, device="cuda")
p = V.default_render_params(query_method=1, tstep=tstep, max_steps=max(500, int(3.5 / tstep) + 1))
r.set_variant("raycast_layout", {"x": "array", "y": "layers_x", "z": "layers_y"}.get(lay, lay))
r.set_view(V.view_matrix(0.0, k * 360.0 / 64))
r.count_samples(True); r.render(out, img, img, p, clear_misses=True); S = r.get_sample_count(); r.count_samples(False)
for _ in range(frames):
    r.render(out, img, img, p, clear_misses=True)
torch.cuda.synchronize()
print("layout", lay, "view", k, "samples", S)
