"""Ray-caster experiments on a decoded VOL^3 volume: unroll x transfer-function path x sampler,
reference constants and the resolution-matched step.  python tools/bench_raycast.py [vol] [img]"""
import os, sys, math
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vrdd_b200 as V

vol = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
img = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
samplers = sys.argv[3].split(",") if len(sys.argv) > 3 else ["texture", "bricked"]
NV = 16

def make(sampler):
    r = V.Renderer(0)
    r.set_stream(torch.cuda.current_stream().cuda_stream)
    r.set_sampler(V.SAMPLER_TEXTURE if sampler == "texture" else V.SAMPLER_BRICKED)
    r.set_volume(vol, vol, vol)
    slab = min(vol, 128)
    buf = torch.empty(slab * vol * vol * 32, dtype=torch.float32, device="cuda")
    for z0 in range(0, vol, slab):
        r.synth_histograms_device(1234, z0, slab, buf)
        r.set_histograms_device(buf, z0, slab)
        r.decode(V.SRC_ORIGINAL, z0, slab)
    r.synchronize(); del buf; torch.cuda.empty_cache()
    return r

out = torch.zeros(img, img, dtype=torch.int32, device="cuda")
for sampler in samplers:
    r = make(sampler)
    for label, over in (("reference step 0.01", {}),
                        ("matched step 2/N", {"tstep": 2.0 / vol, "max_steps": int(math.ceil(2 * math.sqrt(3) * vol / 2)) + 1})):
        p = V.default_render_params(query_method=1, **over)
        r.count_samples(True)
        for k in range(NV):
            r.set_view(V.view_matrix(0.0, k * 360.0 / 64)); r.render(out, img, img, p, clear_misses=True)
        S = r.get_sample_count(); r.count_samples(False)
        for tf in (("texture", "smem") if sampler == "texture" else ("smem",)):
            for U in ("1", "2", "4", "8"):
                r.set_variant("raycast_tf", tf); r.set_variant("raycast_unroll", U)
                for k in range(3):
                    r.set_view(V.view_matrix(0.0, k * 360.0 / 64)); r.render(out, img, img, p, clear_misses=True)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for k in range(NV):
                    r.set_view(V.view_matrix(0.0, k * 360.0 / 64)); r.render(out, img, img, p, clear_misses=True)
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / NV
                print(f"vol {vol} img {img} {sampler:8s} tf {tf:7s} U={U} {label:20s}: {ms:8.4f} ms/frame {S/NV/ms/1e6:8.1f} Gsamples/s ({S/NV/1e6:.1f} M samples/frame)", flush=True)
    r.close()
