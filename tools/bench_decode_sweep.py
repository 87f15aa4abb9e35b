"""BASELINE.json configs[3]: decode-only throughput sweep over volume sizes on one GPU (the 2048^3 point needs
8 GPUs: bench.py --workload sortlast reports it).  Whole volumes, slab by slab, histograms generated on the device;
both output layouts (3-D arrays for the texture unit, 4x4x4-bricked planes) and both sources.
    python tools/bench_decode_sweep.py > profiles/decode_sweep_<tag>.json"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vrdd_b200 as V

PEAK = 6553.6
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def ev():
    return torch.cuda.Event(enable_timing=True)


def sweep_hist(edge, sink, reps=3):
    r = V.Renderer(0); r.set_stream(torch.cuda.current_stream().cuda_stream)
    if sink == "bricked":
        r.set_sampler(V.SAMPLER_BRICKED)
    r.set_volume(edge, edge, edge)
    slab = min(edge, max(1, (32 << 30) // (edge * edge * 128)))
    buf = torch.empty(slab * edge * edge * 32, dtype=torch.float32, device="cuda")
    ms = 0.0
    for z0 in range(0, edge, slab):
        nz = min(slab, edge - z0)
        r.synth_histograms_device(1234, z0, nz, buf); r.set_histograms_device(buf, z0, nz)
        r.decode(V.SRC_ORIGINAL, z0, nz); torch.cuda.synchronize()
        e0, e1 = ev(), ev(); e0.record()
        for _ in range(reps):
            r.decode(V.SRC_ORIGINAL, z0, nz)
        e1.record(); torch.cuda.synchronize()
        ms += e0.elapsed_time(e1) / reps
    r.close(); del buf; torch.cuda.empty_cache()
    n = edge ** 3
    return {"source": "raw histograms", "kernel": "decode_hist_tma_kernel", "edge": edge, "layout": sink, "ms": ms,
            "gbs": n * 140 / ms / 1e6, "frac_of_hbm_peak": n * 140 / ms / 1e6 / PEAK, "gvoxels_per_s": n / ms / 1e6,
            "input_gb": n * 128 / 1e9, "slab_z": slab}


def sweep_fractal(edge, reps=3):
    T, max_ne = 622, 8
    r = V.Renderer(0); r.set_stream(torch.cuda.current_stream().cuda_stream)
    r.set_volume(edge, edge, edge)
    slab = min(edge, 256)
    nvs = slab * edge * edge
    cb = torch.empty(nvs * 4, dtype=torch.int32, device="cuda")
    er = torch.empty(nvs * max_ne * 2, dtype=torch.float32, device="cuda")
    off = torch.empty(nvs // V.ERR_CHUNK + 1, dtype=torch.int64, device="cuda")
    tm = torch.empty(T * 32, dtype=torch.float32, device="cuda")
    ms, nbytes = 0.0, 0
    for z0 in range(0, edge, slab):
        tot = r.synth_fractal_device(1234, T, max_ne, z0, slab, cb, er, off, tm)
        r.set_fractal_device(cb, er, off, tm, T, z0, slab)
        r.decode(V.SRC_FRACTAL, z0, slab); torch.cuda.synchronize()
        e0, e1 = ev(), ev(); e0.record()
        for _ in range(reps):
            r.decode(V.SRC_FRACTAL, z0, slab)
        e1.record(); torch.cuda.synchronize()
        ms += e0.elapsed_time(e1) / reps
        nbytes += nvs * 28 + tot * 8
    r.close()
    n = edge ** 3
    return {"source": "fractal codes", "kernel": "decode_fractal_moments2_kernel", "edge": edge, "layout": "texture",
            "ms": ms, "gbs": nbytes / ms / 1e6, "frac_of_hbm_peak": nbytes / ms / 1e6 / PEAK, "gvoxels_per_s": n / ms / 1e6,
            "bytes_per_voxel": nbytes / n}


for edge in (256, 512, 1024):
    for sink in ("texture", "bricked"):
        print(json.dumps(sweep_hist(edge, sink)), flush=True)
    print(json.dumps(sweep_fractal(edge)), flush=True)
