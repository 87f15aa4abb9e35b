"""Fractal-decode kernel experiments: time per variant on one slab and cross-check of the variants
against the dense kernel (which follows the reference's order of operations).
    python tools/bench_fractal.py [edge] [nz] [variants...]"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vrdd_b200 as V

edge = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
nz = int(sys.argv[2]) if len(sys.argv) > 2 else 256
variants = sys.argv[3:] or ["dense", "moments_global", "moments", "moments768"]
T, max_ne = 622, 8
n = edge * edge * nz
dev = "cuda"
cb = torch.empty(n * 4, dtype=torch.int32, device=dev)
er = torch.empty(n * max_ne * 2, dtype=torch.float32, device=dev)
off = torch.empty((n + V.ERR_CHUNK - 1) // V.ERR_CHUNK + 1, dtype=torch.int64, device=dev)
tm = torch.empty(T * 32, dtype=torch.float32, device=dev)
ref = None
for sink in ("tex", "linear"):
    for variant in variants:
        r = V.Renderer(0)
        r.set_stream(torch.cuda.current_stream().cuda_stream)
        if sink == "linear":
            r.set_sampler(V.SAMPLER_LINEAR)
        r.set_volume(edge, edge, nz)
        name, *opts = variant.split("+")                      # e.g. moments2+pf12+generic
        r.set_variant("decode_fractal", name)
        for o in opts:
            if o.startswith("pf"):
                r.set_variant("decode_fractal_prefetch", o[2:])
            elif o in ("generic", "auto"):
                r.set_variant("decode_fractal_sink", o)
        tot = r.synth_fractal_device(1234, T, max_ne, 0, nz, cb, er, off, tm)
        r.set_fractal_device(cb, er, off, tm, T, 0, nz)
        r.decode(V.SRC_FRACTAL); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record()
        for _ in range(reps):
            r.decode(V.SRC_FRACTAL)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        nbytes = n * 28 + tot * 8
        line = f"edge {edge} nz {nz} sink {sink:6s} {variant:15s}: {ms:8.3f} ms  {nbytes / ms / 1e6:7.1f} GB/s  {nbytes / ms / 1e6 / 6553.6 * 100:5.1f}% of measured peak  {n / ms / 1e6:6.2f} Gvoxel/s"
        if sink == "linear":
            planes = r.get_decoded_planes_device(V.SRC_FRACTAL)
            got = torch.stack([V.as_torch(p, (n,), device=dev).clone() for p in planes])
            if ref is None:
                ref = got
            else:
                d = (got - ref).abs()
                tol = 2e-6 + 2e-5 * ref.abs()
                line += f"  max|diff vs {variants[0]}| {d.max().item():.3e}  outside tol: {(d > tol).sum().item()}"
        print(line, flush=True)
        r.close()
