"""Decode-kernel experiments: GB/s of the raw-histogram decode vs slab size, variant, tile order
and output sink.  python tools/bench_decode.py [edge] """
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vrdd_b200 as V

edge = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
def run(nz, variant, order, sink, reps=3):
    r = V.Renderer(0)
    r.set_stream(torch.cuda.current_stream().cuda_stream)
    if sink == "brick": r.set_sampler(V.SAMPLER_BRICKED)
    if sink == "linear+tex": r.keep_linear_planes(True)
    r.set_volume(edge, edge, nz)
    r.set_variant("decode_hist", variant); r.set_variant("decode_order", order)
    n = edge * edge * nz
    buf = torch.empty(n * 32, dtype=torch.float32, device="cuda")
    r.synth_histograms_device(1, 0, nz, buf)
    r.set_histograms_device(buf, 0, nz)
    r.decode(V.SRC_ORIGINAL); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): r.decode(V.SRC_ORIGINAL)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    r.close(); del buf; torch.cuda.empty_cache()
    return ms, n * 140 / ms / 1e6
for nz in (16, 64, 128, 256):
    for variant, order in (("tma", "interleaved"), ("tma", "chunked"), ("ldg", "interleaved")):
        for sink in ("tex", "brick"):
            ms, gbs = run(nz, variant, order, sink)
            print(f"edge {edge} nz {nz:4d} ({edge*edge*nz*128/2**30:6.1f} GiB in) {variant:4s} {order:12s} sink {sink:6s}: {ms:8.3f} ms  {gbs:7.1f} GB/s  {gbs/6553.6*100:5.1f}% of measured peak", flush=True)
