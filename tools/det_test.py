import os, sys, torch
sys.path.insert(0, '/root/repo')
import vrdd_b200 as V
vol = 1024
r = V.Renderer(0); r.set_stream(torch.cuda.current_stream().cuda_stream)
r.keep_linear_planes(True)
r.set_volume(vol, vol, vol)
sl = vol * vol
nz = 512
buf = torch.empty(nz * sl * 32, dtype=torch.float32, device="cuda")
r.synth_histograms_device(1234, 0, nz, buf); r.set_histograms_device(buf, 0, nz)
planes = [V.as_torch(p, (vol ** 3,)) for p in r.get_decoded_planes_device(V.SRC_ORIGINAL)]
ref = None
for it in range(12):
    r.decode(V.SRC_ORIGINAL, 0, nz); r.synchronize()
    cur = [p[:nz * sl].clone() for p in planes]
    if ref is None:
        ref = cur
    else:
        for c in range(3):
            ne = (cur[c].view(torch.int32) != ref[c].view(torch.int32))
            n = int(ne.sum())
            if n:
                idx = torch.nonzero(ne)[:5].flatten().tolist()
                print("iter", it, "plane", c, "differs in", n, "voxels, first", idx, [(float(cur[c][i]), float(ref[c][i])) for i in idx[:3]])
print("determinism test done")
# also: ldg variant vs tma bitwise
r.set_variant("decode_hist", "ldg"); r.decode(V.SRC_ORIGINAL, 0, nz); r.synchronize()
for c in range(3):
    d = (planes[c][:nz * sl] - ref[c]).abs().max()
    print("ldg vs tma plane", c, "max abs diff", float(d))
