"""Fourth texture probe: linear filtering with UN-normalised coordinates (flexBlockTex)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vrdd_b200 as V
out = {}
for N in (4, 11, 64, 500):
    r = V.Renderer(0); r.keep_linear_planes(True); r.set_volume(N, 2, 2)
    planes = r.get_decoded_planes_device(V.SRC_ORIGINAL)
    ramp = (torch.arange(N, dtype=torch.float32, device="cuda")).repeat(4)
    for p in planes: V.as_torch(p, (4 * N,)).copy_(ramp)
    r.commit_planes(V.SRC_ORIGINAL, 0, 2)
    rng = np.random.default_rng(N)
    x = np.concatenate([np.arange(-256, (N + 1) * 512) / 512.0, rng.uniform(-1, N + 1, 30000)]).astype(np.float32)
    x = np.concatenate([x, np.nextafter(x[:4000], np.float32(1e9)), np.nextafter(x[:4000], np.float32(-1e9))]).astype(np.float32)
    uvw = np.stack([x, np.full_like(x, 0.5), np.full_like(x, 0.5)], 1)
    d_uvw = torch.from_numpy(np.ascontiguousarray(uvw)).cuda()
    d_out = torch.empty(x.shape[0], dtype=torch.float32, device="cuda")
    r.debug_sample_texture_unnorm(V.SRC_ORIGINAL, 0, d_uvw, x.shape[0], d_out); r.synchronize()
    out[f"x_{N}"] = x; out[f"v_{N}"] = d_out.cpu().numpy()
    r.close()
np.savez_compressed("gpurun_out/texprobe4.npz", **out)
print("ok")
