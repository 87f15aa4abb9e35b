"""Probe the B200 texture unit's linear filter: sample a ramp T[i] = i (3-D array N x 1 x 1,
linear / normalised / clamp) at finely spaced u and dump hardware outputs, so the oracle's
filter model can be fitted to the hardware rather than to the programming guide's prose."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vrdd_b200 as V  # noqa: E402

out = {}
for N in (2, 3, 7, 50, 256, 1024):
    r = V.Renderer(0)
    r.keep_linear_planes(True)
    r.set_volume(N, 1, 1)
    planes = r.get_decoded_planes_device(V.SRC_ORIGINAL)
    assert all(planes)
    V.as_torch(planes[0], (N,)).copy_(torch.arange(N, dtype=torch.float32, device="cuda"))
    V.as_torch(planes[1], (N,)).copy_(torch.arange(N, dtype=torch.float32, device="cuda") ** 2)
    V.as_torch(planes[2], (N,)).zero_()
    r.commit_planes(V.SRC_ORIGINAL, 0, 1)
    # sweep: texel pitch around a few texels with 4096 sub-steps per texel
    sub = 4096
    texels = [0, 1, N // 2, N - 2] if N > 3 else list(range(N))
    us = []
    for t in texels:
        us.append((t + np.arange(-sub // 2, sub + sub // 2) / sub) / N)
    u = np.concatenate(us).astype(np.float32)
    rng = np.random.default_rng(N)
    u = np.concatenate([u, rng.uniform(-0.05, 1.05, 20000).astype(np.float32)])
    uvw = np.stack([u, np.full_like(u, 0.5), np.full_like(u, 0.5)], 1)
    d_uvw = torch.from_numpy(np.ascontiguousarray(uvw)).cuda()
    d_out = torch.empty(u.shape[0], dtype=torch.float32, device="cuda")
    r.debug_sample_texture(V.SRC_ORIGINAL, 0, d_uvw, u.shape[0], d_out)
    r.synchronize()
    out[f"u_{N}"] = u
    out[f"ramp_{N}"] = d_out.cpu().numpy()
    r.debug_sample_texture(V.SRC_ORIGINAL, 1, d_uvw, u.shape[0], d_out)
    r.synchronize()
    out[f"sq_{N}"] = d_out.cpu().numpy()
    r.close()

# 3-D: random texels, random coordinates, for fitting the combination order
dims = (5, 4, 3)
r = V.Renderer(0)
r.keep_linear_planes(True)
r.set_volume(*dims)
planes = r.get_decoded_planes_device(V.SRC_ORIGINAL)
n = dims[0] * dims[1] * dims[2]
rng = np.random.default_rng(0)
tex = rng.random(n).astype(np.float32)
for p in planes:
    V.as_torch(p, (n,)).copy_(torch.from_numpy(tex).cuda())
r.commit_planes(V.SRC_ORIGINAL, 0, dims[2])
uvw = rng.uniform(0.0, 1.0, (50000, 3)).astype(np.float32)
d_uvw = torch.from_numpy(uvw).cuda()
d_out = torch.empty(uvw.shape[0], dtype=torch.float32, device="cuda")
r.debug_sample_texture(V.SRC_ORIGINAL, 0, d_uvw, uvw.shape[0], d_out)
r.synchronize()
out["tex3"] = tex; out["uvw3"] = uvw; out["hw3"] = d_out.cpu().numpy()
r.close()
os.makedirs("gpurun_out", exist_ok=True)
np.savez_compressed("gpurun_out/texprobe.npz", **out)
print("saved", {k: v.shape for k, v in out.items()})
