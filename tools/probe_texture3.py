"""Third texture probe: POINT filtering of a normalised 3-D texture at coordinates on and around
texel boundaries (what mode 7 of the reference does with its block-index texture), and the
transfer-function texture at NaN / inf coordinates."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vrdd_b200 as V
out = {}
for N in (2, 3, 7, 10, 50, 64, 100, 1000, 1024):
    r = V.Renderer(0); r.keep_linear_planes(True); r.set_volume(N, 1, 1)
    planes = r.get_decoded_planes_device(V.SRC_ORIGINAL)
    for p in planes: V.as_torch(p, (N,)).copy_(torch.arange(N, dtype=torch.float32, device="cuda"))
    r.commit_planes(V.SRC_ORIGINAL, 0, 1)
    k = np.arange(0, N + 1)
    us = [np.float32(k) / np.float32(N)]                                   # exactly what floor(x*N)/N produces in fp32
    for d in (-3, -2, -1, 1, 2, 3):
        us.append(np.nextafter(us[0], np.float32(10 * d), dtype=np.float32) if abs(d) == 1 else us[0] + np.float32(d) * np.float32(2.0 ** -22))
    rng = np.random.default_rng(N)
    us.append(rng.uniform(-0.1, 1.1, 20000).astype(np.float32))
    u = np.concatenate(us).astype(np.float32)
    uvw = np.stack([u, np.full_like(u, 0.5), np.full_like(u, 0.5)], 1)
    d_uvw = torch.from_numpy(np.ascontiguousarray(uvw)).cuda()
    d_out = torch.empty(u.shape[0], dtype=torch.float32, device="cuda")
    r.debug_sample_texture_point(V.SRC_ORIGINAL, 0, d_uvw, u.shape[0], d_out); r.synchronize()
    out[f"u_{N}"] = u; out[f"pt_{N}"] = d_out.cpu().numpy()
    r.close()
r = V.Renderer(0)
u = np.array([np.nan, np.inf, -np.inf, 0.0, 1.0, 0.5], np.float32)
d_u = torch.from_numpy(u).cuda(); d_o = torch.empty(6, 4, dtype=torch.float32, device="cuda")
r.debug_sample_transfer_function(d_u, 6, d_o); r.synchronize()
out["tf_special_u"] = u; out["tf_special"] = d_o.cpu().numpy()
print(out["tf_special"])
# linear 3-D texture at NaN coordinate
r.keep_linear_planes(True); r.set_volume(4, 4, 4)
planes = r.get_decoded_planes_device(V.SRC_ORIGINAL)
for p in planes: V.as_torch(p, (64,)).copy_(torch.arange(64, dtype=torch.float32, device="cuda"))
r.commit_planes(V.SRC_ORIGINAL, 0, 4)
uvw = np.array([[np.nan, 0.5, 0.5], [0.5, np.nan, 0.5], [np.nan, np.nan, np.nan], [0.5, 0.5, 0.5]], np.float32)
d_uvw = torch.from_numpy(uvw).cuda(); d_o = torch.empty(4, dtype=torch.float32, device="cuda")
r.debug_sample_texture(V.SRC_ORIGINAL, 0, d_uvw, 4, d_o); r.synchronize()
print("linear3d NaN:", d_o.cpu().numpy())
r.debug_sample_texture_point(V.SRC_ORIGINAL, 0, d_uvw, 4, d_o); r.synchronize()
print("point3d NaN:", d_o.cpu().numpy())
np.savez_compressed("gpurun_out/texprobe3.npz", **out)
