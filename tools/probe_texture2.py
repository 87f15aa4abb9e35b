"""Second texture probe: one-hot volumes give the eight hardware trilinear weights directly."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vrdd_b200 as V

def sample(dims, vol, uvw):
    r = V.Renderer(0); r.keep_linear_planes(True); r.set_volume(*dims)
    planes = r.get_decoded_planes_device(V.SRC_ORIGINAL)
    n = dims[0]*dims[1]*dims[2]
    for p in planes: V.as_torch(p, (n,)).copy_(torch.from_numpy(vol.astype(np.float32).ravel()).cuda())
    r.commit_planes(V.SRC_ORIGINAL, 0, dims[2])
    d_uvw = torch.from_numpy(np.ascontiguousarray(uvw, np.float32)).cuda()
    d_out = torch.empty(uvw.shape[0], dtype=torch.float32, device="cuda")
    r.debug_sample_texture(V.SRC_ORIGINAL, 0, d_uvw, uvw.shape[0], d_out); r.synchronize()
    o = d_out.cpu().numpy(); r.close(); return o

out = {}
rng = np.random.default_rng(1)
# interior of a 2x2x2: coordinates between texel centres (0.25..0.75)
uvw = rng.uniform(0.25, 0.75, (20000, 3)).astype(np.float32)
# plus structured: two axes pinned at centres, one swept
sw = np.linspace(0.25, 0.75, 4097).astype(np.float32)
for ax in range(3):
    s = np.full((4097, 3), 0.25, np.float32); s[:, ax] = sw; uvw = np.concatenate([uvw, s])
# plus two axes swept diagonally
s = np.stack([sw, sw, np.full_like(sw, 0.25)], 1); uvw = np.concatenate([uvw, s])
s = np.stack([sw, sw, sw], 1); uvw = np.concatenate([uvw, s])
out["uvw"] = uvw
for c in range(8):
    vol = np.zeros((2, 2, 2), np.float32); vol[(c >> 2) & 1, (c >> 1) & 1, c & 1] = 1.0   # [z][y][x]
    out[f"w{c}"] = sample((2, 2, 2), vol, uvw)
# 2-D case: 2x2x1
uv2 = uvw.copy(); uv2[:, 2] = 0.5
for c in range(4):
    vol = np.zeros((1, 2, 2), np.float32); vol[0, (c >> 1) & 1, c & 1] = 1.0
    out[f"v{c}"] = sample((2, 2, 1), vol, uv2)
# value precision: 2x1x1 with values (0, big/small)
for name, vals in (("p1", (0.0, 1.0)), ("p3", (1.0, 1.0 + 2**-20)), ("p4", (0.1234567, 0.7654321))):
    vol = np.array(vals, np.float32).reshape(1, 1, 2)
    u1 = np.stack([sw, np.full_like(sw, 0.5), np.full_like(sw, 0.5)], 1)
    out[name] = sample((2, 1, 1), vol, u1)
out["sw"] = sw
np.savez_compressed("gpurun_out/texprobe2.npz", **out)
print("ok")
