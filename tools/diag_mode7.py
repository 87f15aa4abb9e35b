"""queryMethod 7: the three fetch paths against the oracle and against one another at small sizes."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vrdd_b200 as V
import torch
from oracle.vrdd_oracle import Oracle
orc = Oracle()
def lsb(a, b):
    a = a.view(np.uint8).reshape(-1, 4).astype(np.int16); b = b.view(np.uint8).reshape(-1, 4).astype(np.int16)
    return np.abs(a - b)
for dims, img, rot in [((32, 16, 8), (256, 192), (20.0, 35.0)), ((64, 64, 4), (200, 200), (0.0, 0.0)), ((16, 32, 5), (160, 120), (-40.0, 100.0)), ((33, 16, 8), (160, 120), (10.0, 80.0)), ((32,16,8),(256,192),(0.0,0.0)), ((32,16,8),(256,256),(20.0,35.0))]:
    hist = orc.synth_histograms(23, dims)
    view = orc.view_matrix(*rot)
    out = torch.zeros(img[1], img[0], dtype=torch.int32, device="cuda")
    p = V.default_render_params(query_method=7)
    res = {}
    for var in ("gather", "texture", "linear"):                  # the array behind a path is chosen at decode time
        r = V.Renderer(0)
        r.enable_interpolated_mean(True)
        r.set_variant("raycast_mode7", var)
        r.set_volume(*dims); r.set_histograms_host(hist); r.decode(V.SRC_ORIGINAL)
        r.set_view(view)
        r.render(out, img[0], img[1], p, clear_misses=True); r.synchronize()
        res[var] = out.cpu().numpy().copy()
        r.close()
    ref, _ = orc.render_mode7(hist, dims, view, image=img)
    ref = np.ascontiguousarray(ref)
    for var in res:
        d = lsb(np.ascontiguousarray(res[var]), ref)
        print(dims, img, rot, var, "vs oracle max", d.max(), "n>1", int((d > 1).sum()), " vs texture equal", np.array_equal(res[var], res["texture"]), flush=True)
