"""Pins the oracle against the REFERENCE's own device code (oracle/_ref/ref_driver: volumeRender_kernel.cu compiled
where it lies behind oracle/ref_shim, run on the GPU with the real texture unit).

    python tools/ref_pin.py gen <dir>       seeded inputs at the reference's fixed configuration -> <dir>/in
    python tools/ref_pin.py run <dir>       oracle/_ref/ref_driver on them (GPU box)              -> <dir>/out
    python tools/ref_pin.py compare <dir>   oracle vs the reference's outputs; writes <dir>/ref_gpu_v1.npz
    python tools/ref_pin.py all <dir>       the three in a row
    python tools/ref_pin.py cuda [fixture]  the CUDA path against the fixture, ray set-up "source" and "nvcc" (GPU)

The fixture (tests/golden/ref_gpu_v1.npz, checked by tests/test_reference_pin.py without a GPU) holds the seeds, a
digest of the inputs and what the reference computed: both decoded volumes and the frames of queryMethod 1..7 for
two views.  Configuration: the reference's hard-wired 50x50x10 blocks x 32 bins, 622 templates
(volumeRender.cpp:86-90), its default render parameters (:129-133) and its 16x16 launch blocks."""
import hashlib
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

DIMS = (50, 50, 10)
T = 622
SEEDS = {"hist": 41, "templates": 5, "fractal": 6}
VIEWS = [(0.0, 0.0), (25.0, 40.0)]          # the self-test view (volumeRender.cpp:1024-1043) and an oblique one
IMAGE = (256, 256)


def inputs(o):
    hist = o.synth_histograms(SEEDS["hist"], DIMS)
    tmpl = o.synth_templates(SEEDS["templates"], T)
    cb, err = o.synth_fractal(SEEDS["fractal"], DIMS, T=T, max_ne=8)
    err = np.ascontiguousarray(err, np.float32)
    k = np.arange(err.shape[1])[None, :]
    err[k >= cb[:, 3:4]] = 0.0               # the reference leaves the entries beyond NE uninitialised: zeros here
    views = np.stack([o.view_matrix(*v) for v in VIEWS]).astype(np.float32)
    return hist, cb, tmpl, err, views


def as_the_reference_build_decodes(cb, tmpl):
    """Codes and templates under which the SOURCE'S algorithm gives what the reference's nvcc 12.9 build gives.

    fractalDecoding() returns a pointer to its local array `decoded` (volumeRender_kernel.cu:196-221), so the compiler
    may drop stores into it: in the PTX of d_basicDataProcessing the unflipped branch keeps the stores of
    original[0..22] only, the flipped branch those of the first 24 reversed elements (original[8..31]); the rest of
    `decoded` stays zero.  That is the same as decoding with template bins 23..31 (unflipped) resp. 0..7 (flipped)
    zeroed: a doubled template table, flipped voxels pointing into its second half."""
    T = tmpl.shape[0]
    a = tmpl.copy(); a[:, 23:] = 0.0
    b = tmpl.copy(); b[:, :8] = 0.0
    cb2 = cb.copy()
    cb2[cb[:, 2] != 0, 0] += T
    return cb2, np.concatenate([a, b]).astype(np.float32)


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def gen(d):
    from oracle.vrdd_oracle import Oracle
    o = Oracle()
    hist, cb, tmpl, err, views = inputs(o)
    os.makedirs(os.path.join(d, "in"), exist_ok=True)
    os.makedirs(os.path.join(d, "out"), exist_ok=True)
    hist.tofile(os.path.join(d, "in", "hist.f32"))
    cb.astype(np.int32).tofile(os.path.join(d, "in", "codebook.i32"))
    tmpl.tofile(os.path.join(d, "in", "templates.f32"))
    err.tofile(os.path.join(d, "in", "errors.f32"))
    views.tofile(os.path.join(d, "in", "views.f32"))
    print("inputs:", digest(hist, cb, tmpl, err, views))


def run(d):
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    r = subprocess.run([exe, d, str(IMAGE[0]), str(IMAGE[1]), str(len(VIEWS))], capture_output=True, text=True, timeout=600)
    tail = (r.stdout + r.stderr).strip().splitlines()
    print("\n".join(tail[-12:]))
    if r.returncode != 0:
        raise SystemExit(f"ref_driver failed with {r.returncode}")


def load_outputs(d):
    n = DIMS[0] * DIMS[1] * DIMS[2]
    orig = np.fromfile(os.path.join(d, "out", "original.f32"), np.float32).reshape(n, 4)
    frac = np.fromfile(os.path.join(d, "out", "fractal.f32"), np.float32).reshape(n, 4)
    imgs = np.zeros((len(VIEWS), 7, IMAGE[1], IMAGE[0]), np.uint32)
    for k in range(len(VIEWS)):
        for qm in range(1, 8):
            imgs[k, qm - 1] = np.fromfile(os.path.join(d, "out", f"img_v{k}_q{qm}.u32"), np.uint32).reshape(IMAGE[1], IMAGE[0])
    return orig, frac, imgs


def oracle_outputs(o):
    hist, cb, tmpl, err, views = inputs(o)
    orig = o.decode_hist(hist)
    frac, bad = o.decode_fractal(cb, err, tmpl)
    imgs = np.zeros((len(VIEWS), 7, IMAGE[1], IMAGE[0]), np.uint32)
    for k in range(len(VIEWS)):
        for qm in range(1, 7):
            imgs[k, qm - 1], _ = o.render(orig, DIMS, views[k], image=IMAGE, query_method=qm, vol_fractal4=frac)
        imgs[k, 6], _ = o.render_mode7(hist, DIMS, views[k], image=IMAGE)
    return orig, frac, imgs, digest(hist, cb, tmpl, err, views)


def report(ref, mine):
    """Prints the differences and returns (worst relative decode error, worst LSB error)."""
    r_orig, r_frac, r_imgs = ref
    m_orig, m_frac, m_imgs = mine[:3]
    worst_dec = 0.0
    for name, a, b in (("original", r_orig, m_orig), ("fractal", r_frac, m_frac)):
        for c, comp in enumerate(("mean", "variance", "entropy")):
            d = np.abs(a[:, c].astype(np.float64) - b[:, c])
            scale = np.maximum(np.abs(a[:, c]), 1e-30)
            print(f"decode {name:8s} {comp:8s}: max |diff| {d.max():.3e}  max rel {np.max(d / np.maximum(scale, 1e-3)):.3e}"
                  f"  range [{a[:, c].min():.4g}, {a[:, c].max():.4g}]  nan ref/oracle {int(np.isnan(a[:, c]).sum())}/{int(np.isnan(b[:, c]).sum())}")
            worst_dec = max(worst_dec, float(np.nanmax(d / np.maximum(scale, 1e-3))))
        print(f"decode {name:8s} w lane   : ref min/max {a[:, 3].min():.3g}/{a[:, 3].max():.3g}")
    worst_lsb = 0
    for k in range(r_imgs.shape[0]):
        for qm in range(1, 8):
            a = r_imgs[k, qm - 1].view(np.uint8).astype(np.int16)
            b = m_imgs[k, qm - 1].view(np.uint8).astype(np.int16)
            d = np.abs(a - b)
            hit = int((r_imgs[k, qm - 1] != 0).sum())
            print(f"view {k} queryMethod {qm}: max LSB diff {int(d.max())}  bytes off by >1: {int((d > 1).sum())}  by 1: {int((d == 1).sum())}"
                  f"  non-zero pixels ref/oracle {hit}/{int((m_imgs[k, qm - 1] != 0).sum())}")
            worst_lsb = max(worst_lsb, int(d.max()))
    return worst_dec, worst_lsb


def compare(d):
    from oracle.vrdd_oracle import Oracle
    o = Oracle()
    ref = load_outputs(d)
    mine = oracle_outputs(o)
    worst = report(ref, mine)
    out = os.path.join(d, "ref_gpu_v1.npz")
    np.savez_compressed(out, dims=np.array(DIMS), templates=np.array(T), image=np.array(IMAGE),
                        views=np.array(VIEWS, np.float32), seeds=np.array([SEEDS["hist"], SEEDS["templates"], SEEDS["fractal"]]),
                        inputs_sha256=np.array(mine[3]), original=ref[0], fractal=ref[1], images=ref[2])
    print("fixture:", out, os.path.getsize(out), "bytes; worst decode rel", worst[0], "worst LSB", worst[1])


def cuda(fixture):
    """The CUDA path (libvrdd.so, needs a GPU) against the fixture, for both roundings of the ray set-up."""
    import torch
    import vrdd_b200 as V
    from oracle.vrdd_oracle import Oracle
    o = Oracle()
    fx = dict(np.load(fixture))
    hist, cb, tmpl, err, views = inputs(o)
    assert digest(hist, cb, tmpl, err, views) == str(fx["inputs_sha256"])
    for setup in ("source", "nvcc"):
        r = V.Renderer(0)
        r.enable_interpolated_mean(True)
        r.set_variant("ray_setup", setup)
        r.set_volume(*DIMS); r.set_histograms_host(hist); r.decode(V.SRC_ORIGINAL)
        out = torch.zeros(IMAGE[1], IMAGE[0], dtype=torch.int32, device="cuda")
        for k in range(len(VIEWS)):
            r.set_view(views[k])
            for qm in (1, 2, 3, 7):
                out.zero_()
                r.render(out, IMAGE[0], IMAGE[1], V.default_render_params(query_method=qm)); r.synchronize()
                d = np.abs(out.cpu().numpy().view(np.uint8).astype(np.int16) - fx["images"][k, qm - 1].view(np.uint8).astype(np.int16))
                print(f"ray_setup {setup:6s} view {k} queryMethod {qm}: max LSB diff {int(d.max())}  bytes off by >1: {int((d > 1).sum())}  by 1: {int((d == 1).sum())}")
        r.close()


if __name__ == "__main__":
    if sys.argv[1] == "cuda":
        cuda(sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "tests", "golden", "ref_gpu_v1.npz"))
        raise SystemExit(0)
    what, d = sys.argv[1], os.path.abspath(sys.argv[2])
    if what in ("gen", "all"):
        gen(d)
    if what in ("run", "all"):
        run(d)
    if what in ("compare", "all"):
        compare(d)
