"""Round-2 preparation: pins the flexible-block chain (dataProcessing(), queryMethod 8/9/0)
against the REFERENCE's own device code, the way tools/ref_pin.py pins the two hot paths.

    python tools/ref_pin_flex.py gen <dir>       synthetic lossless span store (tests/flex_synth.py) of a 64^3 raw
                                                 volume for block size 6, padded to the sizes initCuda hard-codes
                                                 (131 072 spans, 64 entries each, 469 templates) -> <dir>/in/flex_*
    python tools/ref_pin_flex.py run <dir>       oracle/_ref/ref_driver <dir> ... flex     (GPU box; the reference scans
                                                 its span tables linearly per thread: minutes, not milliseconds)
    python tools/ref_pin_flex.py compare <dir>   oracle (flex_process, render_flex) vs what the reference computed;
                                                 writes <dir>/ref_gpu_flex_v1.npz
The raw inputs of tools/ref_pin.py are written as well (the driver always loads them).  The 134 MB of padded tables
are regenerated on the box; keep <dir> outside gpurun_out/ and copy only <dir>/out and the .npz back.

NOT RUN YET (written after the GPU budget of round 1 was spent).  Expect the same kind of finding as in the fixed
path: flexibleFractalDecoding() also returns a pointer to a local array (volumeRender_kernel.cu:224-251)."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_pin as R

NSPAN, NENT, NTMPL, BINS, RAW, BLOCK, SEED = 64 * 64 * 32, 64, 469, 64, 64, 6, 17
IMAGE = (256, 256)


def tables():
    import flex_synth
    return flex_synth.make_tables(SEED, RAW, block=BLOCK)


def _pad(a, n, fill):
    out = np.full((n,) + a.shape[1:], fill, a.dtype)
    assert a.shape[0] <= n, (a.shape, n)
    out[:a.shape[0]] = a
    return out


def gen(d):
    R.gen(d)
    t = tables()
    w = lambda name, a: np.ascontiguousarray(a).tofile(os.path.join(d, "in", name))
    # padding spans can never match: the chain asks for 1-based boxes >= 1 (fractal) or 0-based boxes >= 0 (simple)
    w("flex_span_low.i32", _pad(t["span_low"], NSPAN, -1)); w("flex_span_high.i32", _pad(t["span_high"], NSPAN, -1))
    w("flex_codebook.i32", _pad(t["codebook"], NSPAN, 0)); w("flex_errors.f32", _pad(t["errors"], NSPAN, 0.0))
    w("flex_simple_low.i32", _pad(t["simple_low"], NSPAN, -1)); w("flex_simple_high.i32", _pad(t["simple_high"], NSPAN, -1))
    w("flex_simple_count.i32", _pad(t["simple_count"], NSPAN, 0)); w("flex_simple_hist.f32", _pad(t["simple_hist"], NSPAN, 0.0))
    w("flex_templates.f32", _pad(t["templates"], NTMPL, 0.0))
    print("flex tables:", t["span_low"].shape[0], "fractal spans,", t["simple_low"].shape[0], "simple spans,", t["templates"].shape[0], "templates")


def run(d):
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    r = subprocess.run([exe, d, str(IMAGE[0]), str(IMAGE[1]), str(len(R.VIEWS)), "flex"], capture_output=True, text=True, timeout=3000)
    print("\n".join((r.stdout + r.stderr).strip().splitlines()[-20:]))
    if r.returncode != 0:
        raise SystemExit(f"ref_driver failed with {r.returncode}")


def compare(d):
    from oracle.vrdd_oracle import Oracle
    o = Oracle()
    t = tables()
    dims = np.fromfile(os.path.join(d, "out", "flex_dims.i32"), np.int32)
    ref_blocks = np.fromfile(os.path.join(d, "out", "flex_blocks.f32"), np.float32).reshape(-1, 4)
    mine, nb, missing = o.flex_process(t, BLOCK)
    print("blocks: reference", dims.tolist(), " oracle", nb, " spans the oracle did not find:", missing)
    n = min(len(ref_blocks), len(mine))
    for c, name in enumerate(("mean", "variance", "entropy")):
        dd = np.abs(ref_blocks[:n, c].astype(np.float64) - mine[:n, c])
        print(f"flex {name:8s}: max |diff| {dd.max():.3e}  max rel {np.max(dd / np.maximum(np.abs(ref_blocks[:n, c]), 1e-3)):.3e}  "
              f"blocks off by > 1e-4 rel: {int((dd / np.maximum(np.abs(ref_blocks[:n, c]), 1e-3) > 1e-4).sum())} of {n}")
    views = np.fromfile(os.path.join(d, "in", "views.f32"), np.float32).reshape(-1, 12)
    imgs = np.zeros((len(views), 3, IMAGE[1], IMAGE[0]), np.uint32)
    for k in range(len(views)):
        for j, qm in enumerate((8, 9, 0)):
            a = np.fromfile(os.path.join(d, "out", f"img_v{k}_q{qm}.u32"), np.uint32).reshape(IMAGE[1], IMAGE[0])
            imgs[k, j] = a
            b, _ = o.render_flex(mine, nb, views[k], image=IMAGE, query_method=qm)
            dd = np.abs(a.view(np.uint8).astype(np.int16) - b.view(np.uint8).astype(np.int16))
            print(f"view {k} queryMethod {qm}: max LSB diff {int(dd.max())}  bytes off by >1: {int((dd > 1).sum())}  by 1: {int((dd == 1).sum())}")
    out = os.path.join(d, "ref_gpu_flex_v1.npz")
    np.savez_compressed(out, seed=np.array(SEED), raw=np.array(RAW), block=np.array(BLOCK), dims=dims, blocks=ref_blocks, images=imgs,
                        image=np.array(IMAGE), views=views)
    print("fixture:", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    what, d = sys.argv[1], os.path.abspath(sys.argv[2])
    if what in ("gen", "all"):
        gen(d)
    if what in ("run", "all"):
        run(d)
    if what in ("compare", "all"):
        compare(d)
