"""Regenerates tests/golden/golden_v2.npz from the CPU oracle (oracle/vrdd_oracle.cpp, built with
-ffp-contract=off) in the rounding of the reference's own build (Oracle.set_reference_build: the FMA pattern of its
PTX and the B200's rsqrt.approx table; v1 was frozen in the source's uncontracted order).

The reference ships no data, no reference image and no unit tests, and cannot be compiled with
CUDA 12.9, so these vectors are NOT outputs of the reference: they are the
oracle's outputs on seeded synthetic inputs, frozen so that (a) the oracle is a checked fixed
point across hosts/compilers and (b) the GPU tests have inputs and answers that do not depend
on any generator code.  Run:  python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.vrdd_oracle import Oracle  # noqa: E402

DIMS = (12, 10, 8)
IMG = (64, 48)
SEED = 20261018
VIEWS = [(0.0, 0.0), (25.0, 40.0), (-35.0, 200.0)]


def main():
    o = Oracle()
    o.set_reference_build(True)
    out = {"dims": np.array(DIMS, np.int32), "img": np.array(IMG, np.int32), "seed": np.array([SEED], np.int64)}
    hist = o.synth_histograms(SEED, DIMS)
    tmpl = o.synth_templates(SEED, 37)
    cb, err = o.synth_fractal(SEED, DIMS, T=37, max_ne=8)
    # hand-made edge cases in the first voxels: shift == B (identity through the single wrap),
    # flip with shift, NE == 0, a clamp to zero, a repeated bin, an all-zero histogram
    cb[0] = (3, 32, 0, 0)
    cb[1] = (5, 31, 1, 2); err[1, 0] = (0.0, -0.9); err[1, 1] = (7.0, 0.25)
    cb[2] = (7, 1, 1, 3); err[2, 0] = (4.0, 0.1); err[2, 1] = (4.0, -0.5); err[2, 2] = (4.0, 0.2)
    hist[3] = 0.0
    hist[4] = 0.0; hist[4, 31] = 1.0
    out["hist"], out["templates"], out["codebook"], out["errors"] = hist, tmpl, cb, err
    dec_o = o.decode_hist(hist)
    dec_f, recon, bad = o.decode_fractal(cb, err, tmpl, want_recon=True)
    assert bad == 0
    out["decoded_original"], out["decoded_fractal"], out["recon"] = dec_o, dec_f, recon
    views = np.stack([o.view_matrix(rx, ry) for rx, ry in VIEWS])
    out["views"] = views
    imgs, counts = [], []
    for vi, view in enumerate(views):
        for qm in (1, 2, 3, 4, 5, 6):
            im, s = o.render(dec_o, DIMS, view, image=IMG, query_method=qm, vol_fractal4=dec_f)
            imgs.append(im); counts.append((vi, qm, s))
    out["images"] = np.stack(imgs)
    out["image_index"] = np.array(counts, np.int64)
    # non-default parameters: coarse step, low threshold, TF window, brightness
    im, s = o.render(dec_o, DIMS, views[1], image=IMG, query_method=1, density=0.2, brightness=1.7,
                     transfer_offset=0.1, transfer_scale=1.6, tstep=0.037, max_steps=40, opacity_threshold=0.6)
    out["image_params"] = im
    out["image_params_samples"] = np.array([s], np.int64)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v2.npz")
    np.savez_compressed(path, **out)
    h = hashlib.sha256(open(path, "rb").read()).hexdigest()
    print(path, os.path.getsize(path), "bytes sha256", h)


if __name__ == "__main__":
    main()
