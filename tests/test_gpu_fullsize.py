"""Parity at BASELINE.json's full sizes.

configs[1] (512^3 volume, 1024x1024 frames): the oracle still finishes in seconds, so the comparison is
direct — decoded slices against the oracle's decode of the same synthetic voxels, and a full 1024x1024
oblique orbit view against the oracle's ray caster within +-1 LSB.
The headline size (1024^3, 137 GB of histograms; configs[2] renders it at 2048x2048 in image-space tiles): the
oracle cannot hold the histograms, but it can decode single z-slices of the same synthetic voxels and ray cast a
band of image rows on ONE decoded plane (4.3 GB) — decoded slices against the oracle's, 256-row bands of
1024x1024 and 2048x2048 frames (whole and assembled from four ranks' tiles) against the oracle's ray caster within
+-1 LSB; and the size-independent properties: the decode does not depend on how the volume is cut into slabs, the
image does not depend on the image-space partition nor on how many march steps are in flight.
configs[4] (sort-last): 2x2x2 bricks of 256^3 (+ ghost) of a 512^3 volume, ranks emulated on one GPU, against the
oracle on the whole volume."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL, ATOL = 2e-5, 2e-6
SEED = 1234


def _lsb_diff(a, b):
    a8 = np.ascontiguousarray(a).view(np.uint8).astype(np.int16)
    b8 = np.ascontiguousarray(b).view(np.uint8).astype(np.int16)
    return np.abs(a8 - b8)


def _decode_synthetic(r, V, edge, slab):
    import torch
    r.set_volume(edge, edge, edge)
    buf = torch.empty(slab * edge * edge * 32, dtype=torch.float32, device="cuda")
    for z0 in range(0, edge, slab):
        nz = min(slab, edge - z0)
        r.synth_histograms_device(SEED, z0, nz, buf)
        r.set_histograms_device(buf, z0, nz)
        r.decode(V.SRC_ORIGINAL, z0, nz)
    r.synchronize()
    del buf
    torch.cuda.empty_cache()


def _render(r, V, w, h, part=None, **params):
    import torch
    out = torch.zeros(h, w, dtype=torch.int32, device="cuda")
    r.render(out, w, h, V.default_render_params(**params), part=part, clear_misses=True)
    r.synchronize()
    return out.cpu().numpy().view(np.uint32)


@pytest.fixture(scope="module")
def vol512():
    import vrdd_b200 as V
    r = V.Renderer(0)
    _decode_synthetic(r, V, 512, 128)
    dec = r.get_decoded_host(V.SRC_ORIGINAL, np.empty((512 ** 3, 4), np.float32))
    yield r, dec
    r.close()


def test_decode_512_slices_match_the_oracle(vol512, oracle):
    _, dec = vol512
    dims, sl = (512, 512, 512), 512 * 512
    for z in (0, 137, 511):
        hist = oracle.synth_histograms(SEED, dims, z0=z, nz=1)
        np.testing.assert_allclose(dec[z * sl:(z + 1) * sl], oracle.decode_hist(hist), rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("rot_y,qm", [(50.625, 1), (0.0, 3)])
def test_render_512_volume_1024_frame_matches_the_oracle(vol512, oracle, rot_y, qm):
    """configs[1]: one view of the 64-view orbit (k = 9) and the frontal view, full 1024x1024 frames."""
    import vrdd_b200 as V
    r, dec = vol512
    view = oracle.view_matrix(0.0, rot_y)
    r.set_view(view)
    r.count_samples(True)
    got = _render(r, V, 1024, 1024, query_method=qm)
    s_got = r.get_sample_count()
    r.count_samples(False)
    ref, s_ref = oracle.render(dec, (512, 512, 512), view, image=(1024, 1024), query_method=qm)
    assert abs(s_got - s_ref) <= max(2, int(1e-4 * s_ref))
    d = _lsb_diff(got, ref)
    assert d.max() <= 1, (int(d.max()), int((d > 1).sum()))
    assert (ref != 0).mean() > 0.3


def test_fractal_decode_512_slices_match_the_oracle(oracle):
    """The fractal half at configs[1]'s size: codes generated on the device slab by slab (compact round-major
    errors), decoded with the shared-memory moments kernel, three slices against the oracle's dense decode."""
    import torch
    import vrdd_b200 as V
    E, T, max_ne, slab = 512, 622, 8, 128
    sl = E * E
    r = V.Renderer(0)
    r.set_volume(E, E, E)
    nvs = slab * sl
    cb = torch.empty(nvs * 4, dtype=torch.int32, device="cuda")
    er = torch.empty(nvs * max_ne * 2, dtype=torch.float32, device="cuda")
    off = torch.empty(nvs // V.ERR_CHUNK + 1, dtype=torch.int64, device="cuda")
    tm = torch.empty(T * 32, dtype=torch.float32, device="cuda")
    for z0 in range(0, E, slab):
        r.synth_fractal_device(SEED, T, max_ne, z0, slab, cb, er, off, tm)
        r.set_fractal_device(cb, er, off, tm, T, z0, slab)
        r.decode(V.SRC_FRACTAL, z0, slab)
    dec = r.get_decoded_host(V.SRC_FRACTAL, np.empty((E ** 3, 4), np.float32))
    r.close()
    tmpl = oracle.synth_templates(SEED, T)
    for z in (0, 300, 511):
        c, e = oracle.synth_fractal(SEED, (E, E, E), T=T, max_ne=max_ne, z0=z, nz=1)
        ref, bad = oracle.decode_fractal(c, e, tmpl)
        assert bad == 0
        np.testing.assert_allclose(dec[z * sl:(z + 1) * sl], ref, rtol=RTOL, atol=ATOL)


@pytest.fixture(scope="module")
def vol1024():
    """The headline volume, decoded once (slab 128) with the linear planes kept next to the sampled arrays."""
    import vrdd_b200 as V
    r = V.Renderer(0)
    r.keep_linear_planes(True)
    _decode_synthetic(r, V, 1024, 128)
    yield r
    r.close()


@pytest.fixture(scope="module")
def mean1024(vol1024):
    """The decoded mean plane of the headline volume on the host (4.3 GB): what the oracle's ray caster samples."""
    import vrdd_b200 as V
    return V.as_torch(vol1024.get_decoded_planes_device(V.SRC_ORIGINAL)[0], (1024 ** 3,)).cpu().numpy()


def test_decode_1024_slices_match_the_oracle(vol1024, oracle):
    import vrdd_b200 as V
    E = 1024
    sl = E * E
    planes = [V.as_torch(p, (E * sl,)) for p in vol1024.get_decoded_planes_device(V.SRC_ORIGINAL)]
    for z in (0, 517, 1023):
        ref = oracle.decode_hist(oracle.synth_histograms(SEED, (E, E, E), z0=z, nz=1))
        for c in range(3):
            got = planes[c][z * sl:(z + 1) * sl].cpu().numpy()
            np.testing.assert_allclose(got, ref[:, c], rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("img,rows", [(1024, (384, 640)), (2048, (896, 1152))])
@pytest.mark.parametrize("rot", [(0.0, 50.625), (0.0, 90.0)])
def test_render_1024_volume_bands_match_the_oracle(vol1024, mean1024, oracle, img, rows, rot):
    """The headline (1024^2) and configs[2] (2048^2, also assembled from the 64x64 tiles of four ranks): a 256-row
    band through the middle of the frame against the oracle's d_render on the decoded mean plane."""
    import vrdd_b200 as V
    E = 1024
    r = vol1024
    mean = mean1024
    view = oracle.view_matrix(*rot)
    r.set_view(view)
    got = _render(r, V, img, img, query_method=1)
    ref, s_ref = oracle.render(None, (E, E, E), view, image=(img, img), query_method=1, plane=mean, rows=rows)
    band = slice(rows[0], rows[1])
    d = _lsb_diff(got[band], ref[band])
    assert d.max() <= 1, (int(d.max()), int((d > 1).sum()))
    assert (ref[band] != 0).mean() > 0.4 and s_ref > 4e6
    if img == 2048:
        acc = np.zeros_like(got)
        for part in range(4):
            acc |= _render(r, V, img, img, part=V.TilePartition(64, 64, part, 4), query_method=1)
        assert np.array_equal(acc, got)


def test_sortlast_bricks_of_256_match_the_oracle(vol512, oracle):
    """configs[4] at a size the oracle holds: the 512^3 volume as 2x2x2 bricks of 256^3 (+ one ghost layer), ranks
    emulated on one GPU (tests/test_gpu_sortlast.py), 1024x1024 frames; a 256-row band against the oracle on the
    whole decoded volume, the full frame against the single-volume render."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_gpu_sortlast import _render_bricks
    import vrdd_b200 as V
    import vrdd_b200.dist as D
    r, dec = vol512
    gdims, grid, img, rows = (512, 512, 512), (2, 2, 2), (1024, 1024), (384, 640)
    view = oracle.view_matrix(20.0, 50.625)
    params = V.default_render_params(query_method=1)
    got, s = _render_bricks(V, D, oracle, gdims, grid, view, img, params, SEED, None, V.SAMPLER_TEXTURE)
    ref, _ = oracle.render(dec, gdims, view, image=img, query_method=1, rows=rows)
    band = slice(rows[0], rows[1])
    d = _lsb_diff(got[band], ref[band])
    assert d.max() <= 1, (int(d.max()), int((d > 1).sum()))
    assert (ref[band] != 0).mean() > 0.4
    r.set_view(view)
    r.count_samples(True)
    one = _render(r, V, img[0], img[1], query_method=1)
    s_one = r.get_sample_count()
    r.count_samples(False)
    assert _lsb_diff(got, one).max() <= 1
    assert abs(s - s_one) <= max(2, s_one // 5000), (s, s_one)


def test_1024_volume_properties():
    """The headline size.  Slab size, image-space partition and march batching must not change one bit."""
    import vrdd_b200 as V
    views = [V.view_matrix(0.0, 9 * 5.625), V.view_matrix(20.0, 90.0)]
    imgs = {}
    for slab in (128, 192):                                      # 192 does not divide 1024: ragged last slab
        r = V.Renderer(0)
        _decode_synthetic(r, V, 1024, slab)
        out = []
        for view in views:
            r.set_view(view)
            out.append(_render(r, V, 1024, 1024, query_method=1))
        imgs[slab] = out
        if slab == 128:
            r.set_view(views[0])
            whole = out[0]
            assert (whole != 0).mean() > 0.3
            # image-space tiles of 4 ranks compose to the whole frame (each pixel by exactly one rank)
            acc = np.zeros_like(whole)
            for part in range(4):
                tile = _render(r, V, 1024, 1024, part=V.TilePartition(64, 64, part, 4), query_method=1)
                assert not (acc & tile).any()
                acc |= tile
            assert np.array_equal(acc, whole)
            # 1, 2 or 8 march steps in flight instead of 4
            for u in ("1", "2", "8"):
                r.set_variant("raycast_unroll", u)
                assert np.array_equal(_render(r, V, 1024, 1024, query_method=1), whole)
            r.set_variant("raycast_unroll", "4")
        r.close()
    for a, b in zip(imgs[128], imgs[192]):
        assert np.array_equal(a, b)
