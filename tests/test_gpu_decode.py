"""GPU parity tests of P1 (decode) through the C ABI, against the CPU oracle and the golden
vectors.  Tolerances: the decode kernels compute in fp32 with MUFU.LG2 where the reference
promotes single terms to double (see csrc/decode_hist.cu); decoded values are O(1) after the
reference's normalisation, and the bound below is ~10 fp32 ulps of that scale:
        |gpu - oracle| <= 2e-6 + 2e-5 * |oracle|
The integer part of the fractal decode (template row, flip/shift permutation, error placement
and order, clamp) is compared BIT-EXACTLY through the reconstructed histogram."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL, ATOL = 2e-5, 2e-6


def _decode_hist(r, V, dims, hist, variant):
    r.set_volume(*dims)
    r.set_variant("decode_hist", variant)
    r.set_histograms_host(hist)
    r.decode(V.SRC_ORIGINAL)
    return r.get_decoded_host(V.SRC_ORIGINAL, np.empty((hist.shape[0], 4), np.float32))


@pytest.mark.parametrize("variant", ["tma", "ldg"])
def test_hist_decode_matches_golden(renderer, golden, variant):
    import vrdd_b200 as V
    dims = tuple(int(v) for v in golden["dims"])
    got = _decode_hist(renderer, V, dims, golden["hist"], variant)
    np.testing.assert_allclose(got, golden["decoded_original"], rtol=RTOL, atol=ATOL)
    assert np.all(got[:, 3] == 0)
    assert np.all(got[3] == 0)                       # the all-zero histogram (p <= 0 branch)


@pytest.mark.parametrize("variant", ["tma", "ldg"])
@pytest.mark.parametrize("dims", [(1, 1, 1), (5, 3, 2), (31, 17, 3), (64, 8, 1), (33, 16, 1), (50, 50, 10), (64, 64, 48)])
def test_hist_decode_matches_oracle_ragged_sizes(renderer, oracle, variant, dims):
    """Sizes around the 512-voxel tile / 32-voxel warp-tile boundaries, the reference's own
    50x50x10, and a multi-tile volume."""
    import vrdd_b200 as V
    hist = oracle.synth_histograms(99, dims)
    got = _decode_hist(renderer, V, dims, hist, variant)
    np.testing.assert_allclose(got, oracle.decode_hist(hist), rtol=RTOL, atol=ATOL)


def test_hist_decode_random_and_degenerate_rows(renderer, oracle):
    import vrdd_b200 as V
    rng = np.random.default_rng(3)
    dims = (40, 13, 3)
    n = dims[0] * dims[1] * dims[2]
    hist = rng.random((n, 32)).astype(np.float32)
    hist[hist < 0.5] = 0
    hist[5] = 0
    hist /= np.maximum(hist.sum(1, keepdims=True), 1e-30)
    hist[7] = 0; hist[7, 0] = 1
    hist[8] = 0; hist[8, 31] = 1
    hist[9] = 1e-30                                  # denormal-scale frequencies
    for variant in ("tma", "ldg"):
        got = _decode_hist(renderer, V, dims, hist, variant)
        np.testing.assert_allclose(got, oracle.decode_hist(hist), rtol=RTOL, atol=ATOL)


def test_slab_decode_from_device_memory(renderer, oracle):
    """Slab-wise streaming (vrdd_set_histograms_device) writes the same volume as one decode."""
    import torch
    import vrdd_b200 as V
    dims = (24, 20, 9)
    hist = oracle.synth_histograms(5, dims)
    r = renderer
    r.set_volume(*dims)
    slice_vox = dims[0] * dims[1]
    for z0, nz in ((0, 4), (4, 3), (7, 2)):
        d = torch.from_numpy(hist[z0 * slice_vox:(z0 + nz) * slice_vox]).cuda()
        r.set_histograms_device(d, z0, nz)
        r.decode(V.SRC_ORIGINAL, z0, nz)
        r.synchronize()
    got = r.get_decoded_host(V.SRC_ORIGINAL, np.empty((hist.shape[0], 4), np.float32))
    np.testing.assert_allclose(got, oracle.decode_hist(hist), rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("sampler", ["texture", "bricked"])
def test_decoded_layouts_agree(renderer, oracle, sampler):
    """cudaArray planes, bricked planes and linear planes all hold the same decoded values."""
    import vrdd_b200 as V
    dims = (21, 10, 7)                               # not a multiple of the 4^3 brick
    hist = oracle.synth_histograms(8, dims)
    r = renderer
    r.set_sampler(V.SAMPLER_TEXTURE if sampler == "texture" else V.SAMPLER_BRICKED)
    r.keep_linear_planes(True)
    r.set_volume(*dims)
    r.set_histograms_host(hist)
    r.decode(V.SRC_ORIGINAL)
    n = hist.shape[0]
    ref = oracle.decode_hist(hist)
    planes = r.get_decoded_planes_device(V.SRC_ORIGINAL)
    assert all(planes)
    r.synchronize()
    for c, p in enumerate(planes):
        np.testing.assert_allclose(_from_device_ptr(p, n), ref[:, c], rtol=RTOL, atol=ATOL)
    # drop the linear planes: get_decoded_host now reads the sampler layout itself
    r2 = V.Renderer(0)
    r2.set_sampler(V.SAMPLER_TEXTURE if sampler == "texture" else V.SAMPLER_BRICKED)
    r2.set_volume(*dims)
    r2.set_histograms_host(hist)
    r2.decode(V.SRC_ORIGINAL)
    got = r2.get_decoded_host(V.SRC_ORIGINAL, np.empty((n, 4), np.float32))
    r2.close()
    np.testing.assert_allclose(got, ref, rtol=RTOL, atol=ATOL)


def _from_device_ptr(ptr, n):
    """n floats behind a raw device pointer, as numpy."""
    import vrdd_b200 as V
    return V.as_torch(ptr, (n,)).cpu().numpy()


@pytest.mark.parametrize("variant", ["moments", "moments2", "moments2r", "moments2b", "moments2br", "moments_global", "dense"])
def test_fractal_decode_matches_golden_and_is_bit_exact_in_its_integer_part(renderer, golden, variant):
    """The golden codes include shift == 32, flip + shift, NE == 0, a clamp to zero and a bin hit
    three times (the moments variant must fall back to the ordered dense route there)."""
    import torch
    import vrdd_b200 as V
    dims = tuple(int(v) for v in golden["dims"])
    n = dims[0] * dims[1] * dims[2]
    r = renderer
    r.set_variant("decode_fractal", variant)
    r.set_volume(*dims)
    r.set_fractal_host(golden["codebook"], golden["errors"], golden["templates"])
    recon = torch.empty(n, 32, dtype=torch.float32, device="cuda")
    r.reconstruct_fractal_device(recon)
    r.synchronize()
    assert np.array_equal(recon.cpu().numpy(), golden["recon"])          # BIT-exact
    r.decode(V.SRC_FRACTAL)
    got = r.get_decoded_host(V.SRC_FRACTAL, np.empty((n, 4), np.float32))
    np.testing.assert_allclose(got, golden["decoded_fractal"], rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("variant", ["moments", "moments2", "moments2r", "moments2b", "moments2br", "moments768", "moments_global", "dense"])
@pytest.mark.parametrize("dims,T,max_ne", [((1, 1, 1), 3, 8), ((30, 9, 2), 622, 8), ((50, 50, 10), 622, 8),
                                            ((64, 32, 5), 100, 32), ((16, 16, 4), 1500, 0), ((96, 64, 40), 622, 8)])
def test_fractal_decode_matches_oracle(renderer, oracle, dims, T, max_ne, variant):
    """Ragged sizes, the reference's 50x50x10 / 622 templates, NE up to 32, NE == 0, a template table too
    large for shared memory, and a volume of more than one sweep of the persistent grid with rows that are
    a multiple of 32 voxels (the incremental-coordinate path of the shared-table kernel)."""
    import torch
    import vrdd_b200 as V
    tmpl = oracle.synth_templates(4, T)
    cb, err = oracle.synth_fractal(4, dims, T=T, max_ne=max_ne)
    rng = np.random.default_rng(11)
    cb[:, 1] = rng.integers(0, 33, cb.shape[0])                       # every legal shift, 32 included
    if max_ne >= 2:                                                   # a few voxels hit one bin twice
        for v in range(0, cb.shape[0], 37):
            if cb[v, 3] >= 2:
                err[v, 1, 0] = err[v, 0, 0]
    ref, ref_recon, bad = oracle.decode_fractal(cb, err, tmpl, want_recon=True)
    assert bad == 0
    n = cb.shape[0]
    r = renderer
    r.set_variant("decode_fractal", variant)
    r.set_volume(*dims)
    r.set_fractal_host(cb, err, tmpl)
    recon = torch.empty(n, 32, dtype=torch.float32, device="cuda")
    r.reconstruct_fractal_device(recon)
    r.synchronize()
    assert np.array_equal(recon.cpu().numpy(), ref_recon)
    r.decode(V.SRC_FRACTAL)
    got = r.get_decoded_host(V.SRC_FRACTAL, np.empty((n, 4), np.float32))
    np.testing.assert_allclose(got, ref, rtol=RTOL, atol=ATOL)


def test_fractal_slabs_from_packed_device_arrays(renderer, oracle):
    """vrdd_pack_fractal_errors + vrdd_set_fractal_device, one z-slab at a time (slab boundaries fall on chunk
    boundaries because W*H is a multiple of 32), equals the whole-volume answer."""
    import torch
    import vrdd_b200 as V
    dims, T = (64, 48, 12), 622
    tmpl = oracle.synth_templates(9, T)
    cb, err = oracle.synth_fractal(9, dims, T=T, max_ne=8)
    ref, _ = oracle.decode_fractal(cb, err, tmpl)
    sl = dims[0] * dims[1]
    r = renderer
    r.set_volume(*dims)
    d_tm = torch.from_numpy(tmpl).cuda()
    keep = []
    for z0, nz in ((0, 5), (5, 4), (9, 3)):
        ent, off = V.pack_fractal_errors(cb[z0 * sl:(z0 + nz) * sl], err[z0 * sl:(z0 + nz) * sl])
        d_cb = torch.from_numpy(np.ascontiguousarray(cb[z0 * sl:(z0 + nz) * sl])).cuda()
        d_er = torch.from_numpy(ent.view(np.uint8).reshape(-1).copy()).cuda()
        d_off = torch.from_numpy(off.view(np.int64)).cuda()
        keep += [d_cb, d_er, d_off]
        r.set_fractal_device(d_cb, d_er, d_off, d_tm, T, z0, nz)
        r.decode(V.SRC_FRACTAL, z0, nz)
    n = sl * dims[2]
    got = r.get_decoded_host(V.SRC_FRACTAL, np.empty((n, 4), np.float32))
    np.testing.assert_allclose(got, ref, rtol=RTOL, atol=ATOL)


def test_fractal_upload_applies_the_reference_guards(renderer, oracle):
    """volumeRender_kernel.cu:781-816 / volumeRender.cpp:611-614: out-of-range codes are an error
    here (VRDD_ERR_RANGE), not a printf."""
    import vrdd_b200 as V
    dims = (4, 4, 2)
    tmpl = oracle.synth_templates(1, 10)
    cb, err = oracle.synth_fractal(1, dims, T=10)
    r = renderer
    r.set_volume(*dims)
    for field, bad in ((0, 10), (0, -1), (1, 33), (3, 33), (3, -2)):
        c2 = cb.copy(); c2[5, field] = bad
        with pytest.raises(V.VrddError) as e:
            r.set_fractal_host(c2, err, tmpl)
        assert e.value.code == V.ERR_RANGE
    c2 = cb.copy(); c2[5, 3] = 1
    e2 = err.copy(); e2[5, 0, 0] = 32.0
    with pytest.raises(V.VrddError) as e:
        r.set_fractal_host(c2, e2, tmpl)
    assert e.value.code == V.ERR_RANGE
    t2 = tmpl.copy(); t2[3, 3] = 1.5
    with pytest.raises(V.VrddError):
        r.set_fractal_host(cb, err, t2)


def test_call_order_errors(renderer):
    import vrdd_b200 as V
    with pytest.raises(V.VrddError) as e:
        renderer.decode(V.SRC_ORIGINAL)
    assert e.value.code == V.ERR_INVALID
    with pytest.raises(V.VrddError) as e:
        renderer.set_volume(8, 8, 8, bins=64)
    assert e.value.code == V.ERR_UNSUPPORTED
    renderer.set_volume(8, 8, 8)
    with pytest.raises(V.VrddError):
        renderer.decode(V.SRC_FRACTAL)


def test_device_synth_is_bit_identical_to_host_synth(renderer, oracle):
    """include/vrdd_synth.h gives the same bits under nvcc (device) and g++ (host)."""
    import torch
    dims, seed, T = (37, 21, 9), 1234, 622
    r = renderer
    r.set_volume(*dims)
    n_slab = dims[0] * dims[1] * 4
    d = torch.empty(n_slab, 32, dtype=torch.float32, device="cuda")
    r.synth_histograms_device(seed, 3, 4, d)
    r.synchronize()
    assert np.array_equal(d.cpu().numpy(), oracle.synth_histograms(seed, dims, z0=3, nz=4))
    n = dims[0] * dims[1] * dims[2]
    import vrdd_b200 as V
    nch = (n + V.ERR_CHUNK - 1) // V.ERR_CHUNK
    cb = torch.empty(n, 4, dtype=torch.int32, device="cuda")
    er = torch.empty(n * 8, 2, dtype=torch.float32, device="cuda")
    off = torch.empty(nch + 1, dtype=torch.int64, device="cuda")
    tm = torch.empty(T, 32, dtype=torch.float32, device="cuda")
    tot = r.synth_fractal_device(seed, T, 8, 0, dims[2], cb, er, off, tm)
    hcb, herr = oracle.synth_fractal(seed, dims, T=T, max_ne=8)
    assert np.array_equal(cb.cpu().numpy(), hcb)
    assert np.array_equal(tm.cpu().numpy(), oracle.synth_templates(seed, T))
    assert tot == int(hcb[:, 3].sum())
    from packing import round_major
    want = round_major(hcb, herr)
    got_e = er.cpu().numpy()[:tot]
    assert np.array_equal(got_e[:, 0].view(np.int32), want["bin"]) and np.array_equal(got_e[:, 1], want["value"])
    offs = off.cpu().numpy()
    cum = np.concatenate([[0], np.cumsum(hcb[:, 3])])
    assert np.array_equal(offs[:-1], cum[0:n:V.ERR_CHUNK]) and offs[-1] == tot
    # and the compact device form decodes to the oracle's answer
    r.set_fractal_device(cb, er, off, tm, T, 0, dims[2])
    r.decode(V.SRC_FRACTAL)
    got = r.get_decoded_host(V.SRC_FRACTAL, np.empty((n, 4), np.float32))
    ref, _ = oracle.decode_fractal(hcb, herr, oracle.synth_templates(seed, T))
    np.testing.assert_allclose(got, ref, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("variant", ["moments", "moments2", "moments2r", "moments_global"])
def test_fractal_errors_that_cancel_the_template(renderer, oracle, variant):
    """Errors that remove all, or all but a thousandth, of a template's mass: the corrected sum of the moment table is
    then rounding noise (about 1e-7 where the reference's tot is exactly 0 and it writes 0, 0, 0), so such voxels must
    take the in-order dense route.  Templates with 3, 8 and 32 occupied bins; every voxel of the volume is one case."""
    import vrdd_b200 as V
    rng = np.random.default_rng(5)
    T, dims = 6, (32, 4, 2)
    n = dims[0] * dims[1] * dims[2]
    tmpl = np.zeros((T, 32), np.float32)
    for t, k in enumerate((3, 3, 8, 8, 32, 32)):
        bins = rng.choice(32, k, replace=False)
        w = rng.random(k).astype(np.float32) + 0.1
        tmpl[t, bins] = (w / w.sum()).astype(np.float32)
    cb = np.zeros((n, 4), np.int32)
    err = np.zeros((n, 32, 2), np.float32)
    for v in range(n):
        t, shift, flip = v % T, int(rng.integers(0, 33)), int(rng.integers(0, 2))
        src = tmpl[t][::-1] if flip else tmpl[t]
        cur = np.roll(src, shift % 32)
        nzb = np.nonzero(cur)[0]
        kind = v % 3
        keep = nzb[0]
        k = 0
        for b in nzb:
            if kind == 0:
                val = -cur[b]                                        # exactly cancelled: tot == 0 in the reference
            elif kind == 1:
                val = -np.float32(1.5) * cur[b]                      # over-cancelled, clamped to 0: tot == 0
            else:
                val = -cur[b] if b != keep else -np.float32(0.999) * cur[b]      # 99.9 % of the bin, the rest gone
            err[v, k] = (b, val); k += 1
        cb[v] = (t, shift, flip, k)
    ref, bad = oracle.decode_fractal(cb, err, tmpl)
    assert bad == 0
    assert (ref[0::3, :3] == 0).all() and (ref[1::3, :3] == 0).all()           # the reference's all-zero histogram
    r = renderer
    r.set_variant("decode_fractal", variant)
    r.set_volume(*dims)
    r.set_fractal_host(cb, err, tmpl)
    r.decode(V.SRC_FRACTAL)
    got = r.get_decoded_host(V.SRC_FRACTAL, np.empty((n, 4), np.float32))
    assert (got[0::3, :3] == 0).all() and (got[1::3, :3] == 0).all()
    np.testing.assert_allclose(got, ref, rtol=RTOL, atol=ATOL)


def test_templates_rewritten_in_place_are_picked_up(renderer, oracle):
    """vrdd_set_fractal_device with the SAME template pointer after the table was rewritten in place: the moment table
    derived from it is rebuilt on every call (it used to be keyed on the address)."""
    import torch
    import vrdd_b200 as V
    dims, T = (32, 8, 4), 50
    n = dims[0] * dims[1] * dims[2]
    cb, err = oracle.synth_fractal(21, dims, T=T, max_ne=8)
    ent, off = V.pack_fractal_errors(cb, err)
    d_cb = torch.from_numpy(cb).cuda()
    d_er = torch.from_numpy(ent.view(np.uint8).reshape(-1).copy()).cuda()
    d_off = torch.from_numpy(off.view(np.int64)).cuda()
    d_tm = torch.empty(T, 32, dtype=torch.float32, device="cuda")
    r = renderer
    r.set_volume(*dims)
    for seed in (3, 4):
        tmpl = oracle.synth_templates(seed, T)
        d_tm.copy_(torch.from_numpy(tmpl))                            # same address, new contents
        torch.cuda.synchronize()
        r.set_fractal_device(d_cb, d_er, d_off, d_tm, T, 0, dims[2])
        r.decode(V.SRC_FRACTAL)
        got = r.get_decoded_host(V.SRC_FRACTAL, np.empty((n, 4), np.float32))
        ref, _ = oracle.decode_fractal(cb, err, tmpl)
        np.testing.assert_allclose(got, ref, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("source", ["original", "fractal"])
def test_decode_stores_into_the_peers_planes(oracle, source):
    """vrdd_set_peer_planes: two handles play two ranks, each decodes its z-slab and the decode kernel itself stores
    every value into the other's linear planes as well (over NVLink on a node; the same pointers here); after
    committing the received slab both hold the whole volume — the all-gather costs no pass of its own."""
    import torch
    import vrdd_b200 as V
    dims, T = (64, 32, 12), 60
    sl = dims[0] * dims[1]
    n = sl * dims[2]
    hist = oracle.synth_histograms(13, dims)
    cb, err = oracle.synth_fractal(13, dims, T=T)
    tmpl = oracle.synth_templates(13, T)
    ref = oracle.decode_hist(hist) if source == "original" else oracle.decode_fractal(cb, err, tmpl)[0]
    src = V.SRC_ORIGINAL if source == "original" else V.SRC_FRACTAL
    slabs = ((0, 7), (7, 5))
    ranks, keep = [], []
    for k in range(2):
        r = V.Renderer(0)
        r.keep_linear_planes(True)
        r.set_volume(*dims)
        ranks.append(r)
    planes = [r.get_decoded_planes_device(src) for r in ranks]
    for k, r in enumerate(ranks):
        z0, nz = slabs[k]
        other = planes[1 - k]
        r.set_peer_planes(src, [(other[0], None, other[2])] if k == 0 else [other])       # rank 0 skips the variance for now
        if source == "original":
            d = torch.from_numpy(hist[z0 * sl:(z0 + nz) * sl]).cuda(); keep.append(d)
            r.set_histograms_device(d, z0, nz)
        else:
            ent, off = V.pack_fractal_errors(cb[z0 * sl:(z0 + nz) * sl], err[z0 * sl:(z0 + nz) * sl])
            d_cb = torch.from_numpy(np.ascontiguousarray(cb[z0 * sl:(z0 + nz) * sl])).cuda()
            d_er = torch.from_numpy(ent.view(np.uint8).reshape(-1).copy()).cuda()
            d_off = torch.from_numpy(off.view(np.int64)).cuda()
            d_tm = torch.from_numpy(tmpl).cuda()
            keep += [d_cb, d_er, d_off, d_tm]
            r.set_fractal_device(d_cb, d_er, d_off, d_tm, T, z0, nz)
        r.decode(src, z0, nz)
    for r in ranks:
        r.synchronize()
    # rank 1 received mean and entropy of slab 0 (variance later); rank 0 received all of slab 1
    ranks[1].commit_planes(src, slabs[0][0], slabs[0][1], plane_mask=5)
    ranks[0].commit_planes(src, slabs[1][0], slabs[1][1])
    got0 = ranks[0].get_decoded_host(src, np.empty((n, 4), np.float32))
    np.testing.assert_allclose(got0, ref, rtol=RTOL, atol=ATOL)
    # the arrays the ray caster samples hold the same: a frame from either rank is the single-volume frame
    view = oracle.view_matrix(15.0, 35.0)
    frames = []
    for r in ranks:
        r.set_view(view)
        out = torch.zeros(96, 128, dtype=torch.int32, device="cuda")
        r.render(out, 128, 96, V.default_render_params(query_method=3 if source == "original" else 4), clear_misses=True)
        r.synchronize()
        frames.append(out.cpu().numpy())
    assert np.array_equal(frames[0], frames[1])
    # the variance of slab 0 follows: decode again with all three planes mapped
    ranks[0].set_peer_planes(src, [planes[1]])
    ranks[0].decode(src, slabs[0][0], slabs[0][1]); ranks[0].synchronize()
    ranks[1].commit_planes(src, slabs[0][0], slabs[0][1], plane_mask=2)
    got1 = ranks[1].get_decoded_host(src, np.empty((n, 4), np.float32))
    np.testing.assert_allclose(got1, ref, rtol=RTOL, atol=ATOL)
    for r in ranks:
        r.close()


def test_tma_decode_is_reproducible():
    """The bulk-copy ring of decode_hist_tma_kernel: a slot may only be refilled once every warp has its rows in
    registers.  Decoding the same 1024x1024x128 slab forty times must give the same bits forty times (a release of the
    slot that did not wait for the shared-memory loads gave torn rows in a handful of voxels per 6 G decoded)."""
    import torch
    import vrdd_b200 as V
    W, H, nz = 1024, 1024, 128
    r = V.Renderer(0)
    r.keep_linear_planes(True)
    r.set_volume(W, H, nz)
    buf = torch.empty(nz * W * H * 32, dtype=torch.float32, device="cuda")
    r.synth_histograms_device(77, 0, nz, buf)
    r.set_histograms_device(buf, 0, nz)
    planes = [V.as_torch(p, (W * H * nz,)) for p in r.get_decoded_planes_device(V.SRC_ORIGINAL)]
    r.decode(V.SRC_ORIGINAL); r.synchronize()
    ref = [p.clone() for p in planes]
    r.set_variant("decode_hist", "ldg")                       # an independent kernel agrees to rounding
    r.decode(V.SRC_ORIGINAL); r.synchronize()
    for c in range(3):
        assert float((planes[c] - ref[c]).abs().max()) < 1e-6
    r.set_variant("decode_hist", "tma")
    for it in range(40):
        r.decode(V.SRC_ORIGINAL); r.synchronize()
        for c in range(3):
            assert torch.equal(planes[c].view(torch.int32), ref[c].view(torch.int32)), (it, c)
    r.close()


def test_mean_raw_plane_is_replicated_for_query_method_7(oracle):
    """vrdd_set_peer_mean_raw: the un-normalised block means of queryMethod 7 travel like the three planes; both
    'ranks' then render the same queryMethod-7 frame as one handle that decoded everything, tiles included."""
    import torch
    import vrdd_b200 as V
    dims, (w, h) = (32, 32, 16), (160, 128)
    sl = dims[0] * dims[1]
    hist = oracle.synth_histograms(21, dims)
    slabs = ((0, 9), (9, 7))
    one = V.Renderer(0)
    one.enable_interpolated_mean(True)
    one.set_volume(*dims); one.set_histograms_host(hist); one.decode(V.SRC_ORIGINAL)
    view = oracle.view_matrix(20.0, 40.0)
    p7 = V.default_render_params(query_method=7)

    def frame(r, part=None):
        out = torch.zeros(h, w, dtype=torch.int32, device="cuda")
        r.set_view(view); r.render(out, w, h, p7, part=part, clear_misses=True); r.synchronize()
        return out.cpu().numpy()
    want = frame(one)
    ranks, keep = [], []
    for k in range(2):
        r = V.Renderer(0)
        r.keep_linear_planes(True); r.enable_interpolated_mean(True)
        r.set_volume(*dims)
        ranks.append(r)
    raws = [r.get_mean_raw_device() for r in ranks]
    for k, r in enumerate(ranks):
        z0, nz = slabs[k]
        d = torch.from_numpy(hist[z0 * sl:(z0 + nz) * sl]).cuda(); keep.append(d)
        r.set_histograms_device(d, z0, nz)
        r.set_peer_planes(V.SRC_ORIGINAL, [[None, None, None]])
        r.set_peer_mean_raw([raws[1 - k]])
        r.decode(V.SRC_ORIGINAL, z0, nz)
    for r in ranks:
        r.synchronize()
    for k, r in enumerate(ranks):
        r.commit_mean_raw(*slabs[1 - k])
    acc = np.zeros_like(want)
    for k, r in enumerate(ranks):
        assert np.array_equal(frame(r), want), k
        acc |= frame(r, part=V.TilePartition(32, 32, k, 2))
    assert np.array_equal(acc, want)
    for r in ranks + [one]:
        r.close()
