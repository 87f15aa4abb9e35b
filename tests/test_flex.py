"""The flexible-block-size query chain: oracle restatement against a direct count on CPU;
GPU kernels (csrc/flex.cu) against the oracle; queryMethod 8/9/0 rendering; the four remaining loader formats;
the legacy initCuda / dataProcessing path with the table sizes the reference hard-codes."""
import struct
import sys
import os

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import flex_synth as F  # noqa: E402


@pytest.fixture(scope="module")
def store16():
    return F.make_tables(3, 16)


# ---- CPU: the oracle's chain -------------------------------------------------------------------------------

@pytest.mark.parametrize("block", [6, 4, 5, 16, 7])
def test_oracle_chain_matches_direct_count(oracle, store16, block):
    """Lossless synthetic codes: the chain's block statistics must equal those counted directly from the raw
    volume over the region the reference's corner/sign quirks really cover (tests/flex_synth.py)."""
    out, dims, missing = oracle.flex_process(store16, block)
    nb = (16 + block - 1) // block
    assert dims == (nb, nb, nb) and missing == 0
    exp = F.expected_blocks(store16, block)
    np.testing.assert_allclose(out[:, :3], exp[:, :3], rtol=5e-6, atol=1e-4)
    assert np.all(out[:, 3] == 0)


def test_oracle_counts_missing_spans(oracle, store16):
    t = dict(store16)
    keep = np.ones(t["span_low"].shape[0], bool); keep[::7] = False
    for k in ("span_low", "span_high", "codebook", "errors"):
        t[k] = store16[k][keep]
    _, _, missing = oracle.flex_process(t, 6)
    assert missing > 0


# ---- CPU: the four remaining loader formats (volumeRender.cpp:709-997) --------------------------------------

def test_flex_formats_follow_the_loader_layout(tmp_path, store16):
    import vrdd_b200 as V
    L = V.lib()
    t = store16
    nf, ns = t["span_low"].shape[0], t["simple_low"].shape[0]
    f = lambda s: str(tmp_path / s).encode()
    # span list, written byte by byte: n, {lowX, highX, lowY, highY, lowZ, highZ} (:744-749)
    b = bytearray(struct.pack("<i", nf))
    for lo, hi in zip(t["span_low"], t["span_high"]):
        b += struct.pack("<6i", lo[0], hi[0], lo[1], hi[1], lo[2], hi[2])
    (tmp_path / "spanList.bin").write_bytes(bytes(b))
    assert L.vrdd_io_span_count(f("spanList.bin")) == nf
    lo2 = np.empty((nf, 4), np.int32); hi2 = np.empty((nf, 4), np.int32)
    assert L.vrdd_io_read_span_list(f("spanList.bin"), nf, lo2.ctypes.data, hi2.ctypes.data) == 0
    assert np.array_equal(lo2, t["span_low"]) and np.array_equal(hi2, t["span_high"])
    # flexible codebook: nTimeSteps, n, {spanId, templateId, shift, bool, NE, int ids[NE], double vals[NE]} (:784-870)
    b = bytearray(struct.pack("<ii", 1, nf))
    for i in range(nf):
        tid, sh, fl, ne = (int(x) for x in t["codebook"][i])
        b += struct.pack("<iii?i", i, tid, sh, bool(fl), ne)
        b += struct.pack(f"<{ne}i", *[int(x) for x in t["errors"][i, :ne, 0]])
        b += struct.pack(f"<{ne}d", *[float(x) for x in t["errors"][i, :ne, 1]])
    (tmp_path / "codebook0.bin").write_bytes(bytes(b))
    assert L.vrdd_io_codebook_blocks(f("codebook0.bin")) == nf
    ids = np.empty(nf, np.int32); cb = np.empty((nf, 4), np.int32); er = np.empty((nf, 64, 2), np.float32)
    assert L.vrdd_io_read_flex_codebook(f("codebook0.bin"), 64, nf, ids.ctypes.data, cb.ctypes.data, er.ctypes.data) == 0
    assert np.array_equal(ids, np.arange(nf)) and np.array_equal(cb, t["codebook"]) and np.array_equal(er, t["errors"])
    # simple histograms: counts {low xyz, high xyz, count}, bin ids, frequencies (:893-935)
    bc = bytearray(struct.pack("<i", ns)); bi = bytearray(); bf = bytearray()
    for i in range(ns):
        lo, hi, c = t["simple_low"][i], t["simple_high"][i], int(t["simple_count"][i])
        bc += struct.pack("<7i", lo[0], lo[1], lo[2], hi[0], hi[1], hi[2], c)
        bi += struct.pack(f"<{c}i", *[int(x) for x in t["simple_hist"][i, :c, 0]])
        bf += struct.pack(f"<{c}d", *[float(x) for x in t["simple_hist"][i, :c, 1]])
    (tmp_path / "nzbCounts0.bin").write_bytes(bytes(bc)); (tmp_path / "nzbBinIds0.bin").write_bytes(bytes(bi))
    (tmp_path / "nzbFreqs0.bin").write_bytes(bytes(bf))
    assert L.vrdd_io_simple_count(f("nzbCounts0.bin")) == ns
    sl = np.empty((ns, 4), np.int32); sh_ = np.empty((ns, 4), np.int32); sc = np.empty(ns, np.int32); shh = np.empty((ns, 64, 2), np.float32)
    assert L.vrdd_io_read_simple(f("nzbCounts0.bin"), f("nzbBinIds0.bin"), f("nzbFreqs0.bin"), 64, ns, sl.ctypes.data,
                                 sh_.ctypes.data, sc.ctypes.data, shh.ctypes.data) == 0
    assert np.array_equal(sl, t["simple_low"]) and np.array_equal(sh_, t["simple_high"]) and np.array_equal(sc, t["simple_count"])
    assert np.array_equal(shh, t["simple_hist"])
    # writers round-trip
    assert L.vrdd_io_write_span_list(f("s2"), nf, t["span_low"].ctypes.data, t["span_high"].ctypes.data) == 0
    assert (tmp_path / "s2").read_bytes() == (tmp_path / "spanList.bin").read_bytes()
    assert L.vrdd_io_write_flex_codebook(f("c2"), 64, nf, ids.ctypes.data, cb.ctypes.data, er.ctypes.data) == 0
    assert (tmp_path / "c2").read_bytes() == (tmp_path / "codebook0.bin").read_bytes()
    assert L.vrdd_io_write_simple(f("a"), f("b"), f("c"), 64, ns, sl.ctypes.data, sh_.ctypes.data, sc.ctypes.data, shh.ctypes.data) == 0
    assert (tmp_path / "a").read_bytes() == bytes(bc) and (tmp_path / "b").read_bytes() == bytes(bi) and (tmp_path / "c").read_bytes() == bytes(bf)
    # flexible templates share the template format with 64 bins (:951-997)
    assert L.vrdd_io_write_templates(f("t"), 64, t["templates"].shape[0], t["templates"].ctypes.data) == 0
    t2 = np.empty_like(t["templates"])
    assert L.vrdd_io_template_count(f("t"), 64) == t["templates"].shape[0]
    assert L.vrdd_io_read_templates(f("t"), 64, t["templates"].shape[0], t2.ctypes.data) == 0 and np.array_equal(t2, t["templates"])


# ---- GPU -----------------------------------------------------------------------------------------------------

gpu = pytest.mark.gpu


@gpu
@pytest.mark.parametrize("block", [6, 4, 5, 16])
def test_gpu_chain_matches_oracle(renderer, oracle, store16, block):
    ref, dims, _ = oracle.flex_process(store16, block)
    renderer.flex_set_tables_host(store16)
    assert renderer.flex_process(block) == 0
    got, gdims = renderer.flex_get_blocks_host()
    assert gdims == dims
    # corner sums are reduced with warp shuffles instead of the reference's unordered atomics: rounding only
    np.testing.assert_allclose(got, ref, rtol=2e-5, atol=2e-4)


@gpu
def test_gpu_chain_reference_configuration(renderer, oracle):
    """The reference's own configuration: 64^3 raw volume, block size 6 (volumeRender_kernel.cu:104, 1737)."""
    t = F.make_tables(5, 64, n_templates=469, block=6)
    ref, dims, _ = oracle.flex_process(t, 6)
    renderer.flex_set_tables_host(t)
    assert renderer.flex_process(6) == 0
    got, gdims = renderer.flex_get_blocks_host()
    assert gdims == dims == (11, 11, 11)
    np.testing.assert_allclose(got, ref, rtol=2e-5, atol=2e-4)
    np.testing.assert_allclose(got[:, :3], F.expected_blocks(t, 6)[:, :3], rtol=2e-5, atol=2e-4)   # and the direct count


@gpu
def test_gpu_chain_counts_missing_spans_like_the_oracle(renderer, oracle, store16):
    t = dict(store16)
    keep = np.ones(t["span_low"].shape[0], bool); keep[::7] = False
    for k in ("span_low", "span_high", "codebook", "errors"):
        t[k] = store16[k][keep]
    keep = np.ones(t["simple_low"].shape[0], bool); keep[::5] = False
    for k in ("simple_low", "simple_high", "simple_count", "simple_hist"):
        t[k] = store16[k][keep]
    ref, _, missing = oracle.flex_process(t, 5)
    renderer.flex_set_tables_host(t)
    assert renderer.flex_process(5) == missing > 0
    got, _ = renderer.flex_get_blocks_host()
    np.testing.assert_allclose(got, ref, rtol=2e-5, atol=2e-4)


@gpu
def test_flex_table_guards(renderer, store16):
    import vrdd_b200 as V
    for field, row, col, bad in (("codebook", 3, 0, 99), ("codebook", 3, 1, 65), ("codebook", 3, 3, 65), ("simple_count", 2, None, 70)):
        t = dict(store16); t[field] = store16[field].copy()
        if col is None:
            t[field][row] = bad
        else:
            t[field][row, col] = bad
        with pytest.raises(V.VrddError) as e:
            renderer.flex_set_tables_host(t)
        assert e.value.code == V.ERR_RANGE
    with pytest.raises(V.VrddError):
        renderer.flex_process(6)                                     # nothing valid was installed


@gpu
@pytest.mark.parametrize("qm,scale", [(8, 1.0), (9, 1.0 / 255.0), (0, 1.0 / 3000.0)])
def test_render_flexible_block_modes(renderer, oracle, store16, qm, scale):
    """queryMethod 8 / 9 / 0 (volumeRender_kernel.cu:654-680): un-normalised linear sampling of the block volume
    inside its zero-filled 500^3 array.  Mean and variance are not normalised by the reference, so the
    transfer-function scale brings them into range (the ',' and '.' keys of the reference, volumeRender.cpp:364-372)."""
    import torch
    import vrdd_b200 as V
    blocks, dims, _ = oracle.flex_process(store16, 3)
    r = renderer
    r.flex_set_tables_host(store16)
    r.flex_process(3)
    w, h = 160, 120
    r.count_samples(True)
    for rot in ((0.0, 0.0), (25.0, 40.0)):
        view = oracle.view_matrix(*rot)
        r.set_view(view)
        ref, s = oracle.render_flex(blocks, dims, view, image=(w, h), query_method=qm, transfer_scale=scale)
        out = torch.zeros(h, w, dtype=torch.int32, device="cuda")
        r.render(out, w, h, V.default_render_params(query_method=qm, transfer_scale=scale))
        assert abs(r.get_sample_count() - s) <= max(2, s // 10000)
        d = np.abs(out.cpu().numpy().view(np.uint8).reshape(h, w, 4).astype(np.int16) - ref.view(np.uint8).reshape(h, w, 4).astype(np.int16))
        assert d.max() <= 1, (qm, rot, int(d.max()))
        assert (ref != 0).mean() > 0.1


@gpu
def test_legacy_initcuda_with_flexible_tables(oracle):
    """initCuda's last nine arguments with the sizes the reference hard-codes (131 072 spans, 469 templates,
    volumeRender_kernel.cu:96-101), dataProcessing() with its block size 6, render_kernel with queryMethod 8."""
    import torch
    import vrdd_b200 as V
    L = V.legacy
    t = F.make_tables(5, 64, n_templates=469, block=6)
    N = 64 * 64 * 32

    def pad(a, fill):
        out = np.full((N,) + a.shape[1:], fill, a.dtype); out[:a.shape[0]] = a; return out
    arrs = [pad(t["span_low"], -1), pad(t["span_high"], -1), pad(t["codebook"], 0), pad(t["errors"], 0),
            pad(t["simple_low"], -1), pad(t["simple_high"], -1), pad(t["simple_count"], 0), pad(t["simple_hist"], 0), t["templates"]]
    dims = (50, 50, 10)
    hist = oracle.synth_histograms(1, dims)
    w = h = 256
    V.lib().initCuda(hist.ctypes.data, V.Extent(*dims), V.Extent(32, 2500, 10), None, V.Extent(*dims), None, V.Extent(32, 0, 1), None,
                     V.Extent(32, 2500, 10), *[a.ctypes.data for a in arrs])
    L.dataProcessing()
    L.basicDataProcessing()
    ref_blocks, bdims, _ = oracle.flex_process(t, 6)
    view = V.view_matrix(10.0, 30.0)
    L.copyInvViewMatrix(view)
    d_out = torch.zeros(h, w, dtype=torch.int32, device="cuda")
    L.render_kernel(d_out, w, h, query_method=8, volume_size=dims)
    torch.cuda.synchronize()
    ref, _ = oracle.render_flex(ref_blocks, bdims, np.array(view, np.float32), image=(w, h), query_method=8)
    d = np.abs(d_out.cpu().numpy().view(np.uint8).reshape(h, w, 4).astype(np.int16) - ref.view(np.uint8).reshape(h, w, 4).astype(np.int16))
    assert d.max() <= 1 and (ref != 0).mean() > 0.1
    L.freeCudaBuffers()
