"""The reference's on-disk formats (include/vrdd_io.h; volumeRender.cpp:538-691), on CPU: files
written byte by byte with struct (an independent statement of the layout) are read back by the
library's loaders, and the library's writers round-trip."""
import struct

import numpy as np
import pytest


@pytest.fixture()
def data(oracle):
    dims = (6, 5, 4)
    hist = oracle.synth_histograms(9, dims)
    tmpl = oracle.synth_templates(9, 11)
    cb, err = oracle.synth_fractal(9, dims, T=11)
    return dims, hist, tmpl, cb, err


def test_readers_follow_the_loader_layout(tmp_path, data):
    import vrdd_b200 as V
    L = V.lib()
    dims, hist, tmpl, cb, err = data
    n = hist.shape[0]
    p = tmp_path / "hist.bin"; p.write_bytes(hist.tobytes())                       # float32[V*B]      (:549-550)
    got = np.empty_like(hist)
    assert L.vrdd_io_read_histograms(str(p).encode(), n, 32, got.ctypes.data) == 0 and np.array_equal(got, hist)
    assert L.vrdd_io_read_histograms(str(p).encode(), n + 1, 32, np.empty((n + 1, 32), np.float32).ctypes.data) == V.ERR_INVALID
    # codebook: nSteps, nBlocks, {spanId, templateId, shift, bool flip, NE, int ids[NE], double vals[NE]}   (:569-637)
    b = bytearray(struct.pack("<ii", 1, n))
    for v in range(n):
        tid, sh, fl, ne = (int(x) for x in cb[v])
        b += struct.pack("<iii?i", v, tid, sh, bool(fl), ne)
        b += struct.pack(f"<{ne}i", *[int(x) for x in err[v, :ne, 0]])
        b += struct.pack(f"<{ne}d", *[float(x) for x in err[v, :ne, 1]])
    p = tmp_path / "codebook.bin"; p.write_bytes(bytes(b))
    assert L.vrdd_io_codebook_blocks(str(p).encode()) == n
    cb2 = np.empty_like(cb); err2 = np.empty_like(err)
    assert L.vrdd_io_read_codebook(str(p).encode(), 32, n, cb2.ctypes.data, err2.ctypes.data) == 0
    assert np.array_equal(cb2, cb) and np.array_equal(err2, err)
    # templates: n, {double limits[6], double freq[B]}                                               (:656-681)
    b = bytearray(struct.pack("<i", tmpl.shape[0]))
    for row in tmpl:
        b += struct.pack("<6d", *range(6)) + struct.pack("<32d", *[float(x) for x in row])
    p = tmp_path / "templates.bin"; p.write_bytes(bytes(b))
    assert L.vrdd_io_template_count(str(p).encode(), 32) == tmpl.shape[0]
    t2 = np.empty_like(tmpl)
    assert L.vrdd_io_read_templates(str(p).encode(), 32, tmpl.shape[0], t2.ctypes.data) == 0 and np.array_equal(t2, tmpl)


def test_guards_and_errors(tmp_path):
    import vrdd_b200 as V
    L = V.lib()
    assert L.vrdd_io_codebook_blocks(str(tmp_path / "missing.bin").encode()) == V.ERR_INVALID
    p = tmp_path / "bad.bin"                                                      # NE > bins: the loader refuses (:611-614)
    p.write_bytes(struct.pack("<ii", 1, 1) + struct.pack("<iii?i", 0, 0, 0, False, 33))
    assert L.vrdd_io_read_codebook(str(p).encode(), 32, 1, np.empty(4, np.int32).ctypes.data,
                                   np.empty(64, np.float32).ctypes.data) == V.ERR_RANGE
    p.write_bytes(struct.pack("<ii", 1, 2) + struct.pack("<iii?i", 0, 0, 0, False, 0))   # truncated
    assert L.vrdd_io_read_codebook(str(p).encode(), 32, 2, np.empty(8, np.int32).ctypes.data,
                                   np.empty(128, np.float32).ctypes.data) == V.ERR_INVALID


def test_writers_round_trip_and_ppm(tmp_path, data):
    import vrdd_b200 as V
    L = V.lib()
    dims, hist, tmpl, cb, err = data
    n = hist.shape[0]
    f = lambda s: str(tmp_path / s).encode()
    assert L.vrdd_io_write_histograms(f("h"), n, 32, hist.ctypes.data) == 0
    assert L.vrdd_io_write_codebook(f("c"), 32, n, cb.ctypes.data, err.ctypes.data) == 0
    assert L.vrdd_io_write_templates(f("t"), 32, tmpl.shape[0], tmpl.ctypes.data) == 0
    h2 = np.empty_like(hist); c2 = np.empty_like(cb); e2 = np.empty_like(err); t2 = np.empty_like(tmpl)
    assert L.vrdd_io_read_histograms(f("h"), n, 32, h2.ctypes.data) == 0
    assert L.vrdd_io_read_codebook(f("c"), 32, n, c2.ctypes.data, e2.ctypes.data) == 0
    assert L.vrdd_io_read_templates(f("t"), 32, tmpl.shape[0], t2.ctypes.data) == 0
    assert np.array_equal(h2, hist) and np.array_equal(c2, cb) and np.array_equal(e2, err) and np.array_equal(t2, tmpl)
    img = (np.arange(7 * 5, dtype=np.uint32) * 0x01030507).reshape(5, 7)
    assert L.vrdd_io_write_ppm(f("i.ppm"), img.ctypes.data, 7, 5) == 0
    raw = (tmp_path / "i.ppm").read_bytes()
    assert raw.startswith(b"P6\n7 5\n255\n") and len(raw) == 11 + 7 * 5 * 3
    rgb = np.empty((5, 7, 3), np.uint8)
    assert L.vrdd_io_read_ppm(f("i.ppm"), rgb.ctypes.data, 7, 5) == 0
    assert np.array_equal(rgb, img.view(np.uint8).reshape(5, 7, 4)[:, :, :3])      # alpha dropped, rows in memory order
