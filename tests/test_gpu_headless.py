"""The headless C++ driver (host/vrdd_headless.cpp) run as a program: it must reproduce the
oracle's image of the reference's own configuration (50x50x10 volume, 512x512, self-test view,
volumeRender.cpp:86, 121, 1024-1043) from synthetic inputs and from files in the reference's
formats, and its --file self-test must pass and fail correctly."""
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ppm(path, w, h):
    raw = open(path, "rb").read()
    return np.frombuffer(raw[len(b"P6\n%d %d\n255\n" % (w, h)):], np.uint8).reshape(h, w, 3)


def test_headless_driver(tmp_path, oracle):
    import vrdd_b200 as V
    L = V.lib()
    dims, seed, w, h = (50, 50, 10), 1234, 512, 512
    hist = oracle.synth_histograms(seed, dims)
    tmpl = oracle.synth_templates(seed, 622)
    cb, err = oracle.synth_fractal(seed, dims)
    ref_o = oracle.decode_hist(hist)
    ref_f, _ = oracle.decode_fractal(cb, err, tmpl)
    out = str(tmp_path / "volume")
    r = subprocess.run([V.HEADLESS_PATH, f"--out={out}", "--iters=3"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "volumeRender, Throughput =" in r.stdout and "MTexels/s" in r.stdout     # the reference's log line (:1066)
    ref, _ = oracle.render(ref_o, dims, oracle.view_matrix(), image=(w, h))
    ref_rgb = ref.view(np.uint8).reshape(h, w, 4)[:, :, :3].astype(np.int16)
    got = _ppm(out + ".ppm", w, h).astype(np.int16)
    assert np.abs(got - ref_rgb).max() <= 1
    # the --file self-test: passes against its own output, fails against a different view
    r = subprocess.run([V.HEADLESS_PATH, f"--out={tmp_path / 'again'}", f"--file={out}.ppm", "--iters=1"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "PASSED" in r.stdout, r.stdout + r.stderr
    r = subprocess.run([V.HEADLESS_PATH, f"--out={tmp_path / 'rot'}", f"--file={out}.ppm", "--iters=1", "--roty=30"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "FAILED" in r.stdout
    # from files in the reference's formats, fractal mode 5, a two-view orbit
    f = lambda s: str(tmp_path / s).encode()
    assert L.vrdd_io_write_histograms(f("h.bin"), hist.shape[0], 32, hist.ctypes.data) == 0
    assert L.vrdd_io_write_codebook(f("c.bin"), 32, cb.shape[0], cb.ctypes.data, err.ctypes.data) == 0
    assert L.vrdd_io_write_templates(f("t.bin"), 32, 622, tmpl.ctypes.data) == 0
    r = subprocess.run([V.HEADLESS_PATH, f"--volume={tmp_path / 'h.bin'}", f"--codebook={tmp_path / 'c.bin'}",
                        f"--templates={tmp_path / 't.bin'}", "--query=5", "--views=2", "--width=256", "--height=192",
                        f"--out={tmp_path / 'orbit'}", "--iters=1"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    for k in range(2):
        ref, _ = oracle.render(ref_o, dims, oracle.view_matrix(0.0, 180.0 * k), image=(256, 192), query_method=5,
                               vol_fractal4=ref_f)
        ref_rgb = ref.view(np.uint8).reshape(192, 256, 4)[:, :, :3].astype(np.int16)
        assert np.abs(_ppm(str(tmp_path / f"orbit_{k}.ppm"), 256, 192).astype(np.int16) - ref_rgb).max() <= 1
