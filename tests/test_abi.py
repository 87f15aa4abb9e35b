"""CPU tests of the drop-in boundary: libvrdd.so loads without a GPU, exports every symbol that
include/*.h declares, fails loudly (no CPU fallback) and its host-only helpers are right."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    body = src[src.index('extern "C" {'):]
    return set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", body))


def test_every_declared_symbol_is_exported():
    import vrdd_b200 as V
    L = C.CDLL(V.LIB_PATH)
    names = _declared("vrdd.h") | _declared("vrdd_legacy.h") | _declared("vrdd_io.h")
    assert {"vrdd_create", "vrdd_decode", "vrdd_render", "initCuda", "basicDataProcessing", "render_kernel",
            "copyInvViewMatrix", "setTextureFilterMode", "freeCudaBuffers", "dataProcessing"} <= names
    missing = [n for n in sorted(names) if not hasattr(L, n)]
    assert not missing, missing
    assert names == set(V.EXPORTS) | set(V.LEGACY_EXPORTS) | set(V.IO_EXPORTS) | set(V.PROBE_EXPORTS)


def test_legacy_signatures_match_reference_declarations():
    """The seven legacy declarations are the ones volumeRender.cpp:156-170 links against."""
    src = open(os.path.join(ROOT, "include", "vrdd_legacy.h")).read()
    for frag in ("void initCuda(void* h_volume, cudaExtent volumeSize, cudaExtent histogramSize, int4* h_codebook",
                 "void render_kernel(dim3 gridSize, dim3 blockSize, unsigned int* d_output, unsigned int imageW",
                 "void copyInvViewMatrix(float* invViewMatrix, size_t sizeofMatrix)",
                 "void setTextureFilterMode(bool bLinearFilter)", "void freeCudaBuffers(void)",
                 "void basicDataProcessing(void)", "void dataProcessing(void)"):
        assert frag in src, frag


def test_no_cpu_fallback_without_device():
    import torch
    import vrdd_b200 as V
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(V.VrddError) as e:
        V.Renderer(0)
    assert e.value.code == V.ERR_NO_DEVICE


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: no statement of the product includes, imports or loads it."""
    pkg = os.path.join(ROOT, "volume-rendering-based-on-distribution-data_b200")
    pat = re.compile(r"^\s*(#\s*include|import|from)\b[^\n]*oracle|(CDLL|dlopen)\([^\n]*oracle", re.M)
    seen = 0
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or f == "Makefile":
                seen += 1
                assert not pat.search(open(os.path.join(dp, f)).read()), f
    assert seen >= 8


def test_view_matrix_host_helper(oracle):
    import vrdd_b200 as V
    assert V.view_matrix() == [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 4]      # volumeRender.cpp:1024-1043
    for rx, ry, t in ((10.0, 20.0, (0.0, 0.0, -4.0)), (-75.0, 310.0, (0.5, -0.25, -2.5))):
        np.testing.assert_array_equal(np.array(V.view_matrix(rx, ry, t), np.float32), oracle.view_matrix(rx, ry, t))


def test_default_render_params_are_the_reference_constants():
    import vrdd_b200 as V
    p = V.default_render_params()
    assert (p.density, p.brightness, p.transfer_offset, p.transfer_scale) == (
        pytest.approx(0.05), 1.0, 0.0, 1.0)                          # volumeRender.cpp:130-133
    assert (p.max_steps, p.query_method) == (500, 1)                 # volumeRender_kernel.cu:276
    assert p.tstep == pytest.approx(0.01) and p.opacity_threshold == pytest.approx(0.95)


def test_pack_fractal_errors_matches_the_documented_layout(oracle):
    """vrdd_pack_fractal_errors (host code, no GPU): round-major entries and per-chunk offsets, against an
    independent numpy restatement; ragged tail chunk, voxels with NE = 0 and NE = bins included."""
    import vrdd_b200 as V
    from packing import round_major, chunk_offsets
    dims = (13, 7, 5)                                           # 455 voxels: 14 full chunks + 7
    cb, err = oracle.synth_fractal(77, dims, T=50, max_ne=8)
    cb = cb.copy(); err = err.copy()
    cb[3, 3] = 0
    cb[40, 3] = 32
    err[40, :, 0] = np.arange(32)[::-1]
    err[40, :, 1] = np.linspace(-0.01, 0.01, 32)
    ent, off = V.pack_fractal_errors(cb, err)
    want = round_major(cb, err)
    assert ent.shape == want.shape and np.array_equal(ent["bin"], want["bin"]) and np.array_equal(ent["value"], want["value"])
    assert np.array_equal(off, chunk_offsets(cb))
    bad = cb.copy(); bad[5, 3] = 33
    with pytest.raises(V.VrddError):
        V.pack_fractal_errors(bad, err)
    bad_e = err.copy(); bad_e[0, 0, 0] = 32.0
    if cb[0, 3] > 0:
        with pytest.raises(V.VrddError):
            V.pack_fractal_errors(cb, bad_e)
