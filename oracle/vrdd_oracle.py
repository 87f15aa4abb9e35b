"""ctypes binding of the CPU oracle (oracle/vrdd_oracle.cpp).

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs
may import this module; the product package never does.  The reference ships no fixtures; the
oracle is pinned by outputs of the reference's own device code run on a B200 (oracle/_ref,
tests/golden/ref_gpu_v1.npz, tests/test_reference_pin.py; see the header of vrdd_oracle.cpp).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)


class RenderParams(C.Structure):
    """struct vrdd_oracle_render_params; defaults are the reference's constants
    (volumeRender.cpp:121, 130-133; volumeRender_kernel.cu:276-278)."""
    _fields_ = [("image_w", C.c_int), ("image_h", C.c_int),
                ("density", C.c_float), ("brightness", C.c_float),
                ("transfer_offset", C.c_float), ("transfer_scale", C.c_float),
                ("query_method", C.c_int), ("tstep", C.c_float), ("max_steps", C.c_int),
                ("opacity_threshold", C.c_float), ("weight_quant", C.c_int),
                ("y0", C.c_int), ("y1", C.c_int)]


class FlexTables(C.Structure):
    """struct vrdd_oracle_flex_tables / struct vrdd_flex_tables (same layout)."""
    _fields_ = [("raw_w", C.c_int), ("raw_h", C.c_int), ("raw_d", C.c_int), ("bins", C.c_int),
                ("n_fractal", C.c_int), ("span_low", C.c_void_p), ("span_high", C.c_void_p), ("codebook", C.c_void_p),
                ("errors", C.c_void_p),
                ("n_simple", C.c_int), ("simple_low", C.c_void_p), ("simple_high", C.c_void_p),
                ("simple_count", C.c_void_p), ("simple_hist", C.c_void_p),
                ("n_templates", C.c_int), ("templates", C.c_void_p)]


def flex_tables_struct(t, cls=FlexTables):
    """dict of numpy arrays (tests/flex_synth.py) -> ctypes struct; keeps the arrays alive on the struct."""
    s = cls()
    s.raw_w, s.raw_h, s.raw_d = t["raw_dims"]
    s.bins = t["bins"]
    s.n_fractal = t["span_low"].shape[0]
    s.n_simple = t["simple_low"].shape[0]
    s.n_templates = t["templates"].shape[0]
    keep = []
    for name in ("span_low", "span_high", "codebook", "errors", "simple_low", "simple_high", "simple_count",
                 "simple_hist", "templates"):
        a = np.ascontiguousarray(t[name])
        keep.append(a)
        setattr(s, name, a.ctypes.data)
    s._keep = keep
    return s


def build(force=False):
    """Compile both flavours of the oracle with oracle/Makefile (g++ only, no CUDA)."""
    need = force or not all(os.path.exists(os.path.join(_HERE, n))
                            for n in ("libvrdd_oracle.so", "libvrdd_oracle_fast.so"))
    if need:
        subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else []),
                              stdout=subprocess.DEVNULL)


_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")


class Oracle:
    def __init__(self, fast=False):
        build()
        self.lib = C.CDLL(os.path.join(_HERE, "libvrdd_oracle_fast.so" if fast else "libvrdd_oracle.so"))
        L = self.lib
        L.vrdd_oracle_num_threads.restype = C.c_int
        L.vrdd_oracle_set_num_threads.argtypes = [C.c_int]
        L.vrdd_oracle_set_fma_contract.argtypes = [C.c_int]
        L.vrdd_oracle_set_fma_contract.restype = C.c_int
        L.vrdd_oracle_set_rsqrt_table.argtypes = [C.c_uint32, C.c_int64, C.c_void_p]
        L.vrdd_oracle_set_rsqrt_table.restype = None
        L.vrdd_oracle_decode_hist.argtypes = [_f32p, C.c_int64, C.c_int, _f32p]
        L.vrdd_oracle_decode_fractal.argtypes = [_i32p, _f32p, _f32p, C.c_int, C.c_int64, C.c_int, _f32p,
                                                 C.c_void_p]
        L.vrdd_oracle_decode_fractal.restype = C.c_int64
        L.vrdd_oracle_default_transfer_function.argtypes = [_f32p]
        L.vrdd_oracle_view_matrix.argtypes = [C.c_float] * 5 + [_f32p]
        L.vrdd_oracle_render.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, _f32p, C.c_int,
                                         _f32p, _u32p, C.POINTER(RenderParams)]
        L.vrdd_oracle_render.restype = C.c_int64
        L.vrdd_oracle_render_mode7.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, C.c_int, _f32p, _u32p,
                                               C.POINTER(RenderParams)]
        L.vrdd_oracle_render_mode7.restype = C.c_int64
        L.vrdd_oracle_point_index.argtypes = [C.c_float, C.c_int]
        L.vrdd_oracle_point_index.restype = C.c_int
        L.vrdd_oracle_flex_process.argtypes = [C.POINTER(FlexTables), C.c_int, _f32p, C.POINTER(C.c_int * 3)]
        L.vrdd_oracle_flex_process.restype = C.c_int64
        L.vrdd_oracle_render_flex.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, _f32p, C.c_int, _f32p, _u32p,
                                              C.POINTER(RenderParams)]
        L.vrdd_oracle_render_flex.restype = C.c_int64
        L.vrdd_oracle_tex3d.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                        C.c_float, C.c_int]
        L.vrdd_oracle_tex3d.restype = C.c_float
        L.vrdd_oracle_tex1d4.argtypes = [_f32p, C.c_int, C.c_float, C.c_int, _f32p]
        L.vrdd_oracle_synth_histograms.argtypes = [C.c_uint32] + [C.c_int] * 6 + [_f32p]
        L.vrdd_oracle_synth_templates.argtypes = [C.c_uint32, C.c_int, C.c_int, _f32p]
        L.vrdd_oracle_synth_fractal.argtypes = [C.c_uint32] + [C.c_int] * 8 + [_i32p, _f32p]

    # -- threads ------------------------------------------------------------------
    def num_threads(self):
        return int(self.lib.vrdd_oracle_num_threads())

    def set_num_threads(self, n):
        self.lib.vrdd_oracle_set_num_threads(int(n))

    def set_fma_contract(self, on):
        """Rounding of the ray set-up and compositing: False (default) = no contraction, the order the CUDA kernels
        reproduce bit for bit; True = nvcc's default contraction of the reference's d_render (vrdd_oracle.cpp,
        g_fma_contract).  Returns the previous setting."""
        return bool(self.lib.vrdd_oracle_set_fma_contract(1 if on else 0))

    def set_reference_build(self, on=True):
        """Rounding of the reference's own nvcc build for the ray set-up and compositing (the default of libvrdd.so,
        variant ray_setup = "nvcc"): FMA contraction as in its PTX plus the B200's rsqrt.approx.f32 from
        tests/golden/rsqrt_approx_b200_v1.npz (tools/rsqrt_dump.cu).  With the table the ray geometry is the
        binary's bit for bit.  Returns the previous contraction setting."""
        if on:
            fx = np.load(os.path.join(_ROOT, "tests", "golden", "rsqrt_approx_b200_v1.npz"))
            delta = np.ascontiguousarray(fx["delta"], np.int8)
            self.lib.vrdd_oracle_set_rsqrt_table(int(fx["lo_bits"]), delta.size, delta.ctypes.data)
        else:
            self.lib.vrdd_oracle_set_rsqrt_table(0, 0, None)
        return self.set_fma_contract(on)

    # -- synthetic inputs -----------------------------------------------------------
    def synth_histograms(self, seed, dims, bins=32, z0=0, nz=None):
        W, H, D = dims
        nz = D - z0 if nz is None else nz
        out = np.empty((nz * H * W, bins), np.float32)
        self.lib.vrdd_oracle_synth_histograms(seed, W, H, D, bins, z0, nz, out)
        return out

    def synth_templates(self, seed, T=622, bins=32):
        out = np.empty((T, bins), np.float32)
        self.lib.vrdd_oracle_synth_templates(seed, T, bins, out)
        return out

    def synth_fractal(self, seed, dims, bins=32, T=622, max_ne=8, z0=0, nz=None):
        W, H, D = dims
        nz = D - z0 if nz is None else nz
        V = nz * H * W
        codebook = np.empty((V, 4), np.int32)
        errors = np.empty((V, bins, 2), np.float32)
        self.lib.vrdd_oracle_synth_fractal(seed, W, H, D, bins, T, max_ne, z0, nz, codebook, errors)
        return codebook, errors

    # -- P1 -------------------------------------------------------------------------
    def decode_hist(self, hist):
        hist = np.ascontiguousarray(hist, np.float32)
        V, B = hist.shape
        out = np.empty((V, 4), np.float32)
        self.lib.vrdd_oracle_decode_hist(hist, V, B, out)
        return out

    def decode_fractal(self, codebook, errors, templates, want_recon=False):
        codebook = np.ascontiguousarray(codebook, np.int32)
        errors = np.ascontiguousarray(errors, np.float32)
        templates = np.ascontiguousarray(templates, np.float32)
        V = codebook.shape[0]
        T, B = templates.shape
        out = np.empty((V, 4), np.float32)
        recon = np.empty((V, B), np.float32) if want_recon else None
        bad = self.lib.vrdd_oracle_decode_fractal(codebook, errors, templates, T, V, B, out,
                                                  recon.ctypes.data if want_recon else None)
        return (out, recon, int(bad)) if want_recon else (out, int(bad))

    # -- P2 -------------------------------------------------------------------------
    def default_transfer_function(self):
        tf = np.empty((9, 4), np.float32)
        self.lib.vrdd_oracle_default_transfer_function(tf)
        return tf

    def view_matrix(self, rot_x=0.0, rot_y=0.0, trans=(0.0, 0.0, -4.0)):
        m = np.empty(12, np.float32)
        self.lib.vrdd_oracle_view_matrix(rot_x, rot_y, trans[0], trans[1], trans[2], m)
        return m

    def render(self, vol4, dims, view, image=(512, 512), query_method=1, tf=None, density=0.05,
               brightness=1.0, transfer_offset=0.0, transfer_scale=1.0, tstep=0.01, max_steps=500,
               opacity_threshold=0.95, weight_quant=3, vol_fractal4=None, rows=None, plane=None):
        """Returns (uint32 image[h][w] pre-cleared to 0, sample count).  plane: float[V] holding the ONE component
        query_method samples (instead of the float4 volumes), for volumes whose float4 form is too large."""
        W, H, D = dims
        iw, ih = image
        tf = self.default_transfer_function() if tf is None else np.ascontiguousarray(tf, np.float32)
        vo = None if vol4 is None else np.ascontiguousarray(vol4, np.float32)
        vf = None if vol_fractal4 is None else np.ascontiguousarray(vol_fractal4, np.float32)
        if plane is not None:
            plane = np.ascontiguousarray(plane, np.float32)
            assert plane.size == W * H * D
            vo, vf = (plane, None) if query_method < 4 else (None, plane)
            query_method += 100
        out = np.zeros((ih, iw), np.uint32)
        y0, y1 = (0, ih) if rows is None else rows
        P = RenderParams(iw, ih, density, brightness, transfer_offset, transfer_scale, query_method, tstep,
                         max_steps, opacity_threshold, weight_quant, y0, y1)
        s = self.lib.vrdd_oracle_render(None if vo is None else vo.ctypes.data,
                                        None if vf is None else vf.ctypes.data, W, H, D, tf, tf.shape[0],
                                        np.ascontiguousarray(view, np.float32), out, C.byref(P))
        return out, int(s)

    def render_mode7(self, hist, dims, view, image=(512, 512), tf=None, density=0.05, brightness=1.0,
                     transfer_offset=0.0, transfer_scale=1.0, tstep=0.01, max_steps=500, opacity_threshold=0.95):
        """queryMethod 7 (interpolated mean) straight from the raw histograms."""
        W, H, D = dims
        iw, ih = image
        hist = np.ascontiguousarray(hist, np.float32)
        tf = self.default_transfer_function() if tf is None else np.ascontiguousarray(tf, np.float32)
        out = np.zeros((ih, iw), np.uint32)
        P = RenderParams(iw, ih, density, brightness, transfer_offset, transfer_scale, 7, tstep, max_steps,
                         opacity_threshold, 3, 0, ih)
        s = self.lib.vrdd_oracle_render_mode7(hist, W, H, D, hist.shape[1], tf, tf.shape[0],
                                              np.ascontiguousarray(view, np.float32), out, C.byref(P))
        return out, int(s)

    def flex_process(self, tables, block):
        """dataProcessing(): returns (float4 blocks [n][4], (nx, ny, nz), spans not found)."""
        st = flex_tables_struct(tables)
        nb = [(d + block - 1) // block for d in tables["raw_dims"]]
        out = np.zeros((nb[0] * nb[1] * nb[2], 4), np.float32)
        dims = (C.c_int * 3)()
        missing = self.lib.vrdd_oracle_flex_process(C.byref(st), block, out, C.byref(dims))
        return out, tuple(dims), int(missing)

    def render_flex(self, blocks4, nb, view, image=(512, 512), query_method=8, tf=None, density=0.05, brightness=1.0,
                    transfer_offset=0.0, transfer_scale=1.0, tstep=0.01, max_steps=500, opacity_threshold=0.95):
        iw, ih = image
        tf = self.default_transfer_function() if tf is None else np.ascontiguousarray(tf, np.float32)
        out = np.zeros((ih, iw), np.uint32)
        P = RenderParams(iw, ih, density, brightness, transfer_offset, transfer_scale, query_method, tstep, max_steps,
                         opacity_threshold, 3, 0, ih)
        s = self.lib.vrdd_oracle_render_flex(np.ascontiguousarray(blocks4, np.float32), nb[0], nb[1], nb[2], tf,
                                             tf.shape[0], np.ascontiguousarray(view, np.float32), out, C.byref(P))
        return out, int(s)

    def point_index(self, u, n):
        return int(self.lib.vrdd_oracle_point_index(u, n))

    def tex3d(self, vol4, dims, comp, u, v, w, weight_quant=3):
        W, H, D = dims
        return float(self.lib.vrdd_oracle_tex3d(np.ascontiguousarray(vol4, np.float32), W, H, D, comp, u, v, w,
                                                weight_quant))

    def tex1d4(self, tf, u, weight_quant=3):
        tf = np.ascontiguousarray(tf, np.float32)
        out = np.empty(4, np.float32)
        self.lib.vrdd_oracle_tex1d4(tf, tf.shape[0], u, weight_quant, out)
        return out
