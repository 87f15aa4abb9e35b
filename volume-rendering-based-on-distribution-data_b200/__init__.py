"""vrdd_b200 — Python binding of libvrdd.so (include/vrdd.h), the B200-native distribution
decode and volume ray caster behind the host-facing surface of
ykou/Volume-Rendering-Based-on-Distribution-Data.

This module is plumbing: it loads the in-tree CUDA library with ctypes and mirrors the C ABI
one to one (`Renderer` methods are named after the vrdd_* functions; the legacy names
initCuda / basicDataProcessing / copyInvViewMatrix / render_kernel / freeCudaBuffers are
exposed through `legacy`).  There is NO CPU fallback and nothing here imports oracle/: if
libvrdd.so is missing or no B200 is present, calls raise.

The directory name contains hyphens, so import it through the `vrdd_b200` shim at the
repository root.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvrdd.so")
CSRC = os.path.join(_HERE, "csrc")

OK, ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE, ERR_RANGE, ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5
SRC_ORIGINAL, SRC_FRACTAL = 0, 1
SAMPLER_TEXTURE, SAMPLER_BRICKED, SAMPLER_LINEAR = 0, 1, 2
ERR_CHUNK = 32

EXPORTS = [
    "vrdd_create", "vrdd_destroy", "vrdd_set_stream", "vrdd_synchronize", "vrdd_last_error",
    "vrdd_kernel_launches", "vrdd_set_volume", "vrdd_set_histograms_host", "vrdd_set_histograms_device",
    "vrdd_set_fractal_host", "vrdd_set_fractal_device", "vrdd_pack_fractal_errors", "vrdd_set_sampler", "vrdd_decode",
    "vrdd_get_decoded_host", "vrdd_get_decoded_planes_device", "vrdd_keep_linear_planes", "vrdd_commit_planes", "vrdd_commit_planes_mask",
    "vrdd_reconstruct_fractal_device", "vrdd_set_transfer_function", "vrdd_set_view",
    "vrdd_default_render_params", "vrdd_render", "vrdd_render_host", "vrdd_render_host_async",
    "vrdd_render_host_wait", "vrdd_render_host_fence", "vrdd_host_register", "vrdd_host_unregister", "vrdd_count_samples",
    "vrdd_get_sample_count", "vrdd_view_matrix", "vrdd_synth_histograms_device", "vrdd_synth_fractal_device",
    "vrdd_set_variant", "vrdd_enable_interpolated_mean",
    "vrdd_flex_divide_blocks", "vrdd_flex_prefix_spans", "vrdd_flex_set_tables_host", "vrdd_flex_process",
    "vrdd_flex_get_blocks_host",
    "vrdd_frame_alloc", "vrdd_frame_free", "vrdd_frame_export", "vrdd_frame_open", "vrdd_frame_close",
    "vrdd_set_frame_signal", "vrdd_stream_wait_flag", "vrdd_stream_post_flag", "vrdd_stream_wait_post_flag", "vrdd_set_peer_planes",
    "vrdd_get_mean_raw_device", "vrdd_set_peer_mean_raw", "vrdd_commit_mean_raw",
    "vrdd_render_brick_alpha_send", "vrdd_render_brick_color_send", "vrdd_pack_frame_slots",
    "vrdd_render_brick_color_send_bands", "vrdd_pack_band_slots",
    "vrdd_render_brick_alpha", "vrdd_compose_alpha_in", "vrdd_compose_alpha_in_rows", "vrdd_render_brick_color", "vrdd_pack_frame",
    "vrdd_synth_histograms_region_device",
]
# texture-unit probes: present when libvrdd.so was built with PROBES=1 (csrc/Makefile; the default of this repository)
PROBE_EXPORTS = ["vrdd_debug_sample_texture", "vrdd_debug_sample_transfer_function", "vrdd_debug_sample_texture_point",
                 "vrdd_debug_sample_texture_unnorm"]
IO_EXPORTS = ["vrdd_io_read_histograms", "vrdd_io_codebook_blocks", "vrdd_io_read_codebook", "vrdd_io_template_count",
              "vrdd_io_read_templates", "vrdd_io_write_histograms", "vrdd_io_write_codebook", "vrdd_io_write_templates",
              "vrdd_io_write_ppm", "vrdd_io_read_ppm", "vrdd_io_span_count", "vrdd_io_read_span_list",
              "vrdd_io_read_flex_codebook", "vrdd_io_simple_count", "vrdd_io_read_simple", "vrdd_io_write_span_list",
              "vrdd_io_write_flex_codebook", "vrdd_io_write_simple"]
HEADLESS_PATH = os.path.join(_HERE, "vrdd_headless")
LEGACY_EXPORTS = ["initCuda", "basicDataProcessing", "dataProcessing", "copyInvViewMatrix", "render_kernel",
                  "setTextureFilterMode", "freeCudaBuffers", "vrdd_legacy_handle"]


class VrddError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"vrdd error {code}: {msg}")
        self.code = code


class RenderParams(C.Structure):
    """struct vrdd_render_params (include/vrdd.h)."""
    _fields_ = [("density", C.c_float), ("brightness", C.c_float), ("transfer_offset", C.c_float),
                ("transfer_scale", C.c_float), ("tstep", C.c_float), ("max_steps", C.c_int),
                ("opacity_threshold", C.c_float), ("query_method", C.c_int)]


class TilePartition(C.Structure):
    """struct vrdd_tile_partition (include/vrdd.h)."""
    _fields_ = [("tile_w", C.c_int), ("tile_h", C.c_int), ("part", C.c_int), ("parts", C.c_int)]


class Brick(C.Structure):
    """struct vrdd_brick (include/vrdd.h)."""
    _fields_ = [("gw", C.c_int), ("gh", C.c_int), ("gd", C.c_int), ("ox", C.c_int), ("oy", C.c_int), ("oz", C.c_int),
                ("lo", C.c_float * 3), ("hi", C.c_float * 3)]


class FlexTables(C.Structure):
    """struct vrdd_flex_tables (include/vrdd.h)."""
    _fields_ = [("raw_w", C.c_int), ("raw_h", C.c_int), ("raw_d", C.c_int), ("bins", C.c_int),
                ("n_fractal", C.c_int), ("span_low", C.c_void_p), ("span_high", C.c_void_p), ("codebook", C.c_void_p),
                ("errors", C.c_void_p),
                ("n_simple", C.c_int), ("simple_low", C.c_void_p), ("simple_high", C.c_void_p),
                ("simple_count", C.c_void_p), ("simple_hist", C.c_void_p),
                ("n_templates", C.c_int), ("templates", C.c_void_p)]


class Extent(C.Structure):
    """cudaExtent"""
    _fields_ = [("width", C.c_size_t), ("height", C.c_size_t), ("depth", C.c_size_t)]


class Dim3(C.Structure):
    _fields_ = [("x", C.c_uint), ("y", C.c_uint), ("z", C.c_uint)]


def build(force=False, verbose=False):
    """Compile libvrdd.so for sm_100a with csrc/Makefile (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", CSRC, "-j8"] + (["-B"] if force else [])
    subprocess.check_call(cmd, stdout=None if verbose else subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    """The loaded library.  Fails loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        vp, i32, f32, u32 = C.c_void_p, C.c_int, C.c_float, C.c_uint32
        sig = {
            "vrdd_create": (i32, [i32, C.POINTER(vp)]),
            "vrdd_destroy": (i32, [vp]),
            "vrdd_set_stream": (i32, [vp, vp]),
            "vrdd_synchronize": (i32, [vp]),
            "vrdd_last_error": (C.c_char_p, [vp]),
            "vrdd_kernel_launches": (C.c_int64, [vp]),
            "vrdd_set_volume": (i32, [vp, i32, i32, i32, i32]),
            "vrdd_set_histograms_host": (i32, [vp, vp]),
            "vrdd_set_histograms_device": (i32, [vp, vp, i32, i32]),
            "vrdd_set_fractal_host": (i32, [vp, vp, vp, vp, i32]),
            "vrdd_set_fractal_device": (i32, [vp, vp, vp, vp, vp, i32, i32, i32]),
            "vrdd_pack_fractal_errors": (i32, [vp, vp, C.c_int64, i32, vp, vp, C.POINTER(C.c_uint64)]),
            "vrdd_set_sampler": (i32, [vp, i32]),
            "vrdd_decode": (i32, [vp, i32, i32, i32]),
            "vrdd_get_decoded_host": (i32, [vp, i32, vp]),
            "vrdd_get_decoded_planes_device": (i32, [vp, i32, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]),
            "vrdd_keep_linear_planes": (i32, [vp, i32]),
            "vrdd_enable_interpolated_mean": (i32, [vp, i32]),
            "vrdd_commit_planes": (i32, [vp, i32, i32, i32]),
            "vrdd_commit_planes_mask": (i32, [vp, i32, i32, i32, i32]),
            "vrdd_reconstruct_fractal_device": (i32, [vp, vp]),
            "vrdd_set_transfer_function": (i32, [vp, vp, i32]),
            "vrdd_set_view": (i32, [vp, vp]),
            "vrdd_default_render_params": (None, [C.POINTER(RenderParams)]),
            "vrdd_render": (i32, [vp, vp, i32, i32, C.POINTER(RenderParams), C.POINTER(TilePartition), i32]),
            "vrdd_render_host": (i32, [vp, vp, i32, i32, C.POINTER(RenderParams)]),
            "vrdd_render_host_async": (i32, [vp, vp, i32, i32, C.POINTER(RenderParams), C.POINTER(TilePartition)]),
            "vrdd_render_host_wait": (i32, [vp]),
            "vrdd_render_host_fence": (i32, [vp, i32]),
            "vrdd_host_register": (i32, [vp, C.c_size_t]),
            "vrdd_host_unregister": (i32, [vp]),
            "vrdd_count_samples": (i32, [vp, i32]),
            "vrdd_get_sample_count": (i32, [vp, C.POINTER(C.c_int64), i32]),
            "vrdd_view_matrix": (None, [f32, f32, f32, f32, f32, vp]),
            "vrdd_synth_histograms_device": (i32, [vp, u32, i32, i32, vp]),
            "vrdd_synth_fractal_device": (i32, [vp, u32, i32, i32, i32, i32, vp, vp, vp, vp, C.POINTER(C.c_uint64)]),
            "vrdd_set_variant": (i32, [vp, C.c_char_p, C.c_char_p]),
            "vrdd_debug_sample_texture": (i32, [vp, i32, i32, vp, i32, vp]),
            "vrdd_debug_sample_transfer_function": (i32, [vp, vp, i32, vp]),
            "vrdd_debug_sample_texture_point": (i32, [vp, i32, i32, vp, i32, vp]),
            "vrdd_debug_sample_texture_unnorm": (i32, [vp, i32, i32, vp, i32, vp]),
            "vrdd_render_brick_alpha": (i32, [vp, vp, i32, i32, C.POINTER(RenderParams), C.POINTER(Brick)]),
            "vrdd_compose_alpha_in": (i32, [vp, vp, i32, i32, i32, i32, i32, i32, vp, i32, i32]),
            "vrdd_compose_alpha_in_rows": (i32, [vp, vp, i32, i32, i32, i32, i32, i32, vp, i32, vp, i32, i32]),
            "vrdd_render_brick_color": (i32, [vp, vp, vp, i32, i32, C.POINTER(RenderParams), C.POINTER(Brick)]),
            "vrdd_pack_frame": (i32, [vp, vp, vp, i32, i32, f32]),
            "vrdd_render_brick_alpha_send": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, i32, C.POINTER(RenderParams), C.POINTER(Brick)]),
            "vrdd_render_brick_color_send": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, i32, C.POINTER(RenderParams), C.POINTER(Brick)]),
            "vrdd_pack_frame_slots": (i32, [vp, vp, i32, vp, i32, vp, i32, i32, f32]),
            "vrdd_render_brick_color_send_bands": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, C.POINTER(RenderParams), C.POINTER(Brick)]),
            "vrdd_pack_band_slots": (i32, [vp, vp, i32, vp, i32, i32, i32, vp, vp, i32, i32, f32]),
            "vrdd_synth_histograms_region_device": (i32, [vp, u32, i32, i32, i32, i32, i32, i32, i32, i32, vp]),
            "vrdd_flex_set_tables_host": (i32, [vp, C.POINTER(FlexTables)]),
            "vrdd_flex_process": (i32, [vp, i32, C.POINTER(C.c_int64)]),
            "vrdd_flex_get_blocks_host": (i32, [vp, vp, C.POINTER(C.c_int * 3)]),
            "vrdd_flex_divide_blocks": (i32, [i32, i32, i32, i32, vp, i32]),
            "vrdd_flex_prefix_spans": (i32, [i32, vp]),
            "vrdd_frame_alloc": (i32, [vp, C.c_size_t, C.POINTER(vp)]),
            "vrdd_frame_free": (i32, [vp, vp]),
            "vrdd_frame_export": (i32, [vp, vp, vp]),
            "vrdd_frame_open": (i32, [vp, vp, C.POINTER(vp)]),
            "vrdd_frame_close": (i32, [vp, vp]),
            "vrdd_set_frame_signal": (i32, [vp, vp]),
            "vrdd_stream_wait_flag": (i32, [vp, vp, u32]),
            "vrdd_stream_post_flag": (i32, [vp, vp]),
            "vrdd_stream_wait_post_flag": (i32, [vp, vp, u32, vp]),
            "vrdd_set_peer_planes": (i32, [vp, i32, i32, vp]),
            "vrdd_get_mean_raw_device": (i32, [vp, C.POINTER(vp)]),
            "vrdd_set_peer_mean_raw": (i32, [vp, i32, vp]),
            "vrdd_commit_mean_raw": (i32, [vp, i32, i32]),
            # on-disk formats (include/vrdd_io.h)
            "vrdd_io_read_histograms": (i32, [C.c_char_p, C.c_size_t, i32, vp]),
            "vrdd_io_codebook_blocks": (C.c_int64, [C.c_char_p]),
            "vrdd_io_read_codebook": (i32, [C.c_char_p, i32, C.c_int64, vp, vp]),
            "vrdd_io_template_count": (i32, [C.c_char_p, i32]),
            "vrdd_io_read_templates": (i32, [C.c_char_p, i32, i32, vp]),
            "vrdd_io_write_histograms": (i32, [C.c_char_p, C.c_size_t, i32, vp]),
            "vrdd_io_write_codebook": (i32, [C.c_char_p, i32, C.c_int64, vp, vp]),
            "vrdd_io_write_templates": (i32, [C.c_char_p, i32, i32, vp]),
            "vrdd_io_write_ppm": (i32, [C.c_char_p, vp, i32, i32]),
            "vrdd_io_span_count": (i32, [C.c_char_p]),
            "vrdd_io_read_span_list": (i32, [C.c_char_p, i32, vp, vp]),
            "vrdd_io_read_flex_codebook": (i32, [C.c_char_p, i32, C.c_int64, vp, vp, vp]),
            "vrdd_io_simple_count": (i32, [C.c_char_p]),
            "vrdd_io_read_simple": (i32, [C.c_char_p, C.c_char_p, C.c_char_p, i32, i32, vp, vp, vp, vp]),
            "vrdd_io_write_span_list": (i32, [C.c_char_p, i32, vp, vp]),
            "vrdd_io_write_flex_codebook": (i32, [C.c_char_p, i32, C.c_int64, vp, vp, vp]),
            "vrdd_io_write_simple": (i32, [C.c_char_p, C.c_char_p, C.c_char_p, i32, i32, vp, vp, vp, vp]),
            "vrdd_io_read_ppm": (i32, [C.c_char_p, vp, i32, i32]),
            # legacy surface (include/vrdd_legacy.h)
            "initCuda": (None, [vp, Extent, Extent, vp, Extent, vp, Extent, vp, Extent] + [vp] * 9),
            "basicDataProcessing": (None, []),
            "dataProcessing": (None, []),
            "copyInvViewMatrix": (None, [vp, C.c_size_t]),
            "render_kernel": (None, [Dim3, Dim3, vp, C.c_uint, C.c_uint, f32, f32, f32, f32, i32, Extent]),
            "setTextureFilterMode": (None, [C.c_bool]),
            "freeCudaBuffers": (None, []),
            "vrdd_legacy_handle": (vp, []),
        }
        for name, (res, args) in sig.items():
            if name in PROBE_EXPORTS and not hasattr(L, name):
                continue                                    # a library built without the probes
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def default_render_params(**over):
    p = RenderParams()
    lib().vrdd_default_render_params(C.byref(p))
    for k, v in over.items():
        setattr(p, k, v)
    return p


def view_matrix(rot_x=0.0, rot_y=0.0, trans=(0.0, 0.0, -4.0)):
    """vrdd_view_matrix: the inverse view matrix of volumeRender.cpp:224-246 as 12 floats."""
    m = (C.c_float * 12)()
    lib().vrdd_view_matrix(rot_x, rot_y, trans[0], trans[1], trans[2], m)
    return list(m)


def _ptr(x):
    """Address of a numpy array / torch tensor / ctypes array / int."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if hasattr(x, "ctypes"):
        return x.ctypes.data
    return C.addressof(x)


class DeviceArray:
    """Raw device pointer published through __cuda_array_interface__, so memory owned by the
    library (e.g. the decoded planes) can be handed to torch.distributed without a copy."""

    def __init__(self, ptr, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3}


def as_torch(ptr, shape, typestr="<f4", device="cuda"):
    import torch
    return torch.as_tensor(DeviceArray(ptr, shape, typestr), device=device)


ERROR_ENTRY_DTYPE = [("bin", "<i4"), ("value", "<f4")]      # vrdd_error_entry


def pack_fractal_errors(codebook, errors_dense, bins=32):
    """vrdd_pack_fractal_errors: the reference's dense float2[V][bins] error table -> the compact form of
    vrdd_set_fractal_device.  Returns (entries as a structured numpy array, chunk_offsets uint64)."""
    import numpy as np
    cb = np.ascontiguousarray(codebook, dtype=np.int32).reshape(-1, 4)
    ed = np.ascontiguousarray(errors_dense, dtype=np.float32).reshape(cb.shape[0], bins, 2)
    n = cb.shape[0]
    off = np.zeros((n + ERR_CHUNK - 1) // ERR_CHUNK + 1, dtype=np.uint64)
    tot = C.c_uint64()
    rc = lib().vrdd_pack_fractal_errors(cb.ctypes.data, ed.ctypes.data, n, bins, None, off.ctypes.data, C.byref(tot))
    if rc != 0:
        raise VrddError(rc, "vrdd_pack_fractal_errors: codebook entry or error bin out of range")
    ent = np.zeros(max(int(tot.value), 1), dtype=ERROR_ENTRY_DTYPE)
    rc = lib().vrdd_pack_fractal_errors(cb.ctypes.data, ed.ctypes.data, n, bins, ent.ctypes.data, off.ctypes.data, C.byref(tot))
    if rc != 0:
        raise VrddError(rc, "vrdd_pack_fractal_errors failed")
    return ent[:int(tot.value)], off


def host_register(ptr, nbytes):
    """vrdd_host_register: page-lock caller memory (address or array) for asynchronous copies."""
    rc = lib().vrdd_host_register(_ptr(ptr), nbytes)
    if rc != 0:
        raise VrddError(rc, "vrdd_host_register failed")


def host_unregister(ptr):
    lib().vrdd_host_unregister(_ptr(ptr))


class Renderer:
    """One vrdd handle.  Methods map 1:1 onto the vrdd_* C functions."""

    def __init__(self, device=-1):
        self._h = C.c_void_p()
        rc = lib().vrdd_create(device, C.byref(self._h))
        if rc != OK:
            raise VrddError(rc, "vrdd_create failed: no usable sm_100 CUDA device (there is no CPU fallback)"
                            if rc == ERR_NO_DEVICE else "vrdd_create failed")

    def _ck(self, rc):
        if rc != OK:
            raise VrddError(rc, lib().vrdd_last_error(self._h).decode())

    def close(self):
        if self._h:
            lib().vrdd_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # lifetime / stream
    def set_stream(self, cuda_stream):
        self._ck(lib().vrdd_set_stream(self._h, cuda_stream))

    def synchronize(self):
        self._ck(lib().vrdd_synchronize(self._h))

    def kernel_launches(self):
        return int(lib().vrdd_kernel_launches(self._h))

    # inputs
    def set_volume(self, w, h, d, bins=32):
        self._ck(lib().vrdd_set_volume(self._h, w, h, d, bins))

    def set_histograms_host(self, hist):
        self._ck(lib().vrdd_set_histograms_host(self._h, _ptr(hist)))

    def set_histograms_device(self, d_hist, z0, nz):
        self._ck(lib().vrdd_set_histograms_device(self._h, _ptr(d_hist), z0, nz))

    def set_fractal_host(self, codebook, errors_dense, templates):
        self._ck(lib().vrdd_set_fractal_host(self._h, _ptr(codebook), _ptr(errors_dense), _ptr(templates),
                                             templates.shape[0]))

    def set_fractal_device(self, d_codebook, d_errors, d_chunk_offsets, d_templates, num_templates, z0, nz):
        self._ck(lib().vrdd_set_fractal_device(self._h, _ptr(d_codebook), _ptr(d_errors), _ptr(d_chunk_offsets),
                                               _ptr(d_templates), num_templates, z0, nz))

    # decode
    def set_sampler(self, sampler):
        self._ck(lib().vrdd_set_sampler(self._h, sampler))

    def enable_interpolated_mean(self, enable=True):
        self._ck(lib().vrdd_enable_interpolated_mean(self._h, int(enable)))

    def keep_linear_planes(self, keep=True):
        self._ck(lib().vrdd_keep_linear_planes(self._h, int(keep)))

    def decode(self, source=SRC_ORIGINAL, z0=0, nz=0):
        self._ck(lib().vrdd_decode(self._h, source, z0, nz))

    def get_decoded_host(self, source, out4):
        self._ck(lib().vrdd_get_decoded_host(self._h, source, _ptr(out4)))
        return out4

    def get_decoded_planes_device(self, source):
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._ck(lib().vrdd_get_decoded_planes_device(self._h, source, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def commit_planes(self, source, z0, nz, plane_mask=7):
        self._ck(lib().vrdd_commit_planes_mask(self._h, source, z0, nz, plane_mask))

    def reconstruct_fractal_device(self, d_out):
        self._ck(lib().vrdd_reconstruct_fractal_device(self._h, _ptr(d_out)))

    # render
    def set_transfer_function(self, tf=None):
        if tf is None:
            self._ck(lib().vrdd_set_transfer_function(self._h, None, 0))
        else:
            self._ck(lib().vrdd_set_transfer_function(self._h, _ptr(tf), tf.shape[0]))

    def set_view(self, m12):
        arr = (C.c_float * 12)(*[float(v) for v in m12])
        self._ck(lib().vrdd_set_view(self._h, arr))

    def render(self, d_output, w, h, params=None, part=None, clear_misses=False):
        self._ck(lib().vrdd_render(self._h, _ptr(d_output), w, h, C.byref(params) if params is not None else None,
                                   C.byref(part) if part is not None else None, int(clear_misses)))

    def render_host(self, h_output, w, h, params=None):
        self._ck(lib().vrdd_render_host(self._h, _ptr(h_output), w, h,
                                        C.byref(params) if params is not None else None))
        return h_output

    def render_host_async(self, h_output, w, h, params=None, part=None):
        self._ck(lib().vrdd_render_host_async(self._h, _ptr(h_output), w, h,
                                              C.byref(params) if params is not None else None,
                                              C.byref(part) if part is not None else None))

    def render_host_fence(self, lag=1):
        self._ck(lib().vrdd_render_host_fence(self._h, lag))

    def render_host_wait(self):
        self._ck(lib().vrdd_render_host_wait(self._h))

    def count_samples(self, enable=True):
        self._ck(lib().vrdd_count_samples(self._h, int(enable)))

    def get_sample_count(self, reset=True):
        v = C.c_int64()
        self._ck(lib().vrdd_get_sample_count(self._h, C.byref(v), int(reset)))
        return int(v.value)

    # synthetic inputs
    def synth_histograms_device(self, seed, z0, nz, d_hist):
        self._ck(lib().vrdd_synth_histograms_device(self._h, seed, z0, nz, _ptr(d_hist)))

    def synth_fractal_device(self, seed, num_templates, max_ne, z0, nz, d_codebook, d_errors, d_chunk_offsets,
                             d_templates):
        tot = C.c_uint64()
        self._ck(lib().vrdd_synth_fractal_device(self._h, seed, num_templates, max_ne, z0, nz, _ptr(d_codebook),
                                                 _ptr(d_errors), _ptr(d_chunk_offsets), _ptr(d_templates),
                                                 C.byref(tot)))
        return int(tot.value)

    # flexible-block chain
    def flex_set_tables_host(self, t):
        """t: dict with raw_dims, bins and the nine tables as C-contiguous numpy arrays."""
        import numpy as np
        s = FlexTables()
        s.raw_w, s.raw_h, s.raw_d = t["raw_dims"]
        s.bins = t["bins"]
        s.n_fractal, s.n_simple, s.n_templates = t["span_low"].shape[0], t["simple_low"].shape[0], t["templates"].shape[0]
        keep = []
        for name in ("span_low", "span_high", "codebook", "errors", "simple_low", "simple_high", "simple_count",
                     "simple_hist", "templates"):
            a = np.ascontiguousarray(t[name])
            keep.append(a)
            setattr(s, name, a.ctypes.data)
        self._ck(lib().vrdd_flex_set_tables_host(self._h, C.byref(s)))

    def flex_process(self, block_size, want_missing=True):
        m = C.c_int64(0)
        self._ck(lib().vrdd_flex_process(self._h, block_size, C.byref(m) if want_missing else None))
        return int(m.value)

    def flex_get_blocks_host(self):
        import numpy as np
        dims = (C.c_int * 3)()
        self._ck(lib().vrdd_flex_get_blocks_host(self._h, None, C.byref(dims)))
        out = np.empty((dims[0] * dims[1] * dims[2], 4), np.float32)
        self._ck(lib().vrdd_flex_get_blocks_host(self._h, out.ctypes.data, C.byref(dims)))
        return out, tuple(dims)

    # peer-visible frames
    def frame_alloc(self, nbytes):
        p = C.c_void_p()
        self._ck(lib().vrdd_frame_alloc(self._h, nbytes, C.byref(p)))
        return p.value

    def frame_free(self, ptr):
        self._ck(lib().vrdd_frame_free(self._h, ptr))

    def frame_export(self, ptr):
        buf = (C.c_ubyte * 64)()
        self._ck(lib().vrdd_frame_export(self._h, ptr, buf))
        return bytes(buf)

    def frame_open(self, handle_bytes):
        buf = (C.c_ubyte * 64).from_buffer_copy(handle_bytes)
        p = C.c_void_p()
        self._ck(lib().vrdd_frame_open(self._h, buf, C.byref(p)))
        return p.value

    def frame_close(self, ptr):
        self._ck(lib().vrdd_frame_close(self._h, ptr))

    # frame-complete signals (no host barrier per frame) and decode-fused replication
    def set_frame_signal(self, d_flag):
        self._ck(lib().vrdd_set_frame_signal(self._h, _ptr(d_flag)))

    def stream_wait_flag(self, d_flag, at_least):
        self._ck(lib().vrdd_stream_wait_flag(self._h, _ptr(d_flag), int(at_least) & 0xffffffff))

    def stream_post_flag(self, d_flag):
        self._ck(lib().vrdd_stream_post_flag(self._h, _ptr(d_flag)))

    def stream_wait_post_flag(self, d_wait, at_least, d_post):
        self._ck(lib().vrdd_stream_wait_post_flag(self._h, _ptr(d_wait), int(at_least) & 0xffffffff, _ptr(d_post)))

    def set_peer_planes(self, source, peers):
        """peers: list (one entry per other rank) of 3 device pointers (mean, variance, entropy planes; None = skip)."""
        n = len(peers)
        arr = (C.c_void_p * max(1, 3 * n))()
        for q, planes in enumerate(peers):
            for i in range(3):
                arr[3 * q + i] = planes[i]
        self._ck(lib().vrdd_set_peer_planes(self._h, source, n, arr))

    def get_mean_raw_device(self):
        p = C.c_void_p()
        self._ck(lib().vrdd_get_mean_raw_device(self._h, C.byref(p)))
        return p.value

    def set_peer_mean_raw(self, peers):
        arr = (C.c_void_p * max(1, len(peers)))(*peers)
        self._ck(lib().vrdd_set_peer_mean_raw(self._h, len(peers), arr))

    def commit_mean_raw(self, z0, nz):
        self._ck(lib().vrdd_commit_mean_raw(self._h, z0, nz))

    # sort-last bricks
    def synth_histograms_region_device(self, seed, gdims, origin, z0, nz, d_hist):
        self._ck(lib().vrdd_synth_histograms_region_device(self._h, seed, gdims[0], gdims[1], gdims[2], origin[0],
                                                           origin[1], origin[2], z0, nz, _ptr(d_hist)))

    def render_brick_alpha(self, d_alpha_seg, w, h, params, brick):
        self._ck(lib().vrdd_render_brick_alpha(self._h, _ptr(d_alpha_seg), w, h, C.byref(params), C.byref(brick)))

    def render_brick_alpha_send(self, seg_tables, flags, brick_index, row0, rows, w, h, params, brick):
        n = len(seg_tables)
        t = (C.c_void_p * n)(*[_ptr(x) for x in seg_tables])
        f = (C.c_void_p * n)(*[_ptr(x) for x in flags])
        self._ck(lib().vrdd_render_brick_alpha_send(self._h, t, f, n, brick_index, row0, rows, w, h, C.byref(params), C.byref(brick)))

    def render_brick_color_send(self, d_alpha_in, root_slots, root_flag, brick_index, row0, rows, w, h, params, brick):
        self._ck(lib().vrdd_render_brick_color_send(self._h, _ptr(d_alpha_in), _ptr(root_slots), _ptr(root_flag), brick_index, row0,
                                                    rows, w, h, C.byref(params), C.byref(brick)))

    def render_brick_color_send_bands(self, d_alpha_in, owner_slots, owner_flags, band_rows, brick_index, row0, rows, w, h, params, brick):
        n = len(owner_slots)
        t = (C.c_void_p * n)(*[_ptr(x) for x in owner_slots])
        f = (C.c_void_p * n)(*[_ptr(x) for x in owner_flags])
        self._ck(lib().vrdd_render_brick_color_send_bands(self._h, _ptr(d_alpha_in), t, f, n, band_rows, brick_index, row0, rows, w, h,
                                                          C.byref(params), C.byref(brick)))

    def pack_band_slots(self, d_slots, nbricks, row0, rows, band_index, band_rows, d_frame, d_frame_flag, w, h, brightness):
        r0 = (C.c_int * nbricks)(*row0)
        self._ck(lib().vrdd_pack_band_slots(self._h, _ptr(d_slots), nbricks, r0, rows, band_index, band_rows, _ptr(d_frame),
                                            _ptr(d_frame_flag), w, h, brightness))

    def pack_frame_slots(self, d_slots, nbricks, row0, rows, d_out, w, h, brightness):
        r0 = (C.c_int * nbricks)(*row0)
        self._ck(lib().vrdd_pack_frame_slots(self._h, _ptr(d_slots), nbricks, r0, rows, _ptr(d_out), w, h, brightness))

    def compose_alpha_in(self, d_alpha_seg_all, grid, q, d_alpha_in, w, h):
        self._ck(lib().vrdd_compose_alpha_in(self._h, _ptr(d_alpha_seg_all), grid[0], grid[1], grid[2], q[0], q[1], q[2],
                                             _ptr(d_alpha_in), w, h))

    def compose_alpha_in_rows(self, d_alpha_seg_rows, grid, q, row0, rows, d_alpha_in, w, h):
        """Row-window form: brick b contributed rows [row0[b], row0[b] + rows) of its pass-1 image."""
        n = grid[0] * grid[1] * grid[2]
        tab = (C.c_int * n)(*[int(v) for v in row0])
        self._ck(lib().vrdd_compose_alpha_in_rows(self._h, _ptr(d_alpha_seg_rows), grid[0], grid[1], grid[2], q[0], q[1],
                                                  q[2], tab, int(rows), _ptr(d_alpha_in), w, h))

    def render_brick_color(self, d_alpha_in, d_partial4, w, h, params, brick):
        self._ck(lib().vrdd_render_brick_color(self._h, _ptr(d_alpha_in), _ptr(d_partial4), w, h, C.byref(params),
                                               C.byref(brick)))

    def pack_frame(self, d_sum4, d_output, w, h, brightness=1.0):
        self._ck(lib().vrdd_pack_frame(self._h, _ptr(d_sum4), _ptr(d_output), w, h, brightness))

    # diagnostics
    def set_variant(self, what, variant):
        self._ck(lib().vrdd_set_variant(self._h, what.encode(), variant.encode()))

    def debug_sample_texture(self, source, comp, d_uvw, n, d_out):
        self._ck(lib().vrdd_debug_sample_texture(self._h, source, comp, _ptr(d_uvw), n, _ptr(d_out)))


    def debug_sample_texture_point(self, source, comp, d_uvw, n, d_out):
        self._ck(lib().vrdd_debug_sample_texture_point(self._h, source, comp, _ptr(d_uvw), n, _ptr(d_out)))

    def debug_sample_texture_unnorm(self, source, comp, d_uvw, n, d_out):
        self._ck(lib().vrdd_debug_sample_texture_unnorm(self._h, source, comp, _ptr(d_uvw), n, _ptr(d_out)))

    def debug_sample_transfer_function(self, d_u, n, d_out4):
        self._ck(lib().vrdd_debug_sample_transfer_function(self._h, _ptr(d_u), n, _ptr(d_out4)))


class legacy:
    """The reference's seven entry points (volumeRender.cpp:156-170), bound as its host code
    would bind them."""

    @staticmethod
    def initCuda(h_volume, volume_size, h_codebook=None, h_templates=None, h_errorsbook=None, bins=32):
        w, h, d = volume_size
        nt = 0 if h_templates is None else h_templates.shape[0]
        lib().initCuda(_ptr(h_volume), Extent(w, h, d), Extent(bins, w * h, d), _ptr(h_codebook), Extent(w, h, d),
                       _ptr(h_templates), Extent(bins, nt, 1), _ptr(h_errorsbook), Extent(bins, w * h, d),
                       *([None] * 9))

    @staticmethod
    def basicDataProcessing():
        lib().basicDataProcessing()

    @staticmethod
    def dataProcessing():
        lib().dataProcessing()

    @staticmethod
    def copyInvViewMatrix(m12):
        arr = (C.c_float * 12)(*[float(v) for v in m12])
        lib().copyInvViewMatrix(arr, C.sizeof(arr))

    @staticmethod
    def render_kernel(d_output, w, h, density=0.05, brightness=1.0, transfer_offset=0.0, transfer_scale=1.0,
                      query_method=1, volume_size=(50, 50, 10), block=(16, 16)):
        grid = Dim3((w + block[0] - 1) // block[0], (h + block[1] - 1) // block[1], 1)
        lib().render_kernel(grid, Dim3(block[0], block[1], 1), _ptr(d_output), w, h, density, brightness,
                            transfer_offset, transfer_scale, query_method, Extent(*volume_size))

    @staticmethod
    def setTextureFilterMode(linear):
        lib().setTextureFilterMode(bool(linear))

    @staticmethod
    def freeCudaBuffers():
        lib().freeCudaBuffers()

    @staticmethod
    def handle():
        return lib().vrdd_legacy_handle()
