"""simple-spectral_b200 — B200-native replacement for simple-spectral's per-pixel Monte-Carlo hot path.

The product is the CUDA library `libssb200.so` (C ABI: include/ssb200.h) built from csrc/.  This
package is a thin ctypes binding plus a host-side mirror of the reference's Scene / Color /
Renderer interface (host.py).  There is no CPU fallback: if the CUDA library is missing or no
GPU is present, calls fail loudly.
"""
import ctypes as C
import os

from . import _abi
from ._abi import *  # noqa: F401,F403  (POD structs + constants)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libssb200.so")
_lib = None


class SsbError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"ssb200 error {code}: {message}")
        self.code = code


def lib():
    """Load libssb200.so (built in-tree by __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    P = C.POINTER
    L.ssb_abi_version.restype = C.c_uint32
    L.ssb_last_error.restype = C.c_char_p
    L.ssb_default_options.argtypes = [P(_abi.ssb_options), C.c_uint32, C.c_uint32, C.c_uint32]
    L.ssb_default_options.restype = None
    L.ssb_create.argtypes = [C.c_int, P(C.c_void_p)]
    L.ssb_destroy.argtypes = [C.c_void_p]
    L.ssb_destroy.restype = None
    L.ssb_upload_scene.argtypes = [C.c_void_p, P(_abi.ssb_scene)]
    L.ssb_upload_scene_async.argtypes = [C.c_void_p, P(_abi.ssb_scene)]
    L.ssb_upload_color.argtypes = [C.c_void_p, P(_abi.ssb_color)]
    L.ssb_render.argtypes = [C.c_void_p, P(_abi.ssb_options)]
    L.ssb_clear.argtypes = [C.c_void_p]
    L.ssb_read_accum.argtypes = [C.c_void_p, P(C.c_double)]
    L.ssb_write_accum.argtypes = [C.c_void_p, P(C.c_double)]
    L.ssb_accum_device.argtypes = [C.c_void_p, P(C.c_void_p), P(C.c_size_t)]
    L.ssb_resolve.argtypes = [C.c_void_p, P(_abi.ssb_options), P(C.c_double), P(C.c_float)]
    L.ssb_resolve_device.argtypes = [C.c_void_p, P(_abi.ssb_options), P(C.c_void_p), P(C.c_void_p)]
    L.ssb_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    L.ssb_render_frame.argtypes = [C.c_void_p, P(_abi.ssb_options), P(C.c_double), P(C.c_float)]
    L.ssb_get_stats.argtypes = [C.c_void_p, P(_abi.ssb_stats)]
    L.ssb_synchronize.argtypes = [C.c_void_p]
    L.ssb_debug_eval_math.argtypes = [C.c_void_p, C.c_uint32, P(C.c_float), C.c_float, P(C.c_float), C.c_size_t]
    L.ssb_debug_trace_samples.argtypes = [C.c_void_p, P(_abi.ssb_options), C.c_uint32, C.c_uint32, P(C.c_float)]
    L.ssb_device_count.argtypes = [P(C.c_int)]
    L.ssb_accum_merge.argtypes = [C.c_void_p, C.c_void_p, P(_abi.ssb_options)]
    L.ssb_debug_intersect.argtypes = [C.c_void_p, P(C.c_float), P(C.c_int32), C.c_uint32, C.c_float, P(C.c_float), C.c_size_t]
    for name in ("ssb_create", "ssb_upload_scene", "ssb_upload_scene_async", "ssb_upload_color", "ssb_render", "ssb_clear", "ssb_read_accum",
                 "ssb_write_accum", "ssb_accum_device", "ssb_resolve", "ssb_resolve_device", "ssb_set_stream",
                 "ssb_render_frame", "ssb_get_stats", "ssb_synchronize", "ssb_debug_eval_math", "ssb_debug_trace_samples",
                 "ssb_device_count", "ssb_accum_merge", "ssb_debug_intersect"):
        getattr(L, name).restype = C.c_int
    _lib = L
    return L


EXPORTED_SYMBOLS = (
    "ssb_abi_version", "ssb_last_error", "ssb_default_options", "ssb_create", "ssb_destroy", "ssb_upload_scene",
    "ssb_upload_scene_async", "ssb_upload_color", "ssb_render", "ssb_clear", "ssb_read_accum", "ssb_write_accum", "ssb_accum_device",
    "ssb_resolve", "ssb_resolve_device", "ssb_set_stream", "ssb_render_frame", "ssb_get_stats", "ssb_synchronize", "ssb_debug_eval_math",
    "ssb_debug_trace_samples", "ssb_device_count", "ssb_accum_merge", "ssb_debug_intersect",
)


def check(code):
    if code != 0:
        raise SsbError(code, lib().ssb_last_error().decode(errors="replace"))


def device_count():
    n = C.c_int()
    check(lib().ssb_device_count(C.byref(n)))
    return n.value


class Context:
    """Owns one ssb_ctx (one per GPU)."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        check(lib().ssb_create(device, C.byref(self._h)))

    def close(self):
        if self._h:
            lib().ssb_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload_scene(self, scene):
        check(lib().ssb_upload_scene(self._h, C.byref(scene)))

    def upload_scene_async(self, scene):
        """Texel copies are only enqueued (they overlap the next render's camera-ray stage): keep the texture buffers
        alive and unmodified until a synchronising call (render_frame / resolve / synchronize) has returned."""
        check(lib().ssb_upload_scene_async(self._h, C.byref(scene)))

    def upload_color(self, color):
        check(lib().ssb_upload_color(self._h, C.byref(color)))

    def render(self, opt):
        check(lib().ssb_render(self._h, C.byref(opt)))

    def clear(self):
        check(lib().ssb_clear(self._h))

    def read_accum(self, width, height):
        import numpy as np
        a = np.empty((height, width, 4), np.float64)
        check(lib().ssb_read_accum(self._h, a.ctypes.data_as(C.POINTER(C.c_double))))
        return a

    def write_accum(self, a):
        import numpy as np
        a = np.ascontiguousarray(a, np.float64)
        check(lib().ssb_write_accum(self._h, a.ctypes.data_as(C.POINTER(C.c_double))))

    def accum_device(self):
        p, n = C.c_void_p(), C.c_size_t()
        check(lib().ssb_accum_device(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def resolve(self, opt, want_xyza=True, want_srgba=True, xyza=None, srgba=None):
        import numpy as np
        if want_xyza and xyza is None:
            xyza = np.empty((opt.height, opt.width, 4), np.float64)
        if want_srgba and srgba is None:
            srgba = np.empty((opt.height, opt.width, 4), np.float32)
        check(lib().ssb_resolve(self._h, C.byref(opt),
                                xyza.ctypes.data_as(C.POINTER(C.c_double)) if want_xyza else None,
                                srgba.ctypes.data_as(C.POINTER(C.c_float)) if want_srgba else None))
        return xyza, srgba

    def resolve_device(self, opt):
        x, s = C.c_void_p(), C.c_void_p()
        check(lib().ssb_resolve_device(self._h, C.byref(opt), C.byref(x), C.byref(s)))
        return x.value, s.value

    def set_stream(self, cuda_stream):
        check(lib().ssb_set_stream(self._h, C.c_void_p(cuda_stream) if cuda_stream else None))

    def render_frame(self, opt, xyza=None, srgba=None, want_xyza=True, want_srgba=True):
        import numpy as np
        if want_xyza and xyza is None:
            xyza = np.empty((opt.height, opt.width, 4), np.float64)
        if want_srgba and srgba is None:
            srgba = np.empty((opt.height, opt.width, 4), np.float32)
        check(lib().ssb_render_frame(self._h, C.byref(opt),
                                     xyza.ctypes.data_as(C.POINTER(C.c_double)) if want_xyza else None,
                                     srgba.ctypes.data_as(C.POINTER(C.c_float)) if want_srgba else None))
        return xyza, srgba

    def stats(self):
        s = _abi.ssb_stats()
        check(lib().ssb_get_stats(self._h, C.byref(s)))
        return s

    def synchronize(self):
        check(lib().ssb_synchronize(self._h))

    def merge_from(self, src, src_opt):
        """ssb_accum_merge: fold the accumulator of `src` (another Context, possibly on another GPU) into this one;
        `src_opt` = the options src rendered its share with (pixel subset: copied; sample range: added)."""
        check(lib().ssb_accum_merge(self._h, src._h, C.byref(src_opt)))

    def intersect(self, rays, ignore=None, scan_mode=0, eps=1e-3):
        """ssb_debug_intersect: closest hits of n rays (n x 6: origin, direction) -> (quad, tri, dist, bary[n,3])."""
        import numpy as np
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        n = rays.shape[0]
        out = np.empty((n, 6), np.float32)
        ign = None if ignore is None else np.ascontiguousarray(ignore, np.int32)
        check(lib().ssb_debug_intersect(self._h, rays.ctypes.data_as(C.POINTER(C.c_float)),
                                        ign.ctypes.data_as(C.POINTER(C.c_int32)) if ign is not None else None,
                                        int(scan_mode), float(eps), out.ctypes.data_as(C.POINTER(C.c_float)), n))
        return out[:, 0].view(np.int32).copy(), out[:, 1].view(np.int32).copy(), out[:, 2].copy(), out[:, 3:6].copy()

    def eval_math(self, fn, x, arg=0.0):
        import numpy as np
        x = np.ascontiguousarray(x, np.float32)
        out = np.empty_like(x)
        check(lib().ssb_debug_eval_math(self._h, fn, x.ctypes.data_as(C.POINTER(C.c_float)), float(arg),
                                        out.ctypes.data_as(C.POINTER(C.c_float)), x.size))
        return out

    def trace_samples(self, opt, px, py):
        import numpy as np
        s1 = opt.sample_end or opt.spp
        out = np.empty((s1 - opt.sample_begin, 4), np.float32)
        check(lib().ssb_debug_trace_samples(self._h, C.byref(opt), px, py, out.ctypes.data_as(C.POINTER(C.c_float))))
        return out
