"""ctypes binding of the C++ host layer (include/ssb200_host.h): the reference-shaped Color / Scene / Renderer
surface.  Loading data files, building scenes and flattening them happens in C++ (csrc/host); rendering happens
in the CUDA path.  Nothing here computes."""
import ctypes as C
import os

from . import _abi, lib as _lib, SsbError

_bound = False


class ssbh_renderer_options(C.Structure):
    _fields_ = [("scene_name", C.c_char_p), ("width", C.c_uint32), ("height", C.c_uint32), ("spp", C.c_uint32),
                ("indirect_only", C.c_uint32), ("output_path", C.c_char_p), ("observer", C.c_int),
                ("upsampling", C.c_uint32), ("explicit_light_sampling", C.c_uint32), ("max_depth", C.c_uint32),
                ("flat_field_correction", C.c_uint32), ("seed", C.c_uint64), ("device", C.c_int),
                ("data_root", C.c_char_p), ("render_mode", C.c_uint32), ("n_wavelengths", C.c_uint32),
                ("prebaked_textures", C.c_uint32), ("progressive", C.c_uint32),
                ("devices", C.POINTER(C.c_int)), ("ndevices", C.c_uint32), ("shard", C.c_uint32), ("band_height", C.c_uint32)]


SSBH_SHARD_TILES, SSBH_SHARD_SAMPLES = 0, 1


HOST_SYMBOLS = (
    "ssbh_last_error", "ssbh_color_init", "ssbh_color_flat", "ssbh_color_query", "ssbh_color_spectrum", "ssbh_color_free",
    "ssbh_scene_new", "ssbh_scene_flat", "ssbh_scene_camera", "ssbh_scene_free", "ssbh_load_png_rgb8", "ssbh_free",
    "ssbh_save_image", "ssbh_renderer_new", "ssbh_renderer_render", "ssbh_renderer_framebuffer", "ssbh_renderer_xyza",
    "ssbh_renderer_stats", "ssbh_renderer_free", "ssbh_renderer_start", "ssbh_renderer_stop", "ssbh_renderer_wait",
    "ssbh_renderer_is_rendering", "ssbh_renderer_snapshot", "ssbh_color_round_trip_srgb", "ssbh_color_round_trip_running_max",
)


def hostlib():
    global _bound
    L = _lib()
    if _bound:
        return L
    P = C.POINTER
    L.ssbh_last_error.restype = C.c_char_p
    L.ssbh_color_init.argtypes = [C.c_char_p, C.c_int, C.c_uint32, P(C.c_void_p)]
    L.ssbh_color_flat.argtypes = [C.c_void_p]
    L.ssbh_color_flat.restype = P(_abi.ssb_color)
    L.ssbh_color_query.argtypes = [C.c_void_p] + [P(C.c_float)] * 5
    L.ssbh_color_spectrum.argtypes = [C.c_void_p, C.c_char_p, P(_abi.ssb_spectrum)]
    L.ssbh_color_round_trip_srgb.argtypes = [C.c_void_p, P(C.c_float), P(C.c_float)]
    L.ssbh_color_round_trip_srgb.restype = C.c_int
    L.ssbh_color_round_trip_running_max.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_float, P(C.c_float), C.c_uint32]
    L.ssbh_color_round_trip_running_max.restype = C.c_int
    L.ssbh_color_free.argtypes = [C.c_void_p]
    L.ssbh_color_free.restype = None
    L.ssbh_scene_new.argtypes = [C.c_char_p, C.c_char_p, C.c_void_p, C.c_int, P(C.c_void_p)]
    L.ssbh_scene_flat.argtypes = [C.c_void_p]
    L.ssbh_scene_flat.restype = P(_abi.ssb_scene)
    L.ssbh_scene_camera.argtypes = [C.c_void_p, P(C.c_double), P(C.c_double), P(C.c_double)]
    L.ssbh_scene_free.argtypes = [C.c_void_p]
    L.ssbh_scene_free.restype = None
    L.ssbh_load_png_rgb8.argtypes = [C.c_char_p, P(P(C.c_uint8)), P(C.c_uint32), P(C.c_uint32)]
    L.ssbh_free.argtypes = [C.c_void_p]
    L.ssbh_free.restype = None
    L.ssbh_save_image.argtypes = [C.c_char_p, P(C.c_float), C.c_uint32, C.c_uint32]
    L.ssbh_renderer_new.argtypes = [P(ssbh_renderer_options), P(C.c_void_p)]
    L.ssbh_renderer_render.argtypes = [C.c_void_p]
    L.ssbh_renderer_framebuffer.argtypes = [C.c_void_p]
    L.ssbh_renderer_framebuffer.restype = P(C.c_float)
    L.ssbh_renderer_xyza.argtypes = [C.c_void_p]
    L.ssbh_renderer_xyza.restype = P(C.c_double)
    L.ssbh_renderer_stats.argtypes = [C.c_void_p, P(_abi.ssb_stats)]
    L.ssbh_renderer_free.argtypes = [C.c_void_p]
    L.ssbh_renderer_free.restype = None
    L.ssbh_renderer_start.argtypes = [C.c_void_p]
    L.ssbh_renderer_start.restype = C.c_int
    L.ssbh_renderer_stop.argtypes = [C.c_void_p]
    L.ssbh_renderer_stop.restype = None
    L.ssbh_renderer_wait.argtypes = [C.c_void_p]
    L.ssbh_renderer_wait.restype = C.c_int
    L.ssbh_renderer_is_rendering.argtypes = [C.c_void_p]
    L.ssbh_renderer_is_rendering.restype = C.c_int
    L.ssbh_renderer_snapshot.argtypes = [C.c_void_p, P(C.c_float)]
    L.ssbh_renderer_snapshot.restype = C.c_uint32
    for n in ("ssbh_color_init", "ssbh_color_query", "ssbh_color_spectrum", "ssbh_scene_new", "ssbh_scene_camera",
              "ssbh_load_png_rgb8", "ssbh_save_image", "ssbh_renderer_new", "ssbh_renderer_render", "ssbh_renderer_stats"):
        getattr(L, n).restype = C.c_int
    _bound = True
    return L


def _check(rc):
    if rc != 0:
        raise SsbError(rc, hostlib().ssbh_last_error().decode(errors="replace"))


def find_data_root():
    """Directory containing the reference's data/ tree ($SSB_DATA_ROOT, ./, <repo>/assets, /root/reference)."""
    here = os.path.dirname(os.path.abspath(__file__))
    for cand in (os.environ.get("SSB_DATA_ROOT"), os.getcwd(), os.path.join(here, "..", "assets"), "/root/reference"):
        if cand and os.path.isfile(os.path.join(cand, "data", "d65-300+5+780.csv")):
            return os.path.abspath(cand)
    raise SsbError(-1, "reference data files not found: set SSB_DATA_ROOT or run __graft_entry__.build() (stages assets/data)")


VARIANTS = {  # the reference's compile-time variants (stdafx.hpp:66,81)
    "ours1931": (1931, _abi.SSB_UPSAMPLE_OURS), "ours2006": (2006, _abi.SSB_UPSAMPLE_OURS),
    "jh": (1931, _abi.SSB_UPSAMPLE_JH), "meng": (1931, _abi.SSB_UPSAMPLE_MENG),
    "rgb": (1931, _abi.SSB_UPSAMPLE_OURS),  # RENDER_MODE_RGB: the colour tables are built but not used by the render
}


class Color:
    """Color::init() result."""

    def __init__(self, data_root=None, observer=1931, upsampling=_abi.SSB_UPSAMPLE_OURS):
        self.data_root = data_root or find_data_root()
        self._h = C.c_void_p()
        _check(hostlib().ssbh_color_init(self.data_root.encode(), observer, upsampling, C.byref(self._h)))
        self.observer, self.upsampling = observer, upsampling
        lm = (C.c_float * 2)()
        self._orig, self._rad = (C.c_float * 3)(), (C.c_float * 3)()
        self._m, self._mi = (C.c_float * 9)(), (C.c_float * 9)()
        _check(hostlib().ssbh_color_query(self._h, lm, self._orig, self._rad, self._m, self._mi))
        self.lambda_min, self.lambda_max = lm[0], lm[1]

    @property
    def flat(self):
        return hostlib().ssbh_color_flat(self._h).contents

    def spectrum(self, name):
        import numpy as np
        s = _abi.ssb_spectrum()
        _check(hostlib().ssbh_color_spectrum(self._h, name.encode(), C.byref(s)))
        return np.ctypeslib.as_array(s.data, shape=(s.n,)).copy(), s.low, s.high

    def round_trip_srgb(self, srgb):
        """Color::round_trip_srgb (color.cpp:290-294)."""
        a, out = (C.c_float * 3)(*srgb), (C.c_float * 3)()
        _check(hostlib().ssbh_color_round_trip_srgb(self._h, a, out))
        return tuple(out)

    def round_trip_running_max(self, r_begin, r_end, start_max=0.0, threads=0):
        """The reference's round-trip self-test (main.cpp:246-262) for red levels [r_begin, r_end)."""
        import numpy as np
        out = np.empty(r_end - r_begin, np.float32)
        _check(hostlib().ssbh_color_round_trip_running_max(self._h, r_begin, r_end, start_max, out.ctypes.data_as(C.POINTER(C.c_float)),
                                                           threads or (os.cpu_count() or 1)))
        return out

    def close(self):
        if self._h:
            hostlib().ssbh_color_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Scene:
    """Scene::get_new_cornell / get_new_cornell_srgb / get_new_plane_srgb."""

    def __init__(self, name, color, explicit_light_sampling=True):
        self.color = color
        self._h = C.c_void_p()
        _check(hostlib().ssbh_scene_new(name.encode(), color.data_root.encode(), color._h, int(explicit_light_sampling), C.byref(self._h)))
        self.name = name

    @property
    def flat(self):
        return hostlib().ssbh_scene_flat(self._h).contents

    def camera_matrices(self):
        import numpy as np
        P, V, I = (np.empty(16, np.float64) for _ in range(3))
        dp = C.POINTER(C.c_double)
        _check(hostlib().ssbh_scene_camera(self._h, P.ctypes.data_as(dp), V.ctypes.data_as(dp), I.ctypes.data_as(dp)))
        return P, V, I

    def close(self):
        if self._h:
            hostlib().ssbh_scene_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def options_for(color, width, height, spp, **kw):
    o = _abi.default_options(width, height, spp)
    o.upsampling = color.upsampling
    o.lambda_min, o.lambda_max = color.lambda_min, color.lambda_max
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def load_png_rgb8(path):
    import numpy as np
    p, w, h = C.POINTER(C.c_uint8)(), C.c_uint32(), C.c_uint32()
    _check(hostlib().ssbh_load_png_rgb8(path.encode(), C.byref(p), C.byref(w), C.byref(h)))
    a = np.ctypeslib.as_array(p, shape=(h.value, w.value, 3)).copy()
    hostlib().ssbh_free(p)
    return a


def save_image(path, srgba):
    import numpy as np
    a = np.ascontiguousarray(srgba, np.float32)
    _check(hostlib().ssbh_save_image(path.encode(), a.ctypes.data_as(C.POINTER(C.c_float)), a.shape[1], a.shape[0]))


class Renderer:
    """Renderer(options): same knobs as the reference's CLI + its compile-time macros; render() = render_start()+render_wait()."""

    def __init__(self, scene_name, width, height, spp, output_path=None, indirect_only=False, variant="ours1931",
                 explicit_light_sampling=True, max_depth=10, flat_field_correction=True, seed=1, device=0, data_root=None,
                 n_wavelengths=4, prebaked_textures=False, progressive=False, devices=None, shard="tiles", band_height=0):
        """devices: list of GPU indices to render on together (None: just `device`); shard: "tiles" (interleaved row
        bands, bit-identical to one GPU) or "samples" (sample ranges, f64 summation order differs)."""
        obs, ups = VARIANTS[variant]
        self._keep = (scene_name.encode(), output_path.encode() if output_path else None, (data_root or find_data_root()).encode())
        o = ssbh_renderer_options(self._keep[0], width, height, spp, int(indirect_only), self._keep[1], obs, ups,
                                  int(explicit_light_sampling), max_depth, int(flat_field_correction), seed, device, self._keep[2],
                                  _abi.SSB_RENDER_RGB if variant == "rgb" else _abi.SSB_RENDER_SPECTRAL, n_wavelengths,
                                  int(prebaked_textures), int(progressive))
        if devices:
            self._devices = (C.c_int * len(devices))(*devices)
            o.devices, o.ndevices = self._devices, len(devices)
        o.shard = {"tiles": SSBH_SHARD_TILES, "samples": SSBH_SHARD_SAMPLES}[shard]
        o.band_height = band_height
        self._h = C.c_void_p()
        self.width, self.height = width, height
        _check(hostlib().ssbh_renderer_new(C.byref(o), C.byref(self._h)))

    # the reference's asynchronous life cycle (renderer.hpp:71-81): start / poll / stop / wait
    def start(self):
        _check(hostlib().ssbh_renderer_start(self._h))

    def stop(self):
        hostlib().ssbh_renderer_stop(self._h)

    def wait(self):
        _check(hostlib().ssbh_renderer_wait(self._h))
        return self._results()

    def is_rendering(self):
        return bool(hostlib().ssbh_renderer_is_rendering(self._h))

    def snapshot(self):
        """(samples per pixel behind the image, sRGBA copy of the framebuffer) — safe while rendering."""
        import numpy as np
        fb = np.empty((self.height, self.width, 4), np.float32)
        done = hostlib().ssbh_renderer_snapshot(self._h, fb.ctypes.data_as(C.POINTER(C.c_float)))
        return done, fb

    def render(self):
        _check(hostlib().ssbh_renderer_render(self._h))
        return self._results()

    def _results(self):
        import numpy as np
        n = self.width * self.height * 4
        fb = np.ctypeslib.as_array(hostlib().ssbh_renderer_framebuffer(self._h), shape=(n,)).reshape(self.height, self.width, 4).copy()
        xyza = np.ctypeslib.as_array(hostlib().ssbh_renderer_xyza(self._h), shape=(n,)).reshape(self.height, self.width, 4).copy()
        return xyza, fb

    def stats(self):
        s = _abi.ssb_stats()
        _check(hostlib().ssbh_renderer_stats(self._h, C.byref(s)))
        return s

    def close(self):
        if self._h:
            hostlib().ssbh_renderer_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
