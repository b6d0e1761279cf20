"""Shared helpers for the parity tests: fixture loading, the oracle (checker) binding, comparisons."""
import ctypes as C
import functools
import importlib
import os
import struct

import numpy as np

import refdump

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
GOLDEN = os.path.join(ROOT, "tests", "golden")
ssb = importlib.import_module("simple-spectral_b200")
abi = ssb._abi

VARIANT_OPTS = {  # compile-time variants of the reference (stdafx.hpp:66,81) as runtime options
    "ours1931": dict(upsampling=abi.SSB_UPSAMPLE_OURS, lambda_min=380.0, lambda_max=780.0),
    "ours2006": dict(upsampling=abi.SSB_UPSAMPLE_OURS, lambda_min=390.0, lambda_max=830.0),
    "jh": dict(upsampling=abi.SSB_UPSAMPLE_JH, lambda_min=380.0, lambda_max=780.0),
    "meng": dict(upsampling=abi.SSB_UPSAMPLE_MENG, lambda_min=380.0, lambda_max=780.0),
    "ours1931_noels": dict(upsampling=abi.SSB_UPSAMPLE_OURS, lambda_min=380.0, lambda_max=780.0, explicit_light_sampling=0),
    "ours1931_noffc": dict(upsampling=abi.SSB_UPSAMPLE_OURS, lambda_min=380.0, lambda_max=780.0, flat_field_correction=0),
    "rgb": dict(render_mode=abi.SSB_RENDER_RGB),  # RENDER_MODE_RGB (stdafx.hpp:62-90)
    # SAMPLE_WAVELENGTHS 3 / 2 (stdafx.hpp:90)
    "ours1931_nw3": dict(upsampling=abi.SSB_UPSAMPLE_OURS, lambda_min=380.0, lambda_max=780.0, n_wavelengths=3),
    "meng_nw2": dict(upsampling=abi.SSB_UPSAMPLE_MENG, lambda_min=380.0, lambda_max=780.0, n_wavelengths=2),
    "ours1931_d3": dict(upsampling=abi.SSB_UPSAMPLE_OURS, lambda_min=380.0, lambda_max=780.0, max_depth=3),  # MAX_DEPTH 3
}


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("samples", "closest_queries", "shadow_queries", "unshadowed",
                                          "bsdf_samples", "texture_lookups", "tri_tests", "double_fallbacks")]


@functools.lru_cache(None)
def oracle():
    """The CPU checker (oracle/ssb_oracle.c).  TEST INFRASTRUCTURE — never used by the product."""
    path = os.path.join(ROOT, "oracle", "libssb_oracle.so")
    if not os.path.exists(path):
        raise RuntimeError("oracle/libssb_oracle.so missing: run __graft_entry__.build()")
    L = C.CDLL(path)
    P = C.POINTER
    L.ssb_oracle_render.argtypes = [P(abi.ssb_scene), P(abi.ssb_color), P(abi.ssb_options), P(C.c_double), P(C.c_float), P(Counters)]
    L.ssb_oracle_render.restype = C.c_int
    L.ssb_oracle_resolve.argtypes = [P(abi.ssb_color), P(abi.ssb_options), P(C.c_double), P(C.c_double), P(C.c_float)]
    L.ssb_oracle_resolve.restype = C.c_int
    L.ssb_oracle_intersect.argtypes = [P(abi.ssb_scene), P(C.c_float), P(C.c_int32), C.c_float, P(C.c_float), C.c_size_t]
    L.ssb_oracle_intersect.restype = C.c_int
    L.ssb_oracle_eval_math.argtypes = [C.c_uint32, P(C.c_float), C.c_float, P(C.c_float), C.c_size_t]
    L.ssb_oracle_eval_math.restype = None
    return L


def data_root():
    return refdump.data_root()


@functools.lru_cache(None)
def lizard_texture():
    root = data_root()
    if root is None:
        return None
    cache = "/tmp/ssb_lizard4096_rgb8.npy"
    if os.path.exists(cache):
        return np.load(cache)
    a = refdump.load_texture_rgb8(os.path.join(root, "data", "scenes", "crystal-lizard-4096.png"))
    try:
        np.save(cache, a)
    except OSError:
        pass
    return a


@functools.lru_cache(None)
def jh_tables():
    root = data_root()
    if root is None:
        return None
    b = open(os.path.join(root, "data", "jakob-and-hanika-2019-srgb.coeff"), "rb").read()
    assert b[:4] == b"SPEC"
    res = struct.unpack("<I", b[4:8])[0]
    scale = np.frombuffer(b, np.float32, res, 8).copy()
    data = np.frombuffer(b, np.float32, 3 * res ** 3 * 3, 8 + 4 * res).copy()
    return scale, data, res


@functools.lru_cache(None)
def meng_tables():
    root = data_root()
    path = os.path.join(root, "data", "meng-et-al-2015-tables.bin") if root else None
    if not path or not os.path.exists(path):
        return None
    b = open(path, "rb").read()
    assert b[:8] == b"SSBMENG1"
    gw, gh, npnt, ns = struct.unpack("<4I", b[8:24])
    m = abi.ssb_meng_tables()
    xy = struct.unpack("<6f", b[24:48])
    for i in range(6):
        m.xy_to_uv[i] = xy[i]
    m.sample_min, m.sample_max = struct.unpack("<2f", b[48:56])
    grid = np.frombuffer(b, np.int32, gw * gh * 8, 56).copy()
    pts = np.frombuffer(b, np.float32, npnt * (5 + ns), 56 + 4 * gw * gh * 8).copy()
    m.grid = grid.ctypes.data_as(C.POINTER(C.c_int32))
    m.points = pts.ctypes.data_as(C.POINTER(C.c_float))
    m.grid_w, m.grid_h, m.npoints, m.nsamples = gw, gh, npnt, ns
    m._keep = (grid, pts)
    return m


def needs_assets(scene, variant):
    return scene != "cornell" or variant in ("jh", "meng", "meng_nw2")


def have_assets():
    return data_root() is not None


def load_flat(scene, variant):
    """Inputs exactly as the real reference built them (committed dump), + texture/JH/Meng data files."""
    t = refdump.parse(os.path.join(GOLDEN, f"tables_{scene}_{variant}.bin"))
    tex = lizard_texture() if scene != "cornell" else None
    jh = jh_tables() if variant == "jh" else None
    meng = meng_tables() if variant.startswith("meng") else None
    return refdump.flat_from_dump(t, tex, jh=jh, meng=meng)


def options(variant, w, h, spp, **kw):
    o = abi.default_options(w, h, spp)
    for k, v in VARIANT_OPTS[variant].items():
        setattr(o, k, v)
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def oracle_render(flat, opt, want_samples=False, counters=False):
    acc = np.zeros((opt.height, opt.width, 4), np.float64)
    s1 = opt.sample_end or opt.spp
    samples = np.zeros((opt.height, opt.width, s1 - opt.sample_begin, 4), np.float32) if want_samples else None
    cnt = Counters() if counters else None
    rc = oracle().ssb_oracle_render(C.byref(flat.scene), C.byref(flat.color), C.byref(opt),
                                    acc.ctypes.data_as(C.POINTER(C.c_double)),
                                    samples.ctypes.data_as(C.POINTER(C.c_float)) if want_samples else None,
                                    C.byref(cnt) if counters else None)
    assert rc == 0, rc
    return acc, samples, cnt


def oracle_resolve(flat, opt, acc):
    xyza = np.empty_like(acc)
    srgba = np.empty(acc.shape, np.float32)
    rc = oracle().ssb_oracle_resolve(C.byref(flat.color), C.byref(opt), acc.ctypes.data_as(C.POINTER(C.c_double)),
                                     xyza.ctypes.data_as(C.POINTER(C.c_double)), srgba.ctypes.data_as(C.POINTER(C.c_float)))
    assert rc == 0
    return xyza, srgba


def oracle_intersect(scene, rays, ignore=None, eps=1e-3):
    """Scene::intersect of the checker for n rays -> (quad, tri, dist, bary[n,3]) like Context.intersect."""
    rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
    n = rays.shape[0]
    out = np.empty((n, 6), np.float32)
    ign = None if ignore is None else np.ascontiguousarray(ignore, np.int32)
    rc = oracle().ssb_oracle_intersect(C.byref(scene), rays.ctypes.data_as(C.POINTER(C.c_float)),
                                       ign.ctypes.data_as(C.POINTER(C.c_int32)) if ign is not None else None, float(eps),
                                       out.ctypes.data_as(C.POINTER(C.c_float)), n)
    assert rc == 0
    return out[:, 0].view(np.int32).copy(), out[:, 1].view(np.int32).copy(), out[:, 2].copy(), out[:, 3:6].copy()


def gpu_context(flat, device=0):
    ctx = ssb.Context(device)
    ctx.upload_scene(flat.scene)
    if not getattr(flat, "rgb", False):  # RGB mode needs no colour tables
        ctx.upload_color(flat.color)
    return ctx


def rel_err(a, b):
    """per-pixel max-channel relative difference of XYZ (the north-star tolerance metric)."""
    d = np.abs(a[..., :3] - b[..., :3]).max(axis=-1)
    m = np.maximum(np.abs(b[..., :3]).max(axis=-1), 1e-300)
    return d / m


def bits_equal(a, b):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    if a.dtype == np.float32:
        return np.array_equal(a.view(np.uint32), b.view(np.uint32)) or np.array_equal(a, b, equal_nan=True)
    return np.array_equal(a.view(np.uint64), b.view(np.uint64)) or np.array_equal(a, b, equal_nan=True)


def smoke():
    """__graft_entry__.smoke(): tiny cornell render on cuda:0 vs the oracle, bit-for-bit."""
    scene = "cornell-srgb" if have_assets() else "cornell"
    flat = load_flat(scene, "ours1931")
    opt = options("ours1931", 32, 24, 4, seed=7)
    acc_o, _, _ = oracle_render(flat, opt)
    xo, so = oracle_resolve(flat, opt, acc_o)
    with gpu_context(flat) as ctx:
        xg, sg = ctx.render_frame(opt)
        st = ctx.stats()
    assert bits_equal(xg, xo), f"GPU XYZA differs from oracle: max rel {rel_err(xg, xo).max()}"
    assert np.allclose(sg, so, rtol=1e-6, atol=1e-7, equal_nan=True)
    ref = np.load(os.path.join(GOLDEN, f"xyza_{scene}_ours1931_32x24_spp4_seed7.npy"))
    assert bits_equal(xg, ref), "GPU XYZA differs from the committed reference fixture"
    print(f"smoke: {scene} 32x24 spp4 bit-identical to oracle and reference fixture; "
          f"{st.samples} samples, {st.device_ms:.3f} ms, {st.launches} launches")
