"""Test helper: parse the table dumps written by the hooked reference (oracle/ref_hooks.hpp,
SSB_DUMP_TABLES) and turn them into the flat POD structs of include/ssb200.h.

This gives the oracle (and the CUDA path) inputs that come *straight from the real reference's
own Color::init / Scene::get_new_* code*, independently of the product's host layer — which is
in turn checked against the same dumps."""
import ctypes as C
import importlib
import os

import numpy as np

_abi = importlib.import_module("simple-spectral_b200._abi")


def parse(path_or_bytes):
    b = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray)) else open(path_or_bytes, "rb").read()
    pos, out = 0, {}
    dts = {"f32": np.float32, "f64": np.float64, "u32": np.uint32, "u8": np.uint8}
    while pos < len(b):
        nl = b.index(b"\n", pos)
        hdr = b[pos:nl].decode().split()
        pos = nl + 1
        assert hdr[0] == "REC", hdr
        dt = dts[hdr[2]]
        n = int(hdr[3])
        size = np.dtype(dt).itemsize * n
        out[hdr[1]] = np.frombuffer(b[pos:pos + size], dtype=dt).copy()
        pos += size
    return out


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class Flat:
    """Owns numpy buffers + the ctypes structs that point into them."""

    def __init__(self):
        self.keep = []
        self.scene = _abi.ssb_scene()
        self.color = _abi.ssb_color()
        self.lambda_min = 380.0
        self.lambda_max = 780.0

    def spectrum(self, data, low, high, filt=0):
        a = np.ascontiguousarray(data, dtype=np.float32)
        self.keep.append(a)
        s = _abi.ssb_spectrum()
        s.data, s.n, s.low, s.high, s.filter = _fp(a), a.size, float(low), float(high), filt
        return s


def load_texture_rgb8(path):
    from PIL import Image
    Image.MAX_IMAGE_PIXELS = None
    im = Image.open(path).convert("RGB")
    a = np.ascontiguousarray(np.asarray(im, dtype=np.uint8))
    return a


def flat_from_dump(t, texture_rgb8=None, jh=None, meng=None):
    """t: parse() result.  texture_rgb8: HxWx3 uint8 (required if a material is textured)."""
    f = Flat()
    col = f.color
    f.rgb = "color.xbar.data" not in t  # dump of the RENDER_MODE_RGB build: no colour tables, RGB material constants
    if not f.rgb:
        for name in ("xbar", "ybar", "zbar"):
            setattr(col, name, f.spectrum(t[f"color.{name}.data"], *t[f"color.{name}.lowhigh"]))
        if "color.basis_r.data" in t:
            for name in ("basis_r", "basis_g", "basis_b"):
                setattr(col, name, f.spectrum(t[f"color.{name}.data"], *t[f"color.{name}.lowhigh"]))
        for i in range(9):
            col.xyz_to_lrgb[i] = float(t["color.matr_xyz_to_lrgb"][i])
        col.d65_rad_Y = float(t["color.D65_rad_XYZ"][1])
    if jh is not None:
        scale, data, res = jh
        f.keep += [scale, data]
        col.jh_scale, col.jh_data, col.jh_res = _fp(scale), _fp(data), res
    if meng is not None:
        f.keep.append(meng)
        col.meng = C.pointer(meng)
    if not f.rgb:
        f.lambda_min = float(t["config.lambda_min_max_step"][0])
        f.lambda_max = float(t["config.lambda_min_max_step"][1])

    sc = f.scene
    for i in range(16):
        sc.camera.pv_inv[i] = float(t["camera.matr_PV_inv"][i])
    for i in range(3):
        sc.camera.pos[i] = float(t["camera.pos"][i])
        sc.camera.dir[i] = float(t["camera.dir"][i])
    q = t["scene.quads"].reshape(-1, 2, 18)
    nq = q.shape[0]
    quads = (_abi.ssb_quad * nq)()
    for qi in range(nq):
        for ti in range(2):
            for vi in range(3):
                for k in range(3):
                    quads[qi].tri[ti].v[vi].pos[k] = float(q[qi, ti, vi * 5 + k])
                for k in range(2):
                    quads[qi].tri[ti].v[vi].st[k] = float(q[qi, ti, vi * 5 + 3 + k])
            for k in range(3):
                quads[qi].tri[ti].normal[k] = float(q[qi, ti, 15 + k])
        quads[qi].material = int(t["scene.quad_material"][qi])
        quads[qi].is_light = int(t["scene.quad_is_light"][qi])
    nm = 0
    while f"material.{nm}.kind_mode" in t:
        nm += 1
    mats = (_abi.ssb_material * nm)()
    textures = []
    for m in range(nm):
        kind, mode = t[f"material.{m}.kind_mode"]
        mats[m].kind, mats[m].albedo_mode = int(kind), int(mode)
        if f.rgb:
            for k in range(3):
                mats[m].emission_rgb[k] = float(t[f"material.{m}.emission_rgb"][k])
                if mode == 0:
                    mats[m].albedo_rgb[k] = float(t[f"material.{m}.albedo_rgb"][k])
        else:
            mats[m].emission = f.spectrum(t[f"material.{m}.emission.data"], *t[f"material.{m}.emission.lowhigh"])
            if mode == 0:
                mats[m].albedo = f.spectrum(t[f"material.{m}.albedo.data"], *t[f"material.{m}.albedo.lowhigh"])
        if mode != 0:
            assert texture_rgb8 is not None, "scene has a textured material: pass texture_rgb8"
            w, h = t[f"material.{m}.texture_res"]
            assert texture_rgb8.shape == (h, w, 3)
            tex = _abi.ssb_texture()
            tex.rgb8 = texture_rgb8.ctypes.data_as(C.POINTER(C.c_uint8))
            tex.width, tex.height = int(w), int(h)
            mats[m].texture = len(textures)
            textures.append(tex)
    texs = (_abi.ssb_texture * max(1, len(textures)))(*textures)
    f.keep += [quads, mats, texs, texture_rgb8]
    sc.quads, sc.nquads = quads, nq
    sc.materials, sc.nmaterials = mats, nm
    sc.textures, sc.ntextures = texs, len(textures)
    return f


def data_root():
    """Directory that contains data/ (the reference's cwd-relative layout)."""
    here = os.path.dirname(os.path.abspath(__file__))
    for cand in (os.environ.get("SSB_DATA_ROOT"), os.path.join(here, "..", "assets"), "/root/reference"):
        if cand and os.path.isdir(os.path.join(cand, "data")):
            return os.path.abspath(cand)
    return None
