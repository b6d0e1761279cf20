"""GPU: runtime options and input shapes the reference's three hard-coded scenes never exercise — CUDA path vs the
oracle, bit for bit (the oracle itself is pinned against builds of the real reference for every option that the
reference can be compiled with: tests/test_oracle_golden.py).

Covered here: MAX_DEPTH 1..7 with and without explicit light sampling, FLAT_FIELD_CORRECTION off, the nearest
spectrum filter, a mirror material under explicit light sampling, more than 32 quads (the intersect stage scans the
list in blocks of 32), a non-planar and a degenerate quad (conservative filter must keep them), odd image sizes with
pixel-rectangle / sample-range subsets, several texture sizes."""
import ctypes as C

import numpy as np
import pytest

import parity_util as pu

pytestmark = pytest.mark.gpu
abi = pu.abi


def _compare(flat, opt, pixels=((0, 0),)):
    acc_o, samp_o, _ = pu.oracle_render(flat, opt, want_samples=True)
    xo, so = pu.oracle_resolve(flat, opt, acc_o)
    with pu.gpu_context(flat) as ctx:
        xg, sg = ctx.render_frame(opt)
        acc_g = ctx.read_accum(opt.width, opt.height)
        for (px, py) in pixels:
            assert pu.bits_equal(ctx.trace_samples(opt, px, py), samp_o[py, px]), f"per-sample mismatch at ({px},{py})"
    assert pu.bits_equal(acc_g, acc_o), f"accumulator differs: max rel {pu.rel_err(acc_g, acc_o).max()}"
    assert pu.bits_equal(xg, xo) and pu.bits_equal(sg, so)
    return xg


def _need_assets():
    if not pu.have_assets():
        pytest.fail("data files not staged on the GPU box (assets/data): run __graft_entry__.build() first")


@pytest.mark.parametrize("depth", [1, 2, 3, 7])
@pytest.mark.parametrize("els", [1, 0])
def test_max_depth(depth, els):
    flat = pu.load_flat("cornell", "ours1931")
    opt = pu.options("ours1931", 24, 20, 3, seed=11, max_depth=depth, explicit_light_sampling=els)
    _compare(flat, opt, pixels=((12, 10), (3, 17)))


def test_flat_field_correction_off():
    """renderer.cpp:262-266: flux = radiance * dot(camera ray, camera dir) — here CUDA vs oracle with per-sample values;
    the frame of the real reference's build of this configuration: tests/test_zz_gpu_prebake_progressive.py."""
    flat = pu.load_flat("cornell", "ours1931")
    on = _compare(flat, pu.options("ours1931", 24, 20, 3, seed=11))
    off = _compare(flat, pu.options("ours1931", 24, 20, 3, seed=11, flat_field_correction=0), pixels=((12, 10),))
    assert (off[..., :3] <= on[..., :3] + 1e-12).all() and (off[..., 1].sum() < on[..., 1].sum())


def test_nearest_spectrum_filter():
    """_Spectrum::set_filter_nearest (spectrum.hpp:44, spectrum.cpp:29-38) on every spectrum of the scene and the observer."""
    _need_assets()
    flat = pu.load_flat("cornell-srgb", "ours1931")
    for name in ("xbar", "ybar", "zbar", "basis_r", "basis_g", "basis_b"):
        getattr(flat.color, name).filter = abi.SSB_FILTER_NEAREST
    for m in range(flat.scene.nmaterials):
        flat.scene.materials[m].emission.filter = abi.SSB_FILTER_NEAREST
        flat.scene.materials[m].albedo.filter = abi.SSB_FILTER_NEAREST
    lin = pu.load_flat("cornell-srgb", "ours1931")
    opt = pu.options("ours1931", 24, 20, 3, seed=11)
    got = _compare(flat, opt, pixels=((12, 10),))
    ref, _ = pu.oracle_resolve(lin, opt, pu.oracle_render(lin, opt)[0])
    assert not pu.bits_equal(got, ref), "nearest filter had no effect"


def test_mirror_material_with_explicit_light_sampling():
    """MaterialMirror (material.cpp:146-167) on the Cornell blocks while explicit light sampling stays on:
    evaluate_bsdf = 0 for the light sample, delta reflection for the path."""
    flat = pu.load_flat("cornell", "ours1931")
    block_mat = flat.scene.quads[flat.scene.nquads - 1].material
    flat.scene.materials[block_mat].kind = abi.SSB_MATERIAL_MIRROR
    _compare(flat, pu.options("ours1931", 32, 24, 4, seed=13), pixels=((12, 8), (20, 6)))
    _compare(flat, pu.options("ours1931", 32, 24, 2, seed=13, explicit_light_sampling=0))


def _with_quads(flat, new_quads):
    """Replace the quad list (ctypes array) of a Flat, keeping materials / camera."""
    arr = (abi.ssb_quad * len(new_quads))()
    for i, q in enumerate(new_quads):
        C.memmove(C.byref(arr[i]), C.byref(q), C.sizeof(abi.ssb_quad))
    flat.keep.append(arr)
    flat.scene.quads, flat.scene.nquads = arr, len(new_quads)
    return flat


def _copy_quad(q):
    c = abi.ssb_quad()
    C.memmove(C.byref(c), C.byref(q), C.sizeof(abi.ssb_quad))
    return c


def _translated(q, dx, dy, dz, scale=1.0, about=(0.0, 0.0, 0.0)):
    c = _copy_quad(q)
    for t in range(2):
        for v in range(3):
            p = c.tri[t].v[v].pos
            for k, d in enumerate((dx, dy, dz)):
                p[k] = np.float32((np.float32(p[k]) - np.float32(about[k])) * np.float32(scale) + np.float32(about[k]) + np.float32(d))
    return c


def _renormal(q):
    """PrimTri ctor: normal = normalize(cross(v1-v0, v2-v0)) in float (geometry.hpp:60-69); GPU and oracle only consume it."""
    for t in range(2):
        p = [np.array(q.tri[t].v[v].pos[:], np.float32) for v in range(3)]
        n = np.cross(p[1] - p[0], p[2] - p[0]).astype(np.float32)
        ln = np.float32(np.sqrt(np.float32(n @ n)))
        n = n / ln if ln > 0 else np.array([0, 0, 1], np.float32)
        for k in range(3):
            q.tri[t].normal[k] = float(n[k])
    return q


def test_more_than_32_quads_nonplanar_and_degenerate():
    """45 quads: the Cornell box + three shrunken copies of the short block + one non-planar quad + one degenerate
    (zero-area) quad.  The intersect stage scans the list in blocks of 32 with per-block candidate masks; the
    conservative filter gives non-planar / degenerate quads an always-pass record."""
    flat = pu.load_flat("cornell", "ours1931")
    quads = [_copy_quad(flat.scene.quads[i]) for i in range(flat.scene.nquads)]
    n0 = len(quads)
    short_block = quads[n0 - 10:n0 - 5]
    for (dx, dz, s) in ((180.0, -40.0, 0.45), (-60.0, 250.0, 0.35), (250.0, 300.0, 0.3)):
        for q in short_block:
            quads.append(_translated(q, dx, 0.0, dz, scale=s, about=(185.0, 0.0, 169.0)))
    # a non-planar quad floating in the room (one corner lifted), and a degenerate one (all vertices on a line)
    bent = _translated(quads[0], 0.0, 0.0, 0.0, scale=0.2, about=(278.0, 0.0, 279.0))
    for t in range(2):
        for v in range(3):
            bent.tri[t].v[v].pos[1] = 260.0
    bent.tri[0].v[1].pos[1] = 300.0  # v10 lifted: tri0 and tri1 are no longer coplanar
    quads.append(_renormal(bent))
    line = _copy_quad(quads[0])
    for t in range(2):
        for v in range(3):
            line.tri[t].v[v].pos[0], line.tri[t].v[v].pos[1], line.tri[t].v[v].pos[2] = 100.0 + 10.0 * v, 50.0, 100.0
    quads.append(_renormal(line))
    assert len(quads) > 32
    _with_quads(flat, quads)
    opt = pu.options("ours1931", 48, 40, 3, seed=17)
    x = _compare(flat, opt, pixels=((24, 20), (10, 8), (40, 30), (30, 12)))
    base = pu.load_flat("cornell", "ours1931")
    xb, _ = pu.oracle_resolve(base, opt, pu.oracle_render(base, opt)[0])
    assert not pu.bits_equal(x, xb), "the extra geometry is not visible"


@pytest.mark.parametrize("w,h,spp,rect,srange", [(1, 1, 5, None, None), (33, 17, 3, (5, 2, 29, 16), (1, 3)), (7, 64, 2, (0, 10, 7, 64), None)])
def test_odd_sizes_and_subsets(w, h, spp, rect, srange):
    flat = pu.load_flat("cornell", "ours1931")
    kw = {}
    if rect:
        kw.update(x0=rect[0], y0=rect[1], x1=rect[2], y1=rect[3])
    if srange:
        kw.update(sample_begin=srange[0], sample_end=srange[1])
    opt = pu.options("ours1931", w, h, spp, seed=19, **kw)
    acc_o, _, _ = pu.oracle_render(flat, opt)
    with pu.gpu_context(flat) as ctx:  # a fresh context starts from a zeroed accumulator (sample_begin > 0 adds to it)
        ctx.render(opt)
        acc_g = ctx.read_accum(w, h)
    assert pu.bits_equal(acc_g, acc_o)


@pytest.mark.parametrize("tw,th", [(1, 1), (3, 5), (64, 16)])
def test_small_textures(tw, th):
    """sRGB_ReflectanceTexture::sample (material.cpp:45-97): clamped nearest texel, any texture size."""
    _need_assets()
    flat = pu.load_flat("plane-srgb", "ours1931")
    rng = np.random.default_rng(tw * 100 + th)
    tex = np.ascontiguousarray(rng.integers(0, 256, (th, tw, 3), dtype=np.uint8))
    flat.keep.append(tex)
    flat.scene.textures[0].rgb8 = tex.ctypes.data_as(C.POINTER(C.c_uint8))
    flat.scene.textures[0].width, flat.scene.textures[0].height = tw, th
    _compare(flat, pu.options("ours1931", 24, 24, 2, seed=23), pixels=((12, 12),))


def test_async_scene_upload_overlaps_and_orders():
    """ssb_upload_scene_async: the texel copy runs on the copy stream while the camera rays are traced; the result must
    equal the synchronous upload's, also when the SAME device texture buffer is overwritten between two frames (the copy
    has to wait for the previous frame's shading) and when the texture size changes."""
    _need_assets()
    flat = pu.load_flat("plane-srgb", "ours1931")
    opt = pu.options("ours1931", 32, 24, 4, seed=29)
    want1, _ = pu.oracle_resolve(flat, opt, pu.oracle_render(flat, opt)[0])
    rng = np.random.default_rng(5)
    tex2 = np.ascontiguousarray(rng.integers(0, 256, (4096, 4096, 3), dtype=np.uint8))  # same size: device buffer is reused
    tex3 = np.ascontiguousarray(rng.integers(0, 256, (40, 24, 3), dtype=np.uint8))      # other size: reallocated
    with pu.gpu_context(flat) as ctx:
        ctx.upload_scene_async(flat.scene)
        got1, _ = ctx.render_frame(opt)
        assert pu.bits_equal(got1, want1)
        for tex in (tex2, tex3):
            flat.keep.append(tex)
            flat.scene.textures[0].rgb8 = tex.ctypes.data_as(C.POINTER(C.c_uint8))
            flat.scene.textures[0].width, flat.scene.textures[0].height = tex.shape[1], tex.shape[0]
            want, _ = pu.oracle_resolve(flat, opt, pu.oracle_render(flat, opt)[0])
            ctx.upload_scene_async(flat.scene)
            got, _ = ctx.render_frame(opt)
            assert pu.bits_equal(got, want) and not pu.bits_equal(got, want1)
        ctx.upload_scene_async(flat.scene)
        ctx.synchronize()  # an upload that no render consumes must still complete here
