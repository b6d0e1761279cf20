"""GPU: the WHOLE path (camera ray -> closest hit -> emission / albedo -> light pick + spherical-triangle sample ->
shadow query -> BSDF sample -> fold -> XYZ accumulate -> sRGB) on RANDOM scenes, CUDA path vs the oracle, bit for bit.

The reference's three scenes are axis-aligned boxes lit by one horizontal rectangle: the spherical-triangle sampler
(Arvo; renderer.cpp:182-218, scene.cpp:417-431, geometry.cpp:141-237) only ever sees that one light from below, the
orthonormal basis only six normals, the conservative intersection filter only box faces.  Here the quads are the
generator of tests/test_gpu_isect_fuzz.py (rectangles, sheared, trapezoid, non-planar, degenerate, slivers,
duplicates, coplanar neighbours) scaled into the Cornell room, with

  * 1-4 lights of arbitrary orientation and shape (slivers and degenerate triangles included: zero solid angle,
    infinite pdf, NaN directions must come out as in the reference's arithmetic),
  * materials drawn at random from the scene's list, a mirror in every third scene,
  * a small random texture on the textured material (cornell-srgb family), all three upsampling methods and the RGB mode,
  * explicit light sampling on and off, MAX_DEPTH 2-6, 2-4 hero wavelengths.

Every case compares the f64 accumulators of the whole (small) frame, the resolved XYZA / sRGBA, and the per-sample
float4 of a few pixels.  The oracle is pinned to builds of the real reference on the reference's scenes and options
(tests/test_oracle_golden.py); this test extends the CUDA-vs-oracle comparison to geometry the fixtures cannot reach."""
import ctypes as C

import numpy as np
import pytest

import parity_util as pu
from test_gpu_isect_fuzz import fuzz_quads

abi = pu.abi
pytestmark = pytest.mark.gpu


def _quad_from_corners(template, corners, material, is_light, st=None):
    """PrimQuad(v00, v10, v11, v01) (geometry.hpp:93-95) with the template's texture coordinates, or st[4][2] for the corners."""
    q = abi.ssb_quad()
    C.memmove(C.byref(q), C.byref(template), C.sizeof(abi.ssb_quad))
    v00, v10, v11, v01 = (corners[k].astype(np.float32) for k in range(4))
    with np.errstate(invalid="ignore", divide="ignore"):
        for ti, (tri, idx) in enumerate((((v00, v10, v11), (0, 1, 2)), ((v00, v11, v01), (0, 2, 3)))):
            for vi in range(3):
                for k in range(3):
                    q.tri[ti].v[vi].pos[k] = float(tri[vi][k])
                if st is not None:
                    q.tri[ti].v[vi].st[0], q.tri[ti].v[vi].st[1] = float(st[idx[vi]][0]), float(st[idx[vi]][1])
            nrm = np.cross((tri[1] - tri[0]).astype(np.float32), (tri[2] - tri[0]).astype(np.float32)).astype(np.float32)
            ln = np.float32(np.sqrt(np.float32(np.dot(nrm, nrm))))
            nrm = nrm / ln if ln > 0 else np.zeros(3, np.float32)
            for k in range(3):
                q.tri[ti].normal[k] = float(nrm[k])
    q.material, q.is_light = material, is_light
    return q


def _random_scene(flat, rng, nquads, nlights, kind, mirror):
    """Replace the quad list of `flat` by random quads inside the room the camera looks into; keep camera and materials."""
    sc = flat.scene
    light_mats = sorted({sc.quads[i].material for i in range(sc.nquads) if sc.quads[i].is_light})
    other_mats = [m for m in range(sc.nmaterials) if m not in light_mats]
    assert light_mats and other_mats
    template = sc.quads[0]
    # a floor-like template quad carries the texture coordinates of a textured scene; any quad does for the others
    for i in range(sc.nquads):
        if sc.materials[sc.quads[i].material].albedo_mode == abi.SSB_ALBEDO_TEXTURE:
            template = sc.quads[i]
            break
    # in front of the camera, wherever the scene put it (cornell: the room's centre is 1080 units down the view axis)
    cam = sc.camera
    centre = np.array([cam.pos[k] + 1080.0 * cam.dir[k] for k in range(3)])
    corners = fuzz_quads(rng, nquads, 300.0, centre, kind)
    quads = []
    light_ids = set(rng.choice(nquads, size=min(nlights, nquads), replace=False).tolist())
    for qi in range(nquads):
        # texture coordinates beyond [0,1] and negative ones: clamped nearest texel (material.cpp:73-84); every third quad keeps the template's
        st = None if qi % 3 == 0 else rng.uniform(-1.5, 2.5, (4, 2)).astype(np.float32)
        if qi in light_ids:
            quads.append(_quad_from_corners(template, corners[qi], int(rng.choice(light_mats)), 1, st))
        else:
            quads.append(_quad_from_corners(template, corners[qi], int(rng.choice(other_mats)), 0, st))
    if mirror:
        m = int(rng.choice(other_mats))
        if sc.materials[m].albedo_mode != abi.SSB_ALBEDO_TEXTURE:
            sc.materials[m].kind = abi.SSB_MATERIAL_MIRROR
    arr = (abi.ssb_quad * nquads)()
    for i, q in enumerate(quads):
        C.memmove(C.byref(arr[i]), C.byref(q), C.sizeof(abi.ssb_quad))
    flat.keep.append(arr)
    sc.quads, sc.nquads = arr, nquads
    if sc.ntextures:
        tw, th = int(rng.integers(1, 48)), int(rng.integers(1, 48))
        tex = np.ascontiguousarray(rng.integers(0, 256, (th, tw, 3), dtype=np.uint8))
        flat.keep.append(tex)
        sc.textures[0].rgb8 = tex.ctypes.data_as(C.POINTER(C.c_uint8))
        sc.textures[0].width, sc.textures[0].height = tw, th
    return flat


CASES = []
_kinds = ["mixed", "rect", "axis", "mixed", "sheared", "trapezoid", "mixed", "nonplanar", "mixed", "mixed"]  # "mixed" includes slivers and degenerate quads
for _i in range(int(__import__("os").environ.get("SSB_RENDER_FUZZ_CASES", "96"))):  # a longer campaign: SSB_RENDER_FUZZ_CASES=3000
    _scene, _variant = [("cornell", "ours1931"), ("cornell-srgb", "ours1931"), ("cornell", "ours2006"), ("cornell-srgb", "meng"),
                        ("cornell-srgb", "jh"), ("plane-srgb", "ours1931"), ("cornell-srgb", "rgb"), ("plane-srgb", "jh")][_i % 8]
    CASES.append((_i, _scene, _variant, _kinds[_i % len(_kinds)]))


@pytest.mark.parametrize("case,scene,variant,kind", CASES)
def test_random_scene_full_path(case, scene, variant, kind):
    if pu.needs_assets(scene, variant) and not pu.have_assets():
        pytest.fail("data files not staged on the GPU box (assets/data): run __graft_entry__.build() first")
    rng = np.random.default_rng(7000 + case)
    flat = pu.load_flat(scene, variant)
    nquads = int(rng.choice([6, 12, 20, 33, 48, 64]))
    flat = _random_scene(flat, rng, nquads, int(rng.integers(1, 5)), kind, mirror=(case % 3 == 2))
    w, h, spp = int(rng.integers(12, 40)), int(rng.integers(12, 32)), int(rng.integers(2, 6))
    opt = pu.options(variant, w, h, spp, seed=1000 + case, max_depth=int(rng.integers(2, 7)),
                     explicit_light_sampling=int(case % 5 != 3), n_wavelengths=4 if variant == "rgb" else int(rng.choice([4, 4, 3, 2])))
    acc_o, samp_o, _ = pu.oracle_render(flat, opt, want_samples=True)
    xo, so = pu.oracle_resolve(flat, opt, acc_o)
    pixels = [(int(rng.integers(w)), int(rng.integers(h))) for _ in range(3)]
    with pu.gpu_context(flat) as ctx:
        xg, sg = ctx.render_frame(opt)
        acc_g = ctx.read_accum(w, h)
        for (px, py) in pixels:
            assert pu.bits_equal(ctx.trace_samples(opt, px, py), samp_o[py, px]), f"case {case}: per-sample mismatch at ({px},{py})"
        # the library's own list-order scan must give the same frame (ssb_options.scan_mode)
        opt_list = pu.options(variant, w, h, spp, seed=1000 + case, max_depth=opt.max_depth, explicit_light_sampling=opt.explicit_light_sampling,
                              n_wavelengths=opt.n_wavelengths, scan_mode=abi.SSB_SCAN_LIST)
        ctx.render(opt_list)
        acc_l = ctx.read_accum(w, h)
    bad = ~((acc_g.view(np.uint64) == acc_o.view(np.uint64)) | (np.isnan(acc_g) & np.isnan(acc_o)))
    assert pu.bits_equal(acc_g, acc_o), f"case {case}: {int(bad.sum())} accumulator words differ, max rel {pu.rel_err(acc_g, acc_o).max()}"
    assert pu.bits_equal(acc_l, acc_o), f"case {case}: list-scan frame differs from the oracle"
    assert pu.bits_equal(xg, xo) and pu.bits_equal(sg, so)
    # the scene must actually be lit and visible, or the comparison proves nothing
    assert np.nansum(np.abs(acc_o[..., 3])) > 0, f"case {case}: no pixel hit anything"


# ---- the same kind of scene through the REAL reference (tests/test_oracle_random_scenes.py): CUDA path vs the reference's own
# f64 XYZA, from the committed fixtures and — where oracle/_ref travelled to the box — from live runs of the hooked binaries
import os

import test_oracle_random_scenes as rs


def _gpu_xyza(flat, variant, w, h, spp, seed):
    opt = pu.options(variant, w, h, spp, seed=seed)
    with pu.gpu_context(flat) as ctx:
        xyza, _ = ctx.render_frame(opt)
    return xyza


@pytest.mark.parametrize("i", rs.FIXTURES)
def test_random_scene_reference_fixture(i):
    scene, variant, kind, nquads, w, h, spp = rs.CASES[i]
    if pu.needs_assets(scene, variant) and not pu.have_assets():
        pytest.fail("data files not staged on the GPU box (assets/data): run __graft_entry__.build() first")
    flat = rs.flat_of(os.path.join(pu.GOLDEN, f"random_{i}_{scene}_{variant}_tables.bin"), scene, variant)
    ref = np.load(os.path.join(pu.GOLDEN, f"random_{i}_{scene}_{variant}_xyza_{w}x{h}_spp{spp}_seed{300 + i}.npy"))
    assert pu.bits_equal(_gpu_xyza(flat, variant, w, h, spp, 300 + i), ref)


@pytest.mark.parametrize("i", range(len(rs.CASES)))
def test_random_scene_live_reference(i, tmp_path):
    scene, variant, kind, nquads, w, h, spp = rs.CASES[i]
    if not pu.have_assets():
        pytest.fail("data files not staged on the GPU box (assets/data): run __graft_entry__.build() first")
    if not os.path.exists(os.path.join(rs.REFDIR, f"simple_spectral_{variant}_hooked")):
        pytest.skip("oracle/_ref not on this box")
    tables = str(tmp_path / "tables.bin")
    ref = rs.run_reference(scene, variant, rs._payload(i), w, h, spp, 300 + i, tables)
    flat = rs.flat_of(tables, scene, variant)
    assert pu.bits_equal(_gpu_xyza(flat, variant, w, h, spp, 300 + i), ref)


@pytest.mark.parametrize("nquads,nlights,scene,variant", [(256, 64, "cornell", "ours1931"), (256, 3, "cornell-srgb", "jh"), (200, 64, "cornell-srgb", "meng"),
                                                          (1, 1, "cornell", "ours1931"), (2, 2, "cornell-srgb", "ours1931"), (2, 1, "plane-srgb", "ours1931")])
def test_extreme_scene_sizes(nquads, nlights, scene, variant):
    """The ABI's limits (SSB_MAX_QUADS = 256 quads, SSB_MAX_LIGHTS = 64 lights) and the smallest scenes (the light alone)."""
    if pu.needs_assets(scene, variant) and not pu.have_assets():
        pytest.fail("data files not staged on the GPU box (assets/data): run __graft_entry__.build() first")
    rng = np.random.default_rng(8800 + nquads * 7 + nlights)
    flat = _random_scene(pu.load_flat(scene, variant), rng, nquads, nlights, "mixed" if nquads > 2 else "rect", mirror=False)
    opt = pu.options(variant, 24, 20, 3, seed=77, max_depth=5)
    acc_o, samp_o, _ = pu.oracle_render(flat, opt, want_samples=True)
    with pu.gpu_context(flat) as ctx:
        ctx.render(opt)
        acc_g = ctx.read_accum(24, 20)
        assert pu.bits_equal(ctx.trace_samples(opt, 12, 10), samp_o[10, 12])
    assert pu.bits_equal(acc_g, acc_o), f"max rel {pu.rel_err(acc_g, acc_o).max()}"
