"""CPU: the C++ host layer (Color::init / Scene::get_new_* restatements, PNG decode, CSV loader) against
table dumps written by the REAL reference (tests/golden/tables_*.bin).  Bar: every float / double bit-identical."""
import ctypes as C
import importlib
import os

import numpy as np
import pytest

import parity_util as pu
import refdump

ssb = importlib.import_module("simple-spectral_b200")
host = importlib.import_module("simple-spectral_b200.host")

CASES = [("cornell", "ours1931"), ("cornell-srgb", "ours1931"), ("plane-srgb", "ours1931"), ("cornell", "ours2006"),
         ("cornell-srgb", "jh"), ("plane-srgb", "meng")]

pytestmark = pytest.mark.skipif(not pu.have_assets(), reason="reference data files not staged (assets/data)")


def _spec(s):
    return np.ctypeslib.as_array(s.data, shape=(s.n,)).copy(), s.low, s.high


@pytest.mark.parametrize("scene,variant", CASES)
def test_host_tables_match_reference_dump(scene, variant):
    t = refdump.parse(os.path.join(pu.GOLDEN, f"tables_{scene}_{variant}.bin"))
    obs, ups = host.VARIANTS[variant]
    color = host.Color(pu.data_root(), obs, ups)
    sc = host.Scene(scene, color)
    fc, fs = color.flat, sc.flat
    # colour tables
    for name in ("xbar", "ybar", "zbar"):
        d, lo, hi = _spec(getattr(fc, name))
        assert pu.bits_equal(d, t[f"color.{name}.data"]) and (lo, hi) == tuple(t[f"color.{name}.lowhigh"])
    if variant.startswith("ours"):
        for name in ("basis_r", "basis_g", "basis_b"):
            d, lo, hi = _spec(getattr(fc, name))
            assert pu.bits_equal(d, t[f"color.{name}.data"]) and (lo, hi) == tuple(t[f"color.{name}.lowhigh"])
    for name in ("D65_orig", "D65_rad"):
        d, lo, hi = color.spectrum(name)
        assert pu.bits_equal(d, t[f"color.{name}.data"]), name
    assert pu.bits_equal(np.array(color._orig[:], np.float32), t["color.D65_orig_XYZ"])
    assert pu.bits_equal(np.array(color._rad[:], np.float32), t["color.D65_rad_XYZ"])
    assert pu.bits_equal(np.array(color._m[:], np.float32), t["color.matr_lrgb_to_xyz"])
    assert pu.bits_equal(np.array(color._mi[:], np.float32), t["color.matr_xyz_to_lrgb"])
    assert pu.bits_equal(np.array(fc.xyz_to_lrgb[:], np.float32), t["color.matr_xyz_to_lrgb"])
    assert (color.lambda_min, color.lambda_max) == tuple(t["config.lambda_min_max_step"][:2])
    # camera
    P, V, I = sc.camera_matrices()
    assert pu.bits_equal(P, t["camera.matr_P"]) and pu.bits_equal(V, t["camera.matr_V"])
    assert pu.bits_equal(I, t["camera.matr_PV_inv"])
    assert pu.bits_equal(np.array(fs.camera.pv_inv[:]), t["camera.matr_PV_inv"])
    assert pu.bits_equal(np.array(fs.camera.pos[:], np.float32), t["camera.pos"])
    assert pu.bits_equal(np.array(fs.camera.dir[:], np.float32), t["camera.dir"])
    # primitives: order, vertices, ST, normals, material ids, light flags
    q = t["scene.quads"].reshape(-1, 2, 18)
    assert fs.nquads == q.shape[0]
    got = np.zeros_like(q)
    for qi in range(fs.nquads):
        for ti in range(2):
            tri = fs.quads[qi].tri[ti]
            for vi in range(3):
                got[qi, ti, vi * 5:vi * 5 + 3] = tri.v[vi].pos[:]
                got[qi, ti, vi * 5 + 3:vi * 5 + 5] = tri.v[vi].st[:]
            got[qi, ti, 15:18] = tri.normal[:]
    assert pu.bits_equal(got, q)
    assert [fs.quads[i].material for i in range(fs.nquads)] == list(t["scene.quad_material"])
    assert [fs.quads[i].is_light for i in range(fs.nquads)] == list(t["scene.quad_is_light"])
    # materials
    m = 0
    while f"material.{m}.kind_mode" in t:
        fm = fs.materials[m]
        assert (fm.kind, fm.albedo_mode) == tuple(t[f"material.{m}.kind_mode"])
        d, lo, hi = _spec(fm.emission)
        assert pu.bits_equal(d, t[f"material.{m}.emission.data"]) and (lo, hi) == tuple(t[f"material.{m}.emission.lowhigh"])
        if fm.albedo_mode == 0:
            d, lo, hi = _spec(fm.albedo)
            assert pu.bits_equal(d, t[f"material.{m}.albedo.data"]) and (lo, hi) == tuple(t[f"material.{m}.albedo.lowhigh"])
        else:
            tex = fs.textures[fm.texture]
            assert (tex.width, tex.height) == tuple(t[f"material.{m}.texture_res"])
        m += 1
    assert fs.nmaterials == m


@pytest.mark.parametrize("scene", ["cornell", "cornell-srgb", "plane-srgb"])
def test_host_rgb_constants_match_rgb_reference_dump(scene):
    """The RENDER_MODE_RGB build of the reference: same geometry / material ids / light flags, RGB constants."""
    t = refdump.parse(os.path.join(pu.GOLDEN, f"tables_{scene}_rgb.bin"))
    color = host.Color(pu.data_root())
    sc = host.Scene(scene, color)  # keep alive: `flat` points into it
    fs = sc.flat
    assert pu.bits_equal(np.array(fs.camera.pv_inv[:]), t["camera.matr_PV_inv"])
    q = t["scene.quads"].reshape(-1, 2, 18)
    assert fs.nquads == q.shape[0]
    for qi in range(fs.nquads):
        for ti in range(2):
            tri = fs.quads[qi].tri[ti]
            got = np.concatenate([np.concatenate([tri.v[vi].pos[:], tri.v[vi].st[:]]) for vi in range(3)] + [tri.normal[:]]).astype(np.float32)
            assert pu.bits_equal(got, q[qi, ti])
    assert [fs.quads[i].material for i in range(fs.nquads)] == list(t["scene.quad_material"])
    assert [fs.quads[i].is_light for i in range(fs.nquads)] == list(t["scene.quad_is_light"])
    m = 0
    while f"material.{m}.kind_mode" in t:
        fm = fs.materials[m]
        assert (fm.kind, fm.albedo_mode) == tuple(t[f"material.{m}.kind_mode"])
        assert pu.bits_equal(np.array(fm.emission_rgb[:], np.float32), t[f"material.{m}.emission_rgb"])
        if fm.albedo_mode == 0:
            assert pu.bits_equal(np.array(fm.albedo_rgb[:], np.float32), t[f"material.{m}.albedo_rgb"])
        m += 1
    assert fs.nmaterials == m


def test_png_decoder_matches_pillow():
    root = pu.data_root()
    for name in ("scenes/test-img.png", "scenes/crystal-lizard-512.png"):
        p = os.path.join(root, "data", name)
        assert np.array_equal(host.load_png_rgb8(p), refdump.load_texture_rgb8(p))


def test_lizard_texture_every_colour_once():
    a = host.load_png_rgb8(os.path.join(pu.data_root(), "data", "scenes", "crystal-lizard-4096.png"))
    assert a.shape == (4096, 4096, 3)
    assert np.array_equal(a, pu.lizard_texture())
    codes = (a[..., 0].astype(np.uint32) << 16) | (a[..., 1].astype(np.uint32) << 8) | a[..., 2]
    assert np.unique(codes).size == 1 << 24  # all 24-bit colours exactly once


def test_error_codes_follow_reference():
    color = host.Color(pu.data_root())
    with pytest.raises(ssb.SsbError) as e:
        host.Scene("no-such-scene", color)
    assert e.value.code == -3  # renderer.cpp:32-37
    with pytest.raises(ssb.SsbError) as e:
        host.Color("/nonexistent")
    assert e.value.code == -1  # spectrum.cpp:179-182
    with pytest.raises(ssb.SsbError) as e:
        host.Color(pu.data_root(), 2006, ssb.SSB_UPSAMPLE_JH)
    assert e.value.code == -3  # stdafx.hpp:106-108


def test_image_writers_roundtrip(tmp_path):
    rng = np.random.default_rng(0)
    img = rng.uniform(0, 1, (6, 5, 4)).astype(np.float32)
    host.save_image(str(tmp_path / "a.png"), img)
    from PIL import Image
    got = np.asarray(Image.open(tmp_path / "a.png").convert("RGBA"))
    want = np.round(np.clip(img * 255.0, 0, 255)).astype(np.uint8)[::-1]
    assert np.array_equal(got, want)
    host.save_image(str(tmp_path / "a.pfm"), img)
    raw = open(tmp_path / "a.pfm", "rb").read()
    assert raw.startswith(b"PF\n5 6\n-1.0\n")
    data = np.frombuffer(raw[len(b"PF\n5 6\n-1.0\n"):], np.float32).reshape(6, 5, 3)
    s = img[::-1, :, :3]
    lin = np.where(s < 0.04045, s / 12.92, ((s + 0.055) / 1.055) ** 2.4)
    assert np.allclose(data, lin, rtol=1e-5)
    host.save_image(str(tmp_path / "a.csv"), img)
    assert len(open(tmp_path / "a.csv").read().strip().split("\n")) == 6
    host.save_image(str(tmp_path / "a.hdr"), img)
    assert open(tmp_path / "a.hdr", "rb").read().startswith(b"#?RADIANCE\n")


def test_image_writers_match_the_reference_files(tmp_path):
    """Framebuffer::save (framebuffer.cpp:39-176) against the files the REAL reference wrote for the small golden case
    (tests/golden/refout_*, generated by make_golden.py): PFM / HDR / CSV byte for byte, PNG pixel for pixel (the
    reference compresses with lodepng, this layer with zlib)."""
    flat = pu.load_flat("cornell", "ours1931")
    opt = pu.options("ours1931", 32, 24, 4, seed=7)
    _, srgba = pu.oracle_resolve(flat, opt, pu.oracle_render(flat, opt)[0])
    for ext in ("pfm", "hdr", "csv", "png"):
        ours = str(tmp_path / f"o.{ext}")
        host.save_image(ours, srgba)
        ref = os.path.join(pu.GOLDEN, f"refout_cornell_ours1931_32x24_spp4_seed7.{ext}")
        if ext == "png":
            from PIL import Image
            a, b = Image.open(ours), Image.open(ref)
            assert a.mode == b.mode and np.array_equal(np.asarray(a), np.asarray(b))
        else:
            assert open(ours, "rb").read() == open(ref, "rb").read(), ext


def test_round_trip_self_test_reproduces_the_reference_and_its_documented_number():
    """The reference's one documented number: the sRGB -> spectrum -> sRGB round trip over all 2^24 colours has a maximum
    error of 1.851469e-5 (main.cpp:242-245).  tests/golden/roundtrip_running_max.json holds what the REAL reference prints
    when that self-test (main.cpp:246-262, compiled out upstream) is switched on: the running maximum after each red level.
    The host layer's Color::init tables + _Spectrum arithmetic + integrate + round_trip_srgb must reproduce it bit for bit:
    the first levels from scratch, and every level at which the maximum rises (continued from the reference's previous
    value).  SSB_FULL_ROUNDTRIP=1 runs all 256 levels (~3 minutes)."""
    import json
    ref = np.array([float(v) for v in json.load(open(os.path.join(pu.GOLDEN, "roundtrip_running_max.json")))["running_max"]], np.float32)
    assert ref.size == 256 and f"{ref[-1]:.6e}" == "1.851469e-05"
    color = host.Color(pu.data_root(), 1931, ssb.SSB_UPSAMPLE_OURS)
    assert pu.bits_equal(np.array(color.round_trip_srgb((1.0, 1.0, 1.0)), np.float32), np.array([0.99999994, 0.9999998, 0.9999998], np.float32))
    if os.environ.get("SSB_FULL_ROUNDTRIP"):
        assert pu.bits_equal(color.round_trip_running_max(0, 256), ref)
        return
    assert pu.bits_equal(color.round_trip_running_max(0, 3), ref[:3])
    rises = [r for r in range(3, 256) if ref[r] != ref[r - 1]]
    assert rises, "the reference's maximum never rises after level 2?"
    for r in rises + [200]:
        got = color.round_trip_running_max(r, r + 1, start_max=float(ref[r - 1]))
        assert pu.bits_equal(got, ref[r:r + 1]), r
