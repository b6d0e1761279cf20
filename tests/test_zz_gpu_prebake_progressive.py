"""GPU: the two "next" rows of SURVEY §8(f4).

* ssb_options.prebaked_textures — Jakob-Hanika coefficient textures: rgb2spec_fetch once per texel at bake time
  (the pre-process the reference describes but does not implement, color.cpp:204-216), rgb2spec_eval_precise per lookup.
  The bar is the strictest one available: the frame must equal the per-lookup form — and therefore the oracle and the
  real reference's fixture — bit for bit, including the reference's NaN texel (a black texel divides by zero,
  rgb2spec.c:88) and after the texels or the tables change.
* the progressive Renderer (the live preview of the reference's window, main.cpp:313-327): sample slices must not change
  the finished frame; previews are the average of the samples so far.
(Named test_zz_* so that it runs after the parity suites.)"""
import ctypes as C
import importlib
import os
import subprocess
import time

import numpy as np
import pytest

import parity_util as pu

pytestmark = pytest.mark.gpu
abi = pu.abi
host = importlib.import_module("simple-spectral_b200.host")


def _need_assets():
    if not pu.have_assets():
        pytest.fail("data files not staged on the GPU box (assets/data): run __graft_entry__.build() first")


def _frames(ctx, opt, pixels):
    x, s = ctx.render_frame(opt)
    acc = ctx.read_accum(opt.width, opt.height)
    st = ctx.stats()
    per = [ctx.trace_samples(opt, px, py) for (px, py) in pixels]
    return x, s, acc, per, st


@pytest.mark.parametrize("scene", ["plane-srgb", "cornell-srgb"])
def test_prebaked_jh_equals_per_lookup_and_reference_fixture(scene):
    _need_assets()
    flat = pu.load_flat(scene, "jh")
    pixels = ((16, 12), (3, 20))
    opt = pu.options("jh", 32, 24, 4, seed=7)
    baked = pu.options("jh", 32, 24, 4, seed=7, prebaked_textures=1)
    with pu.gpu_context(flat) as ctx:
        x0, s0, a0, p0, st0 = _frames(ctx, opt, pixels)
        x1, s1, a1, p1, st1 = _frames(ctx, baked, pixels)
        x2, _, _, _, st2 = _frames(ctx, baked, pixels)   # second baked frame: nothing to bake
        x3, _, _, _, _ = _frames(ctx, opt, pixels)       # and back to the per-lookup form
    assert pu.bits_equal(a1, a0) and pu.bits_equal(x1, x0) and pu.bits_equal(s1, s0)
    for a, b in zip(p0, p1):
        assert pu.bits_equal(a, b)
    assert pu.bits_equal(x2, x0) and pu.bits_equal(x3, x0)
    ref = np.load(os.path.join(pu.GOLDEN, f"xyza_{scene}_jh_32x24_spp4_seed7.npy"))  # the real reference's frame
    assert pu.bits_equal(x1, ref)
    acc_o, _, _ = pu.oracle_render(flat, opt)
    assert pu.bits_equal(a1, acc_o)


def test_prebaked_jh_nan_texel_and_texture_replacement():
    """A black texel makes rgb2spec_fetch divide by zero (scale = 63/0, x = 0*inf = NaN): the baked coefficients carry the
    same NaNs as the per-lookup form.  Replacing the texels (synchronously or on the copy stream) or the tables must
    invalidate the baked coefficients."""
    _need_assets()
    flat = pu.load_flat("plane-srgb", "jh")
    opt = pu.options("jh", 24, 24, 3, seed=23)
    baked = pu.options("jh", 24, 24, 3, seed=23, prebaked_textures=1)
    rng = np.random.default_rng(41)
    texA = np.ascontiguousarray(rng.integers(0, 256, (5, 3, 3), dtype=np.uint8))
    texA[2, 1] = 0  # the NaN texel
    texA[0, 0] = 255
    texB = np.ascontiguousarray(rng.integers(0, 256, (5, 3, 3), dtype=np.uint8))    # same size: coefficient buffer reused
    texC = np.ascontiguousarray(rng.integers(0, 256, (16, 64, 3), dtype=np.uint8))  # other size: reallocated
    frames = []
    with pu.gpu_context(flat) as ctx:
        for k, tex in enumerate((texA, texB, texC)):
            flat.keep.append(tex)
            flat.scene.textures[0].rgb8 = tex.ctypes.data_as(C.POINTER(C.c_uint8))
            flat.scene.textures[0].width, flat.scene.textures[0].height = tex.shape[1], tex.shape[0]
            want = pu.oracle_render(flat, opt)[0]
            if k == 1:
                ctx.upload_scene_async(flat.scene)  # the bake has to wait for the copy stream
            else:
                ctx.upload_scene(flat.scene)
            ctx.render(baked)
            got = ctx.read_accum(24, 24)
            assert pu.bits_equal(got, want), f"texture {k}"
            frames.append(got)
        assert np.isnan(frames[0]).any(), "the black texel was never hit: the test does not cover the NaN path"
        assert not pu.bits_equal(frames[0], frames[1]) and not pu.bits_equal(frames[1], frames[2])
        ctx.upload_color(flat.color)  # same tables again: must re-bake, not crash or go stale
        ctx.render(baked)
        assert pu.bits_equal(ctx.read_accum(24, 24), frames[2])


def test_prebaked_flag_is_ignored_outside_jh():
    _need_assets()
    for variant in ("ours1931", "meng", "rgb"):
        flat = pu.load_flat("plane-srgb", variant)
        opt = pu.options(variant, 24, 20, 2, seed=5, prebaked_textures=1)
        want = pu.oracle_render(flat, opt)[0]
        with pu.gpu_context(flat) as ctx:
            ctx.render(opt)
            assert pu.bits_equal(ctx.read_accum(24, 20), want), variant


def test_renderer_progressive_and_prebaked_equal_the_single_call_frame():
    """host layer, C++ Renderer through ssbh_*: one call / progressive slices / progressive + prebaked JH textures."""
    _need_assets()
    kw = dict(variant="jh", seed=7, data_root=pu.data_root())
    r = host.Renderer("plane-srgb", 40, 32, 11, **kw)
    x0, f0 = r.render()
    n0 = r.stats().launches
    r.close()
    r = host.Renderer("plane-srgb", 40, 32, 11, progressive=True, prebaked_textures=True, **kw)
    r.start()
    seen = set()
    while r.is_rendering():
        seen.add(r.snapshot()[0])
        time.sleep(0.0002)
    x1, f1 = r.wait()
    st = r.stats()
    assert r.snapshot()[0] == 11
    r.close()
    assert seen <= {0, 1, 2, 4, 8, 11}, seen
    assert st.samples == 40 * 32 * 11 and st.launches > n0  # five slices (+ the bake) against one
    assert pu.bits_equal(x1, x0) and pu.bits_equal(f1, f0)
    # the preview after the first slice is the spp = 1 frame
    r1 = host.Renderer("plane-srgb", 40, 32, 1, **kw)
    _, f_first = r1.render()
    r1.close()
    r = host.Renderer("plane-srgb", 40, 32, 1 << 12, progressive=True, **kw)
    r.start()
    t0 = time.time()
    while r.snapshot()[0] < 1 and time.time() - t0 < 60:
        pass
    done, snap = r.snapshot()
    r.stop()
    r.wait()
    stopped_at = r.snapshot()[0]
    r.close()
    if done == 1:
        assert pu.bits_equal(snap, f_first)
    assert 1 <= stopped_at <= 1 << 12


def test_cli_progressive_preview(tmp_path):
    _need_assets()
    cli = os.path.join(pu.ROOT, "simple-spectral_b200", "simple_spectral_b200")
    base = ["--scene=cornell-srgb", "-w=48", "-h=40", "-spp=9", "--seed=3", f"--data-root={pu.data_root()}", "--variant=jh"]
    a, b, prev = str(tmp_path / "a.pfm"), str(tmp_path / "b.pfm"), str(tmp_path / "preview.pfm")
    r = subprocess.run([cli, *base, f"--output={a}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([cli, *base, f"--output={b}", "--prebake", f"--preview={prev}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "Render completed in" in r.stdout
    assert open(a, "rb").read() == open(b, "rb").read()
    if os.path.exists(prev):  # at least one preview was caught by the polling loop
        assert os.path.getsize(prev) == os.path.getsize(b)


def test_spectra_on_different_grids_take_the_general_path():
    """The observer's three spectra normally share one grid and are sampled with shared index arithmetic in the fold
    stage (spec_sample3); tables on DIFFERENT grids (other length / range / filter; the basis' too) must take the
    per-spectrum form and still equal the oracle."""
    _need_assets()
    flat = pu.load_flat("cornell-srgb", "ours1931")
    opt = pu.options("ours1931", 24, 20, 3, seed=31)
    same, _ = pu.oracle_resolve(flat, opt, pu.oracle_render(flat, opt)[0])
    for name, cut, filt in (("ybar", 3, 0), ("basis_g", 5, 0), ("zbar", 0, abi.SSB_FILTER_NEAREST)):
        s = getattr(flat.color, name)
        data = np.ctypeslib.as_array(s.data, shape=(s.n,)).copy()
        step = (s.high - s.low) / (s.n - 1)
        setattr(flat.color, name, flat.spectrum(data[:s.n - cut], s.low, s.high - cut * step, filt))
    acc_o = pu.oracle_render(flat, opt)[0]
    xo, so = pu.oracle_resolve(flat, opt, acc_o)
    assert not pu.bits_equal(xo, same)
    with pu.gpu_context(flat) as ctx:
        xg, sg = ctx.render_frame(opt)
        acc_g = ctx.read_accum(24, 20)
    assert pu.bits_equal(acc_g, acc_o) and pu.bits_equal(xg, xo) and pu.bits_equal(sg, so)


@pytest.mark.parametrize("scene,variant", [("cornell", "ours2006"), ("plane-srgb", "jh"), ("cornell-srgb", "meng")])
def test_config1_size_sha_of_the_other_baseline_variants(scene, variant):
    """128x128 spp16 frames of the variants BASELINE.json configs[2..4] are quoted on: sha256 of the XYZA buffer the REAL
    reference produced (tests/golden/golden_index.json; the oracle is pinned to the same hashes on the CPU).  JH: also
    with prebaked coefficient textures."""
    import hashlib
    import json
    _need_assets()
    want = json.load(open(os.path.join(pu.GOLDEN, "golden_index.json")))[f"{scene}_{variant}_128x128_spp16_seed1"]["sha256"]
    flat = pu.load_flat(scene, variant)
    with pu.gpu_context(flat) as ctx:
        for prebaked in ((0, 1) if variant == "jh" else (0,)):
            x, _ = ctx.render_frame(pu.options(variant, 128, 128, 16, seed=1, prebaked_textures=prebaked))
            assert hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest() == want, (variant, prebaked)


@pytest.mark.parametrize("scene", ["cornell", "cornell-srgb"])
def test_flat_field_correction_off_equals_the_reference_build(scene):
    """FLAT_FIELD_CORRECTION compiled out (renderer.cpp:262-266): the frame of the real reference's build of that
    configuration (tests/golden/xyza_*_ours1931_noffc_*; oracle/build_ref.py explains the one declaration the reference
    needs to compile it — outside the render path)."""
    if scene != "cornell":
        _need_assets()
    flat = pu.load_flat(scene, "ours1931_noffc")
    opt = pu.options("ours1931_noffc", 32, 24, 4, seed=7)
    ref = np.load(os.path.join(pu.GOLDEN, f"xyza_{scene}_ours1931_noffc_32x24_spp4_seed7.npy"))
    with pu.gpu_context(flat) as ctx:
        x, _ = ctx.render_frame(opt)
    assert pu.bits_equal(x, ref)
