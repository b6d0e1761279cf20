"""GPU: the CUDA path (through the C ABI) against the oracle and the committed reference fixtures.
Bar: bit-identical per-sample float4 and per-pixel double XYZA; sRGB within 1e-6 relative."""
import os

import numpy as np
import pytest

import parity_util as pu

pytestmark = pytest.mark.gpu

CASES = [("cornell", "ours1931"), ("cornell-srgb", "ours1931"), ("plane-srgb", "ours1931"),
         ("cornell", "ours2006"), ("cornell-srgb", "ours2006"),
         ("cornell-srgb", "jh"), ("plane-srgb", "jh"), ("cornell-srgb", "meng"), ("plane-srgb", "meng"),
         ("plane-srgb", "ours1931_noels"), ("cornell", "ours1931_noels"),
         ("cornell", "rgb"), ("cornell-srgb", "rgb"), ("plane-srgb", "rgb"),  # RENDER_MODE_RGB build of the reference
         ("cornell-srgb", "ours1931_nw3"), ("plane-srgb", "ours1931_nw3"), ("cornell-srgb", "meng_nw2"),  # SAMPLE_WAVELENGTHS 3 / 2
         ("cornell-srgb", "ours1931_d3")]  # MAX_DEPTH 3


def _skip_if_no_assets(scene, variant):
    if pu.needs_assets(scene, variant) and not pu.have_assets():
        pytest.fail("data files not staged on the GPU box (assets/data): run __graft_entry__.build() first")


@pytest.mark.parametrize("fn,lo,hi", [(0, -119.0, 119.0), (1, -119.0, 119.0), (2, -1.0, 1.0), (4, -119.0, 119.0), (5, -119.0, 119.0)])
def test_device_math_bit_exact(fn, lo, hi):
    """sinf/cosf/acosf on the device == the host libm the reference links (glibc), bit for bit."""
    rng = np.random.default_rng(123 + fn)
    x = np.concatenate([rng.uniform(lo, hi, 1 << 20), rng.uniform(-7.0, 7.0, 1 << 20).clip(lo, hi),
                        np.array([0.0, -0.0, lo, hi, 1e-8, -1e-8, 0.5, -0.5, 0.78539816, 1.0, -1.0])]).astype(np.float32)
    flat = pu.load_flat("cornell", "ours1931")
    with pu.gpu_context(flat) as ctx:
        got = ctx.eval_math(fn, x)
    import ctypes as C
    want = np.empty_like(x)
    host_fn = {4: 0, 5: 1}.get(fn, fn)  # 4/5 = the paired sincos the kernels use; checked against libm sinf / cosf
    pu.oracle().ssb_oracle_eval_math(host_fn, x.ctypes.data_as(C.POINTER(C.c_float)), 0.0, want.ctypes.data_as(C.POINTER(C.c_float)), x.size)
    assert pu.bits_equal(got, want), f"{(got.view(np.uint32) != want.view(np.uint32)).sum()} mismatches"


def test_device_paired_acosf_equals_scalar_exhaustively():
    """The packed two-argument acosf the light-sampling code calls (ssbm::acosf_exact2: FMUL2/FADD2/FFMA2, hand-expanded
    division and square root) against the scalar acosf_exact — itself pinned to glibc over all 2^32 inputs on the CPU
    (tests/test_math_exhaustive.py) and on samples above — for ALL 2^32 arguments in either lane, evaluated on the device."""
    flat = pu.load_flat("cornell", "ours1931")
    with pu.gpu_context(flat) as ctx:
        bad = ctx.eval_math(8, np.zeros(1 + 3 * 8, np.float32))
        assert bad[0] == 0.0, f"{bad[0]:.0f} lanes differ, e.g. (argument, paired, scalar) bits: " + \
            ", ".join(f"({a:#010x}, {b:#010x}, {c:#010x})" for a, b, c in bad[1:].view(np.uint32).reshape(-1, 3) if a or b or c)
        # and a direct spot check of both lanes against the host libm
        rng = np.random.default_rng(77)
        x = np.concatenate([rng.uniform(-1.0, 1.0, 1 << 18), np.array([0.0, -0.0, 1.0, -1.0, 0.5, -0.5, 1e-20, -1e-20, 0.49999997, 0.50000006])]).astype(np.float32)
        import ctypes as C
        want = np.empty_like(x)
        pu.oracle().ssb_oracle_eval_math(2, x.ctypes.data_as(C.POINTER(C.c_float)), 0.0, want.ctypes.data_as(C.POINTER(C.c_float)), x.size)
        for fn in (6, 7):
            assert pu.bits_equal(ctx.eval_math(fn, x), want)


@pytest.mark.parametrize("y", [2.4, 1.0 / 2.4])
def test_device_powf_bit_exact(y):
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(0.0, 1.0, 1 << 20), np.exp(rng.uniform(-20, 20, 1 << 20)), np.arange(256) / 255.0]).astype(np.float32)
    x = x[x > 1e-30]
    flat = pu.load_flat("cornell", "ours1931")
    with pu.gpu_context(flat) as ctx:
        got = ctx.eval_math(3, x, np.float32(y))
    import ctypes as C
    want = np.empty_like(x)
    pu.oracle().ssb_oracle_eval_math(3, x.ctypes.data_as(C.POINTER(C.c_float)), np.float32(y), want.ctypes.data_as(C.POINTER(C.c_float)), x.size)
    assert pu.bits_equal(got, want)


@pytest.mark.parametrize("scene,variant", CASES)
def test_gpu_matches_oracle_and_reference_fixture(scene, variant):
    _skip_if_no_assets(scene, variant)
    flat = pu.load_flat(scene, variant)
    opt = pu.options(variant, 32, 24, 4, seed=7)
    acc_o, samp_o, _ = pu.oracle_render(flat, opt, want_samples=True)
    xo, so = pu.oracle_resolve(flat, opt, acc_o)
    with pu.gpu_context(flat) as ctx:
        xg, sg = ctx.render_frame(opt)
        acc_g = ctx.read_accum(opt.width, opt.height)
        for (px, py) in ((0, 0), (16, 12), (31, 23), (5, 20)):
            sm = ctx.trace_samples(opt, px, py)
            assert pu.bits_equal(sm, samp_o[py, px]), f"per-sample mismatch at pixel ({px},{py})"
    assert pu.bits_equal(acc_g, acc_o), f"accumulator differs: max rel {pu.rel_err(acc_g, acc_o).max()}"
    assert pu.bits_equal(xg, xo)
    ref = np.load(os.path.join(pu.GOLDEN, f"xyza_{scene}_{variant}_32x24_spp4_seed7.npy"))
    assert pu.bits_equal(xg, ref), "differs from the real reference's XYZA fixture"
    ok = np.isclose(sg, so, rtol=1e-6, atol=1e-7, equal_nan=True)
    assert ok.all()
    assert pu.bits_equal(sg, so), "sRGB tonemap not bit-identical"


def test_gpu_config1_bit_exact():
    """BASELINE.json configs[0]: cornell-srgb 128x128 spp16 — CUDA == oracle == reference (sha)."""
    import hashlib, json
    _skip_if_no_assets("cornell-srgb", "ours1931")
    flat = pu.load_flat("cornell-srgb", "ours1931")
    opt = pu.options("ours1931", 128, 128, 16, seed=1)
    with pu.gpu_context(flat) as ctx:
        xg, _ = ctx.render_frame(opt)
    idx = json.load(open(os.path.join(pu.GOLDEN, "golden_index.json")))["cornell-srgb_ours1931_128x128_spp16_seed1"]
    assert hashlib.sha256(xg.tobytes()).hexdigest() == idx["sha256"]


def test_gpu_rgb_config1_bit_exact():
    """RENDER_MODE_RGB, cornell-srgb 128x128 spp16: CUDA == the RGB build of the real reference (sha of the l-RGB+alpha buffer)."""
    import hashlib, json
    _skip_if_no_assets("cornell-srgb", "rgb")
    flat = pu.load_flat("cornell-srgb", "rgb")
    opt = pu.options("rgb", 128, 128, 16, seed=1)
    with pu.gpu_context(flat) as ctx:
        avg, _ = ctx.render_frame(opt)
    idx = json.load(open(os.path.join(pu.GOLDEN, "golden_index.json")))["cornell-srgb_rgb_128x128_spp16_seed1"]
    assert hashlib.sha256(avg.tobytes()).hexdigest() == idx["sha256"]


def test_gpu_spectral_scene_needs_spectra():
    """A scene uploaded with RGB constants only (no spectra) must be refused in spectral mode, not rendered black."""
    import importlib
    ssb = importlib.import_module("simple-spectral_b200")
    flat = pu.load_flat("cornell", "rgb")
    with pu.gpu_context(flat) as ctx:
        with pytest.raises(ssb.SsbError) as e:
            ctx.render(pu.options("ours1931", 16, 16, 1))
        assert e.value.code == -2


def test_gpu_rejects_unsupported_wavelength_counts():
    import importlib
    ssb = importlib.import_module("simple-spectral_b200")
    flat = pu.load_flat("cornell", "ours1931")
    with pu.gpu_context(flat) as ctx:
        for n in (1, 5):
            with pytest.raises(ssb.SsbError) as e:
                ctx.render(pu.options("ours1931", 16, 16, 1, n_wavelengths=n))
            assert e.value.code == -3


def test_gpu_subsets_compose():
    """tile rectangles x sample ranges accumulate to the full-frame result (the multi-GPU sharding contract);
    keep_accumulator lets several rectangles of the first sample range share one accumulator."""
    flat = pu.load_flat("cornell", "ours1931")
    full = pu.options("ours1931", 40, 30, 6, seed=3)
    with pu.gpu_context(flat) as ctx:
        ctx.render(full)
        want = ctx.read_accum(40, 30)
        ctx.clear()
        for (s0, s1) in ((0, 2), (2, 6)):
            for (x0, x1) in ((0, 17), (17, 40)):
                ctx.render(pu.options("ours1931", 40, 30, 6, seed=3, x0=x0, x1=x1, sample_begin=s0, sample_end=s1, keep_accumulator=1))
        got = ctx.read_accum(40, 30)
        ctx.render(pu.options("ours1931", 40, 30, 6, seed=3, x0=0, x1=17, sample_begin=0, sample_end=2))  # default: clears first
        part = ctx.read_accum(40, 30)
    assert pu.bits_equal(got, want)
    assert (part[:, 17:] == 0).all() and part[:, :17].any()


def test_gpu_full_size_properties():
    """BASELINE configs[1] size (512x512 spp64): size-independent properties — alpha is the hit fraction,
    corners miss the box, image-mean within Monte-Carlo noise of the config-1 reference mean."""
    import json
    _skip_if_no_assets("cornell-srgb", "ours1931")
    flat = pu.load_flat("cornell-srgb", "ours1931")
    opt = pu.options("ours1931", 512, 512, 64, seed=1)
    with pu.gpu_context(flat) as ctx:
        xg, sg = ctx.render_frame(opt)
        st = ctx.stats()
    assert st.samples == 512 * 512 * 64
    assert np.isfinite(xg).all()
    a = xg[..., 3]
    assert (a >= 0).all() and (a <= 1.0 + 1e-6).all()
    assert a[0, 0] == 0 and a[-1, -1] == 0 and abs(a[256, 256] - 1.0) < 1e-6
    ref_mean = np.array(json.load(open(os.path.join(pu.GOLDEN, "golden_index.json")))["cornell-srgb_ours1931_128x128_spp16_seed1"]["mean"])
    got_mean = xg.mean(axis=(0, 1))
    assert np.allclose(got_mean, ref_mean, rtol=0.02), (got_mean, ref_mean)


def test_gpu_config2_bit_exact_vs_oracle():
    """BASELINE.json configs[1] at full size: cornell-srgb 512x512 spp64 (16.8 M samples) — the CUDA accumulators equal
    the oracle's bit for bit (the oracle runs multi-threaded on the box's host cores)."""
    _skip_if_no_assets("cornell-srgb", "ours1931")
    flat = pu.load_flat("cornell-srgb", "ours1931")
    opt = pu.options("ours1931", 512, 512, 64, seed=1)
    with pu.gpu_context(flat) as ctx:
        ctx.render(opt)
        acc_g = ctx.read_accum(512, 512)
    acc_o, _, _ = pu.oracle_render(flat, opt)
    assert pu.bits_equal(acc_g, acc_o), f"max rel {pu.rel_err(acc_g, acc_o).max()}"


def test_gpu_config3_full_frame_bit_exact_vs_oracle():
    """BASELINE.json configs[2] at FULL size and over the WHOLE frame: cornell (measured spectra) 512x512 spp256 under the
    CIE 2006 observer, 67 M path samples — every pixel's f64 accumulator equals the oracle's (multi-threaded on the host
    cores; the windowed check of the parametrised test below covers C4 / C5, whose CPU side would take minutes to hours)."""
    flat = pu.load_flat("cornell", "ours2006")
    opt = pu.options("ours2006", 512, 512, 256, seed=1)
    with pu.gpu_context(flat) as ctx:
        ctx.render(opt)
        acc_g = ctx.read_accum(512, 512)
    acc_o, _, _ = pu.oracle_render(flat, opt)
    assert pu.bits_equal(acc_g, acc_o), f"max rel {pu.rel_err(acc_g, acc_o).max()}"


def test_gpu_multipass_equals_single_pass(monkeypatch):
    """Frames whose path state does not fit the per-pass memory budget are rendered in several sample passes
    (configs 3-5 of BASELINE.json need this): the accumulators must not depend on the pass split."""
    flat = pu.load_flat("cornell", "ours1931")
    opt = pu.options("ours1931", 64, 48, 8, seed=5)
    with pu.gpu_context(flat) as ctx:
        ctx.render(opt)
        want = ctx.read_accum(64, 48)
        assert ctx.stats().launches < 60
    monkeypatch.setenv("SSB_WAVE_BUDGET_MB", "4")  # 64*48 pixels * ~780 B -> 1 sample per pass
    with pu.gpu_context(flat) as ctx:
        ctx.render(opt)
        got = ctx.read_accum(64, 48)
        assert ctx.stats().launches > 100  # several passes
    assert pu.bits_equal(got, want)


@pytest.mark.parametrize("name,scene,variant,w,h,spp", [
    ("C3", "cornell", "ours2006", 512, 512, 256),
    ("C4", "plane-srgb", "jh", 1024, 1024, 1024),
    ("C5", "cornell-srgb", "meng", 2048, 2048, 4096),
])
def test_gpu_full_size_configs_spot_checked(name, scene, variant, w, h, spp):
    """BASELINE.json configs[2..4] at their FULL sizes (0.07 / 1.07 / 17.2 G path samples, rendered in many wavefront
    passes): the whole frame is rendered on the GPU; the oracle renders three 6x6-pixel windows of the same job (all spp
    samples of those pixels) and the f64 accumulators of the windows must agree bit for bit.  Frame-wide properties:
    finite, alpha = hit fraction in [0,1], every pixel touched exactly once per sample (alpha*spp is an integer)."""
    _skip_if_no_assets(scene, variant)
    flat = pu.load_flat(scene, variant)
    opt = pu.options(variant, w, h, spp, seed=1)
    with pu.gpu_context(flat) as ctx:
        ctx.render(opt)
        acc_g = ctx.read_accum(w, h)
        st = ctx.stats()
    assert st.samples == w * h * spp
    if variant == "jh":
        # the texture holds one black texel, for which rgb2spec_fetch yields NaN coefficients (scale = inf * 0,
        # rgb2spec.c:84-90): the reference's pixels that sample it are NaN too
        nan_px = np.isnan(acc_g[..., :3]).any(axis=-1).sum()
        assert nan_px < 0.01 * w * h and np.isfinite(acc_g[..., 3]).all()
    else:
        assert np.isfinite(acc_g).all()
    hits = acc_g[..., 3] / np.float64(np.float32(0.001))  # each hit adds double(1.0f * 0.001f)
    assert (np.abs(hits - np.round(hits)) < 1e-6 * spp).all() and hits.min() >= 0 and hits.max() <= spp * (1 + 1e-9)
    for (x0, y0) in ((w // 2 - 3, h // 2 - 3), (w // 5, h // 3), (w - 40, h - 50)):
        o = pu.options(variant, w, h, spp, seed=1, x0=x0, y0=y0, x1=x0 + 6, y1=y0 + 6)
        acc_o, _, _ = pu.oracle_render(flat, o)
        assert pu.bits_equal(acc_g[y0:y0 + 6, x0:x0 + 6], acc_o[y0:y0 + 6, x0:x0 + 6]), f"{name}: window at ({x0},{y0}) differs"
