"""GPU: multi-GPU rendering inside the library (the reference's Renderer spreads its 8x8 tiles over every
core of the box, renderer.cpp:396-430, framebuffer.hpp:14-21).  Device layer: ssb_options.band_* (interleaved row bands)
+ ssb_accum_merge (gather / f64 add into another context's accumulator, peer loads over NVLink between GPUs); host layer:
RendererOptions::devices / shard, the CLI's --devices.  Row bands must reproduce the single-GPU frame BIT FOR BIT (disjoint
pixels, per-sample seeding); sample ranges differ only by the order of the f64 additions (<= 1e-14 relative).
Every case runs with two contexts on device 0 (so it also runs on a one-GPU box) and, when the box has them, on all GPUs."""
import importlib
import os
import subprocess

import numpy as np
import pytest

import parity_util as pu

abi = pu.abi
host = importlib.import_module("simple-spectral_b200.host")
pytestmark = pytest.mark.gpu


def _device_sets():
    n = pu.ssb.device_count()
    sets = [[0, 0], [0, 0, 0]]
    if n >= 2:
        sets += [list(range(n)), [n - 1, 0]]
    return sets


def test_band_shards_merge_bit_identically_through_the_c_abi():
    flat = pu.load_flat("cornell", "ours1931")
    W, H, SPP = 64, 52, 6  # 52 rows: the last band of height 8 is partial
    opt = pu.options("ours1931", W, H, SPP, seed=3)
    with pu.gpu_context(flat) as ctx:
        ctx.render(opt)
        want = ctx.read_accum(W, H)
    for devs in _device_sets():
        for band_h in (8, 1, 5, 64):
            ctxs = [pu.gpu_context(flat, d) for d in devs]
            parts = []
            for i, c in enumerate(ctxs):
                o = pu.options("ours1931", W, H, SPP, seed=3, band_height=band_h, band_count=len(devs), band_index=i)
                c.render(o)
                parts.append(o)
            total = pu.ssb.Context(devs[0])
            for c, o in zip(ctxs, parts):
                total.merge_from(c, o)
            got = total.read_accum(W, H)
            for c in ctxs + [total]:
                c.close()
            assert pu.bits_equal(got, want), (devs, band_h)


def test_sample_and_rectangle_shards_through_the_c_abi():
    flat = pu.load_flat("cornell", "ours1931")
    W, H, SPP = 48, 40, 9
    opt = pu.options("ours1931", W, H, SPP, seed=5)
    with pu.gpu_context(flat) as ctx:
        ctx.render(opt)
        want = ctx.read_accum(W, H)
    for devs in _device_sets():
        n = len(devs)
        ctxs = [pu.gpu_context(flat, d) for d in devs]
        # sample ranges of the whole frame: added
        total = pu.ssb.Context(devs[0])
        for i, c in enumerate(ctxs):
            o = pu.options("ours1931", W, H, SPP, seed=5, sample_begin=SPP * i // n, sample_end=SPP * (i + 1) // n, keep_accumulator=1)
            c.render(pu.options("ours1931", W, H, SPP, seed=5, x0=W, x1=W))  # allocate + clear, trace nothing
            c.render(o)
            total.merge_from(c, o)
        got = total.read_accum(W, H)
        assert pu.rel_err(got, want).max() <= 1e-14 and np.allclose(got[..., 3], want[..., 3], rtol=1e-14, atol=0)
        # pixel rectangles (column strips): copied, bit-exact
        total2 = pu.ssb.Context(devs[0])
        for i, c in enumerate(ctxs):
            o = pu.options("ours1931", W, H, SPP, seed=5, x0=W * i // n, x1=max(W * (i + 1) // n, 1))
            if o.x1 <= o.x0:
                continue
            c.render(o)
            total2.merge_from(c, o)
        got2 = total2.read_accum(W, H)
        for c in ctxs + [total, total2]:
            c.close()
        assert pu.bits_equal(got2, want), devs


@pytest.mark.skipif(not pu.have_assets(), reason="data files not staged (assets/data)")
def test_renderer_on_several_devices_equals_one_device(tmp_path):
    W, H, SPP = 96, 80, 8
    one = host.Renderer("cornell-srgb", W, H, SPP, seed=2)
    x1, f1 = one.render()
    s1 = one.stats()
    one.close()
    for devs in _device_sets():
        for progressive in (False, True):
            r = host.Renderer("cornell-srgb", W, H, SPP, seed=2, devices=devs, shard="tiles", progressive=progressive)
            x, f = r.render()
            st = r.stats()
            r.close()
            assert st.samples == s1.samples == W * H * SPP
            assert pu.bits_equal(x, x1) and pu.bits_equal(f, f1), (devs, progressive)
        r = host.Renderer("cornell-srgb", W, H, SPP, seed=2, devices=devs, shard="samples")
        x, f = r.render()
        r.close()
        assert pu.rel_err(x, x1).max() <= 1e-14
    # the command-line front end: --devices, same file as one device
    cli = os.path.join(pu.ROOT, "simple-spectral_b200", "simple_spectral_b200")
    base = [cli, "--scene=cornell-srgb", f"-w={W}", f"-h={H}", f"-spp={SPP}", "--seed=2", f"--data-root={pu.data_root()}"]
    a, b = str(tmp_path / "one.pfm"), str(tmp_path / "many.pfm")
    n = pu.ssb.device_count()
    for out, extra in ((a, []), (b, [f"--devices={'0-%d' % (n - 1) if n >= 2 else '0,0'}"])):
        r = subprocess.run([*base, f"--output={out}", *extra], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    assert open(a, "rb").read() == open(b, "rb").read()
