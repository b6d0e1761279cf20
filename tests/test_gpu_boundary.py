"""GPU: the drop-in boundary COMPILED INTO THE REAL REFERENCE.  oracle/build_ref.py (tool "ssb200") builds the reference's
own main.cpp / Color::init / Scene::get_new_* / Framebuffer::save with integration/renderer_ssb200.cpp — the stub of
INTEGRATION.md — in place of Renderer::render_start / _render_threadwork / _render_pixel / _render_sample
(renderer.cpp:103-422), linked against simple-spectral_b200/libssb200.so.  That binary's image files must equal, byte for
byte, the files the reference's own CPU loop writes at the same per-sample seeds: the committed fixtures
(tests/golden/refout_*) and live runs of the hooked reference (oracle/_ref/*_hooked, CPU, small frames)."""
import os
import subprocess

import numpy as np
import pytest

import parity_util as pu

pytestmark = pytest.mark.gpu
REFDIR = os.path.join(pu.ROOT, "oracle", "_ref")


def _run(binary, scene, w, h, spp, out, seed=7, extra_env=None):
    exe = os.path.join(REFDIR, binary)
    if not os.path.exists(exe):
        pytest.skip(f"{exe} not built (run __graft_entry__.build() where /root/reference exists)")
    env = dict(os.environ, SSB_SEED=str(seed), **(extra_env or {}))
    r = subprocess.run([exe, f"--scene={scene}", f"-w={w}", f"-h={h}", f"-spp={spp}", f"--output={out}"], cwd=pu.data_root(), env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-500:], r.stderr[-500:])
    return r


@pytest.mark.skipif(not pu.have_assets(), reason="data files not staged (assets/data)")
def test_reference_binary_with_the_stub_writes_the_reference_files(tmp_path):
    from PIL import Image
    for ext in ("pfm", "hdr", "csv", "png"):
        out = str(tmp_path / f"o.{ext}")
        _run("simple_spectral_ssb200", "cornell", 32, 24, 4, out)
        ref = os.path.join(pu.GOLDEN, f"refout_cornell_ours1931_32x24_spp4_seed7.{ext}")
        if ext == "png":
            assert np.array_equal(np.asarray(Image.open(out)), np.asarray(Image.open(ref)))
        else:
            assert open(out, "rb").read() == open(ref, "rb").read(), ext


@pytest.mark.skipif(not pu.have_assets(), reason="data files not staged (assets/data)")
@pytest.mark.parametrize("stub,hooked,scene,w,h,spp", [
    ("simple_spectral_ssb200", "simple_spectral_ours1931_hooked", "cornell-srgb", 64, 48, 8),
    ("simple_spectral_ssb200", "simple_spectral_ours1931_hooked", "plane-srgb", 40, 56, 5),
    ("simple_spectral_ssb200_jh", "simple_spectral_jh_hooked", "cornell-srgb", 48, 32, 6),
])
def test_stub_binary_equals_the_hooked_reference_live(tmp_path, stub, hooked, scene, w, h, spp):
    a, b = str(tmp_path / "gpu.pfm"), str(tmp_path / "cpu.pfm")
    _run(stub, scene, w, h, spp, a, seed=11)
    _run(hooked, scene, w, h, spp, b, seed=11)
    assert open(a, "rb").read() == open(b, "rb").read()
