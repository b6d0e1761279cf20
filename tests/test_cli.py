"""The command-line front end mirrors the reference's flags (src/main.cpp:33-162) and error behaviour."""
import os
import subprocess

import numpy as np
import pytest

import parity_util as pu

CLI = os.path.join(pu.ROOT, "simple-spectral_b200", "simple_spectral_b200")


def _run(*args):
    return subprocess.run([CLI, *args], capture_output=True, text=True)


def test_cli_usage_and_error_codes():
    assert os.path.exists(CLI), "build with __graft_entry__.build()"
    r = _run()  # missing required arguments -> usage, return -1 (main.cpp:172-177)
    assert r.returncode == 255 and "Required argument `--scene`" in r.stderr and "--samples" in r.stdout
    r = _run("--scene=nope", "-w=8", "-h=8", "-spp=1", "--output=/tmp/x.png")
    assert r.returncode == 255 and "Unrecognized scene" in r.stderr
    r = _run("--scene=cornell", "-w=0", "-h=8", "-spp=1", "--output=/tmp/x.png")
    assert r.returncode == 255 and "Invalid width or height" in r.stderr
    r = _run("--scene=cornell", "-w=8", "-h=8", "-spp=1", "--output=/tmp/x.png", "-io=1")
    assert r.returncode == 255 and "does not take a value" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["ours1931", "rgb"])
def test_cli_renders_the_oracle_image(tmp_path, variant):
    """`simple_spectral_b200 --scene=cornell ...` writes the same PFM the oracle's framebuffer gives
    (host-layer scene + CUDA path vs reference-dumped scene + oracle), in spectral and in RGB mode."""
    import importlib
    host = importlib.import_module("simple-spectral_b200.host")
    out = str(tmp_path / "o.pfm")
    r = _run("--scene=cornell", "-w=32", "-h=24", "-spp=4", f"--output={out}", "--seed=7", f"--data-root={pu.data_root()}", f"--variant={variant}")
    assert r.returncode == 0, r.stderr
    assert "Render completed in" in r.stdout
    flat = pu.load_flat("cornell", variant)
    opt = pu.options(variant, 32, 24, 4, seed=7)
    acc, _, _ = pu.oracle_render(flat, opt)
    _, srgba = pu.oracle_resolve(flat, opt, acc)
    want = str(tmp_path / "w.pfm")
    host.save_image(want, srgba)
    assert open(out, "rb").read() == open(want, "rb").read()
    if variant == "ours1931":  # ... and the file the real reference wrote for this command line (tests/golden/refout_*)
        assert open(out, "rb").read() == open(os.path.join(pu.GOLDEN, "refout_cornell_ours1931_32x24_spp4_seed7.pfm"), "rb").read()
