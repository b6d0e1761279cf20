"""CPU: the Renderer facade of the C++ host layer (csrc/host/host_renderer.cpp; reference src/renderer.hpp:71-81,
renderer.cpp:396-430, main.cpp:313-327) with the device layer answered by the oracle (tests/stub/ssb_device_stub.c —
test infrastructure, never shipped).  What is checked is the HOST logic: the worker thread's life cycle
(start / is_rendering / stop / wait), the progressive sample slices and their previews, that slicing does not change the
finished frame by a bit, error propagation out of the worker, and the image written at the end.
The same life cycle runs against the CUDA path in tests/test_zz_gpu_prebake_progressive.py."""
import ctypes as C
import glob
import os
import subprocess
import time

import numpy as np
import pytest

import parity_util as pu
from importlib import import_module

host = import_module("simple-spectral_b200.host")
ROOT = pu.ROOT


@pytest.fixture(scope="module")
def stub(tmp_path_factory):
    if not pu.have_assets():
        pytest.skip("data files not staged (assets/data)")
    out = str(tmp_path_factory.mktemp("stub") / "libssbh_stub.so")
    hdir = os.path.join(ROOT, "simple-spectral_b200", "csrc", "host")
    odir = os.path.join(ROOT, "oracle")
    obj = out + ".stub.o"
    subprocess.run(["gcc", "-std=c11", "-O1", "-fPIC", "-c", os.path.join(ROOT, "tests", "stub", "ssb_device_stub.c"), "-o", obj], check=True)
    subprocess.run(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-o", out,
                    *sorted(glob.glob(os.path.join(hdir, "*.cpp"))), obj, "-L", odir, "-lssb_oracle", "-lz",
                    f"-Wl,-rpath,{odir}"], check=True)
    # the command-line front end, linked against the same stub (test_cli_stub_*)
    subprocess.run(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-pthread", "-o", out[:-3] + "_cli",
                    os.path.join(ROOT, "simple-spectral_b200", "csrc", "cli", "main.cpp"), out, f"-Wl,-rpath,{os.path.dirname(out)}",
                    f"-Wl,-rpath,{odir}"], check=True)
    L = C.CDLL(out)
    L._cli = out[:-3] + "_cli"
    P = C.POINTER
    L.ssbh_last_error.restype = C.c_char_p
    L.ssbh_renderer_new.argtypes = [P(host.ssbh_renderer_options), P(C.c_void_p)]
    for n in ("ssbh_renderer_render", "ssbh_renderer_start", "ssbh_renderer_wait", "ssbh_renderer_is_rendering"):
        getattr(L, n).argtypes = [C.c_void_p]
        getattr(L, n).restype = C.c_int
    L.ssbh_renderer_stop.argtypes = [C.c_void_p]
    L.ssbh_renderer_stop.restype = None
    L.ssbh_renderer_snapshot.argtypes = [C.c_void_p, P(C.c_float)]
    L.ssbh_renderer_snapshot.restype = C.c_uint32
    L.ssbh_renderer_framebuffer.argtypes = [C.c_void_p]
    L.ssbh_renderer_framebuffer.restype = P(C.c_float)
    L.ssbh_renderer_xyza.argtypes = [C.c_void_p]
    L.ssbh_renderer_xyza.restype = P(C.c_double)
    L.ssbh_renderer_stats.argtypes = [C.c_void_p, P(pu.abi.ssb_stats)]
    L.ssbh_renderer_free.argtypes = [C.c_void_p]
    L.ssbh_renderer_free.restype = None
    return L


W, H = 16, 12


def _new(L, spp, progressive, output=None, scene=b"cornell", devices=None, shard=0, band_height=0):
    keep = (scene, output.encode() if output else None, pu.data_root().encode(), (C.c_int * len(devices))(*devices) if devices else None)
    o = host.ssbh_renderer_options(keep[0], W, H, spp, 0, keep[1], 1931, pu.abi.SSB_UPSAMPLE_OURS, 1, 10, 1, 7, 0, keep[2],
                                   pu.abi.SSB_RENDER_SPECTRAL, 4, 0, int(progressive))
    if devices:
        o.devices, o.ndevices = keep[3], len(devices)
    o.shard, o.band_height = shard, band_height
    h = C.c_void_p()
    rc = L.ssbh_renderer_new(C.byref(o), C.byref(h))
    assert rc == 0, L.ssbh_last_error()
    h._keep = keep
    return h


def _results(L, h):
    n = W * H * 4
    fb = np.ctypeslib.as_array(L.ssbh_renderer_framebuffer(h), shape=(n,)).reshape(H, W, 4).copy()
    xyza = np.ctypeslib.as_array(L.ssbh_renderer_xyza(h), shape=(n,)).reshape(H, W, 4).copy()
    st = pu.abi.ssb_stats()
    assert L.ssbh_renderer_stats(h, C.byref(st)) == 0
    return xyza, fb, st


def test_progressive_slices_do_not_change_the_frame(stub):
    L = stub
    spp = 11  # slices [0,1) [1,2) [2,4) [4,8) [8,11)
    one = _new(L, spp, False)
    assert L.ssbh_renderer_render(one) == 0, L.ssbh_last_error()
    x1, f1, s1 = _results(L, one)
    L.ssbh_renderer_free(one)
    assert s1.launches == 1 and s1.samples == W * H * spp
    prog = _new(L, spp, True)
    assert L.ssbh_renderer_start(prog) == 0
    seen = set()
    snap = np.empty((H, W, 4), np.float32)
    while L.ssbh_renderer_is_rendering(prog):
        seen.add(L.ssbh_renderer_snapshot(prog, snap.ctypes.data_as(C.POINTER(C.c_float))))
        time.sleep(0.0005)
    assert L.ssbh_renderer_wait(prog) == 0, L.ssbh_last_error()
    assert L.ssbh_renderer_snapshot(prog, None) == spp
    x2, f2, s2 = _results(L, prog)
    L.ssbh_renderer_free(prog)
    assert seen <= {0, 1, 2, 4, 8, 11}, seen  # only whole slices are ever shown
    assert s2.launches == 5 and s2.samples == W * H * spp
    assert pu.bits_equal(x1, x2) and pu.bits_equal(f1, f2)
    # ... and the frame is what the checker computes directly from the reference's dumped tables
    flat = pu.load_flat("cornell", "ours1931")
    opt = pu.options("ours1931", W, H, spp, seed=7)
    xo, so = pu.oracle_resolve(flat, opt, pu.oracle_render(flat, opt)[0])
    assert pu.bits_equal(x2, xo) and pu.bits_equal(f2, so)


def test_preview_is_the_average_of_the_samples_so_far(stub):
    """After the slice ending at sample n the framebuffer is the spp = n frame (renderer.cpp:296 with spp = n)."""
    L = stub
    full = _new(L, 2, True)  # slices [0,1) [1,2): stop after the first one is visible
    first = _new(L, 1, False)
    assert L.ssbh_renderer_render(first) == 0
    _, f_first, _ = _results(L, first)
    L.ssbh_renderer_free(first)
    assert L.ssbh_renderer_start(full) == 0
    snap = np.empty((H, W, 4), np.float32)
    got_first = False
    while L.ssbh_renderer_is_rendering(full):
        if L.ssbh_renderer_snapshot(full, snap.ctypes.data_as(C.POINTER(C.c_float))) == 1:
            got_first = True
            assert pu.bits_equal(snap, f_first)
            break
        time.sleep(0.0002)
    assert L.ssbh_renderer_wait(full) == 0
    L.ssbh_renderer_free(full)
    if not got_first:
        pytest.skip("the two slices finished between two polls")


def test_stop_ends_the_render_after_the_slice_in_flight_and_saves(stub, tmp_path):
    L = stub
    out = str(tmp_path / "aborted.pfm")
    spp = 1 << 14  # would take minutes on the CPU checker
    h = _new(L, spp, True, output=out)
    assert L.ssbh_renderer_start(h) == 0
    assert L.ssbh_renderer_start(h) == pu.abi.SSB_ERR_ARG  # one render at a time
    t0 = time.time()
    while L.ssbh_renderer_snapshot(h, None) < 2 and time.time() - t0 < 60:
        time.sleep(0.001)
    L.ssbh_renderer_stop(h)
    assert L.ssbh_renderer_wait(h) == 0, L.ssbh_last_error()
    assert not L.ssbh_renderer_is_rendering(h)
    done = L.ssbh_renderer_snapshot(h, None)
    assert 2 <= done < spp and done & (done - 1) == 0
    x, f, st = _results(L, h)
    assert st.samples == W * H * done
    # the aborted render is saved with what it has (renderer.cpp:388-394), and that is the spp = done frame
    ref = _new(L, done, False)
    assert L.ssbh_renderer_render(ref) == 0
    xr, fr, _ = _results(L, ref)
    L.ssbh_renderer_free(ref)
    assert pu.bits_equal(x, xr) and pu.bits_equal(f, fr)
    assert os.path.getsize(out) > W * H * 12
    # the renderer can be started again
    L.ssbh_renderer_stop(h)
    L.ssbh_renderer_free(h)


def test_multi_device_tiles_are_bit_identical_and_samples_differ_only_by_summation_order(stub):
    """The Renderer's multi-GPU path (RendererOptions::devices; reference renderer.cpp:396-430 spreads its tiles over all
    cores): one context + host thread per device, interleaved row bands or sample ranges, merged on the first device.
    Row bands (disjoint pixels, per-sample seeding) must give the single-device frame bit for bit — also progressively,
    with more devices than bands, and with a device listed twice; sample ranges add partial f64 sums in device order."""
    L = stub
    spp = 6
    one = _new(L, spp, False)
    assert L.ssbh_renderer_render(one) == 0, L.ssbh_last_error()
    x1, f1, s1 = _results(L, one)
    L.ssbh_renderer_free(one)
    for devices, band_h, progressive in (([0, 1, 2], 0, False), ([0, 1, 2, 3], 5, True), ([1, 1], 1, False), ([0, 1, 2, 3], 4, False)):
        h = _new(L, spp, progressive, devices=devices, shard=host.SSBH_SHARD_TILES, band_height=band_h)
        assert L.ssbh_renderer_render(h) == 0, L.ssbh_last_error()
        x, f, st = _results(L, h)
        L.ssbh_renderer_free(h)
        assert st.samples == W * H * spp, (devices, st.samples)
        assert pu.bits_equal(x, x1) and pu.bits_equal(f, f1), (devices, band_h, progressive)
    for devices, progressive in (([0, 1, 2], False), ([0, 1, 2, 3], True)):
        h = _new(L, spp, progressive, devices=devices, shard=host.SSBH_SHARD_SAMPLES)
        assert L.ssbh_renderer_render(h) == 0, L.ssbh_last_error()
        x, f, st = _results(L, h)
        L.ssbh_renderer_free(h)
        assert st.samples == W * H * spp
        assert pu.rel_err(x, x1).max() <= 1e-14
    bad = _new(L, spp, False, devices=[0, 1], shard=host.SSBH_SHARD_TILES)
    L.ssbh_renderer_free(bad)
    # a device the box does not have: the constructor fails and nothing leaks
    keep = (b"cornell", None, pu.data_root().encode(), (C.c_int * 2)(0, 9))
    o = host.ssbh_renderer_options(keep[0], W, H, spp, 0, keep[1], 1931, pu.abi.SSB_UPSAMPLE_OURS, 1, 10, 1, 7, 0, keep[2],
                                   pu.abi.SSB_RENDER_SPECTRAL, 4, 0, 0)
    o.devices, o.ndevices = keep[3], 2
    h = C.c_void_p()
    assert L.ssbh_renderer_new(C.byref(o), C.byref(h)) == pu.abi.SSB_ERR_ARG and not h


def test_worker_errors_surface_in_wait(stub):
    L = stub
    h = _new(L, 0x7fffffff, False)  # the stub's ssb_render fails for this spp
    assert L.ssbh_renderer_start(h) == 0
    rc = L.ssbh_renderer_wait(h)
    assert rc == pu.abi.SSB_ERR_DATA and b"injected" in L.ssbh_last_error()
    assert not L.ssbh_renderer_is_rendering(h)
    L.ssbh_renderer_free(h)


def test_cli_writes_the_files_the_reference_writes(stub, tmp_path):
    """The whole drop-in chain above the device layer on the CPU: command line -> host scene / colour construction ->
    Renderer worker -> (oracle instead of the GPU) -> Framebuffer::save, against the files the REAL reference wrote for
    the same command line at the same seed (tests/golden/refout_*): PFM / HDR / CSV byte for byte, PNG pixel for pixel.
    Progressive slices + a preview file must not change them."""
    from PIL import Image
    base = [stub._cli, "--scene=cornell", "-w=32", "-h=24", "-spp=4", "--seed=7", f"--data-root={pu.data_root()}"]
    for ext, extra in (("pfm", []), ("hdr", []), ("csv", ["--progressive"]), ("png", [f"--preview={tmp_path}/prev.png"])):
        out = str(tmp_path / f"o.{ext}")
        r = subprocess.run([*base, f"--output={out}", *extra], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert "Render completed in" in r.stdout
        ref = os.path.join(pu.GOLDEN, f"refout_cornell_ours1931_32x24_spp4_seed7.{ext}")
        if ext == "png":
            assert np.array_equal(np.asarray(Image.open(out)), np.asarray(Image.open(ref)))
        else:
            assert open(out, "rb").read() == open(ref, "rb").read(), ext
