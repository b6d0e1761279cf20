"""CPU: the oracle (oracle/ssb_oracle.c) against the committed fixtures generated from the REAL
reference (tests/golden/make_golden.py).  Bar: bit-identical double XYZA per pixel."""
import hashlib
import json
import os

import numpy as np
import pytest

import parity_util as pu

SMALL = [("cornell", "ours1931"), ("cornell-srgb", "ours1931"), ("plane-srgb", "ours1931"),
         ("cornell", "ours2006"), ("cornell-srgb", "ours2006"),
         ("cornell-srgb", "jh"), ("plane-srgb", "jh"), ("cornell-srgb", "meng"), ("plane-srgb", "meng"),
         ("plane-srgb", "ours1931_noels"), ("cornell", "ours1931_noels"),
         ("cornell", "ours1931_noffc"), ("cornell-srgb", "ours1931_noffc"),  # FLAT_FIELD_CORRECTION off
         ("cornell", "rgb"), ("cornell-srgb", "rgb"), ("plane-srgb", "rgb"),  # RENDER_MODE_RGB
         ("cornell-srgb", "ours1931_nw3"), ("plane-srgb", "ours1931_nw3"), ("cornell-srgb", "meng_nw2"),  # SAMPLE_WAVELENGTHS 3 / 2
         ("cornell-srgb", "ours1931_d3")]  # MAX_DEPTH 3


@pytest.mark.parametrize("scene,variant", SMALL)
def test_oracle_matches_reference_fixture(scene, variant):
    if pu.needs_assets(scene, variant) and not pu.have_assets():
        pytest.skip("data files (texture / JH / Meng tables) not staged")
    flat = pu.load_flat(scene, variant)
    opt = pu.options(variant, 32, 24, 4, seed=7)
    acc, _, _ = pu.oracle_render(flat, opt)
    xyza, _ = pu.oracle_resolve(flat, opt, acc)
    ref = np.load(os.path.join(pu.GOLDEN, f"xyza_{scene}_{variant}_32x24_spp4_seed7.npy"))
    assert pu.bits_equal(xyza, ref), f"max rel err {pu.rel_err(xyza, ref).max()}"


def test_oracle_indirect_only():
    if not pu.have_assets():
        pytest.skip("texture not staged")
    flat = pu.load_flat("cornell-srgb", "ours1931")
    opt = pu.options("ours1931", 32, 24, 4, seed=7, indirect_only=1)
    acc, _, _ = pu.oracle_render(flat, opt)
    xyza, _ = pu.oracle_resolve(flat, opt, acc)
    ref = np.load(os.path.join(pu.GOLDEN, "xyza_cornell-srgb_ours1931_indirect_32x24_spp4_seed7.npy"))
    assert pu.bits_equal(xyza, ref)


@pytest.mark.parametrize("scene", ["cornell", "cornell-srgb"])
def test_oracle_config1_sha(scene):
    """BASELINE.json configs[0]: 128x128 spp16 — sha256 of the reference's XYZA buffer."""
    if pu.needs_assets(scene, "ours1931") and not pu.have_assets():
        pytest.skip("texture not staged")
    idx = json.load(open(os.path.join(pu.GOLDEN, "golden_index.json")))[f"{scene}_ours1931_128x128_spp16_seed1"]
    flat = pu.load_flat(scene, "ours1931")
    opt = pu.options("ours1931", 128, 128, 16, seed=1)
    acc, _, cnt = pu.oracle_render(flat, opt, counters=True)
    xyza, _ = pu.oracle_resolve(flat, opt, acc)
    assert hashlib.sha256(xyza.tobytes()).hexdigest() == idx["sha256"]
    # path statistics of SURVEY.md §6 (5.30 closest-hit and 3.14 shadow queries per sample)
    assert abs(cnt.closest_queries / cnt.samples - 5.30) < 0.05
    assert abs(cnt.shadow_queries / cnt.samples - 3.14) < 0.05


def test_oracle_rgb_config1_sha():
    """RENDER_MODE_RGB at BASELINE configs[0]'s size: sha256 of the reference's l-RGB+alpha buffer."""
    if not pu.have_assets():
        pytest.skip("texture not staged")
    idx = json.load(open(os.path.join(pu.GOLDEN, "golden_index.json")))["cornell-srgb_rgb_128x128_spp16_seed1"]
    flat = pu.load_flat("cornell-srgb", "rgb")
    opt = pu.options("rgb", 128, 128, 16, seed=1)
    acc, _, _ = pu.oracle_render(flat, opt)
    avg, _ = pu.oracle_resolve(flat, opt, acc)
    assert hashlib.sha256(avg.tobytes()).hexdigest() == idx["sha256"]


@pytest.mark.parametrize("scene,variant", [("cornell", "ours2006"), ("plane-srgb", "jh"), ("cornell-srgb", "meng")])
def test_oracle_config1_size_sha_of_the_other_baseline_variants(scene, variant):
    """The variants BASELINE.json configs[2..4] are quoted on (CIE 2006 observer, Jakob-Hanika, Meng et al.) at
    configs[0]'s size, 128x128 spp16: sha256 of the real reference's XYZA buffer."""
    if pu.needs_assets(scene, variant) and not pu.have_assets():
        pytest.skip("data files not staged")
    idx = json.load(open(os.path.join(pu.GOLDEN, "golden_index.json")))[f"{scene}_{variant}_128x128_spp16_seed1"]
    flat = pu.load_flat(scene, variant)
    opt = pu.options(variant, 128, 128, 16, seed=1)
    acc, _, _ = pu.oracle_render(flat, opt)
    xyza, _ = pu.oracle_resolve(flat, opt, acc)
    assert hashlib.sha256(xyza.tobytes()).hexdigest() == idx["sha256"]


def test_sample_subsets_compose():
    """Tiles and sample ranges accumulate to the same buffer as one full render (multi-GPU sharding)."""
    flat = pu.load_flat("cornell", "ours1931")
    full = pu.options("ours1931", 16, 12, 6, seed=3)
    acc_full, _, _ = pu.oracle_render(flat, full)
    acc = np.zeros_like(acc_full)
    import ctypes as C
    for (x0, x1) in ((0, 7), (7, 16)):
        for (s0, s1) in ((0, 2), (2, 6)):
            o = pu.options("ours1931", 16, 12, 6, seed=3, x0=x0, x1=x1, sample_begin=s0, sample_end=s1)
            rc = pu.oracle().ssb_oracle_render(C.byref(flat.scene), C.byref(flat.color), C.byref(o),
                                               acc.ctypes.data_as(C.POINTER(C.c_double)), None, None)
            assert rc == 0
    assert pu.bits_equal(acc, acc_full)
