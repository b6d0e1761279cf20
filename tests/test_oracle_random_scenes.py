"""CPU: the oracle pinned against the REAL reference on RANDOM scenes.

tests/test_oracle_golden.py pins the oracle (oracle/ssb_oracle.c) to builds of the real reference on the reference's own
three scenes: axis-aligned boxes under one horizontal rectangular light.  This file closes the gap the GPU render fuzz
(tests/test_gpu_render_fuzz.py) leans on: the hooked reference binaries (oracle/_ref/*_hooked, built from the unmodified
sources by oracle/build_ref.py) accept a quad list through SSB_SCENE_QUADS (oracle/ref_hooks.hpp: replace_scene — the
quads are built by the reference's own PrimQuad constructor), so random geometry — lights of any orientation and shape,
slivers, non-planar and degenerate quads, duplicates, coplanar neighbours, random materials — goes through the reference's
Scene::intersect, get_rand_toward_light / spherical-triangle sampling and BSDF code; the reference dumps the scene as it
built it (SSB_DUMP_TABLES) and its f64 XYZA accumulators (SSB_DUMP_XYZA), and the oracle, fed that dump, must reproduce
the accumulators BIT FOR BIT.

  * live runs (wherever oracle/_ref and the data files exist: this container, the GPU box),
  * committed fixtures of four such runs (tests/golden/random_*.bin / .npy, written by `python tests/test_oracle_random_scenes.py`),
    so that the pin also holds where the reference binaries are absent."""
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np
import pytest

import parity_util as pu
import refdump
from test_gpu_isect_fuzz import fuzz_quads

REFDIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")

#        scene            variant           kind        nquads  w   h  spp
CASES = [("cornell", "ours1931", "mixed", 24, 28, 20, 3),
         ("cornell", "ours1931", "axis", 12, 24, 24, 2),
         ("cornell", "ours2006", "sheared", 9, 20, 16, 4),
         ("cornell", "ours1931_noels", "mixed", 16, 24, 18, 3),     # no explicit light sampling: emission on every hit
         ("cornell", "rgb", "trapezoid", 14, 22, 16, 3),
         ("cornell-srgb", "ours1931", "mixed", 33, 26, 18, 2),      # textured material, > 32 quads
         ("cornell-srgb", "ours1931", "nonplanar", 10, 18, 22, 3),
         ("cornell-srgb", "jh", "mixed", 20, 20, 20, 2),
         ("cornell-srgb", "meng", "rect", 18, 24, 14, 3),
         ("cornell-srgb", "ours1931_nw3", "mixed", 12, 16, 16, 4),  # three hero wavelengths
         ("cornell-srgb", "ours1931_d3", "mixed", 40, 30, 16, 2),   # MAX_DEPTH 3
         ("plane-srgb", "ours1931", "mixed", 15, 20, 18, 3),
         ("plane-srgb", "ours1931_noels", "mixed", 15, 18, 18, 3),  # the plane scene's mirror material
         ("plane-srgb", "jh", "axis", 8, 16, 24, 2)]
FIXTURES = [0, 5, 8, 12]  # indices of CASES committed under tests/golden/
# a longer campaign against the real reference: SSB_REF_FUZZ_CASES=300 appends that many generated cases
_combos = [("cornell", "ours1931"), ("cornell-srgb", "ours1931"), ("cornell", "ours2006"), ("cornell-srgb", "jh"), ("cornell-srgb", "meng"),
           ("plane-srgb", "ours1931"), ("cornell", "ours1931_noels"), ("cornell-srgb", "rgb"), ("plane-srgb", "ours1931_noels"),
           ("cornell-srgb", "ours1931_nw3"), ("cornell-srgb", "meng_nw2"), ("cornell-srgb", "ours1931_d3"), ("cornell", "ours1931_noffc"), ("plane-srgb", "jh")]
_kinds = ["mixed", "rect", "axis", "mixed", "sheared", "trapezoid", "nonplanar"]
for _k in range(int(os.environ.get("SSB_REF_FUZZ_CASES", "0"))):
    _r = np.random.default_rng(555 + _k)
    CASES.append((*_combos[_k % len(_combos)], _kinds[_k % len(_kinds)], int(_r.choice([3, 6, 12, 20, 33, 48, 64])),
                  int(_r.integers(12, 32)), int(_r.integers(12, 28)), int(_r.integers(2, 5))))


def scene_file_bytes(flat, rng, nquads, kind):
    """Random quads in front of the scene's camera; returns the SSB_SCENE_QUADS payload."""
    sc = flat.scene
    light_mats = sorted({sc.quads[i].material for i in range(sc.nquads) if sc.quads[i].is_light})
    other_mats = [m for m in range(sc.nmaterials) if m not in light_mats]
    cam = sc.camera
    centre = np.array([cam.pos[k] + 1080.0 * cam.dir[k] for k in range(3)])
    corners = fuzz_quads(rng, nquads, 300.0, centre, kind)
    nlights = int(rng.integers(1, 4))
    light_ids = set(rng.choice(nquads, size=min(nlights, nquads), replace=False).tolist())
    st = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float32)
    out = [struct.pack("<I", nquads)]
    for qi in range(nquads):
        mat = int(rng.choice(light_mats if qi in light_ids else other_mats))
        rec = np.concatenate([np.concatenate([corners[qi, k].astype(np.float32), st[k] * np.float32(rng.uniform(-1.5, 3.0))]) for k in range(4)])
        out.append(struct.pack("<I", mat) + rec.astype("<f4").tobytes())
    return b"".join(out)


def run_reference(scene, variant, payload, w, h, spp, seed, tables_path):
    exe = os.path.join(REFDIR, f"simple_spectral_{variant}_hooked")
    with tempfile.TemporaryDirectory() as tmp:
        qf = os.path.join(tmp, "quads.bin")
        open(qf, "wb").write(payload)
        env = dict(os.environ, SSB_SEED=str(seed), SSB_DUMP_XYZA=os.path.join(tmp, "xyza.bin"), SSB_DUMP_TABLES=tables_path,
                   SSB_SCENE_QUADS=qf, SSB_THREADS="4")
        r = subprocess.run([exe, f"--scene={scene}", f"-w={w}", f"-h={h}", f"-spp={spp}", f"--output={tmp}/o.pfm"], cwd=pu.data_root(),
                           env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, (r.stdout[-400:], r.stderr[-400:])
        return np.fromfile(os.path.join(tmp, "xyza.bin"), dtype=np.float64).reshape(h, w, 4)


def flat_of(tables_path, scene, variant):
    t = refdump.parse(tables_path)
    tex = pu.lizard_texture() if scene != "cornell" else None
    jh = pu.jh_tables() if variant == "jh" else None
    meng = pu.meng_tables() if variant.startswith("meng") else None
    return refdump.flat_from_dump(t, tex, jh=jh, meng=meng)


def oracle_xyza(flat, variant, w, h, spp, seed):
    opt = pu.options(variant, w, h, spp, seed=seed)
    acc, _, _ = pu.oracle_render(flat, opt)
    xyza, _ = pu.oracle_resolve(flat, opt, acc)
    return xyza


def _payload(i):
    scene, variant, kind, nquads, w, h, spp = CASES[i]
    rng = np.random.default_rng(9100 + i)
    return scene_file_bytes(pu.load_flat(scene, variant), rng, nquads, kind)


@pytest.mark.parametrize("i", range(len(CASES)))
def test_oracle_equals_the_real_reference_on_a_random_scene_live(i, tmp_path):
    scene, variant, kind, nquads, w, h, spp = CASES[i]
    if not pu.have_assets():
        pytest.skip("data files not staged (assets/data)")
    if not os.path.exists(os.path.join(REFDIR, f"simple_spectral_{variant}_hooked")):
        pytest.skip("oracle/_ref not built (run __graft_entry__.build() where /root/reference exists)")
    tables = str(tmp_path / "tables.bin")
    ref = run_reference(scene, variant, _payload(i), w, h, spp, 300 + i, tables)
    flat = flat_of(tables, scene, variant)
    assert flat.scene.nquads == nquads
    got = oracle_xyza(flat, variant, w, h, spp, 300 + i)
    assert np.nansum(np.abs(ref[..., 3])) > 0, "nothing visible: the case proves nothing"
    assert pu.bits_equal(got, ref), f"{int((got.view(np.uint64) != ref.view(np.uint64)).sum())} words differ, max rel {pu.rel_err(got, ref).max()}"


@pytest.mark.parametrize("i", FIXTURES)
def test_oracle_equals_the_committed_random_scene_fixture(i):
    scene, variant, kind, nquads, w, h, spp = CASES[i]
    if pu.needs_assets(scene, variant) and not pu.have_assets():
        pytest.skip("data files not staged (assets/data)")
    tables = os.path.join(pu.GOLDEN, f"random_{i}_{scene}_{variant}_tables.bin")
    ref = np.load(os.path.join(pu.GOLDEN, f"random_{i}_{scene}_{variant}_xyza_{w}x{h}_spp{spp}_seed{300 + i}.npy"))
    got = oracle_xyza(flat_of(tables, scene, variant), variant, w, h, spp, 300 + i)
    assert pu.bits_equal(got, ref)


if __name__ == "__main__":  # regenerate the committed fixtures from the real reference
    for i in FIXTURES:
        scene, variant, kind, nquads, w, h, spp = CASES[i]
        tables = os.path.join(pu.GOLDEN, f"random_{i}_{scene}_{variant}_tables.bin")
        ref = run_reference(scene, variant, _payload(i), w, h, spp, 300 + i, tables)
        np.save(os.path.join(pu.GOLDEN, f"random_{i}_{scene}_{variant}_xyza_{w}x{h}_spp{spp}_seed{300 + i}.npy"), ref)
        print("wrote fixture", i, scene, variant, os.path.getsize(tables), "bytes of tables")
    sys.exit(0)
