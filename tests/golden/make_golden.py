#!/usr/bin/env python3
"""Generate the committed golden fixtures from the REAL reference (oracle/_ref/*_hooked, built by
oracle/build_ref.py from /root/reference).  Run here (the reference cannot travel); commit the output.

  tables_<scene>_<variant>.bin         Color::data + camera + flattened scene as the reference built them
                                       (ref_hooks.hpp dump_tables) — inputs for oracle and CUDA path
  xyza_<scene>_<variant>_<W>x<H>_spp<N>_seed<S>.npy
                                       per-pixel double XYZA of the reference at per-sample seeding
  golden_index.json                    sha256 of larger reference renders (BASELINE config 1)
  roundtrip_running_max.json           the reference's OWN round-trip self-test (main.cpp:184-264, `#if 0` upstream,
                                       enabled by `oracle/build_ref.py roundtrip`): running maximum of the
                                       sRGB -> spectrum -> sRGB error after each of the 256 red levels; its last value
                                       is the one number the reference documents (1.851469e-5, main.cpp:242-245).
                                       Takes ~8 minutes: only with `make_golden.py roundtrip`
  refout_cornell_ours1931_32x24_spp4_seed7.{pfm,hdr,csv,png}
                                       the image FILES the reference itself writes (Framebuffer::save,
                                       framebuffer.cpp:39-176) for the small case, one per supported extension
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
REFBIN = os.path.join(ROOT, "oracle", "_ref")
DATA_ROOT = os.environ.get("SSB_REFERENCE_ROOT", "/root/reference")

CASES = [  # (scene, variant)
    ("cornell", "ours1931"), ("cornell-srgb", "ours1931"), ("plane-srgb", "ours1931"),
    ("cornell", "ours2006"), ("cornell-srgb", "ours2006"),
    ("cornell-srgb", "jh"), ("plane-srgb", "jh"),
    ("cornell-srgb", "meng"), ("plane-srgb", "meng"),
    # EXPLICIT_LIGHT_SAMPLING compiled out: MaterialMirror on the plane (scene.cpp:346-355), emission on every hit
    ("plane-srgb", "ours1931_noels"), ("cornell", "ours1931_noels"),
    # FLAT_FIELD_CORRECTION compiled out (renderer.cpp:262-266; see oracle/build_ref.py on how that build is made to compile)
    ("cornell", "ours1931_noffc"), ("cornell-srgb", "ours1931_noffc"),
    # RENDER_MODE_RGB: the three-channel comparison renderer (the stored "xyza" is the l-RGB+alpha average)
    ("cornell", "rgb"), ("cornell-srgb", "rgb"), ("plane-srgb", "rgb"),
    # SAMPLE_WAVELENGTHS 3 (OURS) and 2 (Meng)
    ("cornell-srgb", "ours1931_nw3"), ("plane-srgb", "ours1931_nw3"), ("cornell-srgb", "meng_nw2"),
    ("cornell-srgb", "ours1931_d3"),  # MAX_DEPTH 3
]
SMALL = dict(w=32, h=24, spp=4, seed=7)
C1 = dict(w=128, h=128, spp=16, seed=1)  # BASELINE.json configs[0]


def run_ref(scene, variant, w, h, spp, seed, tables=None, indirect_only=False):
    exe = os.path.join(REFBIN, f"simple_spectral_{variant}_hooked")
    with tempfile.TemporaryDirectory() as tmp:
        env = dict(os.environ, SSB_SEED=str(seed), SSB_DUMP_XYZA=os.path.join(tmp, "xyza.bin"))
        if tables:
            env["SSB_DUMP_TABLES"] = tables
        cmd = [exe, f"--scene={scene}", f"-w={w}", f"-h={h}", f"-spp={spp}", f"--output={tmp}/o.pfm"]
        if indirect_only:
            cmd.append("--indirect-only")
        subprocess.run(cmd, cwd=DATA_ROOT, env=env, check=True, stdout=subprocess.DEVNULL)
        return np.fromfile(os.path.join(tmp, "xyza.bin"), dtype=np.float64).reshape(h, w, 4)


def run_ref_outputs():
    """The reference's own output files (--output=<path>, format by extension) at per-sample seeding."""
    exe = os.path.join(REFBIN, "simple_spectral_ours1931_hooked")
    for ext in ("pfm", "hdr", "csv", "png"):
        out = os.path.join(HERE, f"refout_cornell_ours1931_{SMALL['w']}x{SMALL['h']}_spp{SMALL['spp']}_seed{SMALL['seed']}.{ext}")
        subprocess.run([exe, "--scene=cornell", f"-w={SMALL['w']}", f"-h={SMALL['h']}", f"-spp={SMALL['spp']}", f"--output={out}"],
                       cwd=DATA_ROOT, env=dict(os.environ, SSB_SEED=str(SMALL["seed"])), check=True, stdout=subprocess.DEVNULL)
        print("wrote", os.path.basename(out), os.path.getsize(out), "bytes")


def run_ref_roundtrip():
    exe = os.path.join(REFBIN, "simple_spectral_roundtrip")
    if not os.path.exists(exe):
        subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "build_ref.py"), "roundtrip"], check=True)
    if os.environ.get("SSB_ROUNDTRIP_LOG"):  # stdout of an earlier run of that binary
        out = open(os.environ["SSB_ROUNDTRIP_LOG"]).read()
    else:
        with tempfile.TemporaryDirectory() as tmp:
            out = subprocess.run([exe, "--scene=cornell", "-w=4", "-h=4", "-spp=1", f"--output={tmp}/o.pfm"], cwd=DATA_ROOT,
                                 check=True, capture_output=True, text=True).stdout
    vals = {}
    for line in out.split("\n"):
        if line.startswith("RUNNING "):
            _, r, v = line.split()
            vals[int(r)] = v
    assert sorted(vals) == list(range(256)), "incomplete round-trip log"
    json.dump({"source": "reference src/main.cpp:246-262 (round-trip self-test), CIE 1931, OURS, g++ -O2 -ffp-contract=off",
               "running_max": [vals[r] for r in range(256)]},
              open(os.path.join(HERE, "roundtrip_running_max.json"), "w"), indent=0)
    print("wrote roundtrip_running_max.json, final", vals[255])


def main():
    if sys.argv[1:] == ["roundtrip"]:
        return run_ref_roundtrip()
    index = {}
    only = sys.argv[1:]  # optional: regenerate only these variants (the others are left untouched)
    if only and os.path.exists(os.path.join(HERE, "golden_index.json")):
        index = json.load(open(os.path.join(HERE, "golden_index.json")))
    for scene, variant in CASES:
        if only and variant not in only:
            continue
        tables = os.path.join(HERE, f"tables_{scene}_{variant}.bin")
        x = run_ref(scene, variant, tables=tables, **SMALL)
        name = f"xyza_{scene}_{variant}_{SMALL['w']}x{SMALL['h']}_spp{SMALL['spp']}_seed{SMALL['seed']}.npy"
        np.save(os.path.join(HERE, name), x)
        print("wrote", name, "mean", x.mean(axis=(0, 1)))
    if not only or "ours1931" in only:
        x = run_ref("cornell-srgb", "ours1931", indirect_only=True, **SMALL)
        np.save(os.path.join(HERE, f"xyza_cornell-srgb_ours1931_indirect_{SMALL['w']}x{SMALL['h']}_spp{SMALL['spp']}_seed{SMALL['seed']}.npy"), x)
    # BASELINE.json configs[0]'s size for its own variant, the RGB build, and the variants of configs[2..4]
    for scene, variant in (("cornell-srgb", "ours1931"), ("cornell", "ours1931"), ("cornell-srgb", "rgb"),
                           ("cornell", "ours2006"), ("plane-srgb", "jh"), ("cornell-srgb", "meng")):
        if only and variant not in only:
            continue
        x = run_ref(scene, variant, **C1)
        key = f"{scene}_{variant}_{C1['w']}x{C1['h']}_spp{C1['spp']}_seed{C1['seed']}"
        index[key] = dict(sha256=hashlib.sha256(x.tobytes()).hexdigest(), mean=list(x.mean(axis=(0, 1))),
                          pixel_64_64=list(x[64, 64]))
        print(key, index[key]["sha256"])
    json.dump(index, open(os.path.join(HERE, "golden_index.json"), "w"), indent=1)
    if not only or "ours1931" in only or "refout" in only:
        run_ref_outputs()


if __name__ == "__main__":
    sys.exit(main())
