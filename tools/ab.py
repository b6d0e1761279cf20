#!/usr/bin/env python3
"""A/B of PREBUILT library variants on the GPU box (no compile time on the box).

  here :  tools/ab.py build name[=KNOB=V,KNOB=V...] ...     -> variants/libssb200_<name>.so   (name "base@<rev>" builds the
                                                               sources of that git revision instead of the work tree)
  box  :  tools/ab.py run [name ...]                         -> per variant: ms/frame of the headline frame (CUDA events,
                                                               median of 7 after 3 warm frames), Msamples/s, and the sha256 of
                                                               the f64 XYZA accumulators — every variant must print the SAME hash
  box  :  tools/ab.py test name [pytest args ...]            -> the GPU test suite (default: tests -m gpu -x -q) with that variant
                                                               in place of the library
The variant is copied over simple-spectral_b200/libssb200.so for its run (each run is its own process); the original is
restored at the end."""
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

LIB = os.path.join(ROOT, "simple-spectral_b200", "libssb200.so")
VDIR = os.path.join(ROOT, "variants")
CHILD = r'''
import sys, importlib, hashlib, ctypes as C
import numpy as np
sys.path.insert(0, %r)
ssb = importlib.import_module("simple-spectral_b200"); host = importlib.import_module("simple-spectral_b200.host")
import os
# the headline frame, or (environment) SSB_AB="scene,variant,width,height,spp" e.g. "cornell-srgb,meng,1024,1024,16"
sc, var, W, H, SPP = (os.environ.get("SSB_AB") or "cornell-srgb,ours1931,512,512,64").split(",")
W, H, SPP = int(W), int(H), int(SPP)
color = host.Color(None, *host.VARIANTS[var]); scene = host.Scene(sc, color)
ctx = ssb.Context(0); ctx.upload_color(color.flat); ctx.upload_scene(scene.flat)
opt = host.options_for(color, W, H, SPP, seed=1, **({'render_mode': ssb.SSB_RENDER_RGB} if var == 'rgb' else {}))
for _ in range(3): ctx.render(opt)
ctx.synchronize()
ms, tr = [], []
for _ in range(7):
    ctx.render(opt); st = ctx.stats(); ms.append(st.device_ms); tr.append(st.trace_ms)
ms.sort(); tr.sort()
acc = ctx.read_accum(W, H)
print("SAMPLES", W * H * SPP)
print("RESULT", ms[len(ms)//2], tr[len(tr)//2], hashlib.sha256(np.ascontiguousarray(acc).tobytes()).hexdigest()[:16])
''' % ROOT


def build(specs):
    os.makedirs(VDIR, exist_ok=True)
    for spec in specs:
        name, _, knobs = spec.partition("=")
        defs = [f"-D{kv}" for kv in knobs.split(",") if kv]
        root = ROOT
        tmp = None
        if "@" in name:
            name, rev = name.split("@")
            tmp = tempfile.mkdtemp()
            subprocess.run(f"git -C {ROOT} archive {rev} simple-spectral_b200/csrc include | tar -x -C {tmp}", shell=True, check=True)
            root = tmp
        src = os.path.join(root, "simple-spectral_b200", "csrc")
        host_cpp = [os.path.join(src, "host", f) for f in sorted(os.listdir(os.path.join(src, "host"))) if f.endswith(".cpp")]
        out = os.path.join(VDIR, f"libssb200_{name}.so")
        t0 = time.time()
        r = subprocess.run(["nvcc", *g.NVCC_FLAGS, *defs, "-o", out, os.path.join(src, "ssb_capi.cu"), *host_cpp, "-lz"], capture_output=True, text=True)
        print(f"{name:24s} {'ok' if r.returncode == 0 else 'FAILED ' + r.stderr[-400:]}  ({time.time() - t0:.0f}s)  {' '.join(defs)}", flush=True)
        if tmp:
            shutil.rmtree(tmp)


def run(names):
    if not names:
        names = sorted(f[len("libssb200_"):-3] for f in os.listdir(VDIR) if f.startswith("libssb200_") and f.endswith(".so"))
    backup = LIB + ".orig"
    shutil.copyfile(LIB, backup)
    try:
        for rep in range(1):
            for n in names:  # (list a name twice to see the drift over the call)
                shutil.copyfile(os.path.join(VDIR, f"libssb200_{n}.so"), LIB)
                out = subprocess.run([sys.executable, "-c", CHILD], capture_output=True, text=True)
                res = [ln for ln in out.stdout.split("\n") if ln.startswith("RESULT")]
                if not res:
                    print(f"{n:24s} RUN FAILED: {out.stderr[-400:]}", flush=True)
                    continue
                ms, tr, sha = res[0].split()[1:4]
                ns = int([ln for ln in out.stdout.split("\n") if ln.startswith("SAMPLES")][0].split()[1])
                print(f"{n:24s} frame {float(ms):8.3f} ms  bounce {float(tr):8.3f} ms  {ns / float(ms) / 1e3:8.1f} Msamples/s  xyza sha256 {sha}", flush=True)
    finally:
        shutil.copyfile(backup, LIB)
        os.remove(backup)


def test(args):
    name, pytest_args = args[0], (args[1:] or ["tests", "-m", "gpu", "-x", "-q"])
    backup = LIB + ".orig"
    shutil.copyfile(LIB, backup)
    try:
        shutil.copyfile(os.path.join(VDIR, f"libssb200_{name}.so"), LIB)
        rc = subprocess.run([sys.executable, "-m", "pytest", *pytest_args], cwd=ROOT).returncode
    finally:
        shutil.copyfile(backup, LIB)
        os.remove(backup)
    sys.exit(rc)


if __name__ == "__main__":
    if len(sys.argv) < 2 or sys.argv[1] not in ("build", "run", "test") or (sys.argv[1] == "test" and len(sys.argv) < 3):
        sys.exit(__doc__)
    {"build": build, "run": run, "test": test}[sys.argv[1]](sys.argv[2:])
