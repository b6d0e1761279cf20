#!/usr/bin/env python3
"""On-box tuning sweep: rebuild libssb200.so with different -D knobs and time the headline frame.
usage: tools/tune.py "SSB_BOUNCE_THREADS=128,SSB_BOUNCE_MIN_BLOCKS=5" "SSB_BOUNCE_THREADS=256,SSB_BOUNCE_MIN_BLOCKS=2" ...
Prints one line per variant: knobs, ms/frame (CUDA events), Msamples/s.  The original library is restored at the end."""
import importlib
import os
import shutil
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

LIB = os.path.join(ROOT, "simple-spectral_b200", "libssb200.so")
CHILD = r'''
import sys, importlib, torch
sys.path.insert(0, %r)
ssb = importlib.import_module("simple-spectral_b200"); host = importlib.import_module("simple-spectral_b200.host")
color = host.Color(None, 1931, ssb.SSB_UPSAMPLE_OURS); scene = host.Scene("cornell-srgb", color)
ctx = ssb.Context(0); ctx.upload_color(color.flat); ctx.upload_scene(scene.flat)
opt = host.options_for(color, 512, 512, 64, seed=1)
for _ in range(3): ctx.render(opt)
ctx.synchronize()
ms = []
for _ in range(5):
    ctx.render(opt); st = ctx.stats(); ms.append(st.device_ms)
ms.sort(); print("RESULT", ms[len(ms)//2], st.trace_ms)
''' % ROOT


def main():
    variants = sys.argv[1:] or [""]
    backup = LIB + ".orig"
    shutil.copyfile(LIB, backup)
    try:
        for v in variants:
            defs = [f"-D{kv}" for kv in v.split(",") if kv]
            src_dir = os.path.join(ROOT, "simple-spectral_b200", "csrc")
            host_cpp = [os.path.join(src_dir, "host", f) for f in sorted(os.listdir(os.path.join(src_dir, "host"))) if f.endswith(".cpp")]
            t0 = time.time()
            r = subprocess.run(["nvcc", *g.NVCC_FLAGS, *defs, "-o", LIB, os.path.join(src_dir, "ssb_capi.cu"), *host_cpp, "-lz"], capture_output=True, text=True)
            if r.returncode != 0:
                print(f"{v or 'default':60s} BUILD FAILED: {r.stderr[-300:]}")
                continue
            out = subprocess.run([sys.executable, "-c", CHILD], capture_output=True, text=True)
            res = [ln for ln in out.stdout.split("\n") if ln.startswith("RESULT")]
            if not res:
                print(f"{v or 'default':60s} RUN FAILED: {out.stderr[-300:]}")
                continue
            ms, tr = map(float, res[0].split()[1:3])
            print(f"{v or chr(100)+"efault":60s} frame {ms:8.3f} ms  bounce {tr:8.3f} ms  {512*512*64/ms/1e3:8.1f} Msamples/s  (build {time.time()-t0:.0f}s)", flush=True)
    finally:
        shutil.copyfile(backup, LIB)
        os.remove(backup)


if __name__ == "__main__":
    main()
