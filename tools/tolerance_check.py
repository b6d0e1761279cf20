#!/usr/bin/env python3
"""Per-pixel difference of the headline frame between the (bit-exact) library and a "tolerance" build of it (prebuilt
variants/libssb200_<name>.so, e.g. compiled with -fmad=true): speed of each and the distribution of the relative XYZ
error per pixel (the north star's tolerance is 1e-4).  On the GPU box:  tools/tolerance_check.py <exact> <variant> ..."""
import json
import os
import shutil
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "simple-spectral_b200", "libssb200.so")
CHILD = r'''
import sys, importlib, numpy as np
sys.path.insert(0, %r)
ssb = importlib.import_module("simple-spectral_b200"); host = importlib.import_module("simple-spectral_b200.host")
color = host.Color(None, 1931, ssb.SSB_UPSAMPLE_OURS); scene = host.Scene("cornell-srgb", color)
ctx = ssb.Context(0); ctx.upload_color(color.flat); ctx.upload_scene(scene.flat)
opt = host.options_for(color, 512, 512, 64, seed=1)
for _ in range(3): ctx.render(opt)
ms = []
for _ in range(7):
    ctx.render(opt); ms.append(ctx.stats().device_ms)
ms.sort()
x, _ = ctx.resolve(opt)
np.save(sys.argv[1], x)
print("MS", ms[len(ms)//2])
''' % ROOT

names = sys.argv[1:]
backup = LIB + ".orig"
shutil.copyfile(LIB, backup)
frames, ms = {}, {}
try:
    for n in names:
        shutil.copyfile(os.path.join(ROOT, "variants", f"libssb200_{n}.so"), LIB)
        out = f"/tmp/tol_{n}.npy"
        r = subprocess.run([sys.executable, "-c", CHILD, out], capture_output=True, text=True)
        if r.returncode != 0:
            print(n, "FAILED", r.stderr[-300:]); continue
        ms[n] = float([ln for ln in r.stdout.split("\n") if ln.startswith("MS")][0].split()[1])
        frames[n] = np.load(out)
finally:
    shutil.copyfile(backup, LIB); os.remove(backup)
ref = frames[names[0]]
for n in names:
    x = frames.get(n)
    if x is None:
        continue
    d = np.abs(x[..., :3] - ref[..., :3]).max(axis=-1) / np.maximum(np.abs(ref[..., :3]).max(axis=-1), 1e-300)
    d = d[ref[..., 3] > 0]
    q = {f"p{p}": float(np.nanpercentile(d, p)) for p in (50, 90, 99, 99.9, 100)}
    print(json.dumps({"variant": n, "ms_per_frame": ms[n], "Msamples_per_s": 512 * 512 * 64 / ms[n] / 1e3,
                      "pixels_over_1e-4": float(np.mean(d > 1e-4)), "pixels_over_1e-3": float(np.mean(d > 1e-3)), "nan_pixels": int(np.isnan(d).sum()), **q}))
