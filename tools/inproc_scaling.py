#!/usr/bin/env python3
"""In-process multi-GPU rendering through the library's own Renderer (RendererOptions::devices; one context + host thread
per GPU, ssb_accum_merge over NVLink, resolve on the first device) — what `simple_spectral_b200 --devices=...` runs.
Prints one JSON line per device count: wall time of render() (uploads excluded: the Renderer is constructed before),
the library's own device time (max over GPUs per slice), Msamples/s, and the sha256 of the f64 XYZA frame.
usage: tools/inproc_scaling.py [--scene S --variant V -w W -h H --spp N --shard tiles|samples --counts 1,2,4,8 --reps 5]"""
import argparse
import hashlib
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

ssb = importlib.import_module("simple-spectral_b200")
host = importlib.import_module("simple-spectral_b200.host")

ap = argparse.ArgumentParser(add_help=False)
ap.add_argument("--scene", default="cornell-srgb"); ap.add_argument("--variant", default="ours1931")
ap.add_argument("-w", type=int, default=512); ap.add_argument("-h", type=int, default=512); ap.add_argument("--spp", type=int, default=64)
ap.add_argument("--shard", default="tiles"); ap.add_argument("--counts", default=None); ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
n_dev = ssb.device_count()
counts = [int(x) for x in a.counts.split(",")] if a.counts else [c for c in (1, 2, 4, 8) if c <= n_dev]
base = None
for n in counts:
    devs = list(range(n)) if n <= n_dev else [i % n_dev for i in range(n)]
    r = host.Renderer(a.scene, a.w, a.h, a.spp, variant=a.variant, seed=1, devices=devs if n > 1 else None, shard=a.shard)
    r.render()  # warm-up: allocations, first launches
    wall, dev = [], []
    for _ in range(a.reps):
        t0 = time.perf_counter()
        xyza, _ = r.render()
        wall.append((time.perf_counter() - t0) * 1e3)
        dev.append(r.stats().device_ms)
    r.close()
    sha = hashlib.sha256(np.ascontiguousarray(xyza).tobytes()).hexdigest()[:16]
    wall.sort(); dev.sort()
    w, d = wall[len(wall) // 2], dev[len(dev) // 2]
    ns = a.w * a.h * a.spp
    base = base or w
    print(json.dumps({"devices": devs, "shard": a.shard, "job": f"{a.scene} {a.w}x{a.h} spp{a.spp} {a.variant}", "wall_ms": w, "device_ms_max": d,
                      "Msamples_per_s_wall": ns / w / 1e3, "Msamples_per_s_device": ns / d / 1e3, "speedup_wall": base / w, "xyza_sha": sha}), flush=True)
