#!/usr/bin/env python3
"""On the GPU box: where does the paired acosf (ssb_debug_eval_math fn 6 / 7) differ from the host libm?
Prints mismatch counts and the first offenders for fn 2 (scalar), 6 and 7 (either lane of the pair), the result and
duration of the exhaustive device-side comparison (fn 8), and whether the position of an argument in the batch matters."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_util as pu  # noqa: E402

rng = np.random.default_rng(77)
specials = np.array([0.0, -0.0, 1.0, -1.0, 0.5, -0.5, 1e-20, -1e-20, 0.49999997, 0.50000006])
x = np.concatenate([rng.uniform(-1.0, 1.0, 1 << 18), specials]).astype(np.float32)
want = np.empty_like(x)
pu.oracle().ssb_oracle_eval_math(2, x.ctypes.data_as(C.POINTER(C.c_float)), 0.0, want.ctypes.data_as(C.POINTER(C.c_float)), x.size)
flat = pu.load_flat("cornell", "ours1931")
with pu.gpu_context(flat) as ctx:
    for fn in (2, 6, 7):
        got = ctx.eval_math(fn, x)
        bad = np.nonzero((got.view(np.uint32) != want.view(np.uint32)) & ~(np.isnan(got) & np.isnan(want)))[0]
        print(f"fn {fn}: {bad.size} of {x.size} differ from host acosf", flush=True)
        for i in bad[:12]:
            p = np.uint32((int(x.view(np.uint32)[i]) * 2654435761) & 0xffffffff)
            print(f"   i={i} x={x[i]!r} ({x.view(np.uint32)[i]:#010x}) partner={p:#010x} got={got[i]!r} ({got.view(np.uint32)[i]:#010x}) want={want[i]!r}")
    # the specials first, then alone
    for name, xs in (("specials first", np.concatenate([x[-10:], x[:-10]])), ("specials alone", x[-10:].copy()), ("one value x32", np.full(32, 0.49999997, np.float32))):
        w = np.empty_like(xs)
        pu.oracle().ssb_oracle_eval_math(2, xs.ctypes.data_as(C.POINTER(C.c_float)), 0.0, w.ctypes.data_as(C.POINTER(C.c_float)), xs.size)
        for fn in (6, 7):
            g = ctx.eval_math(fn, xs)
            bad = np.nonzero((g.view(np.uint32) != w.view(np.uint32)) & ~(np.isnan(g) & np.isnan(w)))[0]
            print(f"{name}, fn {fn}: {bad.size} differ; first: {[(int(i), float(xs[i]), float(g[i]), float(w[i])) for i in bad[:4]]}", flush=True)
    t0 = time.time()
    r = ctx.eval_math(8, np.zeros(1 + 3 * 8, np.float32))
    print(f"fn 8 (exhaustive, on the device): {r[0]:.0f} lanes differ, {time.time() - t0:.3f} s; examples {r[1:].view(np.uint32)[:9]}", flush=True)
