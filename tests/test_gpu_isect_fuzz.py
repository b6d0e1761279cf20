"""GPU: the conservative filter of the device's Scene::intersect (simple-spectral_b200/csrc/ssb_isect.cuh: packed-fp32
plane / rectangle / diagonal filter with rcp.approx, nearest-candidate-first exact tests, in-plane pre-reject) fuzzed ON
THE DEVICE against the checker's plain list scan (oracle/ssb_oracle.c scene_intersect; reference scene.cpp:433-445,
geometry.cpp:12-139): random scenes of 1-256 quads uploaded through the C ABI — planar, sheared, non-planar, degenerate,
slivers, duplicated and coplanar-overlapping quads, coordinates from 1e-3 to 1e6, far from the origin — and about 7 M
rays in total (random, surface-to-surface with the `ignore` quad, aimed at vertices / edges / diagonals, axis-aligned, lying
in a quad's plane).  Every hit record (quad, triangle, distance bits, barycentric bits) must be identical, for the filtered
scan AND for the library's own list-scan switch (ssb_options.scan_mode = SSB_SCAN_LIST).  tests/test_isect_host.py
checks the same logic compiled for the host, where rcp.approx and FFMA2 are emulated; this one runs the real instructions."""
import ctypes as C
import zlib

import numpy as np
import pytest

import parity_util as pu

abi = pu.abi
pytestmark = pytest.mark.gpu


def _scene(quads_v):
    """quads_v: (n, 4, 3) float32 corners v00, v10, v11, v01 -> ssb_scene with one Lambertian material (geometry.hpp:93-95)."""
    n = quads_v.shape[0]
    quads = (abi.ssb_quad * n)()
    with np.errstate(invalid="ignore", divide="ignore"):
        for qi in range(n):
            v00, v10, v11, v01 = (quads_v[qi, k].astype(np.float32) for k in range(4))
            for ti, tri in enumerate(((v00, v10, v11), (v00, v11, v01))):
                for vi in range(3):
                    for k in range(3):
                        quads[qi].tri[ti].v[vi].pos[k] = float(tri[vi][k])
                nrm = np.cross((tri[1] - tri[0]).astype(np.float32), (tri[2] - tri[0]).astype(np.float32)).astype(np.float32)
                ln = np.float32(np.sqrt(np.float32(np.dot(nrm, nrm))))
                nrm = nrm / ln if ln > 0 else np.zeros(3, np.float32)
                for k in range(3):
                    quads[qi].tri[ti].normal[k] = float(nrm[k])
            quads[qi].material = 0
            quads[qi].is_light = 1 if qi == 0 else 0
    spec = np.array([0.5, 0.5], np.float32)
    mats = (abi.ssb_material * 1)()
    for s in (mats[0].albedo, mats[0].emission):
        s.data = spec.ctypes.data_as(C.POINTER(C.c_float)); s.n = 2; s.low, s.high = 380.0, 780.0
    sc = abi.ssb_scene()
    sc.quads, sc.nquads, sc.materials, sc.nmaterials = quads, n, mats, 1
    sc._keep = (quads, mats, spec)
    return sc


def _rand_quads(rng, n, scale, offset, kind):
    q = np.empty((n, 4, 3), np.float64)
    for i in range(n):
        c = rng.uniform(-1, 1, 3)
        a = rng.normal(size=3); a /= np.linalg.norm(a)
        b = np.cross(a, rng.normal(size=3)); b /= np.linalg.norm(b)
        ea, eb = rng.uniform(0.05, 0.8), rng.uniform(0.05, 0.8)
        k = kind if kind != "mixed" else rng.choice(["rect", "sheared", "trapezoid", "nonplanar", "degenerate", "sliver", "axis"])
        if k == "axis":  # axis-aligned boxes' faces, like the reference's scenes
            ax = rng.integers(3); a = np.eye(3)[(ax + 1) % 3]; b = np.eye(3)[(ax + 2) % 3]
        if k == "sliver":
            eb = ea * 10.0 ** rng.uniform(-6, -3)
        v = [c - ea * a - eb * b, c + ea * a - eb * b, c + ea * a + eb * b, c - ea * a + eb * b]
        if k == "sheared":
            sh = rng.uniform(-0.5, 0.5) * ea; v[2] = v[2] + sh * a; v[3] = v[3] + sh * a
        if k == "trapezoid":
            t = rng.uniform(0.1, 0.9); v[2] = c + t * ea * a + eb * b; v[3] = c - t * ea * a + eb * b
        if k == "nonplanar":
            v[rng.integers(4)] += np.cross(a, b) * rng.uniform(-0.2, 0.2) * ea
        if k == "degenerate":
            m = rng.integers(3)
            if m == 0: v[1] = v[0].copy()            # tri0 collapses
            elif m == 1: v[3] = v[2].copy()          # tri1 collapses
            else: v = [v[0], v[0].copy(), v[0].copy(), v[0].copy()]  # a point
        q[i] = np.array(v)
    return (q * scale + offset).astype(np.float32)


def fuzz_quads(rng, nq, scale, offset, kind):
    qv = _rand_quads(rng, nq, scale, offset, kind)
    if kind != "degenerate" and nq >= 8:
        # ties: a duplicated quad, a coplanar overlapping one, a coplanar neighbour sharing an edge (the ceiling pieces of the Cornell box)
        qv[nq - 1] = qv[0]
        qv[nq - 2] = qv[1] + (qv[1][1] - qv[1][0]) * np.float32(0.5)
        qv[nq - 3] = qv[2] + (qv[2][1] - qv[2][0])
    return qv


def _rays(rng, quads_v, n, in_plane=False):
    """A mix of ray families (in_plane: mostly family 4); returns rays (n,6) float32 and ignore (n,) int32."""
    lo, hi = quads_v.reshape(-1, 3).min(0).astype(np.float64), quads_v.reshape(-1, 3).max(0).astype(np.float64)
    ext = np.maximum(hi - lo, 1e-30)
    nq = quads_v.shape[0]
    o = np.empty((n, 3)); d = np.empty((n, 3)); ign = np.full(n, -1, np.int32)
    fam = rng.integers(0, 6, n)
    if in_plane:
        fam = np.where(rng.uniform(size=n) < 0.85, 4, fam)
    qi = rng.integers(0, nq, n)
    s, t = rng.uniform(0, 1, n), rng.uniform(0, 1, n)
    special = np.array([0.0, 1.0, 0.5, 1e-7, 1 - 1e-7, 1e-4, 1 - 1e-4, 0.25])
    pick = rng.uniform(size=n) < 0.4
    s = np.where(pick, special[rng.integers(0, 8, n)], s); t = np.where(rng.uniform(size=n) < 0.4, special[rng.integers(0, 8, n)], t)
    diag = rng.uniform(size=n) < 0.15
    t = np.where(diag, s, t)  # on the shared diagonal v00-v11
    Q = quads_v[qi].astype(np.float64)
    target = (Q[:, 0] * ((1 - s) * (1 - t))[:, None] + Q[:, 1] * (s * (1 - t))[:, None] + Q[:, 2] * (s * t)[:, None] + Q[:, 3] * ((1 - s) * t)[:, None])
    rnd = rng.normal(size=(n, 3)); rnd /= np.linalg.norm(rnd, axis=1)[:, None]
    # 0: random origin in the (enlarged) box, random direction
    o[:] = lo + (rng.uniform(-0.3, 1.3, (n, 3))) * ext; d[:] = rnd
    # 1: random origin aimed at a point of a quad (corners / edges / diagonal included)
    m = fam == 1; d[m] = target[m] - o[m]
    # 2: from a point on one quad to a point on another, ignoring the first (what path and shadow rays do)
    m = fam == 2
    q2 = rng.integers(0, nq, n); Q2 = quads_v[q2].astype(np.float64); s2, t2 = rng.uniform(0, 1, n), rng.uniform(0, 1, n)
    src = (Q2[:, 0] * ((1 - s2) * (1 - t2))[:, None] + Q2[:, 1] * (s2 * (1 - t2))[:, None] + Q2[:, 2] * (s2 * t2)[:, None] + Q2[:, 3] * ((1 - s2) * t2)[:, None])
    o[m] = src[m]; d[m] = target[m] - src[m]; ign[m] = q2[m]
    # 3: axis-aligned directions (zero components: the division by zero / inf paths of the watertight set-up)
    m = fam == 3; ax = np.eye(3)[rng.integers(0, 3, n)] * rng.choice([-1.0, 1.0], n)[:, None]; d[m] = ax[m]
    m2 = m & (rng.uniform(size=n) < 0.5); o[m2] = target[m2] - ax[m2] * ext.max() * rng.uniform(0.1, 2.0, n)[m2, None]
    # 4: rays lying (almost) in the plane of a quad, starting on it or beside it (the ill-conditioned filter path)
    m = fam == 4
    e1 = Q[:, 1] - Q[:, 0]; e2 = Q[:, 3] - Q[:, 0]
    inpl = e1 * rng.uniform(-1, 1, n)[:, None] + e2 * rng.uniform(-1, 1, n)[:, None]
    nrm = np.cross(e1, e2); nn = np.linalg.norm(nrm, axis=1); nrm = nrm / np.where(nn > 0, nn, 1)[:, None]
    tilt = 10.0 ** rng.uniform(-9, -1, n) * rng.choice([-1, 0, 1], n, p=[0.15, 0.7, 0.15] if in_plane else None)
    d[m] = (inpl + nrm * (tilt * np.linalg.norm(inpl, axis=1))[:, None])[m]
    o[m] = (target - inpl * rng.uniform(0.7 if in_plane else 0.0, 3.0 if in_plane else 2.0, n)[:, None])[m]
    m4 = m & (rng.uniform(size=n) < 0.5); ign[m4] = qi[m4]
    # 5: from a quad's surface into a random direction, ignoring it
    m = fam == 5; o[m] = target[m]; ign[m] = qi[m]
    ln = np.linalg.norm(d, axis=1); bad = ~(ln > 0); d[bad] = [0, 0, 1]; ln[bad] = 1
    d = d / ln[:, None]
    return np.concatenate([o, d], axis=1).astype(np.float32), ign


CASES = [  # name, quads, kind, scale, offset, rays
    ("single quad", 1, "rect", 1.0, 0.0, 300_000),
    ("12 planar", 12, "rect", 1.0, 0.0, 1_000_000),
    ("19 axis boxes x550", 19, "axis", 550.0, 275.0, 1_000_000),
    ("16 mixed", 16, "mixed", 1.0, 0.0, 1_000_000),
    ("31 mixed x1e-3", 31, "mixed", 1e-3, 0.0, 500_000),
    ("24 mixed x1e6", 24, "mixed", 1e6, 0.0, 500_000),
    ("20 mixed far from origin", 20, "mixed", 10.0, 5e4, 500_000),
    ("45 mixed (> 32 entries)", 45, "mixed", 100.0, 0.0, 500_000),
    ("256 mixed (SSB_MAX_QUADS)", 256, "mixed", 1.0, 0.0, 300_000),
    ("14 degenerate", 14, "degenerate", 1.0, 0.0, 300_000),
    ("10 slivers x1000", 10, "sliver", 1000.0, 0.0, 300_000),
    ("18 trapezoids", 18, "trapezoid", 300.0, 100.0, 500_000),
    # rays IN the planes of the quads, large coordinates: the watertight test's U, V, W are rounding noise there and the
    # reference accepts "hits" beside the triangles (caught an in-plane pre-reject in round 2)
    ("in-plane rays, 18 trapezoids x300", 18, "trapezoid", 300.0, 100.0, 1_000_000),
    ("in-plane rays, 19 axis boxes x550", 19, "axis", 550.0, 275.0, 1_000_000),
    ("in-plane rays, 16 mixed x1e4", 16, "mixed", 1e4, 0.0, 500_000),
]


@pytest.mark.parametrize("name,nq,kind,scale,offset,nrays", CASES, ids=[c[0].replace(" ", "_") for c in CASES])
def test_device_scan_equals_list_scan_on_random_scenes(name, nq, kind, scale, offset, nrays):
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    qv = fuzz_quads(rng, nq, scale, offset, kind)
    sc = _scene(qv)
    rays, ign = _rays(rng, qv, nrays, in_plane=name.startswith("in-plane"))
    eps = np.float32(1e-3 if scale >= 1 else 1e-6)
    want = pu.oracle_intersect(sc, rays, ign, eps)
    with pu.ssb.Context(0) as ctx:
        ctx.upload_scene(sc)
        for mode in (abi.SSB_SCAN_FILTERED, abi.SSB_SCAN_LIST):
            got = ctx.intersect(rays, ign, scan_mode=mode, eps=eps)
            hit = want[0] >= 0
            bad = (got[0] != want[0]) | (hit & ((got[1] != want[1]) | (got[2].view(np.uint32) != want[2].view(np.uint32)) |
                                                (got[3].view(np.uint32) != want[3].view(np.uint32)).any(axis=1)))
            assert not bad.any(), (name, mode, int(bad.sum()), int(np.flatnonzero(bad)[0]), rays[np.flatnonzero(bad)[0]], ign[np.flatnonzero(bad)[0]],
                                   [g[np.flatnonzero(bad)[0]] for g in got], [w[np.flatnonzero(bad)[0]] for w in want])
    assert hit.mean() > 0.02 or name.startswith("in-plane"), (name, hit.mean())  # the rays do hit things


def test_render_with_the_list_scan_switch_is_bit_identical():
    """ssb_options.scan_mode = SSB_SCAN_LIST bypasses the filter in the render kernels themselves: same accumulators."""
    flat = pu.load_flat("cornell", "ours1931")
    opt = pu.options("ours1931", 48, 36, 6, seed=11)
    with pu.gpu_context(flat) as ctx:
        ctx.render(opt)
        a = ctx.read_accum(48, 36)
        opt.scan_mode = abi.SSB_SCAN_LIST
        ctx.render(opt)
        b = ctx.read_accum(48, 36)
    assert pu.bits_equal(a, b)
    acc_o, _, _ = pu.oracle_render(flat, pu.options("ours1931", 48, 36, 6, seed=11))
    assert pu.bits_equal(a, acc_o)
