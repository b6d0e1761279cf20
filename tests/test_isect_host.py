"""CPU: the device's scene_intersect (simple-spectral_b200/csrc/ssb_isect.cuh — packed conservative filter over the
filter entries of ssb_blob.hpp, nearest-candidate-first exact tests) compiled for the HOST by tools/isect_check.cpp and
compared, hit record by hit record and bit for bit, with the reference's plain list scan (Scene::intersect,
scene.cpp:433-445) on ~6 M rays: random, surface-to-surface, edge / corner / diagonal targeted, grazing, axis-aligned,
tied (duplicated / coplanar quads), on the reference's own scenes (quads from the golden table dumps) and on synthetic
ones (non-planar, degenerate, more than 32 filter entries, tiny / huge / far-from-origin coordinates).

The filter's packed-fp32 arithmetic and the approximate reciprocal differ between host and device only in rounding, and
by construction no rounding of the filter reaches the hit record: what this test pins down is the LOGIC — the order
independence of the nearest-first phase, the tie-breaks, the `ignore` rule, the split of non-planar quads, the in-order
fallback.  The device build itself is pinned by the bit-exact render comparisons of the GPU tests."""
import ctypes as C
import importlib
import os
import struct
import subprocess

import parity_util as pu
import refdump

_abi = importlib.import_module("simple-spectral_b200._abi")


def _write_quads(path, tables):
    q = tables["scene.quads"].reshape(-1, 2, 18)
    nq = q.shape[0]
    quads = (_abi.ssb_quad * nq)()
    for qi in range(nq):
        for ti in range(2):
            for vi in range(3):
                for k in range(3):
                    quads[qi].tri[ti].v[vi].pos[k] = float(q[qi, ti, vi * 5 + k])
                for k in range(2):
                    quads[qi].tri[ti].v[vi].st[k] = float(q[qi, ti, vi * 5 + 3 + k])
            for k in range(3):
                quads[qi].tri[ti].normal[k] = float(q[qi, ti, 15 + k])
    with open(path, "wb") as f:
        f.write(struct.pack("<I", nq))
        f.write(bytes(quads))
    return nq


def test_device_scene_intersect_equals_list_scan(tmp_path):
    exe = str(tmp_path / "isect_check")
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", os.path.join(pu.ROOT, "tools", "isect_check.cpp"), "-o", exe], check=True)
    files = []
    for name in ("cornell-srgb_ours1931", "plane-srgb_ours1931"):
        t = refdump.parse(os.path.join(pu.ROOT, "tests", "golden", f"tables_{name}.bin"))
        path = str(tmp_path / f"{name}.quads")
        assert _write_quads(path, t) in (19, 7)
        files.append(path)
    assert C.sizeof(_abi.ssb_quad) == 152
    r = subprocess.run([exe, "250000", *files], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    assert "mismatches 0" in r.stdout.splitlines()[-1]
    # the nearest-first phase must actually be the one that runs on the reference's scenes
    for line in r.stdout.splitlines()[:2]:
        assert "nearest-first 0." in line and float(line.split("nearest-first ")[1].split()[0]) > 0.5, line


def test_host_build_of_the_filter_on_the_device_fuzz_scenes(tmp_path):
    """The random scenes and ray families of tests/test_gpu_isect_fuzz.py (in-plane rays, slivers, degenerate and
    coplanar-overlapping quads, 1e-3 ... 1e6 coordinates) through the HOST build of the same scene_intersect
    (tools/isect_host_lib.cpp), against the checker's list scan — fewer rays than the device run, same generators."""
    import zlib
    import numpy as np
    import test_gpu_isect_fuzz as fz
    so = str(tmp_path / "libisect_host.so")
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", so, os.path.join(pu.ROOT, "tools", "isect_host_lib.cpp")], check=True)
    L = C.CDLL(so)
    L.isect_host.argtypes = [C.POINTER(_abi.ssb_quad), C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_int32), C.c_uint32, C.c_float, C.POINTER(C.c_float), C.c_size_t]
    for name, nq, kind, scale, offset, nrays in fz.CASES:
        rng = np.random.default_rng(zlib.crc32(name.encode()))
        qv = fz.fuzz_quads(rng, nq, scale, offset, kind)
        sc = fz._scene(qv)
        rays, ign = fz._rays(rng, qv, min(nrays, 400_000), in_plane=name.startswith("in-plane"))
        eps = np.float32(1e-3 if scale >= 1 else 1e-6)
        want = pu.oracle_intersect(sc, rays, ign, eps)
        out = np.empty((rays.shape[0], 6), np.float32)
        L.isect_host(sc.quads, nq, rays.ctypes.data_as(C.POINTER(C.c_float)), ign.ctypes.data_as(C.POINTER(C.c_int32)), 0, float(eps),
                     out.ctypes.data_as(C.POINTER(C.c_float)), rays.shape[0])
        got = (out[:, 0].view(np.int32), out[:, 1].view(np.int32), out[:, 2], out[:, 3:6])
        hit = want[0] >= 0
        bad = (got[0] != want[0]) | (hit & ((got[1] != want[1]) | (got[2].view(np.uint32) != want[2].view(np.uint32)) |
                                            (np.ascontiguousarray(got[3]).view(np.uint32) != want[3].view(np.uint32)).any(axis=1)))
        assert not bad.any(), (name, int(bad.sum()), rays[np.flatnonzero(bad)[0]], ign[np.flatnonzero(bad)[0]])
