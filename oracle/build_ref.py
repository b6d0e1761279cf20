#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY — builds the real reference into oracle/_ref/.

Compiles the reference's own C++ sources *from where they lie* under /root/reference/src
(no reference source is copied into this repository; patched temporaries live in a
scratch directory under $TMPDIR and only binaries are written to oracle/_ref/).

The reference cannot be built as shipped here: GLM is an external, un-vendored,
unpinned dependency (cmake/FindGLM.cmake) and is not installed.  We compile against
oracle/glm_shim (our restatement of GLM 0.9.9 scalar semantics for the ~20 functions
the reference uses).  The reference's CMake build is NOT run.

Outputs (per variant v in ours1931, ours2006, meng, jh, ours1931_noels, ours1931_noffc, rgb, ours1931_nw3, meng_nw2, ours1931_d3):
  oracle/_ref/simple_spectral_<v>          pristine sources, -O3 -march=x86-64-v3 -DNDEBUG
                                            (CPU timing arm; multi-threaded, nondeterministic)
  oracle/_ref/simple_spectral_<v>_hooked   sources + oracle/ref_hooks.hpp spliced into
                                            renderer.cpp, -O2 -DNDEBUG -ffp-contract=off
                                            (parity oracle; per-sample seeding + dumps,
                                             inert unless SSB_* env vars are set)
Variants are the reference's compile-time macros (stdafx.hpp:66,81), selected by
editing the scratch copy of stdafx.hpp exactly as the reference's author intends.
"""
import os
import re
import shutil
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SSB_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
SHIM = os.path.join(HERE, "glm_shim")

VARIANTS = {
    "ours1931": dict(alg=1, observer=1931),
    "ours2006": dict(alg=1, observer=2006),
    "meng": dict(alg=2, observer=1931),
    "jh": dict(alg=3, observer=1931),
    # EXPLICIT_LIGHT_SAMPLING compiled out (stdafx.hpp:44): exercises MaterialMirror (scene.cpp:346-355) and emission
    # on every hit (renderer.cpp:167-175).  (FLAT_FIELD_CORRECTION cannot be compiled out in this mode: the reference's
    # own color.cpp:275-279 then refers to an undeclared `flux`.)
    "ours1931_noels": dict(alg=1, observer=1931, no_els=True),
    # FLAT_FIELD_CORRECTION compiled out (stdafx.hpp:55): flux = radiance * dot(camera ray, camera dir), renderer.cpp:262-266.
    # The reference does not compile as shipped in this configuration — Color::round_trip_lrgb (color.cpp:275-279, a
    # self-test helper that is NOT on the render path) then uses an undeclared `flux` —, so the scratch copy of color.cpp
    # gets the one declaration that function lacks; renderer.cpp is untouched.
    "ours1931_noffc": dict(alg=1, observer=1931, no_ffc=True),
    # RENDER_MODE_RGB (stdafx.hpp:62-90 `#if 1` -> `#if 0`): the three-channel comparison renderer
    "rgb": dict(alg=1, observer=1931, rgb=True),
    # SAMPLE_WAVELENGTHS (stdafx.hpp:90) other than 4: glm::vec<N,float> exists for N = 2, 3 (not 1 without extra headers)
    "ours1931_nw3": dict(alg=1, observer=1931, nw=3),
    "meng_nw2": dict(alg=2, observer=1931, nw=2),
    # MAX_DEPTH (stdafx.hpp:47) 3 instead of 10
    "ours1931_d3": dict(alg=1, observer=1931, max_depth=3),
}
# Not a render variant: the reference's own round-trip self-test (main.cpp:184-264, compiled out with `#if 0` upstream),
# switched on in the scratch copy.  Built only on request (`build_ref.py roundtrip`); it prints the running maximum of the
# sRGB -> spectrum -> sRGB error after every red level — the source of tests/golden/roundtrip_running_max.json and of the
# one number the reference documents, 1.851469e-5 (main.cpp:242-245).
TOOLS = {"roundtrip": dict(alg=1, observer=1931)}
# Not a CPU variant either: THE COMPILED BOUNDARY.  The real reference with integration/renderer_ssb200.cpp (the stub of
# INTEGRATION.md) added and Renderer::render_start / render_wait of renderer.cpp compiled out, linked against
# simple-spectral_b200/libssb200.so: main.cpp, Color::init, Scene::get_new_*, Framebuffer::save are the reference's own
# code, the per-pixel loop is the library.  Needs a GPU to RUN (tests/test_gpu_boundary.py); built here like the others.
BOUNDARY = {"ssb200": dict(alg=1, observer=1931), "ssb200_jh": dict(alg=3, observer=1931)}
ROOT = os.path.dirname(HERE)
CXX_SOURCES = [
    "main.cpp", "renderer.cpp", "scene.cpp", "geometry.cpp", "material.cpp", "spectrum.cpp",
    "framebuffer.cpp", "util/color.cpp", "util/random.cpp", "util/spherical-tri.cpp",
]
C_SOURCES = ["jakob-and-hanika-2019/rgb2spec.c"]
LODEPNG = "util/lodepng/lodepng.cpp"

FLAGS_PRISTINE = ["-O3", "-march=x86-64-v3", "-DNDEBUG"]
FLAGS_HOOKED = ["-O2", "-DNDEBUG", "-ffp-contract=off"]


def run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise SystemExit("oracle/build_ref.py: command failed")
    return r


def sub_once(text, pattern, repl, what):
    new, n = re.subn(pattern, repl, text, count=1, flags=re.M)
    if n != 1:
        raise SystemExit(f"oracle/build_ref.py: patch point not found: {what}")
    return new


def patch_variant(src_dir, alg, observer, no_els=False, no_ffc=False, rgb=False, nw=4, max_depth=10):
    p = os.path.join(src_dir, "stdafx.hpp")
    t = open(p, encoding="utf-8-sig").read()
    if no_els:
        t = sub_once(t, r"^#define EXPLICIT_LIGHT_SAMPLING$", "//#define EXPLICIT_LIGHT_SAMPLING", "ELS")
    if no_ffc:
        t = sub_once(t, r"^#define FLAT_FIELD_CORRECTION$", "//#define FLAT_FIELD_CORRECTION", "FFC")
        pc = os.path.join(src_dir, "util", "color.cpp")
        tc = open(pc, encoding="utf-8-sig").read()
        tc = sub_once(tc, r"(#else\n\t\t)assert\(false\);(\n\t#endif\n\n\t//\tViewer-perceived XYZ)", r"\1SpectralRadiantFlux const& flux = radiance;\2", "FFC (round_trip_lrgb)")
        open(pc, "w", encoding="utf-8").write(tc)
    if max_depth != 10:
        t = sub_once(t, r"#define MAX_DEPTH 10u", f"#define MAX_DEPTH {max_depth}u", "MAX_DEPTH")
    if nw != 4:
        t = sub_once(t, r"#define SAMPLE_WAVELENGTHS 4_zu", f"#define SAMPLE_WAVELENGTHS {nw}_zu", "SAMPLE_WAVELENGTHS")
    if rgb:
        t = sub_once(t, r"#if 1(\s+#define RENDER_MODE_SPECTRAL\n)", r"#if 0\1", "RENDER_MODE")
    t = sub_once(t, r"#define RENDER_MODE_SPECTRAL_ALGNUM 1", f"#define RENDER_MODE_SPECTRAL_ALGNUM {alg}", "ALGNUM")
    if observer == 2006:
        t = sub_once(t, r"#if 1(\s+#define CIE_OBSERVER 1931)", r"#if 0\1", "CIE_OBSERVER")
    open(p, "w", encoding="utf-8").write(t)


def patch_hooks(src_dir):
    shutil.copy(os.path.join(HERE, "ref_hooks.hpp"), os.path.join(src_dir, "ssb_ref_hooks.hpp"))
    # the hooks read a few private members
    for name in ("spectrum.hpp", "material.hpp"):
        p = os.path.join(src_dir, name)
        t = open(p, encoding="utf-8-sig").read().replace("private:", "public:")
        open(p, "w", encoding="utf-8").write(t)
    p = os.path.join(src_dir, "renderer.cpp")
    t = open(p, encoding="utf-8-sig").read()
    t = sub_once(t, r'(#include "scene.hpp"\n)', r'\1#include "ssb_ref_hooks.hpp"\n', "include")
    # thread count override (renderer.cpp:45)
    t = sub_once(t, r"_threads\.resize\(std::thread::hardware_concurrency\(\)\);",
                 "_threads.resize(ssb_hooks::threads(std::thread::hardware_concurrency()));\n"
                 "\tssb_hooks::replace_scene(scene);\n"
                 "\tssb_hooks::dump_tables(scene);", "threads")
    # per-sample seeding + sample dump (renderer.cpp:293-295, spectral branch)
    t = sub_once(t, r"avg \+= _render_sample\(rng, i,j\) \* 0\.001f;",
                 "{ ssb_hooks::seed_sample(rng,i,j,k); auto ssb_s = _render_sample(rng, i,j); "
                 "ssb_hooks::record_sample(i,j,k,ssb_s); avg += ssb_s * 0.001f; }", "sample")
    # ... and the RGB branch (renderer.cpp:301-303)
    t = sub_once(t, r"avg \+= _render_sample\(rng, i,j\);",
                 "{ ssb_hooks::seed_sample(rng,i,j,k); auto ssb_s = _render_sample(rng, i,j); "
                 "ssb_hooks::record_sample(i,j,k,ssb_s); avg += ssb_s; }", "sample (rgb)")
    t = sub_once(t, r"(avg /= static_cast<double>\(options\.spp\);\n)",
                 r"\1\t\tssb_hooks::record_pixel(i,j,avg);\n", "pixel (rgb)")
    # pixel dump (after renderer.cpp:296)
    t = sub_once(t, r"(avg \*= 1000\.0 / static_cast<double>\(options\.spp\);\n)",
                 r"\1\t\tssb_hooks::record_pixel(i,j,avg);\n", "pixel")
    # begin / finish
    t = sub_once(t, r"(\t_num_tiles_start = _tiles\.size\(\);)",
                 r"\tssb_hooks::begin(options.res[0],options.res[1],options.spp);\n\1", "begin")
    t = sub_once(t, r"(\t\tframebuffer\.save\(options\.output_path\);)", r"\t\tssb_hooks::finish();\n\1", "finish")
    open(p, "w", encoding="utf-8").write(t)


def patch_roundtrip(src_dir):
    p = os.path.join(src_dir, "main.cpp")
    t = open(p, encoding="utf-8-sig").read()
    t = sub_once(t, r"#if 0 && defined RENDER_MODE_SPECTRAL", "#if 1 && defined RENDER_MODE_SPECTRAL", "round-trip block")
    t = sub_once(t, r"#if 0(\s+float max_error = 0\.0f;)", r"#if 1\1", "round-trip loop")
    old_print = 'printf("\\r%d (%e)   ",r,static_cast<double>(max_error));'
    if t.count(old_print) != 1:
        raise SystemExit("oracle/build_ref.py: patch point not found: round-trip print")
    t = t.replace(old_print, 'printf("RUNNING %d %.9e\\n",r,static_cast<double>(max_error)); fflush(stdout);')
    open(p, "w", encoding="utf-8").write(t)


def patch_boundary(src_dir):
    """renderer.cpp: compile out the two functions the stub replaces; the stub reads a few private members."""
    shutil.copy(os.path.join(ROOT, "integration", "renderer_ssb200.cpp"), os.path.join(src_dir, "renderer_ssb200.cpp"))
    for name in ("spectrum.hpp", "material.hpp"):
        p = os.path.join(src_dir, name)
        t = open(p, encoding="utf-8-sig").read().replace("private:", "public:")
        open(p, "w", encoding="utf-8").write(t)
    p = os.path.join(src_dir, "renderer.cpp")
    t = open(p, encoding="utf-8-sig").read()
    t = sub_once(t, r"^void Renderer::render_start\(\) \{$", "#ifndef SSB200_BOUNDARY\nvoid Renderer::render_start() {", "render_start")
    t = sub_once(t, r"(^void Renderer::render_wait \(\) \{\n(?:.*\n)*?^\}\n)", r"\1#endif\n", "render_wait")
    open(p, "w", encoding="utf-8").write(t)


def build_one(tmp, name, alg, observer, hooked, lodepng_obj, roundtrip=False, boundary=False, **variant_kw):
    tag = name + ("_hooked" if hooked else "")
    src_dir = os.path.join(tmp, tag, "src")
    shutil.copytree(os.path.join(REF, "src"), src_dir, ignore=shutil.ignore_patterns("lodepng*"))
    for root, _, files in os.walk(src_dir):
        os.chmod(root, 0o755)
        for f in files:
            os.chmod(os.path.join(root, f), 0o644)
    patch_variant(src_dir, alg, observer, **variant_kw)
    if hooked:
        patch_hooks(src_dir)
    if roundtrip:
        patch_roundtrip(src_dir)
    if boundary:
        patch_boundary(src_dir)
    flags = FLAGS_HOOKED if (hooked or roundtrip or boundary) else FLAGS_PRISTINE
    if boundary:
        flags = [*flags, "-DSSB200_BOUNDARY", "-I", os.path.join(ROOT, "include")]
    inc = ["-I", SHIM, "-I", os.path.join(REF, "src")]  # lodepng.h is found through the original tree
    objs = []
    for s in CXX_SOURCES + (["renderer_ssb200.cpp"] if boundary else []):
        o = os.path.join(tmp, tag, s.replace("/", "_") + ".o")
        # util/*.cpp include "lodepng/lodepng.h" relative to util/, which is not copied: add that dir
        run(["g++", "-std=c++17", "-w", *flags, *inc, "-I", os.path.join(REF, "src", "util"),
             "-c", os.path.join(src_dir, s), "-o", o])
        objs.append(o)
    for s in C_SOURCES:
        o = os.path.join(tmp, tag, s.replace("/", "_") + ".o")
        run(["gcc", "-w", *flags, "-c", os.path.join(src_dir, s), "-o", o])
        objs.append(o)
    out = os.path.join(OUT, "simple_spectral_" + tag)
    link = []
    if boundary:  # oracle/_ref/<binary> finds the library two directories up, wherever the repo lies
        lib_dir = os.path.join(ROOT, "simple-spectral_b200")
        if not os.path.exists(os.path.join(lib_dir, "libssb200.so")):
            raise SystemExit("oracle/build_ref.py: build simple-spectral_b200/libssb200.so first (__graft_entry__.build())")
        link = ["-L", lib_dir, "-lssb200", "-Wl,-rpath,$ORIGIN/../../simple-spectral_b200"]
    run(["g++", *objs, lodepng_obj, "-o", out, "-pthread", "-lm", *link])
    return out


def main():
    if not os.path.isdir(os.path.join(REF, "src")):
        print("oracle/build_ref.py: reference not present at", REF, "- keeping prebuilt oracle/_ref/")
        return 0
    which = sys.argv[1:] or list(VARIANTS)
    os.makedirs(OUT, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="ssb_ref_build_")
    try:
        lodepng_obj = os.path.join(tmp, "lodepng.o")
        # lodepng (vendored by the reference) is compiled straight from /root/reference
        run(["g++", "-std=c++17", "-w", "-O2", "-DNDEBUG", "-c", os.path.join(REF, "src", LODEPNG), "-o", lodepng_obj])
        jobs = []
        with ThreadPoolExecutor(max_workers=4) as ex:
            for name in which:
                if name in TOOLS:
                    jobs.append(ex.submit(build_one, tmp, name, TOOLS[name]["alg"], TOOLS[name]["observer"], False, lodepng_obj, roundtrip=True))
                    continue
                if name in BOUNDARY:
                    jobs.append(ex.submit(build_one, tmp, name, BOUNDARY[name]["alg"], BOUNDARY[name]["observer"], False, lodepng_obj, boundary=True))
                    continue
                v = VARIANTS[name]
                for hooked in (False, True):
                    kw = {k: v[k] for k in ("no_els", "no_ffc", "rgb", "nw", "max_depth") if k in v}
                    jobs.append(ex.submit(build_one, tmp, name, v["alg"], v["observer"], hooked, lodepng_obj, **kw))
            for j in jobs:
                print("built", os.path.relpath(j.result(), os.path.dirname(HERE)))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
