"""CPU: the C-ABI library loads and exports every symbol include/*.h declares (no compute without a GPU)."""
import ctypes as C
import importlib
import os
import re

import pytest

ssb = importlib.import_module("simple-spectral_b200")
host = importlib.import_module("simple-spectral_b200.host")
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(ssbh?_[a-z0-9_]+)\s*\(", text))


def test_every_declared_symbol_is_exported():
    L = ssb.lib()
    for header, bound in (("ssb200.h", ssb.EXPORTED_SYMBOLS), ("ssb200_host.h", host.HOST_SYMBOLS)):
        declared = _declared(header)
        assert declared, header
        for name in declared:
            assert hasattr(L, name), f"{name} declared in include/{header} but not exported by libssb200.so"
        assert declared == set(bound), (declared ^ set(bound))


def test_abi_version_and_defaults():
    L = ssb.lib()
    assert L.ssb_abi_version() == 4  # 4: row bands, scan mode, ssb_accum_merge, ssb_device_count, ssb_debug_intersect
    o = ssb.ssb_options()
    L.ssb_default_options(C.byref(o), 512, 512, 64)
    d = ssb.default_options(512, 512, 64)
    for f, _ in ssb.ssb_options._fields_:
        assert getattr(o, f) == getattr(d, f), f
    assert (o.max_depth, o.eps, o.lambda_min, o.lambda_max) == (10, pytest.approx(0.001), 380.0, 780.0)


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device the product must fail loudly, not route around the GPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ssb.SsbError) as e:
        ssb.Context(0)
    assert e.value.code == -1 and "no CPU path" in str(e.value)


def test_struct_sizes_match_c():
    # sizes the C compiler sees (include/ssb200.h), checked against the ctypes mirror
    import subprocess, tempfile, textwrap
    src = textwrap.dedent("""
        #include <stdio.h>
        #include "ssb200.h"
        int main(void){ printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(ssb_vertex), sizeof(ssb_tri), sizeof(ssb_quad),
          sizeof(ssb_spectrum), sizeof(ssb_material), sizeof(ssb_texture), sizeof(ssb_camera), sizeof(ssb_scene), sizeof(ssb_color), sizeof(ssb_options)); return 0; }
    """)
    with tempfile.TemporaryDirectory() as tmp:
        open(os.path.join(tmp, "s.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(tmp, "s.c"), "-o", os.path.join(tmp, "s")], check=True)
        out = subprocess.run([os.path.join(tmp, "s")], capture_output=True, text=True, check=True).stdout.split()
    names = ("ssb_vertex", "ssb_tri", "ssb_quad", "ssb_spectrum", "ssb_material", "ssb_texture", "ssb_camera", "ssb_scene", "ssb_color", "ssb_options")
    for n, s in zip(names, out):
        assert C.sizeof(getattr(ssb, n)) == int(s), n


def test_host_struct_size_matches_c():
    import subprocess, tempfile
    with tempfile.TemporaryDirectory() as tmp:
        open(os.path.join(tmp, "s.c"), "w").write('#include <stdio.h>\n#include "ssb200_host.h"\nint main(void){ printf("%zu\\n", sizeof(ssbh_renderer_options)); return 0; }\n')
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(tmp, "s.c"), "-o", os.path.join(tmp, "s")], check=True)
        out = subprocess.run([os.path.join(tmp, "s")], capture_output=True, text=True, check=True).stdout.split()
    assert C.sizeof(host.ssbh_renderer_options) == int(out[0])
