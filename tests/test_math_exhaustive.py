"""CPU: the libm restatements the kernels use (simple-spectral_b200/csrc/ssb_math.cuh: sinf, cosf, the paired sincos,
acosf, powf at the two sRGB exponents) against the host libm the reference links — over ALL 2^32 float inputs per
function (tools/check_math_exhaustive.cpp; about a minute on 8 cores).  The device build of the same header is compared
with libm on random samples on the GPU (tests/test_gpu_parity.py::test_device_math_bit_exact)."""
import os
import subprocess

import parity_util as pu


def test_libm_restatements_exhaustive(tmp_path):
    exe = str(tmp_path / "check_math")
    subprocess.run(["g++", "-O2", "-fopenmp", "-ffp-contract=off", "-I", os.path.join(pu.ROOT, "simple-spectral_b200", "csrc"),
                    os.path.join(pu.ROOT, "tools", "check_math_exhaustive.cpp"), "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "all functions bit-identical to libm over all 2^32 inputs" in r.stdout
    assert r.stdout.count("mismatches: 0") == 7
