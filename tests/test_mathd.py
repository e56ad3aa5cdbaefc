"""device/mathd.cuh (sin / cos rounded once from double-double values, used by the panorama cameras) compiled as HOST code:
correctly rounded against mpmath, and equal to the host's libm except where that misrounds (glibc: ~0.1 % of results)."""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

from tests import common as T

SRC = os.path.join(os.path.dirname(T.HERE), "mallie_b200", "csrc", "device")


@pytest.fixture(scope="module")
def sincos(tmp_path_factory):
    d = tmp_path_factory.mktemp("mathd")
    src = d / "w.cc"
    src.write_text('#include "mathd.cuh"\n'
                   'extern "C" void sincos_rn_batch(const double *x, long n, double *s, double *c) {\n'
                   '  for (long i = 0; i < n; i++) mb200::mathd::sincos_rn(x[i], s[i], c[i]);\n}\n')
    lib = d / "libw.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-I", SRC, "-o", str(lib), str(src)])
    L = C.CDLL(str(lib))

    def run(x):
        x = np.ascontiguousarray(x, np.float64)
        s, c = np.empty_like(x), np.empty_like(x)
        L.sincos_rn_batch(x.ctypes.data_as(C.c_void_p), C.c_long(len(x)), s.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p))
        return s, c
    return run


def angles():
    rng = np.random.default_rng(1)
    return np.concatenate([rng.uniform(0, 2 * math.pi, 300_000), rng.uniform(-50, 50, 50_000),
                           math.pi * rng.integers(0, 2048, 30_000) / 1024.0,        # theta of pixel centres
                           2 * math.pi * rng.integers(0, 4096, 30_000) / 4096.0,    # phi of pixel centres
                           rng.uniform(-8e5, 8e5, 5_000), 10.0 ** rng.uniform(-300, -1, 5_000)])


def test_sincos_rn_is_correctly_rounded_on_a_sample(sincos):
    mp = pytest.importorskip("mpmath")
    mp.mp.prec = 200
    x = angles()[::40]
    s, c = sincos(x)
    bad = sum((float(mp.sin(mp.mpf(float(a)))) != b) + (float(mp.cos(mp.mpf(float(a)))) != d) for a, b, d in zip(x, s, c))
    assert bad == 0, f"{bad} of {2 * len(x)} results are not the correctly rounded value"


def test_sincos_rn_equals_the_host_libm_where_that_rounds_correctly(sincos):
    x = angles()
    s, c = sincos(x)
    gs, gc = np.sin(x), np.cos(x)
    for got, want in ((s, gs), (c, gc)):
        diff = np.abs(got.view(np.int64) - want.view(np.int64))
        assert diff.max() <= 1                                 # never more than the neighbouring double
        assert (diff != 0).mean() <= 3e-3                      # glibc 2.39 misrounds ~1.3e-3 of its results


def test_sincos_rn_special_values(sincos):
    s, c = sincos([0.0, -0.0, 1e-300, math.pi / 2, math.pi, 1e6, -1e6, math.inf, math.nan])
    assert s[0] == 0.0 and math.copysign(1.0, s[1]) == -1.0 and c[0] == c[1] == 1.0 and s[2] == 1e-300
    assert (s[3], c[3]) == (1.0, 6.123233995736766e-17) and (s[4], c[4]) == (1.2246467991473532e-16, -1.0)
    assert s[5] == math.sin(1e6) and c[6] == math.cos(-1e6)    # beyond the reduction's range: the library functions
    assert np.isnan(s[7:]).all() and np.isnan(c[7:]).all()
