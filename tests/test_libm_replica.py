"""Host build of csrc/glibc_math.cuh (the source the kernels compile) against the host libm the reference
calls (powf at NeuCor.cpp:672,678,741; exp at NeuCor.cpp:695,710-711).  Bit-exact is the bar."""
import ctypes as C
import struct

import numpy as np
import pytest


def bits(f):
    return struct.unpack("<I", struct.pack("<f", f))[0]


@pytest.fixture(scope="module")
def M(native_libs):
    L = C.CDLL(native_libs[2])
    L.nc_mathhost_check_powf_range.restype = C.c_uint64
    L.nc_mathhost_check_powf_range.argtypes = [C.c_float, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_float)]
    L.nc_mathhost_check_exp_scaled_range.restype = C.c_uint64
    L.nc_mathhost_check_exp_scaled_range.argtypes = [C.c_double, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_float)]
    L.nc_mathhost_check_exp_gauss_range.restype = C.c_uint64
    L.nc_mathhost_check_exp_gauss_range.argtypes = [C.c_float, C.c_float, C.c_double, C.c_uint32, C.c_uint32, C.c_uint32]
    L.nc_mathhost_check_exp_random.restype = C.c_uint64
    L.nc_mathhost_check_exp_random.argtypes = [C.c_double, C.c_double, C.c_uint64, C.c_uint64, C.POINTER(C.c_double)]
    L.nc_mathhost_check_powf_random.restype = C.c_uint64
    L.nc_mathhost_check_powf_random.argtypes = [C.c_float] * 4 + [C.c_uint64, C.c_uint64]
    L.nc_mathhost_powf.restype = C.c_float
    L.nc_mathhost_powf.argtypes = [C.c_float, C.c_float]
    return L


@pytest.mark.parametrize("base", [0.5, 0.75, 0.65])
def test_powf_reference_bases_every_float_exponent(M, base):
    """recharge 0.5 (NeuCor.cpp:378), trace decays 0.75 / 0.65 (NeuCor.cpp:26-27); every 7th float in [2^-20, 1024)."""
    bad = C.c_float()
    assert M.nc_mathhost_check_powf_range(base, bits(2.0 ** -20), bits(1024.0), 7, C.byref(bad)) == 0, bad.value


def test_powf_step_sized_exponents_exhaustive(M):
    """every float dT in [2^-12, 0.25]: the range a sweep-mode step actually produces."""
    for base in (0.5, 0.75, 0.65):
        assert M.nc_mathhost_check_powf_range(base, bits(2.0 ** -12), bits(0.25), 1, None) == 0


def test_powf_random_bases(M):
    assert M.nc_mathhost_check_powf_random(0.01, 4.0, -200.0, 200.0, 5_000_000, 4) == 0


def test_powf_special_exponents(M):
    inf, nan = float("inf"), float("nan")
    assert M.nc_mathhost_powf(0.75, inf) == 0.0          # lastSpikeArrival = -inf before the first delivery (NeuCor.cpp:469)
    assert M.nc_mathhost_powf(0.65, 0.0) == 1.0          # fire and delivery at the same instant
    assert np.isnan(M.nc_mathhost_powf(0.65, nan))       # lastFire = NaN before the first fire (NeuCor.cpp:392)
    assert M.nc_mathhost_powf(1.0, nan) == 1.0
    assert M.nc_mathhost_powf(2.0, inf) == inf and M.nc_mathhost_powf(2.0, -inf) == 0.0 and M.nc_mathhost_powf(0.5, -inf) == inf


def test_exp_charge_argument(M):
    """exp(0.3702 * dT) for every 5th float dT in [2^-20, 64) (NeuCor.cpp:695)."""
    assert M.nc_mathhost_check_exp_scaled_range(0.3702, bits(2.0 ** -20), bits(64.0), 5, None) == 0


def test_exp_action_potential_arguments(M):
    """The two Gaussians of Neuron::AP (NeuCor.cpp:710-711) for every 3rd float t in (0, 2]."""
    d1 = float(np.float32(0.3)) * 2.0 * float(np.float32(0.3))
    d2 = float(np.float32(0.6)) * 2.0 * float(np.float32(0.6))
    assert M.nc_mathhost_check_exp_gauss_range(1.0, 0.0, d1, bits(2.0 ** -20), bits(2.0), 3) == 0
    assert M.nc_mathhost_check_exp_gauss_range(1.0, 1.16, d2, bits(2.0 ** -20), bits(2.0), 3) == 0


def test_exp_random(M):
    bx = C.c_double()
    assert M.nc_mathhost_check_exp_random(-500.0, 500.0, 10_000_000, 1, C.byref(bx)) == 0, bx.value
