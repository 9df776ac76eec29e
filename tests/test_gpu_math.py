"""GPU suite: accuracy of the kernels' branch-free FP64 elementary functions
(csrc/cloudsc2_math.cuh) against numpy/libm on the argument ranges CLOUDSC2 produces.  They replace
the Fortran intrinsics EXP / TANH / COSH / SQRT and "/" of cloudsc2.F90, cloudsc2tl.F90,
cloudsc2ad.F90; the bounds asserted here are what the field tolerances in test_gpu_nl/tl/ad rely on.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
N = 200_000


def _ulps(got, ref):
    return np.abs(got - ref) / np.spacing(np.abs(ref))


def test_rcp(gpu_nl):
    rng = np.random.default_rng(1)
    x = np.concatenate([10.0 ** rng.uniform(-30, 30, N), -10.0 ** rng.uniform(-12, 12, N // 4),
                        rng.uniform(1.0, 2.0, N)])
    got = gpu_nl.math_probe(0, x)
    assert _ulps(got, 1.0 / x).max() <= 1.0


def test_exp(gpu_nl):
    rng = np.random.default_rng(2)
    x = np.concatenate([rng.uniform(-700, 700, N), rng.uniform(-30, 8, N), rng.uniform(-1e-3, 1e-3, N),
                        np.array([0.0, -0.0, 700.0, -700.0, np.log(2) / 2, -np.log(2) / 2])])
    got = gpu_nl.math_probe(1, x)
    assert _ulps(got, np.exp(x)).max() <= 2.0
    assert gpu_nl.math_probe(1, np.zeros(4)).tolist() == [1.0] * 4
    # clamped variant: arguments far below -700 give exp(-700) (stands in for the underflow to 0)
    y = gpu_nl.math_probe(2, np.array([-1e5, -800.0, -3.0, 0.0]))
    assert y[0] == y[1] and 0 < y[0] < 1e-300
    assert _ulps(y[2:], np.exp(np.array([-3.0, 0.0]))).max() <= 2.0


def test_sqrt(gpu_nl):
    rng = np.random.default_rng(3)
    x = np.concatenate([10.0 ** rng.uniform(-40, 40, N), rng.uniform(0, 4, N), np.zeros(3)])
    got = gpu_nl.math_probe(3, x)
    assert _ulps(got[:-3], np.sqrt(x[:-3])).max() <= 1.0
    assert (got[-3:] == 0.0).all()


def test_tanh_and_sech2(gpu_nl):
    rng = np.random.default_rng(4)
    a = rng.uniform(-20, 12, N)          # 0.17 * (T - RLPTRC) for T in 150 .. 340 K
    t = gpu_nl.math_probe(4, a)
    # reference without the cancellation of tanh(a) + 1 for a << 0: tanh(a) + 1 = 2 / (1 + exp(-2a))
    ref = 2.0 / (1.0 + np.exp(-2.0 * a))
    assert _ulps(t, ref).max() <= 6.0
    # and against the expression the reference evaluates, to its (absolute) accuracy
    assert np.abs(t - (np.tanh(a) + 1.0)).max() <= 4 * np.spacing(1.0)
    s = gpu_nl.math_probe(5, a)
    assert _ulps(s, 1.0 / np.cosh(a) ** 2).max() <= 8.0
