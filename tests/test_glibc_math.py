"""Bit-exactness of the libm restatements (ctcdecode_b200/csrc/glibc_math.cuh) that make float32 scores -- and
therefore the integer outputs -- identical to the reference's.  Host sweep here; exhaustive device sweep on
the GPU box against that box's own libm."""
import os
import subprocess

import numpy as np
import pytest

from ctcdecode_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_sweep_strided():
    """Same source compiled for the host, every 97th float of each domain, against libm."""
    out = os.path.join(ROOT, "tests", "native", "_build", "sweep_glibc_math")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-mfma", "-ffp-contract=off", "-pthread",
                    os.path.join(ROOT, "tests", "native", "sweep_glibc_math.cpp"), "-o", out, "-lm"], check=True)
    r = subprocess.run([out, "97"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "expf 0/" in r.stdout and "logf 0/" in r.stdout and "logprob 0/" in r.stdout and "f64chain 0/" in r.stdout


def _device_vs_libm(cport, which, name, bits_lo, bits_hi, step=1, chunk=1 << 24):
    lib = _native.load()
    bad = 0
    total = 0
    for start in range(bits_lo, bits_hi + 1, chunk * step):
        u = np.arange(start, min(bits_hi + 1, start + chunk * step), step, dtype=np.uint64).astype(np.uint32)
        x = u.view(np.float32)
        y = np.empty_like(x)
        _native.check(lib.ctcdec_selftest_math(which, x.ctypes.data, None, y.ctypes.data, x.size, 0))
        ref = getattr(cport, name)(x)
        bad += int((ref.view(np.int32) != y.view(np.int32)).sum())
        total += x.size
    return bad, total


@pytest.mark.gpu
def test_device_logf_exhaustive(cport):
    bad, total = _device_vs_libm(cport, 1, "logf", 0x3F800000, 0x40000000)
    assert total == 8388609 and bad == 0


@pytest.mark.gpu
def test_device_expf_exhaustive(cport):
    """every float in [-17.5, -0.0] -- the whole domain log_sum_exp<float> can pass to expf"""
    hi = int(np.array([-17.5], np.float32).view(np.uint32)[0])
    bad, total = _device_vs_libm(cport, 0, "expf", 0x80000000, hi)
    assert total > 1_090_000_000 and bad == 0


@pytest.mark.gpu
def test_device_input_log_exhaustive(cport):
    """float(log(double(p) + FLT_MIN)) for every float p in [0, 1] (reference decoder_utils.cpp:40-43)"""
    bad, total = _device_vs_libm(cport, 2, "logprob", 0, 0x3F800000)
    assert total == 0x3F800000 + 1 and bad == 0


@pytest.mark.gpu
def test_device_lse_random(cport):
    lib = _native.load()
    rng = np.random.default_rng(5)
    x = (-50 * rng.random(1 << 22)).astype(np.float32)
    y = (x + rng.normal(0, 6, x.size)).astype(np.float32)
    x[:1000] = -3.4028235e38
    y[500:1500] = -3.4028235e38
    z = np.empty_like(x)
    _native.check(lib.ctcdec_selftest_math(3, x.ctypes.data, y.ctypes.data, z.ctypes.data, x.size, 0))
    assert np.array_equal(z.view(np.int32), cport.lse(x, y).view(np.int32))


@pytest.mark.gpu
def test_device_f64_chain_of_the_vocabulary_cut(cport):
    """the serial replay of the reference's double cum_prob chain (decoder_utils.cpp:26-31) in the prune kernel: glibc's
    double exp (arguments <= 0), log and log_sum_exp<double>, device against this box's libm, 2^22 random arguments
    each over the domains the chain can produce"""
    lib = _native.load()
    rng = np.random.default_rng(11)
    n = 1 << 22
    x = np.concatenate([-40.0 * rng.random(n // 2), -rng.random(n // 4), -np.ldexp(rng.random(n // 4), -rng.integers(0, 60, n // 4))])
    x[:4] = [0.0, -0.0, -1e-300, -40.0]
    y = np.empty_like(x)
    _native.check(lib.ctcdec_selftest_math_f64(0, x.ctypes.data, None, y.ctypes.data, x.size, 0))
    assert np.array_equal(y.view(np.int64), cport.f64(0, x).view(np.int64))
    s = 1.0 + rng.random(n)                                   # exp(0) + exp(<= 0) lies in [1, 2]
    _native.check(lib.ctcdec_selftest_math_f64(1, s.ctypes.data, None, y.ctypes.data, s.size, 0))
    assert np.array_equal(y.view(np.int64), cport.f64(1, s).view(np.int64))
    cum = np.log(2.0) * rng.random(n)
    p = rng.integers(1, 0x3F800001, n, dtype=np.int64).astype(np.uint32).view(np.float32).astype(np.float64)
    term = cport.f64(1, p)
    term[::7] = -0.37 * rng.integers(0, 3000, term[::7].size)  # log input: any non-positive number, far tails included
    term[:3] = [-np.inf, -1.7976931348623157e308, 0.0]
    _native.check(lib.ctcdec_selftest_math_f64(2, cum.ctypes.data, term.ctypes.data, y.ctypes.data, cum.size, 0))
    assert np.array_equal(y.view(np.int64), cport.f64(2, cum, term).view(np.int64))
