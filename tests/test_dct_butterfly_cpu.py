"""The 32 / 64-point DCT register network of combat_b200/csrc/dct_butterfly.cuh (even/odd recursion, 16- and 32-point
DCT-IV odd parts through a complex FFT) compiled for the HOST and held to scipy's orthonormal DCT-II / DCT-III, i.e. to
what utils/dct.py:13-82 of the reference computes.  No GPU needed; the CUDA kernels run the same header."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import scipy.fft

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("dct_host") / "libdct_butterfly_host.so")
    # -ffp-contract=off: only the explicit fmaf() calls fuse, exactly like the device code
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC",
                           os.path.join(HERE, "dct_butterfly_host.cpp"), "-o", so])
    lib = ctypes.CDLL(so)
    lib.dct_rows.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long]
    lib.dct_rows.restype = ctypes.c_int
    return lib


def run(lib, x, inverse):
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    assert lib.dct_rows(x.shape[1], int(inverse), x.ctypes.data, out.ctypes.data, x.shape[0]) == 0
    return out


@pytest.mark.parametrize("n", [32, 64])
def test_butterfly_matches_scipy(host_lib, n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal((4096, n)).astype(np.float32)
    for inverse, ref_fn in ((0, scipy.fft.dct), (1, scipy.fft.idct)):
        ref = ref_fn(x.astype(np.float64), norm="ortho", axis=1)
        got = run(host_lib, x, inverse)
        # a few float32 ulp of the largest coefficient: the network is made of rotations and +/- butterflies only
        assert np.abs(got - ref).max() / np.abs(ref).max() < 5e-7


@pytest.mark.parametrize("n", [32, 64])
def test_butterfly_basis_vectors_and_round_trip(host_lib, n):
    eye = np.eye(n, dtype=np.float32)
    D = run(host_lib, eye, 0)                     # rows = transforms of the unit vectors = D^T
    ref = scipy.fft.dct(np.eye(n), norm="ortho", axis=1)
    assert np.abs(D - ref).max() < 2e-7
    Dinv = run(host_lib, eye, 1)
    assert np.abs(Dinv @ D - np.eye(n)).max() < 1e-6      # inverse network really inverts
    # uint8-valued rows (the detector leg feeds 0..255): DC term up to 255 * sqrt(n)
    q = np.random.default_rng(1).integers(0, 256, (512, n)).astype(np.float32)
    back = run(host_lib, run(host_lib, q, 0), 1)
    assert np.abs(back - q).max() < 2e-4


def test_unsupported_size_is_rejected(host_lib):
    x = np.zeros((1, 16), dtype=np.float32)
    assert host_lib.dct_rows(16, 0, x.ctypes.data, x.ctypes.data, 1) == -1
