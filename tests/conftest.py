import importlib
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def wc():
    """The product package (hyphenated directory name -> importlib)."""
    return importlib.import_module("wavelet-compression_b200")


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    from oracle.pyoracle import Ref, have_ref
    if not have_ref():
        pytest.skip("oracle/_ref/libwcref.so not built (needs /root/reference)")
    return Ref()


class Golden:
    def __init__(self, path):
        self.z = np.load(path)
        self.manifest = json.loads(bytes(self.z["manifest"]))
        self.cases = self.manifest["cases"]

    def arrays(self, i):
        return {k: self.z[f"c{i}_{k}"] for k in ("in", "coef", "runs", "vals", "ser", "recon", "rmse")}


@pytest.fixture(scope="session")
def golden():
    return Golden(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"))


@pytest.fixture(scope="session")
def ctx(wc):
    """One wc_ctx on cuda:0 for the GPU tests.  No fallback: fails if the library or GPU is missing."""
    c = wc.Context(0)
    yield c
    c.close()


def same_bits(a, b):
    """Bit-exact equality.  The one exception: where BOTH values are NaN the payload/sign bits are
    not compared — x86 SSE produces the 'real indefinite' 0xFFC00000 for inf-inf, the GPU the
    canonical 0x7FFFFFFF; neither the reference nor its file format gives NaN payloads a meaning
    (NaN coefficients are never kept, so they never reach the packed stream)."""
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    if a.dtype != b.dtype or a.shape != b.shape:
        return False
    if a.dtype.kind == "f":
        na, nb = np.isnan(a), np.isnan(b)
        if not np.array_equal(na, nb):
            return False
        if na.any():
            a = np.where(na, 0, a).astype(a.dtype)
            b = np.where(nb, 0, b).astype(b.dtype)
    return a.tobytes() == b.tobytes()


def smooth_box(dims, rng, noise=1e-3, dtype=np.float32, sym=False):
    X, Y, Z = dims
    i, j, k = np.meshgrid(np.arange(X), np.arange(Y), np.arange(Z), indexing="ij")
    f = (0.0 if sym else 300.0) + 50 * np.sin(0.1 * i + 0.3) * np.cos(0.07 * j) * np.sin(0.05 * k + 0.1)
    f = f + noise * rng.standard_normal(f.shape)
    return np.ascontiguousarray(f.transpose(2, 1, 0)).astype(dtype)
