"""CPU: the C-ABI library loads and exports every symbol include/wcgpu.h declares; POD layouts match
the numpy mirrors; without a GPU the library reports WC_ERR_NO_DEVICE instead of falling back."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "wcgpu.h")).read()
    return sorted(set(re.findall(r"WC_API\s+[\w\s\*]+?\b(wc_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported_and_bound(wc):
    syms = declared_symbols()
    assert len(syms) >= 30
    lib = wc.capi.load()
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/wcgpu.h but not exported by libwcgpu.so"
    assert set(syms) == set(wc.capi.SIGNATURES), set(syms) ^ set(wc.capi.SIGNATURES)
    assert lib.wc_version() == 200
    assert wc.capi.strerror(0) == "ok" and "fallback" in wc.capi.strerror(3)


def test_pod_layouts_match_c(wc):
    code = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "wcgpu.h"
    int main(void) {
        printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(wc_box_desc), sizeof(wc_box_out), sizeof(wc_pair),
               sizeof(wc_packed), offsetof(wc_packed, pairs), offsetof(wc_box_desc, nx), offsetof(wc_packed, npairs));
        return 0;
    }'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(code)
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                               os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])
        out = subprocess.check_output([os.path.join(d, "t")]).decode().split()
    assert [int(x) for x in out] == [24, 24, 8, 32, 24, 12, 16]
    assert wc.capi.BOX_DESC.itemsize == 24 and wc.capi.PACKED.itemsize == 32
    assert wc.capi.PACKED.fields["pairs"][1] == 24 and wc.capi.BOX_DESC.fields["nx"][1] == 12


def test_serialize_header_is_host_only(wc):
    lib = wc.capi.load()
    rec = np.zeros(1, wc.capi.PACKED)
    rec[0]["shape"] = (16, 32, 64)
    rec[0]["ncoef"] = 32768
    rec[0]["npairs"] = 4096
    out = np.zeros(20, np.uint8)
    assert lib.wc_serialize_header(rec.ctypes.data, out.ctypes.data) == 0
    assert list(out.view("<i4")) == [16, 32, 64, 32768, 4096]


def test_no_gpu_means_error_not_fallback(wc):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = wc.capi.load()
    h = C.c_void_p()
    assert lib.wc_create(C.byref(h), 0) == 3  # WC_ERR_NO_DEVICE
    with pytest.raises(wc.WcError):
        wc.Context(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "wavelet-compression_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pyoracle" not in txt and "wc_oracle" not in txt and "libwcref" not in txt, f
