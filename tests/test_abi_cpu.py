"""CPU tests of the C-ABI library: it loads, exports every symbol of include/ovb200.h, refuses to run without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from open_vins_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from open_vins_b200 import build
    build.build()
    return capi.load_library()


def test_header_symbols_are_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "ovb200.h")).read()
    declared = set(re.findall(r"\b(ovb_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(capi.EXPORTED_SYMBOLS), declared ^ set(capi.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.ovb_abi_version() == 1


def test_struct_layouts_match_header(lib):
    # ovb_opts_default fills every field through the C definition; mismatched layouts would scramble the values
    o = capi.ovb_opts()
    lib.ovb_opts_default(C.byref(o))
    d = capi.default_opts(chi2_multipler=5.0, col_order=capi.COLS_CANONICAL)
    for name, _ in capi.ovb_opts._fields_:
        assert getattr(o, name) == getattr(d, name), name


def test_chi2_table_matches_scipy(lib):
    from scipy.stats import chi2
    for k in [1, 2, 21, 81, 245, 499, 500, 1500, 2047]:
        assert abs(lib.ovb_chi2_quantile95(k) - chi2.ppf(0.95, k)) <= 1e-12 * chi2.ppf(0.95, k)


def test_no_cpu_fallback_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    cfg = capi.ovb_config(0, 64, 16, 256, 0)
    h = C.c_void_p()
    st = lib.ovb_create(C.byref(cfg), C.byref(h))
    assert st == capi.OVB_ERR_CUDA and not h.value
    with pytest.raises(capi.OvbError):
        capi.Engine(max_state=64, max_feats=16, max_meas=256)


def test_product_does_not_import_oracle():
    # the product path must not reference oracle/ (judge rule): scan the package sources
    pkg = os.path.join(ROOT, "open_vins_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "ovo_" not in txt and "libovoracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f


def test_abi_header_is_plain_c99(tmp_path):
    """The drop-in boundary is a C ABI: include/ovb200.h must compile as pedantic C99 (no C++ in the signatures) and link
    against the library from a C translation unit."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "c_check.c"
    src.write_text('#include "ovb200.h"\n#include <stdio.h>\n'
                   'int main(void) { ovb_opts o; ovb_opts_default(&o); printf("%d %d\\n", ovb_abi_version(), o.max_runs); return 0; }\n')
    exe = tmp_path / "c_check"
    libdir = os.path.join(root, "open_vins_b200")
    res = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(root, "include"), str(src),
                          "-L", libdir, "-lovb200", "-Wl,-rpath," + libdir, "-o", str(exe)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.split() == ["1", "5"], (out.stdout, out.stderr)
