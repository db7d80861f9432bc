"""CPU-side checks of the C-ABI library: it loads, exports exactly what include/blmm_b200.h declares,
and refuses to run without a B200 (there is no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "blmm_b200.h")).read()
    return sorted(set(re.findall(r"BLMM_API\s+[\w\s\*]+?\b(blmm_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_surface():
    syms = declared_symbols()
    for must in ("blmm_create", "blmm_destroy", "blmm_last_error", "blmm_kinship", "blmm_decompose", "blmm_rotate",
                 "blmm_bulkscan", "blmm_grid_loglik", "blmm_fit_h2", "blmm_scan_perms", "blmm_scan_null",
                 "blmm_lod2log10p", "blmm_thresholds", "blmm_weight_kinship", "blmm_create_multi",
                 "blmm_device_count", "blmm_last_gather_ms"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from blmm_b200 import _lib
    lib = _lib.load()
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/blmm_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert lib.blmm_abi_version() == _lib.ABI_VERSION == 3


def test_no_cpu_fallback():
    """Without a usable sm_100 device the engine fails loudly; nothing routes to the oracle."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from blmm_b200 import BlmmError, Engine, _lib
    with pytest.raises(BlmmError) as e:
        Engine(0)
    assert e.value.code == _lib.E_NO_DEVICE
    assert _lib.load().blmm_last_error(None) == b"context is NULL"


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "bulklmm.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".jl", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "blmm_oracle" not in txt, f"{f} references the oracle"


def test_struct_layouts_match_header():
    """blmm_problem / blmm_opts as the shims lay them out, checked against a C compile of the header."""
    import subprocess
    import tempfile
    from blmm_b200 import _lib
    assert C.sizeof(_lib.Problem) == 10 * 8
    assert C.sizeof(_lib.Opts) == 4 + 4 + 8 + 8 + 8 + 4 + 4 + 4 + 4 + 8 + 4 + 4 + 8
    assert _lib.Opts.h2_grid.offset == 24 and _lib.Opts.ld_out.offset == 48 and _lib.Opts.log10p_out.offset == 64
    src = """#include <stdio.h>
#include <stddef.h>
#include "blmm_b200.h"
int main(void) { printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(blmm_problem), sizeof(blmm_opts), offsetof(blmm_problem, obs_weights),
  offsetof(blmm_opts, h2_grid), offsetof(blmm_opts, ld_out), offsetof(blmm_opts, log10p_out)); return 0; }"""
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")], check=True)
        out = subprocess.run([os.path.join(d, "t")], capture_output=True, text=True, check=True).stdout.split()
    assert [int(x) for x in out] == [C.sizeof(_lib.Problem), C.sizeof(_lib.Opts), _lib.Problem.obs_weights.offset,
                                     _lib.Opts.h2_grid.offset, _lib.Opts.ld_out.offset, _lib.Opts.log10p_out.offset]
