"""The C-ABI without Python: tests/cabi_smoke.c is compiled with gcc against include/blmm_b200.h, linked to
libblmm_b200.so, and (on the GPU box) run on the committed fixture tests/golden/cabi_smoke.bin."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "bulklmm.jl_b200", "blmm_b200", "lib")
EXE = os.path.join(LIBDIR, "cabi_smoke")


def build():
    cmd = ["gcc", "-O1", "-Wall", "-Werror", "-std=c99", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cabi_smoke.c"), "-o", EXE, "-L", LIBDIR, "-lblmm_b200",
           f"-Wl,-rpath,{LIBDIR}", "-Wl,-rpath,/usr/local/cuda/lib64", "-lm"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return EXE


def test_cabi_smoke_compiles_and_links_as_c99():
    """the header is valid C99 and every entry point the program uses resolves at link time"""
    assert os.path.exists(build())


@pytest.mark.gpu
def test_cabi_smoke_runs_without_python():
    exe = build()
    r = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "cabi_smoke.bin")], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "cabi_smoke ok" in r.stdout


@pytest.mark.gpu
def test_cabi_smoke_multi_gpu_context():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    exe = build()
    r = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "cabi_smoke.bin"), "2"], capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
