"""The drop-in boundary exercised from PLAIN C (tests/c_abi/readme_example.c): gcc links the
program against libdexb200.so with nothing but include/*.h — no Python binding, no torch.

* without a GPU (`-m "not gpu"`): the program must report DEX_ERR_CUDA from the compute entry
  point (exit code 3) — the library has no CPU fallback;
* on a GPU (`-m gpu`): the README example of the reference evaluates to the closed form."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "dynamicexpressions.jl_b200", "lib")


def _build(tmp_path, name="readme_example"):
    exe = str(tmp_path / name)
    subprocess.check_call(["gcc", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_abi", name + ".c"), "-o", exe,
                           "-L", LIBDIR, "-ldexb200", "-lm", f"-Wl,-rpath,{LIBDIR}"])
    return exe


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_c_program_fails_loudly_without_a_gpu(tmp_path):
    if _has_gpu():
        pytest.skip("a GPU is present: covered by the gpu test")
    r = subprocess.run([_build(tmp_path)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 3, (r.returncode, r.stdout, r.stderr)
    assert "no CUDA device" in r.stdout


@pytest.mark.gpu
def test_readme_example_from_plain_c(tmp_path):
    r = subprocess.run([_build(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "complete=1" in r.stdout


def test_julia_extension_replay_builds_and_fails_loudly_without_a_gpu(tmp_path):
    """tests/c_abi/julia_ext_replay.c = the call sequence of ext/DynamicExpressionsB200Ext.jl in plain
    C (the Julia file cannot run here).  Compiling it checks the struct layout asserts and every
    prototype it uses; without a GPU the first compute call must fail with DEX_ERR_CUDA."""
    exe = _build(tmp_path, "julia_ext_replay")
    if _has_gpu():
        pytest.skip("a GPU is present: covered by the gpu test")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 3, (r.returncode, r.stdout, r.stderr)


@pytest.mark.gpu
def test_julia_extension_call_sequence_from_plain_c(tmp_path):
    r = subprocess.run([_build(tmp_path, "julia_ext_replay")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "all checks passed" in r.stdout


def test_julia_extension_file_binds_only_exported_symbols():
    """Every `ccall((:dex_..., LIB), ...)` of the extension names a function that include/dexb200.h
    declares and the library exports, with the right number of arguments."""
    import ctypes
    import re
    src = open(os.path.join(ROOT, "ext", "DynamicExpressionsB200Ext.jl")).read()
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "dexb200.h")).read(), flags=re.S)
    lib = ctypes.CDLL(os.path.join(LIBDIR, "libdexb200.so"))
    calls = re.findall(r"ccall\(\(:(dex_\w+), LIB\), \w+,\s*\(([^)]*)\)", src)
    assert len(calls) >= 15
    for name, argtypes in calls:
        assert hasattr(lib, name), name
        m = re.search(r"\b" + name + r"\s*\(([^;]*?)\)\s*;", hdr, flags=re.S)
        assert m, f"{name} is not declared in include/dexb200.h"
        n_decl = 0 if m.group(1).strip() in ("", "void") else m.group(1).count(",") + 1
        n_call = len([a for a in argtypes.split(",") if a.strip()])
        assert n_decl == n_call, (name, n_decl, n_call)
