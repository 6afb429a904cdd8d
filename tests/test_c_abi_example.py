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


def _build(tmp_path):
    exe = str(tmp_path / "readme_example")
    subprocess.check_call(["gcc", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_abi", "readme_example.c"), "-o", exe,
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
