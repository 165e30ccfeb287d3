"""The C ABI from plain C: tests/c_abi/abi_vs_oracle.c links libfootsies_b200.so, the CUDA runtime (for the caller's
allocations) and the CPU oracle, and compares every field of every battle after every step -- no Python and no torch
in that process.  Device buffers (fg_bind / fg_step), host buffers (fg_step_host), the packed 16-byte records decoded in
C, masked RESET + SEED, fused frame skip, by_example.  Without a GPU the same binary must fail loudly (no CPU path)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_abi", "abi_vs_oracle.c")
EXE = os.path.join(ROOT, "tests", "c_abi", "abi_vs_oracle")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def build_program():
    import oracle_binding
    from footsies_gym_b200 import build as libbuild
    oracle_binding.build()
    lib = libbuild.build()
    libdir, orcdir = os.path.dirname(lib), os.path.join(ROOT, "oracle")
    cmd = ["gcc", "-O2", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), "-I", orcdir,
           "-I", os.path.join(CUDA, "include"), SRC, "-o", EXE, "-L", libdir, "-lfootsies_b200", "-L", orcdir,
           "-lfootsies_oracle", "-L", os.path.join(CUDA, "lib64"), "-lcudart",
           f"-Wl,-rpath,{libdir}", f"-Wl,-rpath,{orcdir}", f"-Wl,-rpath,{os.path.join(CUDA, 'lib64')}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return EXE


def test_c_program_builds_against_the_header_and_fails_loudly_without_a_gpu():
    """include/footsies_b200.h is valid C11 (-Wall -Wextra -Werror), every entry point the program uses links, and on a
    machine without a CUDA device fg_create refuses with FG_ERR_NO_DEVICE instead of falling back to anything."""
    import torch
    exe = build_program()
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the parity run is the gpu-marked test")
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert res.returncode == 3, (res.returncode, res.stdout, res.stderr)
    assert "no CUDA device" in res.stderr and "no CPU fallback" in res.stderr


@pytest.mark.gpu
def test_c_abi_parity_against_the_oracle_from_plain_c():
    exe = build_program()
    res = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "ABI PARITY OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("identical") == 4
