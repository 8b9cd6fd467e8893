"""The C-ABI library loads without a GPU and exports exactly what include/astre_gpu.h declares."""
import ctypes as C
import os
import re
import subprocess

import numpy as np

import astre_b200
from astre_b200 import _lib, gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "astre_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"ASTRE_API\s+[\w\s\*]+?\b(astre_\w+)\s*\(", text)))


def test_header_declares_the_bound_symbols():
    assert header_symbols() == sorted(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol(gpu_lib):
    out = subprocess.run(["nm", "-D", "--defined-only", astre_b200.library_path()], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT\s+(astre_\w+)", out))
    assert exported == set(header_symbols())
    for name in header_symbols():
        assert getattr(gpu_lib, name) is not None


def test_no_torch_or_cxx_types_in_the_abi():
    text = open(os.path.join(ROOT, "include", "astre_gpu.h")).read()
    for banned in ("torch", "at::", "std::", "cudaStream_t stream", "template"):
        assert banned not in re.sub(r"/\*.*?\*/", "", text, flags=re.S)


def test_version_and_host_helpers_work_without_gpu(gpu_lib, oracle):
    assert b"sm_100a" in gpu_lib.astre_gpu_version()
    off, tot = gpu.global_offsets(np.array([[3, 1], [2, 5], [0, 7]], np.uint32), rank=1)
    assert tot.tolist() == [5, 13] and off.tolist() == [3, 5 + 1]


def test_product_never_touches_the_oracle():
    """A product path that routes through the oracle voids every parity claim: nothing
    under astre_b200/ may import, link or execute anything under oracle/ or tests/."""
    pkg = os.path.join(ROOT, "astre_b200")
    banned = ("oracle_lib", "astre_oracle", "libastre_oracle", "oracle/", "orc_", "np_reference", "refstore",
              "libastre_ref", "ref_lib", "_ref/", "ref_world_", "refglm")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")) or f == "Makefile":
                text = open(os.path.join(dirpath, f)).read()
                for b in banned:
                    assert b not in text, (os.path.join(dirpath, f), b)
    out = subprocess.run(["ldd", astre_b200.library_path()], capture_output=True, text=True).stdout
    assert "oracle" not in out and "astre_ref" not in out
