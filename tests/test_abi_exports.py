"""CPU: libtsgpu.so loads without a GPU and exports every function include/tsgpu.h declares; without a CUDA device the
product fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "tsgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(tsgpu_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_every_declared_symbol_is_exported(tsgpu):
    lib = tsgpu.lib()
    names = declared_functions()
    assert len(names) > 70
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in tsgpu.h but not exported: {missing}"
    assert lib.tsgpu_abi_version() == 1


def test_no_cpu_fallback(tsgpu):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(tsgpu.TwistAndShoutError) as e:
        tsgpu.Context(0)
    assert e.value.variant == "ProofGeneration"


def test_product_does_not_link_or_import_the_oracle():
    """the shipped package must not reference oracle/ (it is test infrastructure)"""
    pkg = os.path.join(ROOT, "multilinear-map-cryptography_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath).startswith("build"):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in text and "import oracle" not in text and "oracle/" not in text.replace("the oracle/", ""), os.path.join(dirpath, f)


def test_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): exactly one JSON line on stdout with the contract's keys; no GPU needed"""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, check=True).stdout
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "twist_prove_ms_at_2^20_ops" and d["unit"] == "ms" and d["higher_is_better"] is False
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["vs_baseline"] is None and d["gpu_launches"] == 0


def test_header_is_plain_c(tmp_path):
    """include/tsgpu.h is the drop-in boundary: it must compile as C99 (and as C++) on its own, and a C program must link against libtsgpu.so"""
    import subprocess
    src = tmp_path / "hdr.c"
    src.write_text('#include "tsgpu.h"\nint main(void) { return tsgpu_abi_version() == 1 ? 0 : 1; }\n')
    inc = os.path.join(ROOT, "include")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc, "-fsyntax-only", str(src)])
    subprocess.check_call(["g++", "-std=c++11", "-Wall", "-Werror", "-I", inc, "-fsyntax-only", "-x", "c++", str(src)])
    libdir = os.path.join(ROOT, "multilinear-map-cryptography_b200")
    exe = tmp_path / "hdr"
    subprocess.check_call(["gcc", "-std=c99", "-I", inc, str(src), "-o", str(exe), "-L", libdir, "-ltsgpu", "-Wl,-rpath," + libdir])
    assert subprocess.call([str(exe)]) == 0


def _build_demo(tmp_path):
    import subprocess
    libdir = os.path.join(ROOT, "multilinear-map-cryptography_b200")
    exe = tmp_path / "demo"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "demo.c"),
                           "-o", str(exe), "-L", libdir, "-ltsgpu", "-Wl,-rpath," + libdir])
    return str(exe)


def test_c_demo_builds_and_fails_loudly_without_a_gpu(tmp_path):
    """examples/demo.c (the reference's examples/demo.rs over the C ABI) compiles warning-free; without a CUDA device it stops at tsgpu_init"""
    import subprocess
    import torch
    exe = _build_demo(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: see test_c_demo_on_gpu")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "tsgpu_init failed (2)" in r.stderr


@pytest.mark.gpu
def test_c_demo_on_gpu(tmp_path):
    """the C demo proves and verifies the reference's demo trace and lookup table, and reports the reference's limit error"""
    import subprocess
    r = subprocess.run([_build_demo(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "Twist::verify -> true" in r.stdout and "Shout::verify -> true" in r.stdout
    assert "error 1: Too many operations" in r.stdout
