"""The C-ABI library: every symbol include/breakscore.h declares is exported, nothing links
torch / Python, and without a GPU the library fails loudly (no CPU fallback)."""
import os
import re
import subprocess

import pytest

from conftest import ROOT
from genomeassembler_dev_b200 import breakscore as B


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "breakscore.h")).read()
    return sorted(set(re.findall(r"BS_API[^;(]*?\b(bs_\w+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(B.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol(product_lib):
    out = subprocess.run(["nm", "-D", "--defined-only", product_lib], check=True, capture_output=True, text=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    assert set(declared_symbols()) <= exported
    # hidden visibility: nothing but the C-ABI (and toolchain symbols) leaks
    assert not [s for s in exported if s.startswith("_ZN2bs")]


def test_library_is_standalone(product_lib):
    deps = subprocess.run(["ldd", product_lib], check=True, capture_output=True, text=True).stdout
    for forbidden in ("torch", "python", "libR", "oracle"):
        assert forbidden not in deps


def test_library_holds_sm100a_code(product_lib):
    out = subprocess.run(["cuobjdump", "-lelf", product_lib], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


def test_loads_and_fails_loudly_without_gpu(product_lib):
    lib = B.load_library(product_lib)
    assert lib.bs_abi_version() == 5
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(B.BreakscoreError) as e:
        B.BreakageScorer(0, product_lib)
    assert e.value.code == 5 and "no CPU fallback" in str(e.value)


def test_missing_library_is_an_error(tmp_path):
    with pytest.raises(FileNotFoundError):
        B.load_library(str(tmp_path / "libbreakscore.so"))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "genomeassembler_dev_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, f
