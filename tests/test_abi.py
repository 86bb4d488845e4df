"""The C-ABI shared library loads and exports every symbol include/b200feat.h declares
(no compute calls: there is no GPU in the CPU tier)."""
import ctypes
import os
import re

import pytest

from lhotse_b200 import engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "b200feat.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200feat_[a-z_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    lib = ctypes.CDLL(engine.lib_path())
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/b200feat.h but not exported"
    assert sorted(engine.EXPORTS) == syms


def test_version_and_struct_layout():
    lib = engine.load_library()
    assert lib.b200feat_version() == 1
    assert ctypes.sizeof(engine.PlanDesc) == 16 * 4 + 4 * 4
    assert ctypes.sizeof(engine.BatchTotals) == 6 * 8
    assert lib.b200feat_meta_words(10) == 42


def test_library_is_sm100a_only():
    import shutil
    import subprocess

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", engine.lib_path()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_gpu_means_loud_failure():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    from lhotse_b200 import B200Fbank

    with pytest.raises(Exception) as ei:
        B200Fbank().extract(__import__("numpy").zeros(16000, dtype="float32"), 16000)
    assert "no CUDA device" in str(ei.value) or "CPU fallback" in str(ei.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "lhotse_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f
                assert "kaldi_oracle" not in txt, f
