"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/ffcuda.h declares, and
refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from ffcuda_lib import HEADER, ROOT, ffcuda


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ffcuda_[a-z0-9_]+)\s*\(", src)))


def test_library_is_built():
    assert os.path.exists(ffcuda.LIB_PATH), "run `make -C freefem-sources_b200` (or __graft_entry__.build())"


def test_every_declared_symbol_is_exported():
    L = ffcuda.lib()
    names = _declared()
    assert len(names) >= 50
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert sorted(ffcuda.SYMBOLS) == names


def test_exports_are_plain_c():
    out = subprocess.run(["nm", "-D", "--defined-only", ffcuda.LIB_PATH], capture_output=True, text=True).stdout
    exported = [ln.split()[-1] for ln in out.splitlines() if " T " in ln]
    ours = [s for s in exported if s.startswith("ffcuda_")]
    assert set(_declared()) <= set(ours)


def test_product_never_links_the_oracle():
    out = subprocess.run(["ldd", ffcuda.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out
    for root, _, files in os.walk(os.path.join(ROOT, "freefem-sources_b200")):
        if "build" in root:
            continue
        for f in files:
            if f.endswith((".cu", ".cuh", ".cpp", ".py", ".h")):
                txt = open(os.path.join(root, f), errors="ignore").read()
                assert "fforacle" not in txt and "oracle_lib" not in txt, f


def test_fails_loudly_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(ffcuda.FfcudaError) as e:
        ffcuda.Context(0)
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)
