"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every
symbol include/blomgpu.h declares, and fails loudly without a GPU."""
import ctypes
import re
from pathlib import Path

import pytest

from blom_b200 import lib as blib

ROOT = Path(__file__).resolve().parents[1]


def _ensure_built():
    if not blib.library_path(False).exists() or not blib.library_path(True).exists():
        from blom_b200.build import build
        build()


def header_symbols():
    text = (ROOT / "include" / "blomgpu.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(blomgpu_\w+)\s*\(", text)))


@pytest.mark.parametrize("parity", [False, True])
def test_exports_every_declared_symbol(parity):
    _ensure_built()
    lib = blib.load_library(parity)
    syms = header_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in blomgpu.h but not exported"
    assert sorted(blib.ABI_SYMBOLS) == syms
    assert lib.blomgpu_parity_build() == (1 if parity else 0)


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    _ensure_built()
    with pytest.raises(blib.BlomGpuError, match="no CUDA device"):
        blib.BlomGpu(24, 20, 5, 2)


def test_product_does_not_reference_oracle():
    for p in (ROOT / "blom_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".F90") and p.is_file():
            txt = p.read_text()
            assert "oracle" not in txt.lower().replace("test oracle", "").replace("the oracle", "") or p.name == "synth.py", p


def test_time_levels():
    assert blib.time_levels(0, 12) == (1, 2, 0, 12, 1, 13)
    assert blib.time_levels(1, 12) == (2, 1, 12, 0, 13, 1)
