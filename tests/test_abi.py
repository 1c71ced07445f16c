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


def test_flavours_are_isolated():
    """Both flavours export the same C++ symbols.  Loaded into one process (as the tests and smoke()
    do) each must keep calling its own code: RTLD_LOCAL + -Bsymbolic.  blomgpu_parity_build() goes
    through an external-linkage C++ function, so interposition would show up as the wrong answer."""
    _ensure_built()
    import subprocess
    for order in ((False, True), (True, False)):
        code = ("import ctypes,sys; sys.path.insert(0, %r); from blom_b200 import lib as b; "
                "a = b.load_library(%r); c = b.load_library(%r); "
                "print(a.blomgpu_parity_build(), c.blomgpu_parity_build())" % (str(ROOT), order[0], order[1]))
        out = subprocess.run(["python", "-c", code], capture_output=True, text=True, check=True).stdout.split()
        assert out == [str(int(order[0])), str(int(order[1]))], (order, out)
    for parity in (False, True):
        dyn = subprocess.run(["readelf", "-d", str(blib.library_path(parity))], capture_output=True, text=True).stdout
        assert "SYMBOLIC" in dyn


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


def test_tools_do_not_reference_oracle():
    """tools/ holds profiling helpers of the product; scripts that need the oracle live in tests/dev/."""
    for p in (ROOT / "tools").glob("*.py"):
        assert "oracle" not in p.read_text().lower(), p


def test_time_levels():
    assert blib.time_levels(0, 12) == (1, 2, 0, 12, 1, 13)
    assert blib.time_levels(1, 12) == (2, 1, 12, 0, 13, 1)
