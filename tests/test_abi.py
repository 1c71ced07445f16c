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


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """The boundary is a C ABI: include/blomgpu.h compiles as C99 (no C++ or torch types) and a C
    program links against the library and calls it."""
    import subprocess
    _ensure_built()
    src = tmp_path / "host.c"
    src.write_text('#include "blomgpu.h"\nint main(void) { return blomgpu_parity_build(); }\n')
    exe = tmp_path / "host"
    libdir = blib.library_path(False).parent
    for lib, expect in (("blomgpu", 0), ("blomgpu_parity", 1)):
        subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", str(ROOT / "include"),
                        str(src), "-o", str(exe), "-L", str(libdir), f"-l{lib}", f"-Wl,-rpath,{libdir}"],
                       check=True, capture_output=True)
        assert subprocess.run([str(exe)]).returncode == expect


def test_fortran_shim_binds_the_abi():
    """blom_b200/fortran/mod_blomgpu.F90 cannot be compiled here (no Fortran compiler), so at least its
    bind(C) names are checked against include/blomgpu.h: every name it binds exists, and everything a
    Fortran host needs (all but the Python-side instrumentation entries) is bound."""
    f90 = (ROOT / "blom_b200" / "fortran" / "mod_blomgpu.F90").read_text()
    bound = set(re.findall(r"bind\(C,\s*name='(blomgpu_\w+)'\)", f90))
    declared = set(header_symbols())
    assert bound <= declared, sorted(bound - declared)
    instrumentation = {"blomgpu_device_ptr", "blomgpu_ktimers_enable", "blomgpu_ktimers_get", "blomgpu_launch_count",
                       "blomgpu_launch_count_reset", "blomgpu_parity_build", "blomgpu_stream", "blomgpu_timers_enable",
                       "blomgpu_timers_get", "blomgpu_timers_reset"}
    assert declared - bound <= instrumentation, sorted(declared - bound - instrumentation)
    # same number of arguments on both sides of every binding
    text = re.sub(r"/\*.*?\*/", "", (ROOT / "include" / "blomgpu.h").read_text(), flags=re.S)
    nargs_c = {}
    for m in re.finditer(r"\b(blomgpu_\w+)\s*\(([^)]*)\)\s*;", text):
        a = m.group(2).strip()
        nargs_c[m.group(1)] = 0 if a in ("void", "") else len(a.split(","))
    joined = re.sub(r"&\s*\n\s*", "", f90)
    nargs_f = {}
    for m in re.finditer(r"function\s+(blomgpu_\w+)\s*\(([^)]*)\)\s*bind\(C", joined):
        a = m.group(2).strip()
        nargs_f[m.group(1)] = 0 if a == "" else len(a.split(","))
    for m in re.finditer(r"procedure\(six_int_entry\),\s*bind\(C,\s*name='(blomgpu_\w+)'\)", joined):
        nargs_f[m.group(1)] = 6
    assert set(nargs_f) == bound
    assert not [(k, nargs_c[k], v) for k, v in nargs_f.items() if nargs_c[k] != v]
    # the reference's entry points keep their names (phy/mod_blom_step.F90:96-227)
    for name in ("init_fluxes", "tmsmt1", "eddtra", "advect", "pbcor1", "diffus", "pgforc", "momtum", "barotp",
                 "pbcor2", "tmsmt2"):
        assert re.search(r"subroutine %s\(" % name, f90), name


def test_tools_do_not_reference_oracle():
    """tools/ holds profiling helpers of the product; scripts that need the oracle live in tests/dev/."""
    for p in (ROOT / "tools").glob("*.py"):
        assert "oracle" not in p.read_text().lower(), p


def test_time_levels():
    assert blib.time_levels(0, 12) == (1, 2, 0, 12, 1, 13)
    assert blib.time_levels(1, 12) == (2, 1, 12, 0, 13, 1)
