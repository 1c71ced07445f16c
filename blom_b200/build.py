"""In-tree build of the CUDA library (sm_100a) with plain nvcc.

Two flavours are produced from the same sources:
  libblomgpu.so         performance build (FMA contraction on)
  libblomgpu_parity.so  parity build (-fmad=false, mirrors the reference's
                        -ffp-contract=off release flags, meson.build:17-19)
Both are git-ignored but travel to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
BUILD = HERE / "_build"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++",
          "-Xptxas", "-v", "--expt-relaxed-constexpr"]

FLAVOURS = {
    "libblomgpu.so": [],
    "libblomgpu_parity.so": ["-fmad=false", "-DBLOM_PARITY_BUILD"],
}


def _newer(src: Path, dst: Path, extra: list[Path]) -> bool:
    if not dst.exists():
        return True
    t = dst.stat().st_mtime
    return any(p.stat().st_mtime > t for p in [src, *extra])


# per-file extra flags (development experiments: BLOM_EXTRA_<STEM>="-Xptxas -dlcm=cg")
def _extra(src: Path) -> list[str]:
    return os.environ.get("BLOM_EXTRA_" + src.stem.upper(), "").split()


def _compile(src: Path, obj: Path, flags: list[str], log: Path) -> None:
    cmd = [NVCC, *ARCH, *COMMON, *flags, *_extra(src), "-c", str(src), "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log.write_text(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError(f"nvcc failed for {src.name}")


def build(verbose: bool = False, force: bool = False) -> list[Path]:
    BUILD.mkdir(exist_ok=True)
    headers = list(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "blomgpu.h"]
    sources = sorted(CSRC.glob("*.cu"))
    outs = []
    jobs = []
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        for lib, flags in FLAVOURS.items():
            tag = lib.replace(".so", "")
            for s in sources:
                obj = BUILD / f"{tag}_{s.stem}.o"
                if force or _newer(s, obj, headers):
                    jobs.append(ex.submit(_compile, s, obj, flags, BUILD / f"{tag}_{s.stem}.log"))
        for j in jobs:
            j.result()
    for lib in FLAVOURS:
        tag = lib.replace(".so", "")
        objs = [str(BUILD / f"{tag}_{s.stem}.o") for s in sources]
        out = HERE / lib
        if force or jobs or not out.exists():
            # -Bsymbolic: references to the library's own symbols bind inside the library, whatever
            # else the process has loaded (e.g. the other flavour of this library)
            cmd = [NVCC, *ARCH, "-shared", "-ccbin", "/usr/bin/g++", "-Xlinker", "-Bsymbolic", "-o", str(out), *objs,
                   "-lcudart", "-ldl"]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError(f"link failed for {lib}")
        outs.append(out)
        if verbose:
            print("built", out)
    return outs


if __name__ == "__main__":
    build(verbose=True, force="--force" in sys.argv)
