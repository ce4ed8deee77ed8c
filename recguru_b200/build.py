"""In-tree build of librecguru_b200.so (sm_100a only) with nvcc.

    python -m recguru_b200.build            # build if sources are newer than the .so
    python -m recguru_b200.build --force

nvcc cross-compiles without a GPU, so this runs in the authoring container; the built
.so is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import argparse
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "lib"
LIB = LIBDIR / "librecguru_b200.so"
OBJDIR = PKG.parent / "build" / "obj"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set $NVCC)")


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _stale(target: Path, deps: list[Path]) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False, debug_knobs: bool | None = None) -> Path:
    """``debug_knobs`` (or ``RECGURU_BUILD_DEBUG_KNOBS=1``): also compile the catalog sweep's bisection
    knobs, clock64 trace globals and experimental instantiations (-DRECGURU_DEBUG_KNOBS).  The
    production library is built without them."""
    if debug_knobs is None:
        debug_knobs = os.environ.get("RECGURU_BUILD_DEBUG_KNOBS", "0") not in ("", "0")
    flags = NVCC_FLAGS + (["-DRECGURU_DEBUG_KNOBS"] if debug_knobs else [])
    stamp = OBJDIR / ("flavour_debug" if debug_knobs else "flavour_prod")
    if not stamp.exists():          # switching flavour rebuilds everything
        force = True
    srcs = sources()
    headers = sorted(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "recguru_b200.h"]
    if not force and not _stale(LIB, srcs + headers):
        return LIB
    nvcc = _nvcc()
    OBJDIR.mkdir(parents=True, exist_ok=True)
    LIBDIR.mkdir(parents=True, exist_ok=True)
    for f in OBJDIR.glob("flavour_*"):
        f.unlink()

    def compile_one(src: Path) -> Path:
        obj = OBJDIR / (src.stem + ".o")
        if force or _stale(obj, [src] + headers):
            cmd = [nvcc, *flags, "-c", str(src), "-o", str(obj)]
            if verbose:
                print(" ".join(cmd), flush=True)
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    cmd = [nvcc, "-shared", "-o", str(LIB), *map(str, objs), "-cudart", "static"]
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    stamp.touch()
    return LIB


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("-v", "--verbose", action="store_true")
    ap.add_argument("--debug-knobs", action="store_true", default=None,
                    help="compile the catalog sweep's debug knobs and experimental variants in")
    a = ap.parse_args()
    print(build(force=a.force, verbose=a.verbose, debug_knobs=a.debug_knobs))


if __name__ == "__main__":
    main()
