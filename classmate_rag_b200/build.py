"""Build libcmrag.so in-tree with nvcc for sm_100a (no JIT cache: the built
library travels with the repo snapshot to the GPU box)."""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libcmrag.so"
SOURCES = ["api.cu", "dense.cu", "dense_mma.cu", "bm25.cu", "bm25_mma.cu", "fuse.cu", "text.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or Path(cand).exists()):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "cmrag.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False, defines=(), out: Path = None) -> Path:
    """Compile and link.  ``defines`` / ``out``: experiment builds (extra -D flags into a
    separately named library, loaded through the CMRAG_LIB environment variable)."""
    lib_out = LIB if out is None else Path(out)
    if out is None and not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    objs = []
    build_dir = PKG / ("build" if out is None else "build_" + lib_out.stem)
    build_dir.mkdir(exist_ok=True)
    log = []
    for src in SOURCES:
        obj = build_dir / (src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-c", str(CSRC / src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log.append(r.stderr)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError(f"nvcc failed on {src}")
        objs.append(str(obj))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(lib_out), *objs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    (build_dir / "ptxas.log").write_text("\n".join(log))
    if verbose:
        print("\n".join(log))
    return lib_out


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs, out=outs[0] if outs else None))
