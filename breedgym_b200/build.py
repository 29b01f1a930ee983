"""In-tree build of libbreedgym_b200.so (hand-written sm_100a CUDA + C ABI).

`python -m breedgym_b200.build [--force] [--verbose]`.  nvcc cross-compiles
without a GPU; the .so is git-ignored but travels with the gpurun snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OUT = PKG / "libbreedgym_b200.so"
OBJ_DIR = PKG / "_obj"
SOURCES = ["api.cu", "meiosis.cu", "gebv.cu", "gebv_tc2.cu", "cross_gebv.cu", "cross_gebv_dyn.cu", "layout.cu", "comm.cu", "topk.cu", "peer.cu", "pairs.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.sep not in c or os.path.exists(c)):
            return c
    return "nvcc"


def _deps_mtime() -> float:
    files = list(CSRC.glob("*")) + [PKG.parent / "include" / "breedgym_b200.h"]
    return max(f.stat().st_mtime for f in files)


def build(force: bool = False, verbose: bool = False) -> Path:
    srcs = [CSRC / s for s in SOURCES if (CSRC / s).exists()]
    if OUT.exists() and not force and OUT.stat().st_mtime >= _deps_mtime():
        return OUT
    OBJ_DIR.mkdir(exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src: Path):
        obj = OBJ_DIR / (src.stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, r

    with ThreadPoolExecutor(max_workers=len(srcs)) as ex:
        results = list(ex.map(compile_one, srcs))
    log = []
    for src, obj, r in results:
        log.append(f"== {src.name}\n{r.stderr}")
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError(f"nvcc failed on {src.name}")
    (OBJ_DIR / "ptxas.log").write_text("\n".join(log))
    if verbose:
        print("\n".join(log))
    link = [nvcc, "-shared", "-o", str(OUT), *[str(o) for _, o, _ in results],
            "-gencode", "arch=compute_100a,code=sm_100a", "-ldl"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
