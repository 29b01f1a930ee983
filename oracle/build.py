"""Build recipe for the oracle's C restatement (gcc only; TEST INFRASTRUCTURE).

The reference (`/root/reference`) is pure Python and its hot-path arithmetic
lives in `chromax`/`jax`, which are not vendored and not installable here, so
there is nothing to compile into `oracle/_ref/`: this recipe builds only our own
restatement `oracle/csrc/oracle.c` -> `oracle/liboracle.so`.
"""
from __future__ import annotations

import subprocess
from pathlib import Path

_HERE = Path(__file__).resolve().parent
SRC = _HERE / "csrc" / "oracle.c"
OUT = _HERE / "liboracle.so"


def build(force: bool = False) -> Path:
    if OUT.exists() and not force and OUT.stat().st_mtime >= SRC.stat().st_mtime:
        return OUT
    cmd = ["gcc", "-O3", "-march=x86-64-v2", "-fopenmp", "-fPIC", "-shared", "-std=c11", "-Wall",
           str(SRC), "-o", str(OUT)]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
