#!/usr/bin/env python
"""Build an experimental copy of the library with extra -D flags (kernel tuning only).

    python scripts/build_variant.py NAME -DT2_SPM_VAL=2 -DT2_R_VAL=3 ...

writes breedgym_b200/_variants/NAME.so (git-ignored, travels with gpurun); run any script with
BG_LIB_PATH=breedgym_b200/_variants/NAME.so to use it.
"""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from breedgym_b200 import build as B  # noqa: E402


def main():
    name, flags = sys.argv[1], sys.argv[2:]
    out_dir = B.PKG / "_variants"
    obj_dir = out_dir / ("_obj_" + name)
    obj_dir.mkdir(parents=True, exist_ok=True)
    procs = []
    for s in B.SOURCES:
        obj = obj_dir / (Path(s).stem + ".o")
        procs.append((obj, subprocess.Popen([B._nvcc(), *B.NVCC_FLAGS, *flags, "-c", str(B.CSRC / s), "-o", str(obj)],
                                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for obj, p in procs:
        out, _ = p.communicate()
        if p.returncode:
            sys.stderr.write(out)
            raise SystemExit(f"nvcc failed for {obj.name}")
    so = out_dir / (name + ".so")
    subprocess.check_call([B._nvcc(), "-shared", "-o", str(so), *[str(o) for o, _ in procs], "-gencode", "arch=compute_100a,code=sm_100a", "-ldl"])
    print(so)


if __name__ == "__main__":
    main()
