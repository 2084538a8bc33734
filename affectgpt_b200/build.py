"""Build libagb.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

The built library stays next to this file (git-ignored, but it travels with gpurun snapshots).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libagb.so"
STAMP = HERE / "csrc" / ".build_stamp"

SOURCES = ["agb_api.cu", "gemm_tcgen05.cu", "layernorm.cu", "attention.cu", "attention_tc2.cu", "vit.cu", "hubert.cu", "conv0_tc.cu",
           "fusion.cu", "qformer.cu", "staging.cu", "splice.cu", "pack.cu", "profile.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--use_fast_math", "-Xcompiler", "-fPIC",
    "-Xcompiler", "-Wall", "-Xptxas", "-v",
] + os.environ.get("AGB_NVCC_EXTRA", "").split()      # A/B experiments, e.g. AGB_NVCC_EXTRA=-DAGB_GELU_MUFU


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (needed to build affectgpt_b200/libagb.so)")


def _fingerprint() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h"))
                    + [HERE.parent / "include" / "agb.h", Path(__file__)]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()


def build_lib(force: bool = False, verbose: bool = False) -> Path:
    fp = _fingerprint()
    if not force and LIB.exists() and STAMP.exists() and STAMP.read_text().strip() == fp:
        return LIB
    nvcc = _nvcc()
    objs = []
    build_dir = HERE / "build"
    build_dir.mkdir(exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = build_dir / (src + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    failed = False
    for src, obj, p in procs:
        out, _ = p.communicate()
        log.append(f"==== {src}\n{out}")
        if p.returncode != 0:
            failed = True
        objs.append(str(obj))
    (build_dir / "nvcc.log").write_text("\n".join(log))
    if failed:
        sys.stderr.write("\n".join(log))
        raise RuntimeError("nvcc failed; see affectgpt_b200/build/nvcc.log")
    if verbose:
        print("\n".join(log))
    link = [nvcc, "-shared", "-o", str(LIB), *objs, "-gencode", "arch=compute_100a,code=sm_100a",
            "-Xcompiler", "-fPIC"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link of libagb.so failed")
    STAMP.write_text(fp)
    return LIB


if __name__ == "__main__":
    path = build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
