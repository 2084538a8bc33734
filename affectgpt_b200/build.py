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
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-Wall", "-Xptxas", "-v",
] + os.environ.get("AGB_NVCC_EXTRA", "").split()      # A/B experiments, e.g. AGB_NVCC_EXTRA=-DAGB_GELU_MUFU
# --use_fast_math (approximate division / sqrt, FTZ) only where the arithmetic is bf16-tolerance work anyway.  The fp32
# SIMT heads (fusion.cu, qformer.cu), the cv2-exact / torchaudio-exact staging kernels and the splice are compiled with
# IEEE division and no FTZ: their results are held to 1e-4 .. bit-exact against the reference.
PRECISE_SOURCES = {"fusion.cu", "qformer.cu", "staging.cu", "splice.cu", "pack.cu"}


def flags_for(src: str):
    return NVCC_FLAGS + ([] if src in PRECISE_SOURCES else ["--use_fast_math"])


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


def is_current() -> bool:
    """True when libagb.so exists and was built from the sources as they are now."""
    try:
        return LIB.exists() and STAMP.exists() and STAMP.read_text().strip() == _fingerprint()
    except OSError:
        return False


def build_lib(force: bool = False, verbose: bool = False) -> Path:
    """Build (if stale) under an exclusive file lock: ranks of a torchrun job that import the package at the same time
    serialise here, the first one builds, the others find a current library.  The link goes to a temporary file that is
    renamed into place, so a concurrent dlopen never sees a truncated libagb.so."""
    import fcntl
    if not force and is_current():
        return LIB
    (HERE / "build").mkdir(exist_ok=True)
    with open(HERE / "build" / ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and is_current():
                return LIB
            return _build_locked(verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose: bool) -> Path:
    fp = _fingerprint()
    nvcc = _nvcc()
    objs = []
    build_dir = HERE / "build"
    build_dir.mkdir(exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = build_dir / (src + ".o")
        cmd = [nvcc, *flags_for(src), "-c", str(CSRC / src), "-o", str(obj)]
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
    tmp = LIB.with_name(f".libagb.{os.getpid()}.so.tmp")
    link = [nvcc, "-shared", "-o", str(tmp), *objs, "-gencode", "arch=compute_100a,code=sm_100a",
            "-Xcompiler", "-fPIC"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        tmp.unlink(missing_ok=True)
        raise RuntimeError("link of libagb.so failed")
    os.replace(tmp, LIB)
    STAMP.write_text(fp)
    return LIB


if __name__ == "__main__":
    path = build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
