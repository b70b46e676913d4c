"""In-tree build of libtopo4d_b200.so with nvcc for sm_100a (cross-compiles without a GPU).

    python -m topo4d_b200.build [-v] [--force] [--tag NAME -DFOO=1 ...]

The library lands in ``topo4d_b200/_build/`` (git-ignored, shipped to the GPU box by gpurun).
Rebuilds only when a source is newer than the library.  ``--tag`` builds an experiment variant
``libtopo4d_b200_<tag>.so`` with extra -D defines (selected at run time with TOPO4D_B200_LIB=<path>).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
OUT_DIR = os.path.join(_HERE, "_build")
LIB_PATH = os.path.join(OUT_DIR, "libtopo4d_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC"]
# (source, extra flags).  gs_preprocess.cu / f3d_render.cu carry the bit-exact index/coverage
# arithmetic and must not contract mul+add into FMA.
SOURCES = [
    ("gs_preprocess.cu", ["--fmad=false"]),
    ("gs_binning.cu", []),
    ("gs_blend.cu", []),                       # -DGS_EXP_MODE=0|1|2 selects the exp flavour (default 2, see gs_blend.cu)
    ("gs_backward.cu", []),
    ("gs_api.cu", []),
    ("f3d_render.cu", ["--fmad=false"]),
    ("t4d_loss.cu", []),
    ("t4d_optim.cu", []),
    ("t4d_dense.cu", []),
    ("t4d_activate.cu", []),
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _source_hash(defines: list[str] | None = None) -> str:
    """Content hash of everything the library is built from (sources, header, this recipe, extra defines).  A content
    hash rather than mtimes: the tree travels to the GPU box as a snapshot whose file times mean nothing."""
    import hashlib
    h = hashlib.sha256()
    deps = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + [os.path.join(_HERE, "..", "include", "topo4d_b200.h"),
                                                                      os.path.abspath(__file__)]
    for d in deps:
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    h.update(" ".join(defines or []).encode())
    return h.hexdigest()


def _stale(lib_path: str, defines: list[str] | None = None) -> bool:
    if not os.path.exists(lib_path):
        return True
    try:
        with open(lib_path + ".srchash") as f:
            return f.read().strip() != _source_hash(defines)
    except OSError:
        return True


def build_library(force: bool = False, verbose: bool = False, tag: str = "", defines: list[str] | None = None) -> str:
    lib_path = LIB_PATH if not tag else os.path.join(OUT_DIR, f"libtopo4d_b200_{tag}.so")
    if not force and not _stale(lib_path, defines):
        return lib_path
    obj_dir = OUT_DIR if not tag else os.path.join(OUT_DIR, tag)
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = _nvcc()
    objs = []
    procs = []
    for src, extra in SOURCES:
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        cmd = [nvcc, *ARCH, *COMMON, *extra, *(defines or []), "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, pr in procs:
        out, _ = pr.communicate()
        if verbose and out:
            print(out, file=sys.stderr)
        if pr.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
    tmp = lib_path + f".tmp{os.getpid()}"
    link = [nvcc, *ARCH, "-shared", "-o", tmp, *objs]
    res = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}")
    os.replace(tmp, lib_path)                 # atomic: a concurrent loader sees the old or the new library, never half of one
    with open(lib_path + ".srchash", "w") as f:
        f.write(_source_hash(defines))
    return lib_path


if __name__ == "__main__":
    argv = sys.argv[1:]
    tag = argv[argv.index("--tag") + 1] if "--tag" in argv else ""
    print(build_library(force="--force" in argv, verbose="-v" in argv, tag=tag,
                        defines=[x for x in argv if x.startswith("-D")]))
