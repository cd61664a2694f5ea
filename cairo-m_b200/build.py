"""In-tree build of libcm31.so (sm_100a) and of the CPU oracle used by the tests.

`python cairo-m_b200/build.py [--force]` or `__graft_entry__.build()`.
nvcc cross-compiles without a GPU; the resulting .so files travel to the GPU box with the repo.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
BUILD = PKG / "build"
LIB = PKG / "libcm31.so"
ORACLE_DIR = ROOT / "oracle"
ORACLE_LIB = ORACLE_DIR / "liboracle.so"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-fopenmp", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _newer(target: Path, deps) -> bool:
    if not target.exists():
        return False
    t = target.stat().st_mtime
    return all(Path(d).stat().st_mtime <= t for d in deps)


def _run(cmd, log: Path | None = None):
    res = subprocess.run(cmd, capture_output=True, text=True)
    if log is not None:
        log.write_text(res.stdout + res.stderr)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("build failed: " + " ".join(map(str, cmd)))
    return res


GEN_DIR = CSRC / "generated"
GEN_TOOL_SRC = ROOT / "tools" / "gen_air_kernels.cpp"


def generate_air_kernels(force: bool = False):
    """AOT-specialised AIR kernels: build + run tools/gen_air_kernels.cpp -> csrc/generated/*.cu
    (the generator only rewrites files whose content changed)."""
    GEN_DIR.mkdir(exist_ok=True)
    BUILD.mkdir(exist_ok=True)
    tool = BUILD / "gen_air_kernels"
    deps = [GEN_TOOL_SRC] + list((CSRC / "host").glob("*.hpp")) + list((CSRC / "air").glob("*.hpp")) + [CSRC / "field.cuh", CSRC / "circle.hpp"]
    if force or not _newer(tool, deps):
        _run(["g++", "-O1", "-std=c++17", "-I", str(CSRC), "-I", str(ROOT / "include"), str(GEN_TOOL_SRC), "-o", str(tool)])
    _run([str(tool), str(GEN_DIR)])


def cuda_sources():
    return sorted(CSRC.glob("*.cu")) + sorted((CSRC / "host").glob("*.cu")) + sorted(GEN_DIR.glob("*.cu"))


def headers():
    hs = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.hpp")) + list((ROOT / "include").glob("*.h"))
    hs += list((CSRC / "host").glob("*.hpp")) + list((CSRC / "air").glob("*.hpp")) + list((CSRC / "cairo").glob("*.hpp"))
    return hs


def build_lib(force: bool = False) -> Path:
    BUILD.mkdir(exist_ok=True)
    generate_air_kernels(force)
    srcs = cuda_sources()
    hdrs = headers()
    objs = []
    jobs = []
    for s in srcs:
        o = BUILD / (s.stem + ".o")
        objs.append(o)
        if force or not _newer(o, [s] + hdrs):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        _run([NVCC, *NVCC_FLAGS, "-I", str(ROOT / "include"), "-I", str(CSRC), "-c", str(s), "-o", str(o)],
             log=BUILD / (s.stem + ".ptxas.log"))

    with ThreadPoolExecutor(max_workers=8) as ex:
        list(ex.map(compile_one, jobs))
    if force or jobs or not _newer(LIB, objs):
        _run([NVCC, "-shared", "-o", str(LIB), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a",
              "-Xcompiler", "-fopenmp", "-lcudart", "-lgomp", "-lnccl"])
    return LIB


def build_oracle(force: bool = False) -> Path:
    srcs = sorted(ORACLE_DIR.glob("*.cpp"))
    hdrs = list(ORACLE_DIR.glob("*.hpp")) + headers()
    if force or not _newer(ORACLE_LIB, srcs + hdrs):
        _run(["g++", "-O3", "-march=x86-64-v3", "-std=c++17", "-fPIC", "-shared", "-fopenmp",
              "-I", str(ROOT / "include"), "-I", str(CSRC), "-I", str(ORACLE_DIR),
              *map(str, srcs), "-o", str(ORACLE_LIB)])
    return ORACLE_LIB


def build_all(force: bool = False):
    return build_lib(force), build_oracle(force)


if __name__ == "__main__":
    libs = build_all("--force" in sys.argv)
    print("\n".join(map(str, libs)))
