"""In-tree build of the native extensions (no JIT cache: the .so files travel with the repo snapshot).

    python -m realhf_b200.ops.build            # build everything
    python -m realhf_b200.ops.build --force    # rebuild from scratch

Two artefacts under `realhf_b200/_C/`:
  * `librealhf_b200_ops.so` — every sm_100a kernel (`csrc/*.cu`, compiled by nvcc with
    `-gencode arch=compute_100a,code=sm_100a -lineinfo`) plus the `TORCH_LIBRARY` bindings
    (`csrc/*.cpp`, compiled by g++).  Loaded with `torch.ops.load_library`.
  * `host_ext.<abi>.so` — pybind11 module with the host-side native code (`csrc/host/*.cpp`):
    balanced partitioning, interval merge, and the MCMC allocation search + simulator.
nvcc cross-compiles without a GPU, so this runs on the CPU dev box.
"""

from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

OPS_DIR = Path(__file__).resolve().parent
CSRC = OPS_DIR / "csrc"
OUT_DIR = OPS_DIR.parent / "_C"
BUILD_DIR = OPS_DIR / "build"
CUDA_HOME = Path(os.environ.get("CUDA_HOME", "/usr/local/cuda"))

OPS_LIB = OUT_DIR / "librealhf_b200_ops.so"
HOST_EXT = OUT_DIR / ("host_ext" + sysconfig.get_config_var("EXT_SUFFIX"))

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr", "--expt-extended-lambda", "-Xcompiler", "-fPIC",
    "-Xptxas", "-v", "--use_fast_math",
]
# kernels that must keep IEEE division / exact expf (optimizer, norms) opt out of fast-math
NO_FAST_MATH = {"adam.cu"}


def _cutlass_include() -> list[str]:
    """CuTe/CUTLASS headers vendored with flashinfer (used by kernels that include <cute/...>)."""
    import importlib.util
    for pkg, rel in (("flashinfer", "data/cutlass/include"), ("tilelang", "3rdparty/cutlass/include")):
        spec = importlib.util.find_spec(pkg)
        if spec and spec.submodule_search_locations:
            p = Path(list(spec.submodule_search_locations)[0]) / rel
            if p.exists():
                return ["-I", str(p)]
    return []


def _run(cmd: list[str], log: Path | None = None) -> str:
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if log is not None:
        log.write_text(" ".join(cmd) + "\n" + p.stdout)
    if p.returncode != 0:
        raise RuntimeError(f"build command failed ({p.returncode}):\n{' '.join(cmd)}\n{p.stdout[-6000:]}")
    return p.stdout


def _stamp(src: Path, flags: list[str]) -> str:
    h = hashlib.sha1()
    h.update(src.read_bytes())
    for hdr in sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")):
        h.update(hdr.read_bytes())
    h.update(" ".join(flags).encode())
    return h.hexdigest()


def _compile(src: Path, cmd_prefix: list[str], flags: list[str], force: bool) -> Path:
    obj = BUILD_DIR / (src.stem + ("_cu" if src.suffix == ".cu" else "_cpp") + ".o")
    stamp_file = obj.with_suffix(".stamp")
    stamp = _stamp(src, cmd_prefix + flags)
    if not force and obj.exists() and stamp_file.exists() and stamp_file.read_text() == stamp:
        return obj
    _run(cmd_prefix + flags + ["-c", str(src), "-o", str(obj)], log=obj.with_suffix(".log"))
    stamp_file.write_text(stamp)
    return obj


def _torch_flags():
    import torch
    from torch.utils import cpp_extension as ce
    inc = []
    for p in ce.include_paths():
        inc += ["-isystem", p]
    inc += ["-isystem", str(CUDA_HOME / "include")]
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    lib_dir = str(Path(torch.__file__).parent / "lib")
    return inc, abi, lib_dir


def build_ops(force: bool = False, verbose: bool = False) -> Path:
    BUILD_DIR.mkdir(parents=True, exist_ok=True)
    OUT_DIR.mkdir(parents=True, exist_ok=True)
    inc, abi, lib_dir = _torch_flags()
    cu_srcs = sorted(CSRC.glob("*.cu"))
    cpp_srcs = sorted(CSRC.glob("*.cpp"))
    cutlass = _cutlass_include()
    jobs = []
    for s in cu_srcs:
        flags = [f for f in NVCC_FLAGS if not (f == "--use_fast_math" and s.name in NO_FAST_MATH)]
        jobs.append((s, [str(CUDA_HOME / "bin" / "nvcc")], flags + ["-I", str(CSRC)] + cutlass))
    gxx_flags = ["-O2", "-std=c++17", "-fPIC", f"-D_GLIBCXX_USE_CXX11_ABI={abi}", "-DTORCH_API_INCLUDE_EXTENSION_H",
                 "-Wno-deprecated-declarations"] + inc
    for s in cpp_srcs:
        jobs.append((s, ["g++"], gxx_flags))
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
        objs = list(ex.map(lambda j: _compile(j[0], j[1], j[2], force), jobs))
    newest = max(o.stat().st_mtime for o in objs)
    if force or not OPS_LIB.exists() or OPS_LIB.stat().st_mtime < newest:
        _run(["g++", "-shared", "-o", str(OPS_LIB)] + [str(o) for o in objs] +
             ["-L", lib_dir, "-lc10", "-ltorch_cpu", "-ltorch", "-lc10_cuda", "-ltorch_cuda",
              "-L", str(CUDA_HOME / "lib64"), "-lcudart", f"-Wl,-rpath,{lib_dir}", f"-Wl,-rpath,{CUDA_HOME / 'lib64'}"])
    if verbose:
        for o in objs:
            log = o.with_suffix(".log")
            if log.exists():
                print(log.read_text())
    return OPS_LIB


def build_host(force: bool = False) -> Path:
    import pybind11
    BUILD_DIR.mkdir(parents=True, exist_ok=True)
    OUT_DIR.mkdir(parents=True, exist_ok=True)
    srcs = sorted((CSRC / "host").glob("*.cpp"))
    flags = ["-O3", "-std=c++17", "-fPIC", "-fvisibility=hidden", "-I", pybind11.get_include(),
             "-I", sysconfig.get_paths()["include"]]
    with ThreadPoolExecutor(max_workers=4) as ex:
        objs = list(ex.map(lambda s: _compile(s, ["g++"], flags, force), srcs))
    newest = max(o.stat().st_mtime for o in objs)
    if force or not HOST_EXT.exists() or HOST_EXT.stat().st_mtime < newest:
        _run(["g++", "-shared", "-o", str(HOST_EXT)] + [str(o) for o in objs])
    return HOST_EXT


def sass_summary() -> dict:
    """Count the Blackwell-native SASS mnemonics in the built library (evidence for profiles/)."""
    out = _run([str(CUDA_HOME / "bin" / "cuobjdump"), "-sass", str(OPS_LIB)])
    keys = ["UTCHMMA", "UTCQMMA", "UTCMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "LDGSTS", "SYNCS", "LDGMC", "UTCBAR"]  # LDGMC = multimem.ld_reduce
    # (multimem.st lowers to an ordinary STG on the multicast address and has no mnemonic of its own)
    import re
    pats = {k: re.compile(r"(?<![A-Z0-9_.])" + k + r"(?![A-Z0-9])") for k in keys}  # whole mnemonic: `HMMA` must not count `UTCHMMA`
    return {k: sum(1 for line in out.splitlines() if pats[k].search(line)) for k in keys}


def build_all(force: bool = False, verbose: bool = False):
    h = build_host(force)
    o = build_ops(force, verbose)
    return h, o


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--sass", action="store_true")
    a = ap.parse_args()
    paths = build_all(a.force, a.verbose)
    print("built:", *paths)
    if a.sass:
        print(json.dumps(sass_summary(), indent=1))
