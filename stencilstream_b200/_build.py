"""Build recipes for the native parts of StencilStream-B200.

Everything is built in-tree with explicit compiler invocations (nvcc for sm_100a) so that the resulting shared objects travel with the repository snapshot to the GPU box:

  stencilstream_b200/libstst_rt.so                 C-ABI device runtime           (csrc/stst_rt.cu)
  stencilstream_b200/libstst_workloads.so          generation loop + example functors, default flags
  stencilstream_b200/libstst_workloads_strict.so   same, -fmad=false (bit-parity diagnosis build)

(The CPU parity oracles are test infrastructure and are built by oracle/recipes.py.)

A target is rebuilt when it is missing or older than any of its inputs.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "stencilstream_b200"
ORACLE = ROOT / "oracle"
REFERENCE = Path(os.environ.get("STST_REFERENCE_DIR", "/root/reference"))

ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_COMMON = [
    "-std=c++20",
    "-O3",
    "-lineinfo",
    "--expt-relaxed-constexpr",
    "-Xcompiler",
    "-fPIC,-fvisibility=hidden",
    "-shared",
    f"-I{ROOT / 'include'}",
    f"-I{PKG / 'include'}",
    f"-I{PKG / 'compat'}",
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).exists():
        raise RuntimeError("nvcc not found: StencilStream-B200 has no CPU fallback and cannot be built")
    return nvcc


def _newest(paths) -> float:
    newest = 0.0
    for p in paths:
        p = Path(p)
        if p.is_dir():
            for q in p.rglob("*"):
                if q.is_file():
                    newest = max(newest, q.stat().st_mtime)
        elif p.exists():
            newest = max(newest, p.stat().st_mtime)
    return newest


def _stale(target: Path, inputs) -> bool:
    return (not target.exists()) or target.stat().st_mtime < _newest(inputs)


def _run(cmd, verbose: bool) -> None:
    cmd = [str(c) for c in cmd]
    if verbose:
        print("+", " ".join(cmd), file=sys.stderr, flush=True)
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"build step failed ({proc.returncode}): {' '.join(cmd)}\n{proc.stdout}")
    if verbose and proc.stdout.strip():
        print(proc.stdout, file=sys.stderr)


def build_runtime(force: bool = False, verbose: bool = False) -> Path:
    target = PKG / "libstst_rt.so"
    inputs = [PKG / "csrc" / "stst_rt.cu", ROOT / "include" / "stst_rt.h"]
    if force or _stale(target, inputs):
        _run(
            [_nvcc(), "-std=c++20", "-O2", *ARCH_FLAGS, "-lineinfo", "-Xcompiler", "-fPIC", "-shared",
             f"-I{ROOT / 'include'}", inputs[0], "-o", target, "-ldl"],
            verbose,
        )
    return target


# The workloads library: the C entry points (workloads.cu) plus one translation unit per group of
# transition functions, so that nvcc compiles the kernel templates of the groups in parallel.
WORKLOAD_UNITS = ["workloads.cu", "workloads_light.cu", "workloads_hotspot.cu", "workloads_fdtd.cu",
                  "workloads_convection.cu", "workloads_kat.cu"]


def _compile_and_link_workloads(target: Path, extra_flags, verbose: bool) -> None:
    from concurrent.futures import ThreadPoolExecutor

    obj_dir = ROOT / "build" / "obj" / target.stem
    obj_dir.mkdir(parents=True, exist_ok=True)
    compile_flags = [f for f in NVCC_COMMON if f != "-shared"]

    def compile_unit(name):
        obj = obj_dir / (Path(name).stem + ".o")
        _run([_nvcc(), *compile_flags, *ARCH_FLAGS, *extra_flags, f"-I{PKG / 'csrc'}", "-c",
              PKG / "csrc" / name, "-o", obj], verbose)
        return obj

    with ThreadPoolExecutor(max_workers=len(WORKLOAD_UNITS)) as pool:
        objects = list(pool.map(compile_unit, WORKLOAD_UNITS))
    _run([_nvcc(), *ARCH_FLAGS, "-shared", *objects, "-o", target, f"-L{PKG}", "-lstst_rt", "-Xlinker",
          "-rpath,$ORIGIN"], verbose)


def build_workloads(strict: bool = False, force: bool = False, verbose: bool = False) -> Path:
    build_runtime(force=force, verbose=verbose)
    target = PKG / ("libstst_workloads_strict.so" if strict else "libstst_workloads.so")
    inputs = [PKG / "csrc", PKG / "include", PKG / "compat", ROOT / "include"]
    if force or _stale(target, inputs):
        _compile_and_link_workloads(target, ["-fmad=false"] if strict else [], verbose)
    return target


def build_variant(tag: str, extra_flags, verbose: bool = False) -> Path:
    """An experimental build of the workloads library with extra nvcc flags (tuning only):
    stencilstream_b200/libstst_workloads_<tag>.so, selected at run time by STST_WORKLOADS_LIB."""
    build_runtime(verbose=verbose)
    target = PKG / f"libstst_workloads_{tag}.so"
    _compile_and_link_workloads(target, list(extra_flags), verbose)
    return target


def _json_include() -> Path | None:
    """nlohmann/json 3.11.3 (the version the reference pins) ships inside cudnn_frontend's headers."""
    import sysconfig

    candidates = [Path(sysconfig.get_paths()["purelib"]) / "include" / "cudnn_frontend" / "thirdparty"]
    for c in candidates:
        if (c / "nlohmann" / "json.hpp").exists():
            return c
    return None


def reference_available() -> bool:
    return (REFERENCE / "StencilStream" / "cpu" / "StencilUpdate.hpp").exists()


def build_all(force: bool = False, verbose: bool = False) -> dict:
    """The product libraries (runtime + the two workloads builds). The CPU oracles are test
    infrastructure with their own recipe, oracle/recipes.py."""
    from concurrent.futures import ThreadPoolExecutor

    out = {"runtime": build_runtime(force=force, verbose=verbose)}
    with ThreadPoolExecutor(max_workers=2) as pool:
        jobs = {
            "workloads": pool.submit(build_workloads, False, force, verbose),
            "workloads_strict": pool.submit(build_workloads, True, force, verbose),
        }
        for name, job in jobs.items():
            out[name] = job.result()
    return out


if __name__ == "__main__":
    built = build_all(force="--force" in sys.argv, verbose=True)
    for name, path in built.items():
        print(f"{name}: {path}")
