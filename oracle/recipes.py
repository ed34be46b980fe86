"""TEST INFRASTRUCTURE — build recipes of the CPU parity oracles (never part of the product path).

  oracle/liboracle_port.so      plain-C restatement of the reference algorithm (oracle/stencil_oracle.c)
  oracle/_ref/liboracle_ref.so  the reference's own cpu backend + example sources, compiled IN PLACE from
                                /root/reference by g++ (no reference source is copied; built only where
                                that tree exists — the prebuilt library travels to the GPU box)
  oracle/liboracle_port_fma.so, oracle/_ref/liboracle_ref_fma.so
                                the same two, compiled with `-ffp-contract=fast -mfma`: g++ then contracts
                                a*b+c into fused multiply-adds like nvcc does by default (and like the
                                reference's own icpx build of its cuda backend, whose default FP model
                                contracts too). The checker for the default (-fmad=true) GPU build.
  oracle/_ref/hotspot_openmp    the reference's independent Rodinia OpenMP HotSpot
                                (examples/hotspot/hotspot_openmp.cpp), second CPU baseline for HotSpot

    python -m oracle.recipes [--force]
"""
from __future__ import annotations

import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from stencilstream_b200._build import (PKG, REFERENCE, _json_include, _run, _stale,  # noqa: E402
                                       reference_available)

ORACLE = ROOT / "oracle"


FP_FLAGS = {False: ["-ffp-contract=off"], True: ["-ffp-contract=fast", "-mfma"]}


def build_oracle_port(force: bool = False, verbose: bool = False, fma: bool = False) -> Path:
    target = ORACLE / ("liboracle_port_fma.so" if fma else "liboracle_port.so")
    inputs = [ORACLE / "stencil_oracle.c", ROOT / "include" / "stst_workloads.h"]
    if force or _stale(target, inputs):
        _run(
            ["gcc", "-std=c11", "-O2", *FP_FLAGS[fma], "-fopenmp", "-fPIC", "-shared",
             f"-I{ROOT / 'include'}", inputs[0], "-o", target, "-lm"],
            verbose,
        )
    return target


_REF_GLOBALS = ["exception_handler", "description", "usage", "write_output", "read_input",
                "save_frame"]


def build_oracle_ref(force: bool = False, verbose: bool = False, fma: bool = False) -> Path | None:
    """Compile the reference's own cpu backend and example functors, in place, into oracle/_ref."""
    target = ORACLE / "_ref" / ("liboracle_ref_fma.so" if fma else "liboracle_ref.so")
    if not reference_available():
        return target if target.exists() else None
    sources = sorted((ORACLE / "ref_harness").glob("*.cpp"))
    inputs = [*sources, ORACLE / "ref_harness", PKG / "compat", ROOT / "include" / "stst_workloads.h"]
    if force or _stale(target, inputs):
        target.parent.mkdir(parents=True, exist_ok=True)
        json_inc = _json_include()
        flags = ["-std=c++20", "-O2", *FP_FLAGS[fma], "-fopenmp", "-fPIC", "-w",
                 "-fvisibility=hidden",
                 "-DSTENCILSTREAM_BACKEND_CPU=1", f"-DSTST_REFERENCE_DIR=\"{REFERENCE}\"",
                 f"-I{PKG / 'compat'}", f"-I{REFERENCE}", f"-I{ROOT / 'include'}",
                 f"-I{ORACLE / 'ref_harness'}"]
        if json_inc is not None:
            flags.append(f"-I{json_inc}")
        objects = []
        for src in sources:
            obj = target.parent / (src.stem + ("_fma.o" if fma else ".o"))
            # The example sources define same-named globals (`exception_handler`, `description`,
            # ...): give the known ones a per-translation-unit name.
            renames = [f"-D{name}={name}_{src.stem}" for name in _REF_GLOBALS]
            _run(["g++", *flags, *renames, "-c", src, "-o", obj], verbose)
            objects.append(obj)
        _run(["g++", "-shared", "-fopenmp", *objects, "-o", target], verbose)
    return target




def build_hotspot_openmp(force: bool = False, verbose: bool = False) -> Path | None:
    """The reference's Rodinia OpenMP HotSpot (examples/hotspot/hotspot_openmp.cpp), compiled in place:
    a second, independent CPU implementation of the HotSpot path."""
    target = ORACLE / "_ref" / "hotspot_openmp"
    source = REFERENCE / "examples" / "hotspot" / "hotspot_openmp.cpp"
    if not source.exists():
        return target if target.exists() else None
    if force or _stale(target, [source]):
        target.parent.mkdir(parents=True, exist_ok=True)
        _run(["g++", "-O2", "-ffp-contract=off", "-fopenmp", "-w", source, "-o", target], verbose)
    return target


def build_all(force: bool = False, verbose: bool = False) -> dict:
    from concurrent.futures import ThreadPoolExecutor

    with ThreadPoolExecutor(max_workers=5) as pool:
        jobs = {
            "oracle_port": pool.submit(build_oracle_port, force, verbose),
            "oracle_ref": pool.submit(build_oracle_ref, force, verbose),
            "oracle_port_fma": pool.submit(build_oracle_port, force, verbose, True),
            "oracle_ref_fma": pool.submit(build_oracle_ref, force, verbose, True),
            "hotspot_openmp": pool.submit(build_hotspot_openmp, force, verbose),
        }
        return {name: job.result() for name, job in jobs.items()}


if __name__ == "__main__":
    for name, path in build_all(force="--force" in sys.argv, verbose=True).items():
        print(f"{name}: {path}")
