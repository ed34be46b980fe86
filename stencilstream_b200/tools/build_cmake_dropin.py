"""Drop-in check through CMake: configure tests/cmake_dropin (the reference's example directories with
their own, unmodified CMakeLists.txt against cmake/StencilStreamB200.cmake) and build the `*_cuda`
targets the reference's CUDA CI job benchmarks (.gitlab-ci.yml:279-304) plus conway_cuda.

    python -m stencilstream_b200.tools.build_cmake_dropin [--build-dir build/cmake_dropin] [--fmad]
"""
from __future__ import annotations

import argparse
import shutil
import subprocess
import sys
from pathlib import Path

from .. import _build

ROOT = _build.ROOT
TARGETS = {  # CMake target -> (example name for the comparison case, path below the build dir)
    "conway_cuda": ("conway", "examples/conway/conway_cuda"),
    "Jacobi5General_cuda": ("jacobi", "examples/jacobi/Jacobi5General_cuda"),
    "hotspot_cuda": ("hotspot", "examples/hotspot/hotspot_cuda"),
    "fdtd_coef_device_cuda": ("fdtd", "examples/fdtd/fdtd_coef_device_cuda"),
    "convection_cuda": ("convection", "examples/convection/convection_cuda"),
}


def cmake_executable() -> str | None:
    return shutil.which("cmake")


def configure(build_dir: Path, fmad: bool = False, reference: Path | None = None) -> None:
    json_inc = _build._json_include()
    if json_inc is None:
        raise RuntimeError("nlohmann/json headers not found")
    cmd = [cmake_executable(), "-S", str(ROOT / "tests" / "cmake_dropin"), "-B", str(build_dir),
           f"-DSTENCILSTREAM_REFERENCE_DIR={reference or _build.REFERENCE}",
           f"-DSTST_JSON_INCLUDE_DIR={json_inc}", f"-DSTST_B200_FMAD={'ON' if fmad else 'OFF'}",
           f"-DSTST_B200_PYTHON={sys.executable}", "-DCMAKE_BUILD_TYPE=Release",
           "-DCMAKE_EXPORT_COMPILE_COMMANDS=ON"]
    _build._run(cmd, verbose=False)


def build(build_dir: Path, targets=None, jobs: int = 6) -> dict:
    targets = list(targets or TARGETS)
    proc = subprocess.run([cmake_executable(), "--build", str(build_dir), "--target", *targets, "-j",
                           str(jobs)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"cmake --build failed:\n{proc.stdout[-4000:]}")
    return {t: build_dir / TARGETS[t][1] for t in targets}


def main() -> None:
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--build-dir", type=Path, default=ROOT / "build" / "cmake_dropin")
    ap.add_argument("--fmad", action="store_true", help="allow FMA contraction (default: -fmad=false, "
                    "so that outputs are bit-identical to the reference's cpu backend)")
    args = ap.parse_args()
    configure(args.build_dir, fmad=args.fmad)
    for target, path in build(args.build_dir).items():
        print(f"{target}: {path}")


if __name__ == "__main__":
    main()
