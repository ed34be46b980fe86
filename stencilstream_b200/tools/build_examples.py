"""Drop-in check: build the reference's UNMODIFIED example applications against this backend.

    python -m stencilstream_b200.tools.build_examples [--source-root /root/reference/examples]
                                                      [--out build/examples] [name ...]

For every example (conway, jacobi, hotspot, fdtd, convection) the sources are copied from the
reference tree into the build directory, passed through the annotator (tools/annotate.py: functions
taking a `Stencil<...> const &` become `STST_HD`), and compiled by nvcc for sm_100a with

    -DSTENCILSTREAM_BACKEND_CUDA=1 -DSTENCILSTREAM_TARGET_CUDA=1   (what the reference's
                                     StencilStream_CUDA CMake target defines, CMakeLists.txt:46-51)
    -I stencilstream_b200/include -I stencilstream_b200/compat -I include   (instead of the reference's
                                     include root and the SYCL headers)

plus the per-example macros of the reference's example CMake files (JACOBI_KERNEL, MATERIAL,
TDVS_TYPE, *_SPLIT_CELL_STRUCT). Nothing from the reference is stored in this repository; the copies
live under build/ (git-ignored). The resulting binaries are what `tests/test_examples_gpu.py` runs
against the same sources built for the reference's cpu backend.
"""
from __future__ import annotations

import argparse
import shutil
import subprocess
import sys
from pathlib import Path

from .. import _build
from .annotate import annotate_file

ROOT = _build.ROOT

# name -> (main source relative to the example dir, extra macros, needs nlohmann/json)
EXAMPLES = {
    "conway": ("conway.cpp", [], False),
    "jacobi": ("jacobi.cpp", ["-DJACOBI_KERNEL=Jacobi5General"], False),
    "hotspot": ("hotspot.cpp", ["-DHOTSPOT_SPLIT_CELL_STRUCT=1"], False),
    "fdtd": ("src/fdtd.cpp", ["-DMATERIAL=0", "-DTDVS_TYPE=0", "-DFDTD_SPLIT_CELL_STRUCT=1"], True),
    "convection": ("convection.cpp", ["-DCONVECTION_SPIT_CELL_STRUCT=1"], True),
    # the other material resolvers / TDV strategies of examples/fdtd/CMakeLists.txt:14-40: a functor
    # that carries a 16-entry coefficient table by value and indexes it per cell (lut), or searches
    # it by distance (render); cells with an ac_int member (lut: not plane-splittable, stays AoS)
    "fdtd_lut": ("src/fdtd.cpp", ["-DMATERIAL=1", "-DTDVS_TYPE=1", "-DFDTD_SPLIT_CELL_STRUCT=1"], True,
                 "fdtd"),
    "fdtd_render": ("src/fdtd.cpp", ["-DMATERIAL=2", "-DTDVS_TYPE=2", "-DFDTD_SPLIT_CELL_STRUCT=1"],
                    True, "fdtd"),
}


def example_dir_name(name: str) -> str:
    """Directory under the reference's examples/ that holds the sources of example `name`."""
    entry = EXAMPLES[name]
    return entry[3] if len(entry) > 3 else name


def stage_sources(example_dir: Path, out_dir: Path) -> int:
    """Copy *.cpp / *.hpp of an example into out_dir, annotated. Returns #annotated functions."""
    total = 0
    for src in sorted(example_dir.rglob("*")):
        if src.suffix not in (".cpp", ".hpp", ".h"):
            continue
        total += annotate_file(src, out_dir / src.relative_to(example_dir))
    return total


def build_example(name: str, source_root: Path, out_root: Path, backend: str = "b200",
                  verbose: bool = True) -> Path:
    main_rel, macros, needs_json = EXAMPLES[name][:3]
    example_dir = source_root / example_dir_name(name)
    if not example_dir.exists():
        raise FileNotFoundError(f"{example_dir} not found")
    pkg = ROOT / "stencilstream_b200"
    json_inc = _build._json_include()
    if needs_json and json_inc is None:
        raise RuntimeError("nlohmann/json headers not found")
    out_root.mkdir(parents=True, exist_ok=True)
    if backend == "b200":
        stage = out_root / f"{name}_src"
        if stage.exists():
            shutil.rmtree(stage)
        n = stage_sources(example_dir, stage)
        target = out_root / f"{name}_b200"
        _build.build_runtime(verbose=False)
        # -fmad=false: like the reference cpu build (-ffp-contract=off) no a*b+c is contracted, so the
        # two builds of an example print identical results and can be compared byte for byte.
        cmd = [_build._nvcc(), "-std=c++20", "-O3", "-lineinfo", "--expt-relaxed-constexpr",
               "-fmad=false", *_build.ARCH_FLAGS, "-x", "cu", "-w",
               "-DSTENCILSTREAM_BACKEND_CUDA=1", "-DSTENCILSTREAM_TARGET_CUDA=1", *macros,
               f"-I{pkg / 'include'}", f"-I{pkg / 'compat'}", f"-I{ROOT / 'include'}",
               f"-I{stage}", f"-I{(stage / main_rel).parent}"]
        if json_inc:
            cmd.append(f"-I{json_inc}")
        cmd += [str(stage / main_rel), "-o", str(target), f"-L{pkg}", "-lstst_rt",
                "-Xlinker", f"-rpath,{pkg}", "-Xlinker", "-rpath,$ORIGIN/../../stencilstream_b200"]
        if verbose:
            print(f"[{name}] {n} function(s) annotated; nvcc -> {target}", file=sys.stderr)
    else:
        # The same, unannotated sources on the reference's own cpu backend (comparison binary).
        target = out_root / f"{name}_refcpu"
        if name == "jacobi":
            raise RuntimeError("examples/jacobi has no cpu-backend branch (jacobi.cpp:22-35)")
        cmd = ["g++", "-std=c++20", "-O2", "-ffp-contract=off", "-fopenmp", "-w",
               "-DSTENCILSTREAM_BACKEND_CPU=1", *macros, f"-I{pkg / 'compat'}",
               f"-I{source_root.parent}", f"-I{(example_dir / main_rel).parent}"]
        if json_inc:
            cmd.append(f"-I{json_inc}")
        cmd += [str(example_dir / main_rel), "-o", str(target)]
    _build._run(cmd, verbose=False)
    return target


# ---- comparison cases: inputs + what the reference's cpu backend prints for them -----------------------

def case_commands(name: str, case_dir: Path, binary: Path, out_dir: Path):
    """(argv, stdin path or None) that runs example `name` on the staged inputs of `case_dir`."""
    if name == "conway":
        return [str(binary), "64", "64", "200"], case_dir / "input.txt"
    if name == "jacobi":
        return [str(binary), "300", "260", "37", str(out_dir / "out.bin"), "0.1", "0.2", "0.3", "0.15",
                "0.25"], None
    if name == "hotspot":
        return [str(binary), "200", "264", "100", str(case_dir / "temp.bin"),
                str(case_dir / "power.bin"), str(out_dir / "out.bin")], None
    if name.startswith("fdtd"):
        return [str(binary), "-c", str(case_dir / "experiment.json"), "-o", str(out_dir)], None
    if name == "convection":
        return [str(binary), str(case_dir / "experiment.json"), str(out_dir)], None
    raise KeyError(name)


def stage_case(name: str, case_dir: Path) -> None:
    """Write the synthetic inputs of the comparison case (nothing is taken from the reference)."""
    import json

    import numpy as np

    from .. import workloads as W

    case_dir.mkdir(parents=True, exist_ok=True)
    if name == "conway":
        soup = W.conway_soup(64, 64, seed=7, density=0.15)
        (case_dir / "input.txt").write_text(
            "\n".join("".join("X" if v else "." for v in row) for row in soup) + "\n")
    elif name == "hotspot":
        cells = W.hotspot_input(200, 264)
        rng = np.random.default_rng(3)
        cells["temp"] += rng.random(cells.shape).astype(np.float32) * 40
        cells["temp"].astype("<f4").tofile(case_dir / "temp.bin")
        cells["power"].astype("<f4").tofile(case_dir / "power.bin")
    elif name.startswith("fdtd"):
        cfg = json.loads(json.dumps(W.FDTD_DEFAULT))
        cfg["time"] = {"t_cutoff": 7.0, "t_detect": 0.05, "t_max": 0.25, "t_snap": 0.1}
        (case_dir / "experiment.json").write_text(json.dumps(cfg, indent=1))
    elif name == "convection":
        cfg = W.convection_benchmark_config(res=48, n_iters=30)
        cfg.update({"nt": 3, "nout": 1, "nerr": 10, "iterMax": 30})
        (case_dir / "experiment.json").write_text(json.dumps(cfg, indent=1))


def run_case(name: str, case_dir: Path, binary: Path, out_dir: Path, timeout: float = 600.0) -> None:
    """Run one example binary on the staged case; stdout goes to out_dir/stdout.txt."""
    if out_dir.exists():
        shutil.rmtree(out_dir)
    out_dir.mkdir(parents=True)
    argv, stdin_path = case_commands(name, case_dir, binary, out_dir)
    with open(out_dir / "stdout.txt", "w") as out:
        stdin = open(stdin_path) if stdin_path else subprocess.DEVNULL
        try:
            subprocess.run(argv, stdin=stdin, stdout=out, stderr=subprocess.STDOUT, check=True,
                           timeout=timeout)
        finally:
            if stdin_path:
                stdin.close()


def main() -> None:
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("names", nargs="*", default=list(EXAMPLES))
    ap.add_argument("--source-root", type=Path, default=_build.REFERENCE / "examples")
    ap.add_argument("--out", type=Path, default=ROOT / "build" / "examples")
    ap.add_argument("--with-reference-cpu", action="store_true",
                    help="also build the examples against the reference's cpu backend (g++)")
    args = ap.parse_args()
    for name in args.names or list(EXAMPLES):
        print(build_example(name, args.source_root, args.out))
        case_dir = args.out / "cases" / name
        stage_case(name, case_dir)
        if args.with_reference_cpu and name != "jacobi":
            binary = build_example(name, args.source_root, args.out, backend="refcpu")
            print(binary)
            run_case(name, case_dir, binary, case_dir / "expected")
            print(f"[{name}] expected outputs: {sorted(p.name for p in (case_dir / 'expected').iterdir())}")


if __name__ == "__main__":
    main()
