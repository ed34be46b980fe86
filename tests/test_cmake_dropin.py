"""The CMake side of the drop-in boundary (BASELINE.json north_star: the examples recompile "by
swapping only the include and the CMake target").

tests/cmake_dropin/CMakeLists.txt add_subdirectory()s the reference's example directories with their
OWN CMakeLists.txt; cmake/StencilStreamB200.cmake stands in for the reference's root CMakeLists.txt
(reference CMakeLists.txt:30-104) and makes `StencilStream_CUDA` mean "this backend". The CPU test
configures the project and inspects what would be compiled; the build itself happens in
`__graft_entry__.build()` (nvcc cross-compiles without a GPU), and the GPU test runs the CMake-built
binaries against the outputs of the same sources on the reference's cpu backend.
"""
import filecmp
import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from stencilstream_b200 import _build  # noqa: E402
from stencilstream_b200.tools import build_cmake_dropin as D  # noqa: E402
from stencilstream_b200.tools import build_examples as B  # noqa: E402

BUILD_DIR = ROOT / "build" / "cmake_dropin"


def test_reference_examples_configure_unmodified_against_the_b200_target(tmp_path):
    if not _build.reference_available() or D.cmake_executable() is None:
        pytest.skip("needs the reference tree and cmake")
    D.configure(tmp_path / "b", fmad=False)
    commands = json.loads((tmp_path / "b" / "compile_commands.json").read_text())
    by_output = {c.get("output", ""): c for c in commands}
    expect = {
        "hotspot_cuda": ("hotspot.cpp", ["-DHOTSPOT_SPLIT_CELL_STRUCT=1"]),            # hotspot/CMakeLists.txt:19-21
        "Jacobi5General_cuda": ("jacobi.cpp", ["-DJACOBI_KERNEL=Jacobi5General"]),     # jacobi/CMakeLists.txt:16-20
        "fdtd_coef_device_cuda": ("fdtd.cpp", ["-DMATERIAL=0", "-DTDVS_TYPE=1"]),      # fdtd/CMakeLists.txt:18-40
        "convection_cuda": ("convection.cpp", ["-DCONVECTION_SPIT_CELL_STRUCT=1"]),    # convection/CMakeLists.txt:16-18
        "conway_cuda": ("conway.cpp", []),
    }
    for target, (source, macros) in expect.items():
        matches = [c for out, c in by_output.items() if f"/{target}.dir/" in out]
        assert len(matches) == 1, (target, [c["file"] for c in matches])
        cmd = matches[0]["command"]
        assert "nvcc" in cmd.split()[0]
        assert "100a" in cmd and "--expt-relaxed-constexpr" in cmd and "-fmad=false" in cmd
        assert "-DSTENCILSTREAM_BACKEND_CUDA=1" in cmd and "-DSTENCILSTREAM_TARGET_CUDA=1" in cmd
        for macro in macros:
            assert macro in cmd, (target, macro)
        staged = Path(matches[0]["file"])
        assert staged.name == source and "b200_src_" in str(staged) and str(tmp_path) in str(staged)
        # the build-tree copy is the reference source plus STST_HD prefixes, nothing else
        original = next((_build.REFERENCE / "examples").rglob(source))
        assert staged.read_text().replace("STST_HD ", "") == original.read_text()
        assert "STST_HD" in staged.read_text() or source in ("jacobi.cpp", "fdtd.cpp")  # functors in headers
    # targets of other backends are declared by the examples' CMake files but stay plain C++
    cpu = [c for out, c in by_output.items() if "/hotspot_cpu.dir/" in out]
    assert cpu and "nvcc" not in cpu[0]["command"].split()[0]


def _built(target):
    path = BUILD_DIR / D.TARGETS[target][1]
    if not path.exists():
        pytest.skip(f"{path} not built (needs the reference tree and cmake at build time)")
    return path


def _strip_timing(text):
    return "\n".join(line for line in text.splitlines()
                     if not any(w in line for w in ("Walltime", "GFlops", "time", "Time", "seconds",
                                                    "Makespan")))


@pytest.mark.gpu
@pytest.mark.parametrize("target", ["conway_cuda", "hotspot_cuda", "convection_cuda"])
def test_cmake_built_example_equals_reference_cpu_backend(target, tmp_path):
    binary = _built(target)
    name = D.TARGETS[target][0]
    case_dir = ROOT / "build" / "examples" / "cases" / name
    expected = case_dir / "expected"
    if not expected.exists():
        pytest.skip("expected outputs not generated")
    B.run_case(name, case_dir, binary, tmp_path / "out")
    produced = sorted(p.name for p in (tmp_path / "out").iterdir())
    assert produced == sorted(p.name for p in expected.iterdir())
    for file_name in produced:
        got, want = tmp_path / "out" / file_name, expected / file_name
        if file_name == "stdout.txt":
            assert _strip_timing(got.read_text()) == _strip_timing(want.read_text())
        else:
            assert filecmp.cmp(got, want, shallow=False), f"{target}: {file_name} differs"


@pytest.mark.gpu
def test_cmake_built_fdtd_device_tdv_variant_equals_inline_variant(tmp_path):
    """fdtd_coef_device_cuda (TDVS_TYPE=1) — the variant the reference's CI benchmarks — prints what
    the inline-TDV variant printed on the reference's cpu backend (the strategies only differ in where
    the source wave is evaluated)."""
    binary = _built("fdtd_coef_device_cuda")
    case_dir = ROOT / "build" / "examples" / "cases" / "fdtd"
    expected = case_dir / "expected"
    if not expected.exists():
        pytest.skip("expected outputs not generated")
    B.run_case("fdtd", case_dir, binary, tmp_path / "out")
    for want in expected.iterdir():
        got = tmp_path / "out" / want.name
        if want.name == "stdout.txt":
            assert _strip_timing(got.read_text()) == _strip_timing(want.read_text())
        else:
            assert filecmp.cmp(got, want, shallow=False), want.name


@pytest.mark.gpu
def test_cmake_built_jacobi_matches_oracle(tmp_path, oracle_best):
    binary = _built("Jacobi5General_cuda")
    B.run_case("jacobi", ROOT / "build" / "examples" / "cases" / "jacobi", binary, tmp_path / "out")
    got = np.fromfile(tmp_path / "out" / "out.bin", dtype="<f4").reshape(300, 260)
    from stencilstream_b200 import workloads as W
    params = W.jacobi5_params((0.1, 0.2, 0.3, 0.15, 0.25))
    want = oracle_best.run("jacobi5", params, 0.0, W.jacobi_input(300, 260), 0, 37)
    assert got.tobytes() == want.tobytes()
