"""The reference's own `unit_test_cuda` sources (tests/Stencil.cpp, tests/cuda/Grid.cpp,
tests/cuda/StencilUpdate.cpp with GridTest.hpp / StencilUpdateTest.hpp / TransFuncs.hpp), built
unmodified against this backend by stencilstream_b200/tools/build_reference_tests.py, must pass on the
B200. The binary is produced in the build container (where /root/reference exists) and travels with
the repository snapshot."""
import subprocess
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
BINARY = Path(__file__).resolve().parent.parent / "build" / "ref_tests" / "unit_test_cuda_b200"


def test_reference_cuda_unit_tests_pass():
    if not BINARY.exists():
        pytest.skip(f"{BINARY} not built (needs the reference tree at build time)")
    proc = subprocess.run([str(BINARY)], capture_output=True, text=True, timeout=600)
    print(proc.stdout)
    assert proc.returncode == 0, proc.stdout + proc.stderr
    lines = proc.stdout.splitlines()
    passed = [line for line in lines if line.startswith("[ OK ]")]
    # Stencil: 3, cuda::Grid: 4, cuda::StencilUpdate: 2 (normal + split cell structure)
    assert len(passed) >= 8, proc.stdout
    assert any("cuda::StencilUpdate (Split cell structure)" in line for line in passed)
    assert any("cuda::Grid::copy_to_buffer" in line for line in passed)


def test_own_cpp_tests_of_the_grid_extensions_pass():
    """tests/cpp/grid_extensions.cu: Grid::max_abs, single-field copies, copies from/to ordinary host
    memory, through the C++ template API (built by build_reference_tests.build_own)."""
    binary = BINARY.parent.parent / "own_tests" / "unit_test_extensions_b200"
    if not binary.exists():
        pytest.skip(f"{binary} not built")
    proc = subprocess.run([str(binary)], capture_output=True, text=True, timeout=600)
    print(proc.stdout)
    assert proc.returncode == 0, proc.stdout + proc.stderr
    assert sum(line.startswith("[ OK ]") for line in proc.stdout.splitlines()) == 4


def test_own_cpp_tests_of_the_update_extensions_pass():
    """tests/cpp/update_extensions.cu: stencil radii beyond the column group, Cell::constant_fields,
    Params::cuda_devices — through the C++ template API, against a host restatement of the sweep."""
    binary = BINARY.parent.parent / "own_tests" / "unit_test_update_extensions_b200"
    if not binary.exists():
        pytest.skip(f"{binary} not built")
    proc = subprocess.run([str(binary)], capture_output=True, text=True, timeout=600)
    print(proc.stdout)
    assert proc.returncode == 0, proc.stdout + proc.stderr
    assert sum(line.startswith("[ OK ]") for line in proc.stdout.splitlines()) == 5
