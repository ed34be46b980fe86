"""The C-ABI libraries load and export exactly what include/*.h declares (no GPU, no compute)."""
import ctypes as C
import re
from pathlib import Path

import pytest

from stencilstream_b200 import _native

ROOT = Path(__file__).resolve().parent.parent
DECL = re.compile(r"^\s*(?:int|const char \*|void)\s*\**\s*(stst_\w+)\s*\(", re.M)


def declared(header):
    text = (ROOT / "include" / header).read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(DECL.findall(text)))


def test_runtime_exports_every_declared_symbol(built):
    names = declared("stst_rt.h")
    assert len(names) >= 40
    lib = _native.runtime_lib()
    for name in names:
        assert hasattr(lib, name), f"libstst_rt.so does not export {name}"
    assert sorted(_native.RT_SYMBOLS) == names
    assert lib.stst_rt_abi_version() == 1


@pytest.mark.parametrize("strict", [False, True])
def test_workloads_exports_every_declared_symbol(built, strict):
    names = declared("stst_workloads.h")
    lib = _native.workloads_lib(strict=strict)
    for name in names:
        assert hasattr(lib, name), f"libstst_workloads does not export {name}"
    assert sorted(_native.WORKLOADS_SYMBOLS) == names
    assert lib.stst_workloads_abi_version() == 3


def test_registry_matches_the_python_mirror(built):
    """Cell and parameter struct sizes seen through the ABI equal the ctypes/numpy mirrors; the traffic
    model is the reference's 2*sizeof(Cell)*n_sub (scripts/benchmark-common.jl:150-151)."""
    from stencilstream_b200 import workload_info, workload_names
    names = workload_names()
    assert set(names) == set(_native.CELL_DTYPES)
    expect = {"conway": (1, 1, 1), "jacobi5": (4, 1, 1), "jacobi_r2": (4, 2, 1), "jacobi_r3": (4, 3, 1),
              "hotspot": (8, 1, 1), "fdtd": (32, 1, 2), "convection_pt": (88, 1, 3),
              "convection_thermal": (88, 1, 2), "kat": (20, 1, 2), "kat_r2": (20, 2, 2)}
    for name in names:
        info = workload_info(name)
        assert info.cell_bytes == _native.CELL_DTYPES[name].itemsize
        assert info.params_bytes == C.sizeof(_native.PARAM_TYPES[name])
        assert info.bytes_per_cell_iteration == 2 * info.cell_bytes * info.n_subiterations
        if name in expect:
            assert (info.cell_bytes, info.stencil_radius, info.n_subiterations) == expect[name]


def test_unknown_workload_and_bad_arguments_are_reported(built):
    from stencilstream_b200 import Grid, Params, StencilUpdate, workload_info
    with pytest.raises(KeyError):
        workload_info("no_such_workload")
    lib = _native.workloads_lib()
    handle = C.c_void_p()
    assert lib.stst_grid_create(b"nope", 4, 4, -1, C.byref(handle)) == -1
    assert b"unknown workload" in lib.stst_workloads_last_error()
    # parameter block of the wrong size -> invalid argument, before anything touches a device
    bad = _native.UpdateParams()
    wrong = _native.HotspotParams()
    bad.transition_function = C.addressof(wrong)
    bad.transition_function_bytes = C.sizeof(wrong)
    assert lib.stst_update_create(b"jacobi5", C.byref(bad), C.byref(handle)) == -2


def test_no_cpu_fallback_without_a_device(built):
    """On a machine without a CUDA device the product path must fail loudly, not compute on the CPU."""
    rt = _native.runtime_lib()
    count = C.c_int(0)
    status = rt.stst_device_count(C.byref(count))
    if status == 0 and count.value > 0:
        pytest.skip("a CUDA device is present")
    from stencilstream_b200 import Grid, StencilStreamError
    with pytest.raises(StencilStreamError):
        Grid("jacobi5", 8, 8)


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under stencilstream_b200/ may import or load it
    (the build recipes in _build.py compile it, which is not using it)."""
    pkg = ROOT / "stencilstream_b200"
    for path in pkg.rglob("*.py"):
        if path.name == "_build.py":
            continue
        text = path.read_text()
        assert "import oracle" not in text and "liboracle" not in text, path


def test_host_copy_workers_copy_exactly(built):
    """stst_host_memcpy (the multi-threaded host copy behind the staged transfer pipeline) is a
    memcpy: odd sizes, sizes below and above the threading threshold, unaligned ends."""
    import numpy as np
    rt = _native.runtime_lib()
    rt.stst_host_memcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    rng = np.random.default_rng(0)
    for n in (0, 1, 4095, (4 << 20) - 1, (4 << 20) + 1, 37_000_003):
        src = rng.integers(0, 256, size=n + 16, dtype=np.uint8)
        dst = np.zeros(n + 16, dtype=np.uint8)
        assert rt.stst_host_memcpy(dst.ctypes.data + 3, src.ctypes.data + 5, n) == 0
        assert dst[3:3 + n].tobytes() == src[5:5 + n].tobytes()
        assert not dst[:3].any() and not dst[3 + n:].any()
    pinned = C.c_int(-1)
    probe = np.zeros(16, dtype=np.uint8)
    assert rt.stst_host_is_pinned(probe.ctypes.data_as(C.c_void_p), C.byref(pinned)) == 0
    assert pinned.value == 0


def test_headers_are_plain_c_and_a_c_client_links_and_fails_loudly_without_a_device(built, tmp_path):
    """tests/cpp/c_abi_client.c, compiled as C11 with -Wall -Wextra -pedantic against include/*.h and
    linked against the two shared libraries: the headers are valid C, the calls link, and without a
    CUDA device the client gets an error code and a message instead of a computed result (with a
    device it runs a small HotSpot update through the C ABI)."""
    import subprocess
    pkg = ROOT / "stencilstream_b200"
    binary = tmp_path / "c_abi_client"
    build = subprocess.run(
        ["gcc", "-std=c11", "-Wall", "-Wextra", "-pedantic", "-Werror", f"-I{ROOT / 'include'}",
         str(ROOT / "tests" / "cpp" / "c_abi_client.c"), "-o", str(binary), f"-L{pkg}",
         "-lstst_workloads", "-lstst_rt", f"-Wl,-rpath,{pkg}"], capture_output=True, text=True)
    assert build.returncode == 0, build.stderr
    run = subprocess.run([str(binary)], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0, run.stdout + run.stderr
    assert run.stdout.startswith("no device: status -4") or run.stdout.startswith("centre temp")


@pytest.mark.gpu
def test_c_client_runs_an_update_through_the_c_abi(built, tmp_path):
    """The same plain-C client on a machine with a CUDA device: grid upload, HotSpot update, download,
    device-side max-norm and statistics, all through the C ABI from C code."""
    import subprocess
    pkg = ROOT / "stencilstream_b200"
    binary = tmp_path / "c_abi_client"
    subprocess.run(["gcc", "-std=c11", f"-I{ROOT / 'include'}", str(ROOT / "tests" / "cpp" / "c_abi_client.c"),
                    "-o", str(binary), f"-L{pkg}", "-lstst_workloads", "-lstst_rt", f"-Wl,-rpath,{pkg}"],
                   check=True)
    run = subprocess.run([str(binary)], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stdout + run.stderr
    assert run.stdout.startswith("centre temp") and "pass-through planes 0x2" in run.stdout
