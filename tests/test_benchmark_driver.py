"""Host logic of scripts/benchmark.py (the counterpart of the reference's Julia benchmark drivers):
grid-size ladder and roofline model, no GPU."""
import importlib.util
import math
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
spec = importlib.util.spec_from_file_location("stst_benchmark_driver", ROOT / "scripts" / "benchmark.py")
driver = importlib.util.module_from_spec(spec)
spec.loader.exec_module(driver)


def test_max_grid_wh_follows_the_reference_rule():
    # scripts/benchmark-common.jl:186-199 for the A100 (40 GiB): three HotSpot grids in memory,
    # clipped to a power of sqrt(2) / of 2
    assert driver.max_grid_wh(8, 40 * 2 ** 30, clip_to_base=math.sqrt(2)) == 32768
    assert driver.max_grid_wh(8, 40 * 2 ** 30, clip_to_base=2) == 32768
    assert driver.max_grid_wh(8, 40 * 2 ** 30) == int(math.sqrt(40 * 2 ** 30 / 3 / 8))
    # 180 GB of HBM: the 2^31-cell indexing limit binds for 4-byte cells
    assert driver.max_grid_wh(4, 180e9) == int(math.sqrt(2 ** 31))


def test_model_runtime_is_bandwidth_or_launch_bound():
    bw = 6.4e12
    big = driver.model_runtime(16384, 16384, 1000, 4, 1, 1, bw)
    assert math.isclose(big, 1000 * 2 * 4 * 16384 ** 2 / bw)
    fused = driver.model_runtime(16384, 16384, 1000, 4, 1, 6, bw)
    assert math.isclose(fused, math.ceil(1000 / 6) * 2 * 4 * 16384 ** 2 / bw)
    tiny = driver.model_runtime(32, 32, 1000, 4, 1, 1, bw)            # launch-latency bound
    assert math.isclose(tiny, 1000 * driver.SCHEDULING_LATENCY_PER_PASS)


def test_operation_counts_are_the_reference_drivers():
    assert driver.OPERATIONS_PER_CELL == {"jacobi5": 9, "hotspot": 15, "fdtd": 24, "convection_pt": 67}
