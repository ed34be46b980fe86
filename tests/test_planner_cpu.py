"""Tile geometry of the launch planner (cuda/internal/Planner.hpp `shape_for`, shared-memory accounting
of TileKernel.hpp), exercised without a GPU: tests/cpp/planner_shapes.cpp is host-only code compiled with
g++. What is pinned: halo = k * n_sub * radius (the reference's rule for its tiled pipeline,
StencilStream/tiling/internal/StencilUpdateKernel.hpp:79-99), column halo rounded to the access width,
the ping/pong footprint, and that planes which pass through need only one tile buffer."""
import json
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
BINARY = ROOT / "build" / "own_tests" / "planner_shapes"
N_SUB = {"f32": 1, "hotspot": 1, "fdtd": 2, "convection": 3}


@pytest.fixture(scope="module")
def probe():
    BINARY.parent.mkdir(parents=True, exist_ok=True)
    pkg = ROOT / "stencilstream_b200"
    subprocess.run(["g++", "-std=c++20", "-O1", "-w", f"-I{pkg / 'include'}", f"-I{pkg / 'compat'}",
                    f"-I{ROOT / 'include'}", "-I/usr/local/cuda/include",
                    str(ROOT / "tests" / "cpp" / "planner_shapes.cpp"), "-o", str(BINARY)], check=True)

    def run(cell, k, block_x, budget, single=0):
        out = subprocess.run([str(BINARY), cell, str(k), str(block_x), str(budget), hex(single)],
                             check=True, capture_output=True, text=True).stdout
        return json.loads(out)
    return run


@pytest.mark.parametrize("cell,block_x,budget", [("f32", 64, 115712), ("hotspot", 64, 115712),
                                                 ("fdtd", 32, 232448), ("convection", 64, 232448)])
def test_tile_geometry_is_consistent(probe, cell, block_x, budget):
    for k in (1, 2, 3, 6):
        s = probe(cell, k, block_x, budget)
        if not s["feasible"]:
            continue
        assert s["halo"] == k * N_SUB[cell]                       # radius 1
        assert s["hpad"] >= s["halo"] and s["hpad"] % s["cw"] == 0
        assert s["cols"] == block_x * s["cw"]
        assert s["tile_w"] == s["cols"] - 2 * s["hpad"]
        assert s["rows"] == s["tile_h"] + 2 * s["halo"] and s["tile_h"] >= 1
        n_buffers = 2 if k * N_SUB[cell] > 1 else 1
        assert s["smem_bytes"] == s["buffer_bytes"] * n_buffers + 256 <= budget
        assert abs(s["efficiency"] - s["tile_h"] * s["tile_w"] / (s["rows"] * s["cols"])) < 1e-6
        # one more row would not have fitted
        per_row = s["buffer_bytes"] // s["rows"] * n_buffers
        assert s["smem_bytes"] + per_row > budget - 4096


def test_deeper_fusion_shrinks_the_tile_and_can_become_infeasible(probe):
    heights = [probe("hotspot", k, 64, 115712)["tile_h"] for k in (1, 2, 4, 6)]
    assert heights == sorted(heights, reverse=True)
    assert not probe("convection", 8, 64, 232448)["feasible"]     # 88-byte cells, halo 24: no room
    assert not probe("f32", 200, 64, 115712)["feasible"]


def test_pass_through_planes_need_one_buffer(probe):
    """HotSpot's `power` (plane 1) and FDTD's four coefficients (planes 4-7) in one tile buffer only:
    the second buffer shrinks by exactly those planes and the tile grows."""
    plain, spec = probe("hotspot", 4, 64, 115712), probe("hotspot", 4, 64, 115712, 0b10)
    assert spec["second_buffer_bytes"] * 2 == spec["buffer_bytes"]
    assert spec["smem_bytes"] == spec["buffer_bytes"] + spec["second_buffer_bytes"] + 256
    assert spec["tile_h"] > plain["tile_h"] and spec["efficiency"] > plain["efficiency"]
    plain, spec = probe("fdtd", 3, 32, 232448), probe("fdtd", 3, 32, 232448, 0xf0)
    assert spec["second_buffer_bytes"] * 2 == spec["buffer_bytes"]
    assert spec["tile_h"] >= plain["tile_h"] + 8
