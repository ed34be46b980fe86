"""The convection application loop against what the reference's own, unmodified `main` prints and
writes (tests/golden/convection_app/, produced by tests/golden/generate_app_outputs.py from
/root/reference/examples/convection/convection.cpp on the reference's cpu backend):

* without a GPU: the oracle-driven loop of tests/cases.py (`oracle_convection`, the checker of the GPU
  application tests) — this pins the checker itself to the reference program;
* on the GPU: `stencilstream_b200.apps.run_convection` in the -fmad=false build.

Compared: per time step the number of pseudo-transient iterations and errV / errP as the reference
prints them (`%1.3e`, convection.cpp:447-448), and every temperature frame as the reference writes it
(default ostream formatting = `%g`, convection.cpp:460-477)."""
import json
import re
from pathlib import Path

import numpy as np
import pytest

import cases

GOLDEN = Path(__file__).resolve().parent / "golden" / "convection_app"
LINE = re.compile(r"it = (\d+) \(iter = (\d+), time = [^)]*\), errV=(\S+), errP=(\S+)")


def expected_steps():
    return [(int(m[1]), int(m[2]), m[3], m[4])
            for m in LINE.finditer((GOLDEN / "stdout.txt").read_text())]


def as_printed(steps):
    return [(it, iterations, "%1.3e" % errV, "%1.3e" % errP) for it, iterations, errV, errP in steps]


def frame_text(T):
    return "".join(",".join("%g" % v for v in row) + "\n" for row in T)


def config():
    return json.loads((GOLDEN / "experiment.json").read_text())


def test_oracle_loop_reproduces_the_reference_application(oracle_port):
    _, steps, frames = cases.oracle_convection(oracle_port, config())
    assert as_printed([s[:4] for s in steps]) == expected_steps()
    assert len(frames) == 3
    for it, T in frames:
        assert frame_text(T) == (GOLDEN / f"{it}.csv").read_text(), f"frame {it}"


@pytest.mark.gpu
def test_gpu_application_reproduces_the_reference_application():
    from stencilstream_b200.apps import run_convection
    frames = []
    _, steps = run_convection(config(), strict=True,
                              on_frame=lambda it, T: frames.append((it, T.copy())))
    assert as_printed([(s.it, s.iterations, s.errV, s.errP) for s in steps]) == expected_steps()
    for it, T in frames:
        assert frame_text(T) == (GOLDEN / f"{it}.csv").read_text(), f"frame {it}"


# ---- FDTD ---------------------------------------------------------------------------------------------

import hashlib  # noqa: E402

FDTD_GOLDEN = Path(__file__).resolve().parent / "golden" / "fdtd_app"


def fdtd_frame_hash(values):
    """SHA-256 of a frame as examples/fdtd/src/fdtd.cpp:137-166 writes it (no newline after the last row)."""
    text = "\n".join(",".join("%g" % float(v) for v in row) for row in values)
    return hashlib.sha256(text.encode()).hexdigest()


def fdtd_setup():
    from stencilstream_b200 import workloads as W
    cfg = json.loads((FDTD_GOLDEN / "experiment.json").read_text())
    return cfg, W.FdtdExperiment(cfg), json.loads((FDTD_GOLDEN / "frames.sha256.json").read_text())


def test_fdtd_experiment_setup_matches_what_the_reference_prints():
    """grid size, number of time steps and snapshot interval as the reference's Parameters class
    derives and prints them (examples/fdtd/src/Parameters.hpp:224-262)."""
    _, exp, hashes = fdtd_setup()
    printed = (FDTD_GOLDEN / "stdout.txt").read_text()
    assert f"grid w/h          = {exp.grid_wh()} cells" in printed
    assert f"n. timesteps      = {exp.n_timesteps()}" in printed
    assert f"n. snap timesteps = {exp.n_snap_timesteps()}" in printed
    assert "dt                = %g s/iteration" % float(exp.dt()) in printed
    snap, total = exp.n_snap_timesteps(), exp.n_timesteps()
    assert sorted(hashes) == sorted([f"hz.{snap * i}.csv" for i in range(1, -(-total // snap) + 1)]
                                    + [f"hz_sum.{total}.csv"])


def test_oracle_snapshot_loop_reproduces_the_reference_fdtd_frames(oracle_port):
    _, exp, hashes = fdtd_setup()
    cells = exp.initial_grid()
    snap, total = exp.n_snap_timesteps(), exp.n_timesteps()
    done = 0
    while done < total:
        cells = oracle_port.run("fdtd", exp.kernel_params(), None, cells, done, snap)
        done += snap
        assert fdtd_frame_hash(cells["hz"]) == hashes[f"hz.{done}.csv"], f"hz after {done} steps"
    assert fdtd_frame_hash(cells["hz_sum"]) == hashes[f"hz_sum.{total}.csv"]


@pytest.mark.gpu
def test_gpu_fdtd_application_reproduces_the_reference_frames():
    from stencilstream_b200.apps import run_fdtd
    cfg, exp, hashes = fdtd_setup()
    got = {}
    run_fdtd(cfg, strict=True,
             on_frame=lambda field, it, values: got.__setitem__(f"{field}.{it}.csv",
                                                                 fdtd_frame_hash(values)))
    assert got == hashes
