#!/usr/bin/env python
"""Regenerate tests/golden/convection_app/ and tests/golden/fdtd_app/: what the reference's UNMODIFIED
convection and FDTD applications (/root/reference/examples/convection/convection.cpp,
/root/reference/examples/fdtd/src/fdtd.cpp, built for the reference's own cpu backend with g++
-ffp-contract=off against the SYCL type shim) print and write for a small experiment each.

    python tests/golden/generate_app_outputs.py        # needs /root/reference

Convection files: experiment.json (the input), stdout.txt (one line per time step: pseudo-transient iterations,
errV, errP — convection.cpp:447-448), <it>.csv (temperature frames, convection.cpp:460-477).
FDTD files: experiment.json, stdout.txt (derived quantities, fdtd/src/Parameters.hpp), frames.sha256.json
(hz after every snapshot interval and hz_sum at the end, fdtd.cpp:114-166, 233-250).
tests/test_apps_golden.py replays the same experiments through the CPU oracle loop (no GPU) and through
stencilstream_b200.apps.run_convection (GPU) and compares with these files."""
import shutil
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from stencilstream_b200 import _build                      # noqa: E402
from stencilstream_b200.tools import build_examples as B   # noqa: E402

out = ROOT / "build" / "examples"
case = out / "cases" / "convection"
binary = B.build_example("convection", _build.REFERENCE / "examples", out, backend="refcpu")
B.stage_case("convection", case)
B.run_case("convection", case, binary, case / "expected")
target = ROOT / "tests" / "golden" / "convection_app"
target.mkdir(parents=True, exist_ok=True)
shutil.copy(case / "experiment.json", target / "experiment.json")
for path in sorted((case / "expected").iterdir()):
    shutil.copy(path, target / path.name)
print(sorted(p.name for p in target.iterdir()))

# FDTD: the frames are 250 KB each, so only their SHA-256 is kept (plus the configuration dump)
import hashlib  # noqa: E402
import json     # noqa: E402

case = out / "cases" / "fdtd"
binary = B.build_example("fdtd", _build.REFERENCE / "examples", out, backend="refcpu")
B.stage_case("fdtd", case)
B.run_case("fdtd", case, binary, case / "expected")
target = ROOT / "tests" / "golden" / "fdtd_app"
target.mkdir(parents=True, exist_ok=True)
shutil.copy(case / "experiment.json", target / "experiment.json")
lines = [l for l in (case / "expected" / "stdout.txt").read_text().splitlines() if "Walltime" not in l]
(target / "stdout.txt").write_text("\n".join(lines) + "\n")
hashes = {p.name: hashlib.sha256(p.read_bytes()).hexdigest()
          for p in sorted((case / "expected").glob("*.csv"))}
(target / "frames.sha256.json").write_text(json.dumps(hashes, indent=1) + "\n")
print(sorted(p.name for p in target.iterdir()))
