"""CPU tests of small host-side tools: NUMA placement (stencilstream_b200/affinity.py), the source
annotator's tree mode (what cmake/StencilStreamB200.cmake runs), the SASS histogram tool."""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_cpulist_parsing_and_binding_without_topology_information():
    from stencilstream_b200 import affinity

    assert affinity._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert affinity._parse_cpulist("") == set()
    before = os.sched_getaffinity(0)
    # no GPU / no NVML here: nothing is known about the device's NUMA node, nothing may change
    report = affinity.bind_to_gpu_numa_node(0)
    assert report["bound"] is False and set(report["cpus"]) == before
    assert os.sched_getaffinity(0) == before


def test_binding_restricts_to_the_node_and_shares_it_between_ranks(monkeypatch):
    from stencilstream_b200 import affinity

    before = sorted(os.sched_getaffinity(0))
    if len(before) < 2:
        return
    node_cpus = set(before[:max(2, len(before) // 2)]) | {10_000}   # a core this process may not use
    monkeypatch.setattr(affinity, "gpu_numa_cpus", lambda device: (1, node_cpus))
    try:
        report = affinity.bind_to_gpu_numa_node(3)
        assert report == {"numa_node": 1, "bound": True, "cpus": sorted(node_cpus & set(before))}
        assert os.sched_getaffinity(0) == node_cpus & set(before)
        os.sched_setaffinity(0, before)
        report = affinity.bind_to_gpu_numa_node(3, rank_on_node=1, ranks_on_node=2)
        usable = sorted(node_cpus & set(before))
        share = len(usable) // 2
        assert report["cpus"] == usable[share:2 * share]
    finally:
        os.sched_setaffinity(0, before)


def test_annotator_tree_mode_only_adds_the_prefix(tmp_path):
    src = tmp_path / "src"
    (src / "sub").mkdir(parents=True)
    (src / "kernel.hpp").write_text(
        "struct K : public BaseTransitionFunction {\n"
        "    using Cell = float;\n"
        "    Cell operator()(Stencil<float, 1> const &stencil) const { return helper(stencil); }\n"
        "    float helper(Stencil<float, 1> const &s) const { return s[0][0]; }\n"
        "    static float halo() { return 0.0f; }\n"
        "    K(int argc, char **argv) {}\n"
        "};\n")
    (src / "sub" / "main.cpp").write_text('#include "../kernel.hpp"\nint main() { return 0; }\n')
    (src / "notes.txt").write_text("not a source file\n")
    out = tmp_path / "out"
    proc = subprocess.run([sys.executable, "-m", "stencilstream_b200.tools.annotate", "--tree", str(src),
                           "-o", str(out), "--also=halo"], cwd=ROOT, capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr
    staged = (out / "kernel.hpp").read_text()
    assert staged.count("STST_HD") == 3                                   # operator(), helper, halo
    assert staged.replace("STST_HD ", "") == (src / "kernel.hpp").read_text()
    assert "STST_HD K(int" not in staged                                 # constructors stay host-only
    assert (out / "sub" / "main.cpp").read_text() == (src / "sub" / "main.cpp").read_text()
    assert not (out / "notes.txt").exists()
    # unchanged files are not rewritten (build-tree timestamps stay put)
    stamp = (out / "kernel.hpp").stat().st_mtime_ns
    subprocess.run([sys.executable, "-m", "stencilstream_b200.tools.annotate", "--tree", str(src), "-o",
                    str(out), "--also=halo"], cwd=ROOT, check=True, capture_output=True)
    assert (out / "kernel.hpp").stat().st_mtime_ns == stamp


def test_sass_histogram_shows_tma_and_vector_stores(built):
    """The shipped Jacobi kernel, read back from the library: TMA box loads, mbarrier waits and
    128-bit stores are in the SASS (what profiles/r02_sass_*.txt records)."""
    proc = subprocess.run([sys.executable, str(ROOT / "scripts" / "sass_histogram.py"), "--match",
                           "Jacobi5Rule"], capture_output=True, text=True, timeout=300)
    assert proc.returncode == 0, proc.stderr
    assert "UTMALDG.2D" in proc.stdout and "SYNCS" in proc.stdout and "STG.E.128" in proc.stdout
    assert "fp32" in proc.stdout and "FFMA" in proc.stdout
