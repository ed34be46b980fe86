"""Build-time source annotator: make a StencilStream user's functors callable from nvcc device code.

nvcc has no "device by default" mode, and the reference's example functors carry no
`__host__ __device__`. The one mechanical rule this tool applies to a copy of
the user's sources in the build tree:

    every function whose first parameter is a `Stencil<...> const &` gets the prefix `STST_HD`
    (= `__host__ __device__` under nvcc, nothing under a host compiler),

which covers the transition function's `operator()` and helpers that receive the stencil (e.g. the
FDTD material resolvers' `get_material_coefficients`). Helpers that do not take the stencil but are
called from device code have to be named (`--also halo`). Nothing else is touched; constructors and
`get_time_dependent_value` stay host-only (the latter is evaluated on the host, reference
StencilStream/cuda/StencilUpdate.hpp:224). The sources under version control remain unmodified —
include path and CMake target stay the only hand-made differences.

    python -m stencilstream_b200.tools.annotate SRC [SRC ...] -o OUT_DIR
    python -m stencilstream_b200.tools.annotate --tree SRC_DIR -o OUT_DIR    (every *.cpp/*.hpp/*.h
                                            below SRC_DIR, same relative layout; what the CMake
                                            module cmake/StencilStreamB200.cmake runs)
"""
from __future__ import annotations

import argparse
import re
from pathlib import Path

# `<indent><return type and qualifiers> name(Stencil<` — the declarator starts a line in the
# reference's clang-format style; the return type may not contain ';', '{', '}' or '('.
_DECL = re.compile(
    r"^(?P<indent>[ \t]*)(?P<head>(?!return\b|else\b|STST_HD\b)[A-Za-z_:][^;{}()\n]*?[\s&*>])"
    r"(?P<name>operator\s*\(\s*\)|[A-Za-z_]\w*)\s*\(\s*(?:const\s+)?(?:stencil::)?Stencil\s*<",
    re.M)


def _also_pattern(names):
    alternatives = "|".join(re.escape(n) for n in names)
    return re.compile(
        r"^(?P<indent>[ \t]*)(?P<head>(?!return\b|else\b|STST_HD\b)[A-Za-z_:][^;{}()\n]*?[\s&*>])"
        r"(?P<name>" + alternatives + r")\s*\((?P<args>[^;{}]*?)\)\s*(?:const\s*)?\{", re.M)


def annotate_text(text: str, also=()) -> tuple[str, int]:
    """Returns the annotated text and the number of functions that were prefixed. `also`: names of
    further functions (defined in the text) that device code calls, e.g. a cell's `halo()` factory."""
    count = 0
    if also:
        def also_repl(m: re.Match) -> str:
            nonlocal count
            count += 1
            return f"{m.group('indent')}STST_HD {m.group(0)[len(m.group('indent')):]}"
        text = _also_pattern(also).sub(also_repl, text)

    def repl(m: re.Match) -> str:
        nonlocal count
        count += 1
        return f"{m.group('indent')}STST_HD {m.group('head')}{m.group('name')}{m.group(0)[m.end('name') - m.start(0):]}"

    return _DECL.sub(repl, text), count


def annotate_file(src: Path, dst: Path, also=()) -> int:
    text, count = annotate_text(src.read_text(), also)
    dst.parent.mkdir(parents=True, exist_ok=True)
    dst.write_text(text)
    return count


def annotate_tree(src_dir: Path, out_dir: Path, also=()) -> int:
    """Annotated copies of every C++ source/header below `src_dir`, same relative layout. Files whose
    annotated text is unchanged are not rewritten (keeps build-tree timestamps stable)."""
    total = 0
    for src in sorted(src_dir.rglob("*")):
        if src.suffix not in (".cpp", ".hpp", ".h") or not src.is_file():
            continue
        dst = out_dir / src.relative_to(src_dir)
        text, count = annotate_text(src.read_text(), also)
        total += count
        if not dst.exists() or dst.read_text() != text:
            dst.parent.mkdir(parents=True, exist_ok=True)
            dst.write_text(text)
    return total


def main() -> None:
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("sources", nargs="*", type=Path)
    ap.add_argument("--tree", type=Path, help="annotate every C++ file below this directory")
    ap.add_argument("-o", "--out-dir", type=Path, required=True)
    ap.add_argument("--also", default="", help="comma-separated names of further functions to annotate")
    args = ap.parse_args()
    also = tuple(n for n in args.also.split(",") if n)
    if args.tree:
        n = annotate_tree(args.tree, args.out_dir, also)
        print(f"{args.tree} -> {args.out_dir}: {n} function(s) annotated")
    for src in args.sources:
        n = annotate_file(src, args.out_dir / src.name, also)
        print(f"{src} -> {args.out_dir / src.name}: {n} function(s) annotated")


if __name__ == "__main__":
    main()
