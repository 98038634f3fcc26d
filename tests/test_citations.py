"""Every `file:line` citation of the reference in this repository (docstrings, headers, kernels, docs) must point
inside the cited file, and the load-bearing ones must point at the code they claim to restate.  Runs only where the
reference checkout exists (/root/reference is not shipped to the GPU box; nothing at run time reads it)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
FILES = {"src/lu.jl": "src/lu.jl", "src/butterflylu.jl": "src/butterflylu.jl", "test/runtests.jl": "test/runtests.jl",
         "runtests.jl": "test/runtests.jl", "butterflylu.jl": "src/butterflylu.jl"}
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "src", "lu.jl")), reason="reference checkout absent")


def ref_lines(rel):
    return open(os.path.join(REF, rel)).read().splitlines()


def citations():
    pat = re.compile(r"(?<![A-Za-z0-9_/])(src/lu\.jl|src/butterflylu\.jl|test/runtests\.jl|runtests\.jl|butterflylu\.jl):([0-9]+(?:-[0-9]+)?(?:, ?:?[0-9]+(?:-[0-9]+)?)*)")
    skip_dirs = {".git", "gpurun_out", "__pycache__", "build", ".pytest_cache", ".hypothesis"}
    for dirpath, dirs, files in os.walk(ROOT):
        dirs[:] = [d for d in dirs if d not in skip_dirs]
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".md", ".jl")) or f in ("SURVEY.md", "VERDICT.md", "ADVICE.md", "PAPERS.md", "SNIPPETS.md"):
                continue
            path = os.path.join(dirpath, f)
            for lineno, line in enumerate(open(path, errors="replace"), 1):
                for m in pat.finditer(line):
                    for rng in re.findall(r"[0-9]+(?:-[0-9]+)?", m.group(2)):
                        a, _, b = rng.partition("-")
                        yield os.path.relpath(path, ROOT), lineno, FILES[m.group(1)], int(a), int(b or a)


def test_every_cited_line_range_exists():
    n = 0
    for where, lineno, rel, a, b in citations():
        total = len(ref_lines(rel))
        assert 1 <= a <= b <= total, f"{where}:{lineno} cites {rel}:{a}-{b} but the file has {total} lines"
        n += 1
    assert n > 300          # the repository leans on these citations; a regex that stops matching must not pass silently


@pytest.mark.parametrize("rel,a,b,needle", [
    ("src/lu.jl", 19, 21, "function lu("), ("src/lu.jl", 67, 83, "function lu!("), ("src/lu.jl", 97, 130, "checknonsingular"),
    ("src/lu.jl", 158, 162, "nsplit"), ("src/lu.jl", 189, 263, "function reckernel!"), ("src/lu.jl", 229, 229, "reckernel!"),
    ("src/lu.jl", 233, 233, "apply_permutation!"), ("src/lu.jl", 235, 235, "ldiv!"), ("src/lu.jl", 240, 240, "schur_complement!"),
    ("src/lu.jl", 246, 246, "apply_permutation!"), ("src/lu.jl", 265, 284, "function schur_complement!"),
    ("src/lu.jl", 290, 338, "function _generic_lufact!"), ("src/lu.jl", 296, 305, "amax"), ("src/lu.jl", 317, 320, "inv("),
    ("src/lu.jl", 164, 188, "function apply_permutation!"), ("src/lu.jl", 27, 32, "NotIPIV"), ("src/lu.jl", 85, 87, "Adjoint"),
    ("src/lu.jl", 148, 154, "m < n"), ("src/butterflylu.jl", 45, 55, "solve!"), ("src/butterflylu.jl", 93, 113, "mul!"),
    ("test/runtests.jl", 14, 31, "function testlu"), ("test/runtests.jl", 59, 64, "check = false"),
])
def test_load_bearing_citations_point_at_the_code_they_name(rel, a, b, needle):
    text = "\n".join(ref_lines(rel)[a - 1:b])
    assert needle.replace(" ", "") in text.replace(" ", ""), f"{rel}:{a}-{b} does not contain {needle!r}"
