"""Every `file.rs:line` citation in the C ABI header, the oracle and the CUDA sources must point inside a file of the reference tree.
Runs only where the reference checkout is present (this container); skipped on the GPU box."""
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
CITE = re.compile(r"([A-Za-z_][A-Za-z0-9_/]*\.(?:rs|ron)):(\d+)(?:-(\d+))?")


def _reference_files():
    out = {}
    for path in glob.glob(os.path.join(REF, "**", "*"), recursive=True):
        if os.path.isfile(path) and path.endswith((".rs", ".ron")):
            with open(path, errors="replace") as f:
                out.setdefault(os.path.basename(path), []).append((path, sum(1 for _ in f)))
    return out


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
def test_cited_lines_exist_in_the_reference():
    ref = _reference_files()
    sources = (glob.glob(os.path.join(ROOT, "include", "*.h")) + glob.glob(os.path.join(ROOT, "oracle", "*.hpp")) + glob.glob(os.path.join(ROOT, "oracle", "*.cpp"))
               + glob.glob(os.path.join(ROOT, "moleengine_b200", "csrc", "*.cu*")) + glob.glob(os.path.join(ROOT, "moleengine_b200", "csrc", "host", "*.hpp"))
               + glob.glob(os.path.join(ROOT, "moleengine_b200", "*.py")) + [os.path.join(ROOT, "INTEGRATION.md")])
    n, bad = 0, []
    for src in sources:
        text = open(src, errors="replace").read()
        for m in CITE.finditer(text):
            name, lo, hi = os.path.basename(m.group(1)), int(m.group(2)), int(m.group(3) or m.group(2))
            if name not in ref:
                bad.append(f"{os.path.relpath(src, ROOT)}: {m.group(0)} (no such file in the reference)")
                continue
            n += 1
            # a file name can occur more than once in the tree (mod.rs, main.rs): the citation must fit at least one of them
            if not any(lo >= 1 and lo <= hi <= lines for _p, lines in ref[name]):
                bad.append(f"{os.path.relpath(src, ROOT)}: {m.group(0)} (file has {max(l for _p, l in ref[name])} lines)")
    assert n > 200, n
    assert not bad, "\n".join(bad[:40])
