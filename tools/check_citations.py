"""Every `file.py:line[-line]` citation of a reference file in this tree must point inside that file.
Build container only (needs the reference tree).   python tools/check_citations.py [reference root]"""
from __future__ import annotations

import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SKIP = {"SURVEY.md", "VERDICT.md", "ADVICE.md", "PAPERS.md", "SNIPPETS.md", "BASELINE.md"}      # not written by this repo
PAT = re.compile(r"([A-Za-z0-9_/\.]*?([A-Za-z0-9_]+\.(?:py|sh))):(\d+)(?:[-–](\d+))?((?:,\s?:?\d+(?:[-–]\d+)?)*)")


def check(ref_root: str):
    """-> (number of citations checked, [(file, citation, length of the cited file)] that point past its end)"""
    index = collections.defaultdict(list)
    for d, _, fs in os.walk(ref_root):
        for f in fs:
            if f.endswith((".py", ".sh")):
                index[f].append(os.path.join(d, f))
    files = subprocess.run(["git", "ls-files", "*.py", "*.md", "*.h", "*.cu", "*.cuh", "*.c"], capture_output=True, text=True,
                           cwd=ROOT, check=True).stdout.split()
    checked, bad = 0, []
    for f in files:
        if f in SKIP:
            continue
        text = open(os.path.join(ROOT, f), errors="ignore").read()
        for m in PAT.finditer(text):
            path, base, a, b, more = m.groups()
            cands = index.get(base, [])
            if not cands:                      # a file of this repository, not of the reference
                continue
            narrowed = [c for c in cands if c.endswith(path)]
            cands = narrowed or cands
            length = max(len(open(c, errors="ignore").read().split("\n")) for c in cands)
            lines = [int(a)] + ([int(b)] if b else []) + [int(x) for x in re.findall(r"\d+", more or "")]
            checked += 1
            if max(lines) > length:
                bad.append((f, m.group(0), length))
    return checked, bad


if __name__ == "__main__":
    n, bad = check(sys.argv[1] if len(sys.argv) > 1 else os.environ.get("DAS_REFERENCE_ROOT", "/root/reference"))
    print(f"{n} citations checked, {len(bad)} out of range")
    for b in bad:
        print(*b)
    sys.exit(1 if bad else 0)
