"""Every `File.java:line` citation in the headers, the oracle and the design documents must point into the reference: the file
exists (by base name, anywhere under the reference's source tree) and has at least that many lines.  The reference is only
present in the build container (/root/reference); elsewhere the test skips."""
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
CITE = re.compile(r"\b([A-Z][A-Za-z]+\.(?:java|xml|md)):(\d+)(?:-(\d+))?")


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference checkout is not on this box")
def test_file_line_citations_point_into_the_reference():
    lengths = {}
    for path in glob.glob(os.path.join(REF, "**", "*.*"), recursive=True):
        if os.path.isfile(path) and path.endswith((".java", ".xml", ".md")):
            with open(path, errors="replace") as f:
                n = sum(1 for _ in f)
            lengths.setdefault(os.path.basename(path), []).append(n)
    docs = [os.path.join(ROOT, d) for d in ("DESIGN.md", "INTEGRATION.md", "README.md")] + \
        glob.glob(os.path.join(ROOT, "include", "*.h")) + [os.path.join(ROOT, "oracle", "raft_oracle.c")] + \
        glob.glob(os.path.join(ROOT, "rafting_b200", "csrc", "*.c*")) + glob.glob(os.path.join(ROOT, "rafting_b200", "csrc", "*.inc"))
    bad, seen = [], 0
    for doc in docs:
        for name, lo, hi in CITE.findall(open(doc, errors="replace").read()):
            if name in ("README.md", "DESIGN.md", "INTEGRATION.md", "SURVEY.md", "BASELINE.md", "VERDICT.md", "ADVICE.md") and name not in lengths:
                continue
            seen += 1
            last = int(hi or lo)
            if name not in lengths:
                bad.append((os.path.basename(doc), f"{name}:{lo}", "no such file in the reference"))
            elif int(lo) < 1 or (hi and int(hi) < int(lo)) or last > max(lengths[name]):
                bad.append((os.path.basename(doc), f"{name}:{lo}" + (f"-{hi}" if hi else ""), f"file has {max(lengths[name])} lines"))
    assert seen > 300, seen
    assert not bad, bad[:20]
