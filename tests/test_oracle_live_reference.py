"""Pin the oracle against the LIVE reference where its checkout exists (the build container; skipped on the GPU
box, where only the committed golden vectors of tests/golden/ are available).  Fresh seeded inputs, every stage
of SURVEY.md §8a: the oracle restates the same ATen op sequences, so the results are bit-identical."""
import json
import os
import subprocess
import sys

import pytest

REF = os.environ.get("NVSR_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present on this machine")
def test_oracle_matches_live_reference_bit_exactly():
    res = subprocess.run([sys.executable, os.path.join(HERE, "golden", "check_live_reference.py")], capture_output=True,
                         text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("LIVE_REFERENCE_JSON ")][-1]
    diffs = json.loads(line[len("LIVE_REFERENCE_JSON "):])
    assert len(diffs) >= 30
    bad = {k: v for k, v in diffs.items() if v != 0.0}
    assert not bad, bad
