"""Randomised differential test (tools/fuzz_vs_oracle.py): 80 seeded random configurations -- block sizes 4/8/16, ranges
0..16, half-pel, 1-4 references, VBS, fast ME, ParallelMode 1/2, table rate control, tie-heavy input -- CUDA path vs the CPU
oracle, split flags / vectors / levels / reconstruction / frame types bit for bit."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [11, 12])
def test_random_configurations_match_oracle(seed):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz_vs_oracle.py"), "40", str(seed)], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-2000:])
    assert "0 mismatches" in r.stdout
