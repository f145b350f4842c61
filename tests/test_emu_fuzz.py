"""CPU: a few fixed seeds of the differential fuzzers of tests/emu (the emulated kernels against the oracle on random
stream lengths / parameters / masks / degenerate inputs, the assessment tallies on edge p-values, the -F a packer on
random white space).  The long runs are done by hand (`python tests/emu/fuzz.py 0 400`, ...); this keeps the fuzzers
themselves working and a handful of odd cases in every CPU run."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")


@pytest.mark.parametrize("script,args", [
    ("fuzz.py", ["4", "3"]),                 # seeds 4..6: biased / periodic / LCG streams, n = 2^20 + 8, 4104, 904960
    ("fuzz_assess.py", ["0", "4"]),
    ("fuzz_ascii.py", ["0", "10"]),
])
def test_fuzz_seeds(script, args):
    from emu_util import emu_library
    emu_library()
    r = subprocess.run([sys.executable, os.path.join(EMU, script)] + args, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True)
    assert r.returncode == 0 and "MISMATCH" not in r.stdout and "FAILED" not in r.stdout, r.stdout[-3000:]
