"""CPU: the C host driver's command line (the reference's flag letters, parse_args.c:426-957) - argument errors are
reported before any CUDA call, and without a device the program stops loudly instead of falling back to the CPU."""
import os
import subprocess

import pytest

from sts_b200 import binding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STSB200 = os.path.join(ROOT, "sts_b200", "host", "stsb200")


@pytest.fixture(scope="module", autouse=True)
def built():
    binding.build_library()
    assert os.path.exists(STSB200)


def run(args, **kw):
    return subprocess.run([STSB200] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, **kw)


@pytest.mark.parametrize("args,needle", [
    (["-m", "x", "f"], "-m mode"),
    (["-F", "z", "f"], "-F format"),
    (["-t", "16", "f"], "-t test"),
    (["-P", "12=3", "f"], "-P num=value"),
    (["-i", "0", "f"], "-i iterations"),
    (["-m", "a"], "-m a requires -d"),
    (["-g", "10"], "-g generator"),
    (["-S", "992", "f"], "bitcount(n): 992 must >= 1000"),                 # parse_args.c:945-947
    (["-S", "1001", "f"], "bitcount(n): 1001 must be a multiple of 8"),    # parse_args.c:939-944
    (["-P", "9=104", "f"], "bitcount(n): 104 must >= 1000"),
])
def test_argument_errors_exit_1_with_the_flag_named(args, needle):
    r = run(args)
    assert r.returncode == 1
    assert needle in r.stderr


def test_usage_lists_the_reference_flags():
    r = run(["-h"])
    assert r.returncode == 0
    for flag in ("-v", "-t", "-P", "-S", "-i", "-I", "-O", "-s", "-w", "-F", "-j", "-m", "-d", "-T"):
        assert flag in r.stderr
    assert run([]).returncode == 1                       # the data file is required outside -m a


def test_missing_pvalues_directory_is_fatal(tmp_path):
    r = run(["-m", "a", "-d", str(tmp_path / "nowhere")])
    assert r.returncode == 1 and "Could not open the directory" in r.stderr


def test_no_device_is_fatal_not_a_cpu_run(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = tmp_path / "zeros.bin"
    p.write_bytes(b"\x00" * 131072)
    r = run(["-i", "1", "-w", str(tmp_path / "w"), str(p)])
    assert r.returncode == 50                            # driver.c's init range; never a result file
    assert "CUDA" in r.stderr or "device" in r.stderr
    assert not os.path.exists(tmp_path / "w" / "sts.0000.1.1048576.pvalues")
