"""CPU: the C-ABI library exists, loads without a GPU, exports every symbol include/stsgpu.h declares,
and fails loudly (no CPU fallback) when asked to compute without a device."""
import ctypes
import os
import re

import pytest

import sts_b200
from sts_b200 import binding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "stsgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(stsgpu_[a-z_0-9]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    binding.build_library()
    lib = ctypes.CDLL(binding.library_path())
    declared = header_functions()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), "libstsgpu.so does not export %s" % name
    assert sorted(binding.EXPORTS) == declared, "binding.EXPORTS out of sync with include/stsgpu.h"
    assert lib.stsgpu_abi_version() == 2


def test_struct_sizes_match_header(tmp_path):
    """the ctypes mirrors against what a C compiler makes of include/stsgpu.h"""
    import subprocess
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "stsgpu.h"\nint main(void) { printf("%zu %zu %zu %zu %zu\\n", '
                   'sizeof(stsgpu_params), sizeof(stsgpu_layout), sizeof(stsgpu_metric), offsetof(stsgpu_params, flags), '
                   'offsetof(stsgpu_layout, fs_off)); return 0; }\n')
    exe = str(tmp_path / "sz")
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", exe])
    sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    assert sizes == [ctypes.sizeof(binding.Params), ctypes.sizeof(binding.Layout), ctypes.sizeof(binding.Metric),
                     binding.Params.flags.offset, binding.Layout.fs_off.offset]
    assert sizes[0] == 72 and sizes[1] == 4 * (6 + 64 + 1 + 32)


def test_default_params_are_the_reference_defaults():
    p = sts_b200.default_params()
    assert (p.n, p.block_frequency_M, p.non_overlapping_m, p.overlapping_m, p.apen_m, p.serial_m,
            p.linear_complexity_M) == (1048576, 16384, 9, 9, 10, 16, 500)      # defs.h:110-125
    assert p.test_mask == 0xFFFE


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(sts_b200.EngineError) as ei:
        sts_b200.Engine(max_streams=1)
    assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)


def test_product_never_imports_the_oracle():
    """only tests/, __graft_entry__.smoke and bench.py's baseline legs may touch oracle/"""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "sts_b200")):
        for f in files:
            if f.endswith((".py", ".c", ".h", ".cu", ".cuh", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.replace("the oracle", "").replace("oracle_layout_for", "") or f == "__init__.py", \
                    "%s mentions the oracle" % os.path.join(dirpath, f)


def test_dft_supported_lengths():
    """host-side FFT planning (no CUDA call): n/2 = two 5-smooth factors takes the four-step kernels, every other even n
    up to 2^25 the Bluestein path (k_dft_any.cu); beyond that only the 5-smooth splits, and a longer stream is refused"""
    import sts_b200 as sts
    for n in (1 << 10, 1 << 16, 1 << 20, 1 << 23, 1 << 25, 10 ** 6, 10 ** 7, 400000, 3 * (1 << 18)):
        assert sts.dft_supported(n), n
    for n in (1000008, 2 * 1000003, 14 * (1 << 16), 8072, (1 << 25) - 8):      # not 5-smooth: Bluestein, M <= 2^25
        assert sts.dft_supported(n), n
    for n in (1 << 27, (1 << 25) + 8, 7 * (1 << 23)):
        assert not sts.dft_supported(n), n


def test_product_binding_refuses_the_emulation_library():
    """the CPU emulation of tests/emu is test infrastructure: pointing the product's binding at it (STSGPU_LIB) is an
    error, not a way to run without a GPU"""
    import subprocess
    import sys
    from emu_util import emu_library
    code = ("import sys; sys.path.insert(0, %r)\n"
            "import sts_b200\n"
            "try:\n"
            "    sts_b200.load_library()\n"
            "    print('LOADED')\n"
            "except sts_b200.EngineError as e:\n"
            "    print('REFUSED', e)\n") % ROOT
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, STSGPU_LIB=emu_library()), stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True)
    assert "REFUSED" in r.stdout and "never a fallback" in r.stdout, r.stdout
