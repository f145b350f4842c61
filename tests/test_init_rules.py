"""CPU: the X_init() self-disable rules (SURVEY 8a "init" column, 8b "X_init keeps every self-disable check").

tests/golden/init_rules.json holds, for 50 stream lengths / parameter sets on both sides of every threshold, which
tests the UNMODIFIED reference binary left enabled and how many p-values per iteration each recorded (read from its
own .pvalues file by tests/golden/make_init_golden.py).  Both the engine's device-free stsgpu_query_layout() - the
same code stsgpu_create() runs - and the oracle's layout must agree with it, and with each other on the full record
layout.  Cases the reference itself refuses (n < 1000 at the command line; -P 2=15, where it dies inside
grow_dyn_array) carry no parity target and are only checked for engine == oracle."""
import json
import os

import pytest

import sts_b200
from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = {1: "block_frequency_M", 2: "non_overlapping_m", 3: "overlapping_m", 4: "apen_m", 5: "serial_m",
         6: "linear_complexity_M"}
CASES = json.load(open(os.path.join(HERE, "golden", "init_rules.json")))


def _id(rec):
    return "n%d" % rec["n"] + "".join("_P%s=%d" % kv for kv in sorted(rec["P"].items()))


@pytest.fixture(scope="module", autouse=True)
def built():
    sts_b200.build_library()
    O.build()


@pytest.mark.parametrize("rec", CASES, ids=_id)
def test_enabled_tests_and_pvalue_counts_match_the_reference(rec):
    kw = {NAMES[int(k)]: v for k, v in rec["P"].items()}
    eng = sts_b200.query_layout(n=rec["n"], **kw)
    orc = O.layout(O.make_params(n=rec["n"], **kw))
    for f in ("enabled_mask", "npv", "nst", "nw", "ntemplates"):
        assert getattr(eng, f) == getattr(orc, f), f
    for f in ("pv_off", "pv_cnt", "st_off", "st_cnt"):
        assert list(getattr(eng, f)) == list(getattr(orc, f)), f
    if (eng.enabled_mask >> 10) & 1:
        assert eng.universal_L == orc.universal_L
    if rec.get("refused"):
        return
    assert eng.enabled_mask == rec["enabled_mask"], "engine %#x, reference %#x" % (eng.enabled_mask, rec["enabled_mask"])
    assert list(eng.pv_cnt) == rec["pv_cnt"]
    assert eng.npv == sum(rec["pv_cnt"])


def test_fixture_covers_both_sides_of_every_threshold():
    """every test is seen both enabled and disabled, and the fixture was made by the reference (not by hand)"""
    seen_on, seen_off = 0, 0
    for rec in CASES:
        if rec.get("refused"):
            continue
        seen_on |= rec["enabled_mask"]
        seen_off |= ~rec["enabled_mask"] & 0xFFFE
    assert seen_on == 0xFFFE
    # Frequency, CumulativeSums, Runs, LongestRun and DFT only disable below n = 1000, which the reference's
    # command line refuses (parse_args.c:945): they can never be seen disabled there
    assert seen_off == 0xFFFE & ~((1 << 1) | (1 << 3) | (1 << 4) | (1 << 5) | (1 << 7))


def test_query_layout_refuses_what_create_refuses():
    for bad in (dict(n=1001), dict(n=0), dict(max_streams=0)):
        with pytest.raises(sts_b200.EngineError):
            sts_b200.query_layout(**bad)
    p = sts_b200.default_params()
    p.abi_version = 99
    with pytest.raises(sts_b200.EngineError):
        sts_b200.query_layout(p)
    with pytest.raises(sts_b200.EngineError):           # nothing left to run
        sts_b200.query_layout(n=1 << 20, test_mask=1 << 9, overlapping_m=10)
