"""`check_primers` against REFERENCE OUTPUT: what the reference's own Primer3 2.6.1 `primer3_core` executable
(od-msspe/bin/primer3_core, spawned by primer.rs:125-140) prints for the Boulder-IO records of primer.rs:113-127, obtained by
running that Mach-O arm64 binary under tools/a64emu (tools/gen_primer3_core_emulated_golden.py ->
tests/golden/primer3_core_emulated.json; the generator first reproduces primer.rs:238-250).  337 primers of 8 .. 32 nt (outside
PRIMER_MIN_SIZE=13 .. PRIMER_MAX_SIZE=27 the executable warns and, under PRIMER_PICK_ANYWAY=1, still reports every number), 166 with
SELF_ANY_TH > 0, 92 with SELF_END_TH > 0, 163 with HAIRPIN_TH > 0.

It pins which arguments `check_primers` hands to oligotm and thal (50 mM monovalent, 1.5 mM divalent, 0.6 mM dNTP, 50 nM DNA,
37 C, loops up to 30, the SantaLucia 1998 table with the Owczarzy 2008 salt correction; self-any = thal ANY, self-end = thal END1
of the primer against itself, negative temperatures printed as 0.00) - read from Primer3's manual in rounds 1-2, observed now."""
import json
import os

import pytest

from conftest import GOLDEN


@pytest.fixture(scope="module")
def primers():
    with open(os.path.join(GOLDEN, "primer3_core_emulated.json")) as f:
        return json.load(f)["primers"]


def test_fixture_shape(primers):
    assert len(primers) == 337 and {len(p["primer"]) for p in primers} >= {8, 12, 13, 20, 28, 32} and primers[0]["primer"] == "AGCCCGTGTAAAC"
    assert {k: primers[0][k] for k in ("TM", "GC_PERCENT", "SELF_ANY_TH", "SELF_END_TH", "HAIRPIN_TH")} == \
        {"TM": "43.727", "GC_PERCENT": "53.846", "SELF_ANY_TH": "0.00", "SELF_END_TH": "0.00", "HAIRPIN_TH": "0.00"}  # primer.rs:238-250
    for k, n in (("SELF_ANY_TH", 160), ("SELF_END_TH", 90), ("HAIRPIN_TH", 160)):
        assert sum(1 for p in primers if float(p[k]) > 0) >= n


def test_oracle_equals_the_reference_executable(oracle_lib, primers):
    O = oracle_lib
    c = O.ThalCond(50, 1.5, 0.6, 50, 37.0, 30, 0)
    for p in primers:
        w = p["primer"]
        got = {"TM": "%.3f" % O.oligotm(w), "GC_PERCENT": "%.3f" % O.gc_percent(w)}
        for key, ttype in (("SELF_ANY_TH", 1), ("SELF_END_TH", 2), ("HAIRPIN_TH", 4)):
            got[key] = "%.2f" % max(0.0, O.thal(w, w, ttype, c).tm)
        assert got == {k: p[k] for k in got}, w


def test_penalty_line_of_the_stand_in_follows_the_executable(oracle_lib, primers):
    """host/primer3_core_shim.cpp prints PRIMER_LEFT_0_PENALTY = |Tm - PRIMER_OPT_TM| + |size - PRIMER_OPT_SIZE| (weights 1, 1,
    everything else 0 for this task); od-msspe does not read the tag, the executable's value confirms the formula anyway."""
    for p in primers:
        w = p["primer"]
        assert "%f" % (abs(oracle_lib.oligotm(w) - 60.0) + abs(len(w) - 20.0)) == p["PENALTY"], w
