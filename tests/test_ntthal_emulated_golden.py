"""The thermodynamic oracle (oracle/thal_oracle.c) against REFERENCE OUTPUT: stdout of the reference's own Primer3 2.6.1
`ntthal` executable (od-msspe/bin/ntthal, the program delta_g.rs:90-108 spawns), produced by running that Mach-O arm64 binary
under the instruction-level interpreter in tools/a64emu (tools/gen_ntthal_emulated_golden.py ->
tests/golden/ntthal_emulated.json; the generator refuses to write unless the interpreter reproduces the five blocks of
delta_g.rs:197-230).  526 invocations: ANY, END1, END2 and HAIRPIN, od-msspe's and Primer3's conditions, random salts,
-maxloop, 8 .. 32 nt oligos, hairpins with 3 .. 9 nt loops, bulges and interior loops, the reference's `-path .. -i` protocol.

This pins what rounds 1-2 had to label "parity unpinned": hairpin Tm > 0, END1 / END2 Tm > 0, bulge loops, the triloop and
tetraloop bonuses, non-default salt corrections, and what ntthal prints for a structure-less pair (NOTHING on stdout for a
dimer; a message on stderr for a hairpin)."""
import json
import os

import pytest

from conftest import GOLDEN

TYPE = {"ANY": 1, "END1": 2, "END2": 3, "HAIRPIN": 4}


@pytest.fixture(scope="module")
def cases():
    with open(os.path.join(GOLDEN, "ntthal_emulated.json")) as f:
        return json.load(f)["cases"]


def parse_args(args):
    o = {"maxloop": 30, "s2": None, "i": False, "mv": 50.0, "dv": 0.0, "n": 0.8, "d": 50.0, "t": 37.0}
    it = iter(args)
    for a in it:
        if a == "-i":
            o["i"] = True
        elif a == "-a":
            o["mode"] = next(it)
        elif a in ("-s1", "-s2", "-path"):
            o[a[1:]] = next(it)
        elif a == "-maxloop":
            o["maxloop"] = int(next(it))
        else:
            o[a[1:]] = float(next(it))
    return o


def header_values(line, hairpin):
    tok = line.split()
    assert tok[:5] == ["Calculated", "thermodynamical", "parameters", "for", "dimer:"]
    if hairpin:
        return tok[5], (tok[8], tok[11], tok[14], tok[17])
    return None, (tok[7], tok[10], tok[13], tok[16])


def oracle_text(O, a, b, o):
    c = O.ThalCond(o["mv"], o["dv"], o["n"], o["d"], o["t"], o["maxloop"], 0)
    r = O.thal(a, b or a, TYPE[o["mode"]], c)
    if r.no_structure:
        return None
    return ("%g" % r.ds, "%g" % r.dh, "%g" % r.dg, "%g" % r.tm)


def test_fixture_shape(cases):
    modes = [parse_args(c["args"])["mode"] for c in cases]
    assert len(cases) == 526 and {m: modes.count(m) for m in TYPE} == {"ANY": 126, "END1": 120, "END2": 120, "HAIRPIN": 160}
    hp = [c for c in cases if parse_args(c["args"])["mode"] == "HAIRPIN"]
    # the outputs no reference-held vector covered before: melting hairpins and END1 / END2 duplexes above 0 C
    assert sum(1 for c in hp if c["stdout"] and float(c["stdout"].split("\n")[0].split()[17]) > 0) >= 60
    e1 = [c for c in cases if parse_args(c["args"])["mode"] in ("END1", "END2") and c["stdout"]]
    assert sum(1 for c in e1 if float(c["stdout"].split("\n")[0].split()[16]) > 0) >= 30
    assert sum(1 for c in cases if c["stdout"] == "") >= 20      # structure-less


def test_oracle_equals_the_reference_executable(oracle_lib, cases):
    O = oracle_lib
    n = 0
    for c in cases:
        o = parse_args(c["args"])
        hairpin = o["mode"] == "HAIRPIN"
        if o["i"]:
            pairs = [l.split(",") for l in c["stdin"].split("\n") if l]
        else:
            pairs = [(o["s1"], o["s2"])]
        blocks = c["stdout"].split("\n")[:-1]
        per = 3 if hairpin else 5
        assert len(blocks) % per == 0
        got = [blocks[i: i + per] for i in range(0, len(blocks), per)]
        want = [(a, b, oracle_text(O, a, b, o)) for a, b in pairs]
        structured = [w for w in want if w[2] is not None]
        # a structure-less pair prints nothing at all on stdout
        assert len(got) == len(structured), (c["args"], c["stdout"])
        for blk, (a, b, vals) in zip(got, structured):
            ln, ref = header_values(blk[0], hairpin)
            assert ref == vals, (c["args"], a, b, ref, vals)
            if hairpin:
                assert int(ln) == len(a) and blk[2] == "STR\t" + a
            n += 1
    assert n >= 480


def draw_duplex(o1, o2, pairing):
    """The four rows ntthal draws for a duplex, as host/ntthal_shim.cpp builds them from the traced pairing: unpaired / paired bases
    of oligo 1, paired / unpaired bases of the reversed oligo 2; blanks before the shorter left end, '-' behind the shorter side of
    a loop and of the right end, the two middle rows padded to the full width."""
    r2 = o2[::-1]
    bp = [(i, pairing[i] - 1) for i in range(len(o1)) if pairing[i]]
    rows = ["", "", "", ""]
    i = j = 0
    for t, (bi, bj) in enumerate(bp):
        u1, u2 = bi - i, bj - j
        w = max(u1, u2)
        if t == 0:
            rows[0] += " " * (w - u1) + o1[i:bi]
            rows[3] += " " * (w - u2) + r2[j:bj]
        else:
            rows[0] += o1[i:bi] + "-" * (w - u1)
            rows[3] += r2[j:bj] + "-" * (w - u2)
        rows[1] += " " * w + o1[bi]
        rows[2] += " " * w + r2[bj]
        rows[0] += " "
        rows[3] += " "
        i, j = bi + 1, bj + 1
    t1, t2 = len(o1) - i, len(r2) - j
    w = max(t1, t2)
    rows[0] += o1[i:] + "-" * (w - t1)
    rows[3] += r2[j:] + "-" * (w - t2)
    rows[1] += " " * w
    rows[2] += " " * w
    return rows


def test_drawn_duplexes_equal_the_reference_executable(oracle_lib, cases):
    """The traced duplex (which bases pair) and the drawing rule of the ntthal stand-in against the executable's SEQ / STR rows.
    One known difference: a self pair has two mirror-image optimal placements.  Primer3 2.6.1 finds their free energies bit-equal
    and keeps the first in scan order; the oracle (and the kernels, bit-identical to it) still add the 1e-6 "SMALL_NON_ZERO" offsets of
    older releases before that comparison, and rounding may then prefer the twin (dropping the offsets gives
    236 of 236).  The numbers are the same and the reference's parser reads line 0 only (delta_g.rs:33-36)."""
    O = oracle_lib
    n = mirrored = 0
    for c in cases:
        o = parse_args(c["args"])
        if o["mode"] not in ("ANY", "END1") or o["i"] or not c["stdout"]:
            continue
        O.thal(o["s1"], o["s2"], TYPE[o["mode"]], O.ThalCond(o["mv"], o["dv"], o["n"], o["d"], o["t"], o["maxloop"], 0))
        rows = draw_duplex(o["s1"], o["s2"], O.thal_last_pairing(len(o["s1"])))
        want = [l.split("\t", 1)[1] for l in c["stdout"].split("\n")[1:5]]
        n += 1
        if rows != want:
            assert o["s1"] == o["s2"], (c["args"], rows, want)
            mirrored += 1
    assert n >= 230 and mirrored <= 1


def test_oracle_equals_the_reference_executable_on_2000_13mer_pairs(oracle_lib):
    """Hot loop #2 at od-msspe's own shape (tests/golden/ntthal_emulated_13mer_pairs.json, tools/gen_ntthal_13mer_pairs_golden.py):
    2000 ordered 13-mer pairs through the reference's ntthal exactly as delta_g.rs:93-110 runs it; line 0 of every block as
    printed, null where the executable printed nothing."""
    O = oracle_lib
    with open(os.path.join(GOLDEN, "ntthal_emulated_13mer_pairs.json")) as f:
        rows = json.load(f)["pairs"]
    c = O.ThalCond(50, 3, 0, 250, 25.0, 30, 0)
    assert len(rows) == 2000 and sum(1 for r in rows if r[2] is None) >= 50
    for a, b, ds, dh, dg, t in rows:
        r = O.thal(a, b, 1, c)
        if ds is None:
            assert r.no_structure == 1, (a, b)
        else:
            assert r.no_structure == 0 and ("%g" % r.ds, "%g" % r.dh, "%g" % r.dg, "%g" % r.tm) == (ds, dh, dg, t), (a, b)


def test_traced_hairpin_folds_equal_the_reference_executable(oracle_lib, cases):
    """Which bases pair in the reported fold: ntthal draws '/' for a base paired downstream, a backslash for one paired upstream, '-'
    otherwise (the SEQ row of HAIRPIN mode).  The oracle's traceback against all structured hairpins of the fixture."""
    O = oracle_lib
    n = 0
    for c in cases:
        o = parse_args(c["args"])
        if o["mode"] != "HAIRPIN" or not c["stdout"]:
            continue
        s = o["s1"]
        O.thal(s, s, 4, O.ThalCond(o["mv"], o["dv"], o["n"], o["d"], o["t"], o["maxloop"], 0))
        partner = O.thal_last_pairing(len(s))
        row = "".join("-" if not q else ("/" if q > i + 1 else "\\") for i, q in enumerate(partner))
        assert c["stdout"].split("\n")[1] == "SEQ\t" + row, (s, row)
        n += 1
    assert n >= 120
