"""End-to-end goldens whose thermodynamic numbers come from the REFERENCE'S OWN executables (tests/golden/refexe_pipeline.json,
written by tools/gen_refexe_pipeline_golden.py): the restated od-msspe pipeline spawned od-msspe/bin/primer3_core and
od-msspe/bin/ntthal (Primer3 2.6.1, run under tools/a64emu) with the reference's argv and stdin and read them with the
reference's parsers (primer.rs:67-111, delta_g.rs:27-59).  Two small inputs (a synthetic alignment with {A,C}-only search
windows, whose primers have no structure with themselves; 24 records x 3 kb of the reference's Zika alignment) x flag sets
(thresholds, self-dimer pairs left out, non-default salts - which reach ntthal only -, k = 15 without the hairpin filter).

CPU: the in-process oracle arithmetic gives the same CSV and report.  GPU (tests/test_zz_reference_executables_gpu.py): so does the
od-msspe CLI on the CUDA engine."""
import gzip
import json
import os
import subprocess

import pytest

from conftest import GOLDEN, ROOT

EXE = os.path.join(ROOT, "open-msspe-design_b200", "bin", "od-msspe")
FLAG = {"window_size": "--window-size", "overlap_size": "--overlap-size", "disable_tm_stddev": "--disable-tm-stddev",
        "disable_min_max_tm": "--disable-min-max-tm", "max_mismatch_segments": "--max-mismatch-segments",
        "delta_g_threshold": "--delta-g-threshold", "check_self_dimers": "--check-self-dimers", "mv_conc": "--mv-conc",
        "dv_conc": "--dv-conc", "dntp_conc": "--dntp-conc", "dna_conc": "--dna-conc", "annealing_temp": "--annealing-temp",
        "kmer_size": "--kmer-size", "max_iterations": "--max-iterations", "check_hairpin": "--check-hairpin"}
BOOLS = {"disable_tm_stddev", "disable_min_max_tm", "check_self_dimers", "check_hairpin"}


def _cases():
    with open(os.path.join(GOLDEN, "refexe_pipeline.json")) as f:
        return json.load(f)["cases"]


def _input(name):
    with gzip.open(os.path.join(GOLDEN, "refexe_%s.fa.gz" % name), "rb") as f:
        return f.read()


def test_fixture_shape():
    cases = _cases()
    assert [(c["input"], c["n_filtered"], c["n_final"]) for c in cases] == [
        ("acgt30", [12, 12], [0, 1]), ("acgt30", [12, 12], [10, 12]), ("zika24", [12, 7], [12, 5]), ("zika24", [12, 7], [9, 2]),
        ("zika24", [12, 7], [12, 6]), ("zika24", [12, 7], [9, 3]), ("zika24", [19, 14], [17, 11])]
    # the first input really has primers that cannot pair with themselves: what ntthal prints for them decides the result
    words = [l.split(",")[2] for l in cases[1]["csv"].split("\n")[1:] if l]
    assert any(set(w) <= set("AC") or set(w) <= set("GT") for w in words)


@pytest.mark.parametrize("i", range(7))
def test_in_process_oracle_equals_the_pipeline_over_the_reference_executables(oracle_lib, i):
    c = _cases()[i]
    r = oracle_lib.run_pipeline(_input(c["input"]), oracle_lib.default_config(**c["config"]))
    assert r.csv == c["csv"] and r.report == c["report"]
    assert [len(x) for x in r.candidates] == c["n_candidates"] and [len(x) for x in r.filtered] == c["n_filtered"]
    r.close()
