#!/usr/bin/env python3
"""End-to-end goldens in which every thermodynamic number comes from the REFERENCE'S OWN executables: the restated od-msspe
pipeline (oracle/kmer_oracle.cpp: k-mer selection, `format_primer3_input` / `parse_primer3_output`, `format_ntthal_input` /
`parse_ntthal_output`, filters, conflict graph, vertex cover, CSV, report) is run in its external-tool mode with
ORACLE_PRIMER3 / ORACLE_NTTHAL pointing at tools/a64emu/bin/{primer3_core,ntthal}, i.e. at od-msspe/bin/primer3_core and
od-msspe/bin/ntthal (Primer3 2.6.1, Mach-O arm64) executed under the interpreter - spawned with the reference's argv and stdin,
read with the reference's parsers.  Structure-less pairs, "%g" / "%.2f" text, the 5-line parser: whatever the real tools do.

Writes tests/golden/refexe_*.fa.gz (two small inputs) and tests/golden/refexe_pipeline.json (CSV + report per flag set).
The in-process oracle arithmetic and the CUDA engine (od-msspe CLI) are then tested against these files.

Run here (needs /root/reference; ~5 minutes on 8 cores):  python tools/gen_refexe_pipeline_golden.py
"""
import gzip
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def acgt30():
    """30 x 1400 nt around one ancestor with two {A,C}-only search windows: primers that cannot pair with themselves (the
    fixture of tests/test_cli.py::test_cli_parser_desync_after_structureless_pairs)."""
    rng = np.random.default_rng(11)
    L = 1400
    base = rng.integers(0, 4, L)
    base[100:150] = rng.integers(0, 2, 50)
    base[600:650] = rng.integers(0, 2, 50)
    lines = []
    for i in range(30):
        s = base.copy()
        mut = rng.random(L) < 0.01
        mut[100:150] = False
        mut[600:650] = False
        s[mut] = rng.integers(0, 4, int(mut.sum()))
        lines.append(">s%d\n%s\n" % (i, "".join("ACGT"[x] for x in s)))
    return "".join(lines).encode()


def zika24():
    """Columns [2000, 5000) of the first 24 records of the reference's own Zika alignment (tests/golden/zika96_aligned.fa.gz)."""
    z = gzip.open(os.path.join(GOLDEN, "zika96_aligned.fa.gz"), "rb").read().decode()
    out = ""
    for r in [r for r in z.split(">") if r][:24]:
        h, *seq = r.split("\n")
        out += ">" + h + "\n" + "".join(seq)[2000:5000] + "\n"
    return out.encode()


SMALL = dict(window_size=200, overlap_size=100, disable_tm_stddev=1, disable_min_max_tm=1, max_mismatch_segments=1)
CASES = [
    ("acgt30", dict(SMALL, delta_g_threshold=-1500.0)),
    ("acgt30", dict(SMALL)),
    ("zika24", dict()),
    ("zika24", dict(delta_g_threshold=-6000.0)),
    ("zika24", dict(check_self_dimers=0)),
    ("zika24", dict(mv_conc=40.0, dv_conc=1.5, dntp_conc=0.2, dna_conc=100.0, annealing_temp=37.0, delta_g_threshold=-5000.0)),
    ("zika24", dict(kmer_size=15, max_iterations=40, disable_tm_stddev=1, disable_min_max_tm=1, check_hairpin=0)),
]


def main():
    O.build()
    inputs = {"acgt30": acgt30(), "zika24": zika24()}
    for name, fa in inputs.items():
        with open(os.path.join(GOLDEN, "refexe_%s.fa.gz" % name), "wb") as f:
            f.write(gzip.compress(fa, mtime=0))
    os.environ["ORACLE_PRIMER3"] = os.path.join(HERE, "a64emu", "bin", "primer3_core")
    os.environ["ORACLE_NTTHAL"] = os.path.join(HERE, "a64emu", "bin", "ntthal")
    out = []
    for name, cfg in CASES:
        r = O.run_pipeline(inputs[name], O.default_config(**cfg))
        out.append({"input": name, "config": cfg, "csv": r.csv, "report": r.report,
                    "n_candidates": [len(x) for x in r.candidates], "n_filtered": [len(x) for x in r.filtered],
                    "n_final": [len(x) for x in r.final]})
        print(name, cfg, out[-1]["n_candidates"], out[-1]["n_filtered"], out[-1]["n_final"], flush=True)
        r.close()
    del os.environ["ORACLE_PRIMER3"], os.environ["ORACLE_NTTHAL"]
    for c in out:   # what the in-process arithmetic gives, for the operator's eyes (tests assert it)
        r = O.run_pipeline(inputs[c["input"]], O.default_config(**c["config"]))
        print("in-process oracle equal:", r.csv == c["csv"] and r.report == c["report"])
        r.close()
    with open(os.path.join(GOLDEN, "refexe_pipeline.json"), "w") as f:
        json.dump({"source": "oracle/kmer_oracle.cpp pipeline with od-msspe/bin/primer3_core and od-msspe/bin/ntthal (under tools/a64emu) as "
                             "its external tools", "cases": out}, f, indent=1)
        f.write("\n")


if __name__ == "__main__":
    main()
