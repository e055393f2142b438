"""the `ntthal` / `primer3_core` protocol shims (plugin seam #1 of section 8b) that let the
unmodified Rust binary run its thermodynamics on the GPU.  Checked against the real tool outputs the reference keeps
in its own tests (delta_g.rs:197-230 via tests/golden/ntthal_delta_g_rs.json, primer.rs:36-66 / 238-250) and against
the reference's own parsers restated here."""
import json
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIMS = os.path.join(ROOT, "open-msspe-design_b200", "bin", "shims")
CONFIG_DIR = os.path.join(ROOT, "open-msspe-design_b200", "primer3_config")


def _ntthal(lines, cond, mode="ANY", extra=()):
    args = [os.path.join(SHIMS, "ntthal"), "-a", mode, "-mv", "%.2f" % cond["mv"], "-dv", "%.2f" % cond["dv"], "-n", "%.2f" % cond["dntp"],
            "-d", "%.2f" % cond["dna"], "-t", "%.2f" % cond["t"], *extra, "-i"]   # argv of delta_g.rs:93-110
    r = subprocess.run(args, input="".join(a + "," + b + "\n" for a, b in lines), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    return r.stdout


def test_ntthal_shim_reproduces_the_reference_test_vectors():
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "ntthal_delta_g_rs.json")))
    for g in gold:
        out = _ntthal([(g["a"], g["b"])], g["cond"]).split("\n")
        assert out[0].startswith("Calculated thermodynamical parameters for dimer:\tdS = ")
        tok = out[0].split()
        # the reference's parser reads token 13 (delta_g.rs:33-36)
        assert tok[13] == g["values"]["dG"]
        assert [tok[7], tok[10], tok[13], tok[16]] == [g["values"]["dS"], g["values"]["dH"], g["values"]["dG"], g["values"]["t"]]
        # the Rust source keeps the rows without their trailing blanks; the executable pads the two middle rows
        assert [l.rstrip(" ") for l in out[1:5]] == [tag + "\t" + body for tag, body in g["lines"]]
        assert len({len(l) for l in out[1:5]}) == 1
        assert out[5:] == [""]


def parse_ntthal_output(inp: str, output: str, threshold: float):
    """delta_g.rs:27-59: per input line read ONE output line, dG = token 13, then skip four lines unconditionally."""
    edges, lines, pos = [], output.split("\n"), 0
    if lines and lines[-1] == "":
        lines.pop()
    for l in inp.strip().split("\n"):
        if pos >= len(lines):
            break
        tok = lines[pos].split()
        pos += 1
        if len(tok) > 13:
            try:
                dg = float(np.float32(float(tok[13])))
                if dg < threshold:
                    edges.append((l, dg))
            except ValueError:
                pass
        pos += 4
    return edges


def test_ntthal_shim_through_the_reference_parser_including_structureless_pairs():
    import msspe_b200 as m
    from msspe_b200 import synth
    k = 13
    codes = synth.random_primers(30, k, 9)
    words = [m.decode_word(int(c), k) for c in codes] + ["ACACACACACACA", "AAAAAAAAAAAAA"]  # the last two: no structure with themselves
    codes = np.array([m.encode_word(w) for w in words], dtype=np.uint64)
    pairs = [(a, b) for a in words for b in words]
    cond = dict(mv=50, dv=3, dntp=0, dna=250, t=25)
    out = _ntthal(pairs, cond)
    eng = m.Engine(k, 500, 250, 50)
    n = len(words)
    res = eng.thal_pairs(np.repeat(codes, n), np.tile(codes, n), m.THAL_ANY, m.ThalCond(50, 3, 0, 250, 25.0, 30, 0))
    want = []
    for p, (a, b) in enumerate(pairs):
        if not res["no_structure"][p]:   # a structure-less pair prints NOTHING (tests/golden/ntthal_emulated.json)
            want.append("Calculated thermodynamical parameters for dimer:\tdS = %g\tdH = %g\tdG = %g\tt = %g" % (res["ds"][p], res["dh"][p], res["dg"][p], res["tm"][p]))
    assert int(res["no_structure"].sum()) >= 2
    got = [l for l in out.split("\n") if l.startswith("Calculated")]
    assert got == want and "No secondary" not in out
    assert len(out.split("\n")) - 1 == 5 * (len(pairs) - int(res["no_structure"].sum()))
    # drawn duplexes: both strands complete, paired columns complementary
    blocks = out.split("\n")
    i = 0
    for p, (a, b) in enumerate(pairs):
        if res["no_structure"][p]:
            continue
        r = [x.split("\t", 1)[1] for x in blocks[i + 1:i + 5]]
        i += 5
        L = max(len(x) for x in r)
        r = [x.ljust(L) for x in r]
        top = "".join(r[0][c] if r[0][c] not in " -" else (r[1][c] if r[1][c] != " " else "") for c in range(L))
        bot = "".join(r[3][c] if r[3][c] not in " -" else (r[2][c] if r[2][c] != " " else "") for c in range(L))
        assert top == a and bot == b[::-1]
        npair = 0
        for c in range(L):
            if r[1][c] != " ":
                assert {r[1][c], r[2][c]} in ({"A", "T"}, {"C", "G"})
                npair += 1
        assert npair == int(res["n_bp"][p])
    # the reference parser over the whole stream runs over the stand-in's stream (blocks only: a structure-less pair prints nothing)
    inp = "".join(a + "," + b + "\n" for a, b in pairs)
    edges = parse_ntthal_output(inp, out, -3000.0)
    assert len(edges) > 0
    eng.close()


def test_ntthal_shim_path_and_errors():
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "ntthal_delta_g_rs.json")))[0]
    base = _ntthal([(gold["a"], gold["b"])], gold["cond"])
    if os.path.isdir(CONFIG_DIR):
        assert _ntthal([(gold["a"], gold["b"])], gold["cond"], extra=("-path", CONFIG_DIR + "/")) == base
    exe = os.path.join(SHIMS, "ntthal")
    assert subprocess.run([exe, "-a", "ANY", "-i"], input="ACGT,ACGTT\n", capture_output=True, text=True).returncode != 0
    assert subprocess.run([exe, "-a", "ANY", "-i"], input="ACGN,ACGT\n", capture_output=True, text=True).returncode != 0
    assert subprocess.run([exe, "-a", "HAIRPIN", "-i"], input="ACGT,ACGT\n", capture_output=True, text=True).returncode != 0
    assert subprocess.run([exe, "-a", "ANY", "-i"], input="", capture_output=True, text=True).stdout == ""
    r = subprocess.run([exe, "-a", "ANY", "-mv", "50", "-dv", "3", "-n", "0", "-d", "250", "-t", "37", "-s1", gold["a"], "-s2", gold["b"]],
                       capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout == base


def format_primer3_input(primers, min_tm, max_tm):
    """primer.rs:125-140"""
    s = ""
    for p in primers:
        s += "SEQUENCE_ID=%s\nSEQUENCE_PRIMER=%s\nPRIMER_TASK=check_primers\nPRIMER_MIN_SIZE=13\nPRIMER_MIN_TM=%.2f\nPRIMER_MAX_TM=%.2f\n" % (p, p, min_tm, max_tm)
        s += "PRIMER_OPT_TM=%.2f\nPRIMER_PICK_ANYWAY=1\n=\n" % max_tm
    return s


def parse_primer3_output(text):
    """primer.rs:67-114"""
    out, cur = [], dict(id="", tm=0.0, gc=0.0, self_any_th=0.0, self_end_th=0.0, hairpin_th=0.0)
    keys = {"PRIMER_LEFT_0_TM": "tm", "PRIMER_LEFT_0_GC_PERCENT": "gc", "PRIMER_LEFT_0_SELF_ANY_TH": "self_any_th",
            "PRIMER_LEFT_0_SELF_END_TH": "self_end_th", "PRIMER_LEFT_0_HAIRPIN_TH": "hairpin_th"}
    for line in text.split("\n"):
        if line == "=":
            if not cur["id"]:
                break
            out.append(cur)
            cur = dict(id="", tm=0.0, gc=0.0, self_any_th=0.0, self_end_th=0.0, hairpin_th=0.0)
            continue
        kv = line.split("=")
        if kv[0] == "SEQUENCE_ID":
            cur["id"] = kv[1]
        elif kv[0] in keys:
            cur[keys[kv[0]]] = float(np.float32(float(kv[1])))
    return out


def test_primer3_core_shim_reference_vector_and_engine_parity():
    import msspe_b200 as m
    from msspe_b200 import synth
    exe = os.path.join(SHIMS, "primer3_core")
    # the documented example of primer.rs:36-66 (no PRIMER_OPT_TM given: default 60)
    inp = "SEQUENCE_ID=example1\nSEQUENCE_PRIMER=AGCCCGTGTAAAC\nPRIMER_TASK=check_primers\nPRIMER_MIN_SIZE=13\nPRIMER_MIN_TM=30.0\nPRIMER_MAX_TM=60.0\n=\n"
    r = subprocess.run([exe], input=inp, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.split("\n")
    assert lines[:6] == inp.split("\n")[:6]
    for want in ["PRIMER_LEFT_NUM_RETURNED=1", "PRIMER_RIGHT_NUM_RETURNED=0", "PRIMER_INTERNAL_NUM_RETURNED=0", "PRIMER_PAIR_NUM_RETURNED=0",
                 "PRIMER_LEFT_0_PENALTY=23.273240", "PRIMER_LEFT_0_SEQUENCE=AGCCCGTGTAAAC", "PRIMER_LEFT_0=0,13", "PRIMER_LEFT_0_TM=43.727",
                 "PRIMER_LEFT_0_GC_PERCENT=53.846", "PRIMER_LEFT_0_SELF_ANY_TH=0.00", "PRIMER_LEFT_0_SELF_END_TH=0.00", "PRIMER_LEFT_0_HAIRPIN_TH=0.00"]:
        assert want in lines, want
    assert lines[-2:] == ["=", ""]
    # od-msspe's own input format through its own parser = the engine's kmer_stats numbers
    k = 13
    codes = synth.random_primers(200, k, 12)
    words = [m.decode_word(int(c), k) for c in codes]
    r = subprocess.run([exe], input=format_primer3_input(words, 30.0, 60.0), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    got = parse_primer3_output(r.stdout)
    eng = m.Engine(k, 500, 250, 50)
    st = eng.kmer_stats(codes)
    assert [g["id"] for g in got] == words
    for name, col in (("tm", "tm"), ("gc", "gc_percent"), ("self_any_th", "self_any_th"), ("self_end_th", "self_end_th"), ("hairpin_th", "hairpin_th")):
        assert np.array_equal(np.array([g[name] for g in got], np.float32), st[col].astype(np.float32)), name
    assert any(g["hairpin_th"] > 0 for g in got) or any(g["self_any_th"] > 0 for g in got)
    eng.close()
    bad = subprocess.run([exe], input="SEQUENCE_ID=x\nSEQUENCE_PRIMER=ACGT\nPRIMER_TASK=generic\n=\n", capture_output=True, text=True)
    assert "PRIMER_ERROR=" in bad.stdout and bad.stdout.endswith("=\n")


@pytest.mark.parametrize("okw", [dict(), dict(delta_g_threshold=-3000.0, max_hairpin_tm=40.0, tm_stddev=1.0, max_mismatch_segments=5),
                                 dict(check_self_dimers=0), dict(kmer_size=15, max_iterations=40, disable_tm_stddev=1, disable_min_max_tm=1)],
                         ids=["defaults", "thresholds", "no_self", "k15"])
def test_reference_host_logic_over_the_shims_equals_the_full_pipeline(zika_fasta, oracle_lib, monkeypatch, okw):
    """The seam end to end: the restated reference pipeline (the oracle's port of main.rs: FASTA, k-mer engine, filters,
    5-line parser, vertex cover, report, CSV) SPAWNS the two shims exactly as the Rust binary spawns `--primer3` /
    `--ntthal` (its input text, its argv, its parsers) instead of doing its own thermodynamics.  CSV and report must
    be byte-identical to the same pipeline with the in-process FP64 oracle arithmetic."""
    want = oracle_lib.run_pipeline(zika_fasta, oracle_lib.default_config(**okw))
    monkeypatch.setenv("ORACLE_PRIMER3", os.path.join(SHIMS, "primer3_core"))
    monkeypatch.setenv("ORACLE_NTTHAL", os.path.join(SHIMS, "ntthal"))
    got = oracle_lib.run_pipeline(zika_fasta, oracle_lib.default_config(**okw))
    assert got.csv == want.csv and got.report == want.report
    assert len(got.csv.split("\n")) > 5      # header + primers: the strict thresholds leave 6 primers on this input
    got.close(); want.close()
