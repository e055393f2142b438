"""The drop-in process boundary: `od-msspe` flags / env (config.rs:11-148), FASTA in, CSV + stdout report out.
CPU tests cover the argument surface; GPU tests compare the produced bytes with the oracle pipeline."""
import os
import sys
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "open-msspe-design_b200", "bin", "od-msspe")


@pytest.fixture(scope="module")
def exe():
    if not os.path.exists(EXE):
        subprocess.run([os.path.join(ROOT, "open-msspe-design_b200", "build.sh")], check=True)
        subprocess.run([os.path.join(ROOT, "open-msspe-design_b200", "host", "build.sh")], check=True)
    return EXE


def run(exe, *args, env=None, cwd=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([exe, *args], capture_output=True, text=True, env=e, cwd=cwd)


def test_version_and_help(exe):
    r = run(exe, "--version")
    assert r.returncode == 0 and r.stdout.strip() == "od-msspe 1.2.0"  # Cargo.toml:3
    r = run(exe, "--help")
    assert r.returncode == 0
    for flag in ("--kmer-size", "--window-size", "--overlap-size", "--max-mismatch-segments", "--max-iterations",
                 "--search-windows-size", "--mv-conc", "--dv-conc", "--dntp-conc", "--dna-conc", "--annealing-temp", "--min-tm",
                 "--max-tm", "--max-self-dimer-any-tm", "--max-self-dimer-end-tm", "--max-hairpin-tm", "--delta-g-threshold",
                 "--keep-all", "--check-cross-dimers", "--check-self-dimers", "--check-hairpin", "--tm-stddev",
                 "--disable-tm-stddev", "--disable-min-max-tm", "--do-align", "--ntthal", "--primer3", "--input", "--output"):
        assert flag in r.stdout, flag


def test_argument_errors(exe, tmp_path):
    assert run(exe).returncode == 2                                          # required -i / -o
    assert run(exe, "-i", "a").returncode == 2
    assert run(exe, "-i", "a", "-o", "b", "--keep-all", "maybe").returncode == 2   # value_parser ["true","false"]
    assert run(exe, "-i", "a", "-o", "b", "--kmer-size", "x").returncode == 2
    assert run(exe, "-i", "a", "-o", "b", "--nope").returncode == 2
    r = run(exe, "-i", "a", "-o", "b")                                       # DO_ALIGN defaults to true: mafft is out of scope
    assert r.returncode == 2 and "do-align=false" in r.stderr
    r = run(exe, "-i", str(tmp_path / "missing.fa"), "-o", "b", "--do-align=false")
    assert r.returncode == 1
    empty = tmp_path / "empty.fa"
    empty.write_text("")
    r = run(exe, "-i", str(empty), "-o", "b", "--do-align", "false")
    assert r.returncode == 101 and "No sequences found" in r.stderr           # main.rs:652-654 panic
    r = run(exe, "-i", str(empty), "-o", "b", env={"DO_ALIGN": "false", "KEEP_ALL": "perhaps"})
    assert r.returncode == 2                                                  # env values go through the same parser


def _oracle_cfg(O, **kw):
    return O.default_config(**kw)


CASES = [
    ("defaults", [], {}),
    ("no_cross", ["--check-cross-dimers=false"], dict(check_cross_dimers=0)),
    ("no_self", ["--check-self-dimers", "false"], dict(check_self_dimers=0)),
    ("keep_all", ["--keep-all=true"], dict(keep_all=1)),
    ("thresholds", ["--delta-g-threshold=-3000", "--max-hairpin-tm", "40", "--tm-stddev=1.0", "--max-mismatch-segments=5"],
     dict(delta_g_threshold=-3000.0, max_hairpin_tm=40.0, tm_stddev=1.0, max_mismatch_segments=5)),
    ("k15", ["--kmer-size=15", "--max-iterations=40", "--disable-tm-stddev=true", "--disable-min-max-tm=true", "--check-hairpin=false"],
     dict(kmer_size=15, max_iterations=40, disable_tm_stddev=1, disable_min_max_tm=1, check_hairpin=0)),
    ("salts", ["--mv-conc=40", "--dv-conc=1.5", "--dntp-conc=0.2", "--dna-conc=100", "--annealing-temp=37", "--window-size=400",
               "--overlap-size=200", "--search-windows-size=60"],
     dict(mv_conc=40.0, dv_conc=1.5, dntp_conc=0.2, dna_conc=100.0, annealing_temp=37.0, window_size=400, overlap_size=200,
          search_windows_size=60)),
]


@pytest.mark.gpu
@pytest.mark.parametrize("name,flags,okw", CASES, ids=[c[0] for c in CASES])
def test_cli_bytes_equal_oracle_pipeline(exe, zika_fasta, oracle_lib, tmp_path, name, flags, okw):
    fa = tmp_path / "in.fa"
    fa.write_bytes(zika_fasta)
    out = tmp_path / "out.csv"
    r = run(exe, "-i", str(fa), "-o", str(out), "--do-align=false", *flags, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr
    want = oracle_lib.run_pipeline(zika_fasta, _oracle_cfg(oracle_lib, **okw))
    assert out.read_text() == want.csv
    assert r.stdout == want.report
    want.close()


@pytest.mark.gpu
def test_cli_env_fallbacks_and_param_dir(exe, zika_fasta, oracle_lib, tmp_path):
    """Every option also reads an env var (config.rs:20-147); ./primer3_config/ is honoured like `ntthal -path`."""
    fa = tmp_path / "in.fa"
    fa.write_bytes(zika_fasta)
    out = tmp_path / "o.csv"
    r = run(exe, "-i", str(fa), "-o", str(out), env={"DO_ALIGN": "false", "KMER_SIZE": "14", "MAX_ITERATIONS": "25", "CHECK_CROSS_DIMERS": "false"})
    assert r.returncode == 0, r.stderr
    want = oracle_lib.run_pipeline(zika_fasta, oracle_lib.default_config(kmer_size=14, max_iterations=25, check_cross_dimers=0))
    assert out.read_text() == want.csv and r.stdout == want.report
    want.close()


@pytest.mark.gpu
def test_cli_param_dir_reaches_ntthal_only(exe, zika_fasta, oracle_lib, tmp_path):
    """delta_g.rs:90-108 gives `-path <cwd>/primer3_config/` to ntthal only; primer3_core is started without a parameter
    path (primer.rs:125-160) and uses Primer3's compiled-in tables.  So a ./primer3_config/ that DIFFERS from the built-in
    tables must not move Tm / GC / SELF_*_TH / HAIRPIN_TH, the filter verdicts or (with cross-dimers off) a byte of the
    CSV; with cross-dimers on it does change the dG edges."""
    import msspe_b200 as m
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_abi import write_param_dir
    fa = tmp_path / "in.fa"
    fa.write_bytes(zika_fasta)
    write_param_dir(m, tmp_path / "primer3_config", stack_ds_shift=-3.0)
    out = tmp_path / "o.csv"
    r = run(exe, "-i", str(fa), "-o", str(out), "--do-align=false", "--check-cross-dimers=false", cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr
    want = oracle_lib.run_pipeline(zika_fasta, oracle_lib.default_config(check_cross_dimers=0))   # embedded (= Primer3 built-in) tables
    assert out.read_text() == want.csv and r.stdout == want.report
    want.close()
    # the library call behind it: primer thermodynamics are unchanged by msspe_set_thal_params, pair dG is not
    eng = m.Engine(13, 500, 250, 50)
    codes = [m.encode_word(w) for w in ("AGCCCGTGTAAAC", "GCGCGCGCATATA", "ACGTACGTACGTA", "GGGGCCCCAAATT")]
    cond = m.ThalCond(50, 3, 0, 250, 25.0, 30, 0)
    t0 = eng.primer_thermo(codes)
    p0 = eng.thal_pairs(codes, codes[::-1], m.THAL_ANY, cond)
    eng.set_thal_params_dir(str(tmp_path / "primer3_config"))
    t1 = eng.primer_thermo(codes)
    p1 = eng.thal_pairs(codes, codes[::-1], m.THAL_ANY, cond)
    for key in ("tm", "gc", "self_any", "self_end", "hairpin"):
        assert t0[key].tobytes() == t1[key].tobytes(), key
    assert p0["dg"].tobytes() != p1["dg"].tobytes()
    eng.close()


@pytest.mark.gpu
def test_cli_golden_files(exe, zika_fasta, tmp_path):
    """tests/golden/snapshot_*: SNAPSHOTS written by this repo's own oracle pipeline (not reference output -- the Rust
    binary and its Mach-O Primer3 cannot run here); they catch unintended drift of either side, nothing more."""
    fa = tmp_path / "in.fa"
    fa.write_bytes(zika_fasta)
    out = tmp_path / "o.csv"
    r = run(exe, "-i", str(fa), "-o", str(out), "--do-align=false")
    g = os.path.join(ROOT, "tests", "golden")
    assert out.read_text() == open(os.path.join(g, "snapshot_zika96_default.csv")).read()
    assert r.stdout == open(os.path.join(g, "snapshot_zika96_default.report.txt")).read()


@pytest.mark.gpu
def test_cli_parser_desync_after_structureless_pairs(exe, oracle_lib, tmp_path):
    """delta_g.rs:31-56: a pair without structure prints nothing (the reference's own ntthal, tests/golden/ntthal_emulated.json), so the
    5-line parser credits every later block to an earlier input line.  A pool with
    {A,C}-only primers (no Watson-Crick partner letters) triggers it; --keep-all=false so the vertex cover runs."""
    import numpy as np
    rng = np.random.default_rng(11)
    L = 1400
    base = rng.integers(0, 4, L)
    base[100:150] = rng.integers(0, 2, 50)       # an A/C-only search window -> A/C-only forward primers
    base[600:650] = rng.integers(0, 2, 50)
    lines = []
    for i in range(30):
        s = base.copy()
        mut = rng.random(L) < 0.01
        mut[100:150] = False
        mut[600:650] = False
        s[mut] = rng.integers(0, 4, int(mut.sum()))
        lines.append(">s%d\n%s\n" % (i, "".join("ACGT"[x] for x in s)))
    fasta = "".join(lines).encode()
    fa = tmp_path / "in.fa"
    fa.write_bytes(fasta)
    out = tmp_path / "o.csv"
    flags = ["--window-size=200", "--overlap-size=100", "--delta-g-threshold=-1500", "--disable-tm-stddev=true",
             "--disable-min-max-tm=true", "--max-mismatch-segments=1"]
    r = run(exe, "-i", str(fa), "-o", str(out), "--do-align=false", *flags)
    assert r.returncode == 0, r.stderr
    want = oracle_lib.run_pipeline(fasta, oracle_lib.default_config(window_size=200, overlap_size=100, delta_g_threshold=-1500.0,
                                                                     disable_tm_stddev=1, disable_min_max_tm=1, max_mismatch_segments=1))
    ac_only = [w for w in want.filtered[0] + want.filtered[1] if set(w) <= set("AC") or set(w) <= set("GT")]
    assert ac_only, "the fixture must contain primers that cannot pair with themselves"
    assert out.read_text() == want.csv and r.stdout == want.report
    want.close()
