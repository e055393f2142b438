"""GPU parity tests, through the C ABI, for the thermodynamic kernels (K4 oligotm, K5 thal dimer, K6 hairpin).
Tolerance stated by the north star: 0.01 C / 0.01 kcal/mol (= 10 cal/mol); the kernels are written to be
bit-identical to the FP64 oracle (both built without FMA contraction), which the tests also record."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL_TM = 0.01      # Celsius
TOL_DG = 10.0      # cal/mol  (0.01 kcal/mol)

NTTHAL_GOLDEN = [  # delta_g.rs:197-230
    ("AGGCCTATATCCA", "GAAGCAGTATTTT", 37.0, "-75.3988", "-25700", "-2315.07", "-35.9834"),
    ("GCACTTGATGTGA", "GAAGCAGTATTTT", 37.0, "-65.3976", "-22500", "-2216.94", "-44.4018"),
    ("CTGAAGCAGTATT", "GCATCTTTCCCTT", 25.0, "-101.596", "-33500", "-3209.05", "-24.1908"),
    ("CTGAAGCAGTATT", "AATTGTGTGGATT", 25.0, "-54.2976", "-17700", "-1511.18", "-70.3113"),
    ("AGTCCTGCGTGAT", "TGGCCTACATCAG", 25.0, "-141.872", "-45800", "-3500.74", "-11.1906"),
]


@pytest.fixture(scope="module")
def eng():
    import msspe_b200 as m
    e = m.Engine(13, 500, 250, 50)
    yield e
    e.close()


def test_ntthal_golden_blocks_on_gpu(eng):
    import msspe_b200 as m
    for a, b, t, ds, dh, dg, tm in NTTHAL_GOLDEN:
        o = eng.thal_pairs([m.encode_word(a)], [m.encode_word(b)], m.THAL_ANY, m.ThalCond(50, 3, 0, 250, t, 30, 0))[0]
        assert o["no_structure"] == 0
        assert ("%g" % o["ds"], "%g" % o["dh"], "%g" % o["dg"], "%g" % o["tm"]) == (ds, dh, dg, tm)


def test_primer3_kat_on_gpu(eng):
    """primer.rs:238-250."""
    import msspe_b200 as m
    r = eng.primer_thermo([m.encode_word("AGCCCGTGTAAAC")])
    assert "%.3f" % r["tm"][0] == "43.727" and "%.3f" % r["gc"][0] == "53.846"
    assert "%.2f" % r["self_any"][0] == "0.00" and "%.2f" % r["self_end"][0] == "0.00" and "%.2f" % r["hairpin"][0] == "0.00"


def _compare(got, want_list, what):
    nbad = 0
    exact = 0
    for g, w in zip(got, want_list):
        assert int(g["no_structure"]) == w.no_structure, what
        if w.no_structure:
            continue
        assert abs(g["tm"] - w.tm) <= TOL_TM and abs(g["dg"] - w.dg) <= TOL_DG and abs(g["dh"] - w.dh) <= TOL_DG, (what, g, (w.ds, w.dh, w.dg, w.tm))
        assert int(g["n_bp"]) == w.n_bp
        exact += (g["tm"] == w.tm and g["dg"] == w.dg and g["ds"] == w.ds and g["dh"] == w.dh)
        nbad += 1
    return exact, nbad


# MSSPE_THAL_KERNEL picks one of the three exact dimer kernels (csrc/thal.cu): "thread" = one thread per pair (what large
# batches of oligos <= 16 nt get), "flat" = one warp per pair with flat candidate enumeration (oligos > 16 nt), "legacy" =
# one warp per pair, lanes over inner rows (small batches); None = the library's own choice.
@pytest.mark.parametrize("kernel", [None, "thread", "flat", "legacy"])
@pytest.mark.parametrize("k,ttype", [(13, 1), (13, 2), (15, 1), (16, 1), (20, 1), (31, 1), (32, 2), (8, 1)])
def test_random_pairs_vs_oracle(eng, oracle_lib, monkeypatch, k, ttype, kernel):
    import msspe_b200 as m
    if kernel:
        monkeypatch.setenv("MSSPE_THAL_KERNEL", kernel)
    else:
        monkeypatch.delenv("MSSPE_THAL_KERNEL", raising=False)
    rng = np.random.default_rng(100 + k + ttype)
    n = 400 if k <= 16 else 150
    a = rng.integers(0, 4, (n, k))
    b = rng.integers(0, 4, (n, k))
    q = n // 4
    b[:q] = 3 - a[:q, ::-1]                       # perfect complements: long helices
    near = 3 - a[q:2 * q, ::-1]
    flip = rng.random((q, k)) < 0.2               # near-complements: bulges and internal loops
    near[flip] = rng.integers(0, 4, int(flip.sum()))
    b[q:2 * q] = near
    a[-8:] = rng.integers(0, 2, (8, k))           # {A,C}-only vs {A,C}-only: no structure
    b[-8:] = rng.integers(0, 2, (8, k))
    words_a = ["".join("ACGT"[x] for x in r) for r in a]
    words_b = ["".join("ACGT"[x] for x in r) for r in b]
    cond = m.ThalCond(50, 3, 0, 250, 25.0, 30, 0)
    got = eng.thal_pairs([m.encode_word(w) for w in words_a], [m.encode_word(w) for w in words_b], ttype, cond, oligo_len=k)
    O = oracle_lib
    want = [O.thal(x, y, ttype, O.ThalCond(50, 3, 0, 250, 25.0, 30, 0)) for x, y in zip(words_a, words_b)]
    exact, total = _compare(got, want, "k=%d type=%d" % (k, ttype))
    assert exact == total, "expected bit-identical FP64 results, %d of %d were" % (exact, total)


def test_symmetric_pairs_use_the_symmetric_concentration_term(eng, oracle_lib):
    import msspe_b200 as m
    pal = ["ACGTACGTACGT", "GGAATTCCGGAATTCC", "ATATATATATAT"]
    cond = m.ThalCond(50, 1.5, 0.6, 50, 37.0, 30, 0)
    for w in pal:
        got = eng.thal_pairs([m.encode_word(w)], [m.encode_word(w)], m.THAL_ANY, cond, oligo_len=len(w))
        want = oracle_lib.thal(w, w, 1, oracle_lib.ThalCond(50, 1.5, 0.6, 50, 37.0, 30, 0))
        _compare(got, [want], w)
        assert got[0]["tm"] == want.tm


@pytest.mark.parametrize("k", [13, 15, 20, 24])
def test_primer_thermo_vs_oracle(eng, oracle_lib, k):
    import msspe_b200 as m
    rng = np.random.default_rng(k)
    n = 300
    seqs = rng.integers(0, 4, (n, k))
    # plant stem-loops so hairpins with Tm > 0 occur
    for i in range(0, n, 3):
        stem = rng.integers(3, 6)
        seqs[i, k - stem:] = 3 - seqs[i, :stem][::-1]
    words = ["".join("ACGT"[x] for x in r) for r in seqs]
    r = eng.primer_thermo([m.encode_word(w) for w in words], oligo_len=k)
    O = oracle_lib
    c = O.ThalCond(50, 1.5, 0.6, 50, 37.0, 30, 0)
    n_hp = 0
    for i, w in enumerate(words):
        assert abs(r["tm"][i] - O.oligotm(w)) <= 1e-9 and abs(r["gc"][i] - O.gc_percent(w)) <= 1e-12
        for key, ttype in (("self_any", 1), ("self_end", 2), ("hairpin", 4)):
            o = O.thal(w, w, ttype, c)
            want = max(0.0, o.tm)
            assert abs(r[key][i] - want) <= TOL_TM, (w, key, r[key][i], want)
            assert r[key][i] == want, (w, key, "not bit-identical")
            n_hp += (ttype == 4 and want > 0)
    assert n_hp > 10  # the hairpin path was really exercised


def test_cross_dimer_matrix_matches_pair_list_and_oracle(eng, oracle_lib):
    """run_ntthal replacement: all ordered pairs incl. self (delta_g.rs:64-78), compaction below a limit, the
    structure-less list; row tiling gives the same lists."""
    import msspe_b200 as m
    from msspe_b200 import synth
    k, n = 13, 96
    codes = synth.random_primers(n - 3, k, 4)
    codes = np.concatenate([codes, [m.encode_word("AACCACACACCAA"), m.encode_word("CACACAACCACAC"), m.encode_word("CCCCCCCCCCCCC")]]).astype(np.uint64)
    cond = m.ThalCond(50, 3, 0, 250, 25.0, 30, 0)
    a = np.repeat(codes, n)
    b = np.tile(codes, n)
    full = eng.thal_pairs(a, b, m.THAL_ANY, cond)
    limit = -1500.0
    edges, nos = eng.cross_dimer(codes, cond, limit)
    want_e = [(i, full["dg"][i]) for i in range(n * n) if not full["no_structure"][i] and full["dg"][i] < limit]
    want_n = [i for i in range(n * n) if full["no_structure"][i]]
    assert edges["pair"].tolist() == [p for p, _ in want_e] and edges["dg"].tolist() == [d for _, d in want_e]
    assert nos.tolist() == want_n and len(want_n) >= 9 and len(want_e) > 20
    e2, n2 = [], []
    for rb in range(0, n, 25):
        e, s = eng.cross_dimer(codes, cond, limit, rb, min(n, rb + 25))
        e2.append(e)
        n2.append(s)
    assert np.concatenate(e2).tobytes() == edges.tobytes() and np.concatenate(n2).tolist() == nos.tolist()
    O = oracle_lib
    words = [m.decode_word(c, k) for c in codes]
    for p in list(range(0, n * n, 131)) + want_n[:5]:
        w = O.thal(words[p // n], words[p % n], 1, O.ThalCond(50, 3, 0, 250, 25.0, 30, 0))
        assert int(full["no_structure"][p]) == w.no_structure
        if not w.no_structure:
            assert full["dg"][p] == w.dg and full["tm"][p] == w.tm


def test_kmer_stats_and_filters_vs_oracle_pipeline(zika_fasta, oracle_lib):
    """get_kmer_stats / filter_kmers (main.rs:408-516): text-rounded f32 values, mean, std, tm_ok, runs, verdict."""
    import msspe_b200 as m
    O = oracle_lib
    want = O.run_pipeline(zika_fasta, O.default_config(check_cross_dimers=0))
    eng = m.Engine(13, 500, 250, 50)
    for d in (0, 1):
        codes = [m.encode_word(w) for w, _, _, _ in want.candidates[d]]
        st = eng.kmer_stats(codes)
        for i, ws in enumerate(want.stats[d]):
            for a, b in (("tm", "tm"), ("gc_percent", "gc"), ("self_any_th", "self_any"), ("self_end_th", "self_end"),
                         ("hairpin_th", "hairpin"), ("mean", "mean"), ("std", "std")):
                assert np.float32(st[a][i]).tobytes() == np.float32(ws[b]).tobytes(), (d, i, a)
            assert bool(st["tm_ok"][i]) == ws["tm_ok"] and bool(st["runs"][i]) == ws["runs"]
        kept = [m.decode_word(c, 13) for c in st["code"][st["keep"] != 0]]
        assert kept == want.filtered[d]
    want.close()
    eng.close()


def test_kmer_stats_both_equals_two_calls():
    """One device batch for both directions (main.rs:723-724) gives byte-identical rows to two msspe_kmer_stats calls,
    including the per-direction mean / standard deviation."""
    import msspe_b200 as m
    from msspe_b200 import synth
    eng = m.Engine(13, 500, 250, 50)
    f = synth.random_primers(301, 13, 21); r = synth.random_primers(97, 13, 22)
    cfg = m.default_filter_cfg()
    a, b = eng.kmer_stats_both(f, r, cfg)
    assert a.tobytes() == eng.kmer_stats(f, cfg).tobytes() and b.tobytes() == eng.kmer_stats(r, cfg).tobytes()
    a, b = eng.kmer_stats_both(f[:0], r, cfg)
    assert len(a) == 0 and b.tobytes() == eng.kmer_stats(r, cfg).tobytes()
    eng.close()


def _vertex_cover_by_degrees(pool, ea, eb):
    """main.rs:776-798 on integer degrees (the oracle's string version, ko.greedy_vertex_cover, is quadratic in Python and
    is what this is checked against in tests/test_graph.py): repeatedly the live node with the most live neighbours (a
    self conflict counts the node itself), ties -> the greatest word; until no live node has a live neighbour."""
    n = len(pool)
    nbr = [set() for _ in range(n)]
    for a, b in zip(ea.tolist(), eb.tolist()):
        nbr[a].add(b)
        nbr[b].add(a)
    alive = np.ones(n, bool)
    deg = np.array([len(x) for x in nbr], dtype=np.int64)
    want = np.zeros(n, np.uint8)
    while True:
        cand = np.where(alive & (deg > 0))[0]
        if len(cand) == 0:
            break
        mx = deg[cand].max()
        v = max((int(pool[c]), int(c)) for c in cand[deg[cand] == mx])[1]
        alive[v] = False
        want[v] = 1
        for u in nbr[v]:
            if alive[u]:
                deg[u] -= 1
        deg[v] = 0
    return want


def test_cfg4_pool_20000_primers_vs_oracle(eng, oracle_lib):
    """BASELINE configs[3] at its stated size: all 4.0e8 ordered pairs of the 20,000-primer pool (delta_g.rs:61-81 builds
    all N^2 pairs incl. self) through the matrix entry points, lists kept on the device and row-tiled as the multi-GPU path
    does.  Checked against the CPU oracle's thal on >= 10^4 sampled pairs (edge or not: dG below / above the limit, bit
    for bit) and on EVERY pair the engine reports as structure-less; the edge list through the host entry point for one
    row block equals the device lists; then the conflict graph + vertex cover at this size (main.rs:754-798) against the
    oracle's restatement of the loop on the same edges."""
    import msspe_b200 as m
    from msspe_b200 import synth, distributed as D
    O = oracle_lib
    n, k = 20_000, 13
    pool = synth.random_primers(n, k, 4)
    cond = m.ThalCond(50, 3, 0, 250, 25.0, 30, 0)
    limit = -9000.0 + 1.0
    parts_e, parts_n = [], []
    for r in range(4):          # four row blocks, as four ranks would
        rb, re_ = D.row_block(n, r, 4)
        e, s = eng.cross_dimer_device(pool, cond, limit, rb, re_, edge_capacity=(re_ - rb) * n // 50, nostruct_capacity=1 << 16)
        parts_e.append(e.cpu().numpy().copy())
        parts_n.append(s.cpu().numpy().copy())
    e = np.concatenate(parts_e)
    e = e[np.argsort(e[:, 0], kind="stable")]
    edges = np.ascontiguousarray(e).view(m.EDGE_DTYPE).reshape(-1)
    nos = np.sort(np.concatenate(parts_n)).view(np.uint64)
    assert 0.002 * n * n < len(edges) < 0.02 * n * n and len(np.unique(edges["pair"])) == len(edges)
    # the host entry point on one row block returns the same (sorted) lists
    rb, re_ = D.row_block(n, 1, 4)
    e1, s1 = eng.cross_dimer(pool, cond, limit, rb, rb + 500, edge_capacity=1 << 20, nostruct_capacity=1 << 16)
    sel = (edges["pair"] >= rb * n) & (edges["pair"] < (rb + 500) * n)
    assert e1.tobytes() == edges[sel].tobytes()
    assert s1.tolist() == nos[(nos >= rb * n) & (nos < (rb + 500) * n)].tolist()
    # oracle on sampled pairs + on every reported edge of a sample + every structure-less pair
    words = [m.decode_word(c, k) for c in pool]
    oc = O.ThalCond(50, 3, 0, 250, 25.0, 30, 0)
    rng = np.random.default_rng(44)
    edge_dg = dict(zip(edges["pair"].tolist(), edges["dg"].tolist()))
    nos_set = set(nos.tolist())
    sample = set(rng.integers(0, n * n, 10_000).tolist()) | set(rng.choice(edges["pair"], 2_000, replace=False).tolist()) | {i * n + i for i in range(0, n, 37)}
    for p in sample:
        w = O.thal(words[p // n], words[p % n], 1, oc)
        if w.no_structure:
            assert p in nos_set
        elif w.dg < limit:
            assert edge_dg.get(p) == w.dg, (p, edge_dg.get(p), w.dg)
        else:
            assert p not in edge_dg and p not in nos_set
    for p in nos.tolist():
        assert O.thal(words[p // n], words[p % n], 1, oc).no_structure == 1
    # conflict graph + vertex cover (main.rs:754-798) on these 20,000 nodes: edges with %g -> f32 dG < threshold
    dg32 = np.array([np.float32(float("%g" % d)) for d in edges["dg"]], dtype=np.float32)
    conf = edges["pair"][dg32 < np.float32(-9000.0)]
    ea, eb = (conf // n).astype(np.uint32), (conf % n).astype(np.uint32)
    deleted = eng.vertex_cover(pool, ea, eb)
    want = _vertex_cover_by_degrees(pool, ea, eb)
    assert deleted.tolist() == want.tolist() and want.sum() > 100


def test_engine_equals_the_reference_executable(eng):
    """The kernels against REFERENCE OUTPUT: stdout of the reference's own Primer3 2.6.1 ntthal executable run under tools/a64emu
    (tests/golden/ntthal_emulated.json) - ANY and END1 pairs of equal length and every HAIRPIN case up to 32 nt, random salts and
    -maxloop included.  Compared as ntthal prints them ("%g", 6 significant digits); a structure-less case has empty stdout."""
    import json
    import os
    import msspe_b200 as m
    from conftest import GOLDEN
    with open(os.path.join(GOLDEN, "ntthal_emulated.json")) as f:
        cases = json.load(f)["cases"]
    ttype = {"ANY": m.THAL_ANY, "END1": m.THAL_END1, "HAIRPIN": m.THAL_HAIRPIN}
    n = {"ANY": 0, "END1": 0, "HAIRPIN": 0}
    n_hot = 0
    for c in cases:
        a = c["args"]
        mode = a[1]
        if mode not in ttype or "-i" in a:
            continue
        o = {"-maxloop": "30"}
        o.update({a[i]: a[i + 1] for i in range(2, len(a), 2)})
        s1, s2 = o["-s1"], o.get("-s2", o["-s1"])
        if len(s1) != len(s2) or len(s1) > 32:
            continue
        cond = m.ThalCond(float(o["-mv"]), float(o["-dv"]), float(o["-n"]), float(o["-d"]), float(o["-t"]), int(o["-maxloop"]), 0)
        g = eng.thal_pairs([m.encode_word(s1)], [m.encode_word(s2)], ttype[mode], cond, oligo_len=len(s1))[0]
        if c["stdout"] == "":
            assert int(g["no_structure"]) == 1, a
        else:
            tok = c["stdout"].split("\n")[0].split()
            ref = (tok[8], tok[11], tok[14], tok[17]) if mode == "HAIRPIN" else (tok[7], tok[10], tok[13], tok[16])
            assert int(g["no_structure"]) == 0, a
            assert ("%g" % g["ds"], "%g" % g["dh"], "%g" % g["dg"], "%g" % g["tm"]) == ref, (a, g)
            n_hot += mode != "ANY" and float(ref[3]) > 0
        n[mode] += 1
    assert n["ANY"] >= 90 and n["END1"] >= 90 and n["HAIRPIN"] >= 150 and n_hot >= 80, (n, n_hot)
