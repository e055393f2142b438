"""Executable specification of K1's packed encoder (open-msspe-design_b200/csrc/kmer_build_fast.cu,
encode_keys_packed_kernel) on the CPU, checked against the oracle's find_kmers (od-msspe/src/main.rs:163-171):

  * a search window of <= 64 bases is three bit planes over its positions (low bit, high bit of the 2-bit base; is-ACGT);
  * a slot is valid when k consecutive positions are ACGT: a run test on the third plane;
  * the word at slot q + d repeats the word at slot q exactly when the window equals itself shifted by d over the k positions
    q .. q + k - 1 -- so `itertools::unique()` (first occurrence kept) needs no comparison of words, only shifts and ANDs;
  * the reverse direction complements and reverses the kept words; duplicates among reverse complements are duplicates
    among the words themselves.

The CUDA kernel does the same with 64-bit registers, one distance per lane and an OR-reduction; tests/test_gpu_kmer.py checks
the kernel itself against the oracle on the GPU."""
import random

from oracle import kmer_oracle as ko

MASK = (1 << 64) - 1
CODE = {"A": 0, "C": 1, "G": 2, "T": 3, "U": 3}


def run_k(m, k):
    """bit p of the result: m has ones at p .. p + k - 1 (doubling, then one overlapping step, as the kernel's lambda)."""
    length = 1
    while 2 * length <= k:
        m &= m >> length
        length *= 2
    if length < k:
        m &= m >> (k - length)
    return m


def packed_window_words(window, k):
    """The kept (slot, word) pairs of one search window, by the kernel's plane arithmetic."""
    w = len(window)
    assert w <= 64 and k <= w
    lo = hi = val = 0
    for p, ch in enumerate(window.upper()):
        if ch in CODE:
            c = CODE[ch]
            lo |= (c & 1) << p
            hi |= (c >> 1) << p
            val |= 1 << p
    slots = w - k + 1
    ok = run_k(val, k)
    dup = 0
    for dist in range(1, slots):          # lane l of the warp takes dist = l + 1 and l + 33
        eq = ~((lo ^ (lo >> dist)) | (hi ^ (hi >> dist))) & val & (val >> dist) & MASK
        dup |= (run_k(eq, k) << dist) & MASK
    keep = ok & ~dup
    return [(q, window[q:q + k].upper().replace("U", "T")) for q in range(slots) if (keep >> q) & 1]


def _oracle_words(window, k):
    # find_kmers keeps a word by its characters; 'U' and 'T' are different characters there but the same base after
    # to_records (main.rs:108-122 upper-cases and maps U to T before any window is cut), so the windows given to both sides
    # carry T only
    return ko.find_kmers(window, k)


def _random_window(rng, w, kind):
    if kind == "random":
        s = [rng.choice("ACGT") for _ in range(w)]
    elif kind == "repeat":                       # a short unit repeated: almost every word repeats an earlier one
        unit = [rng.choice("ACGT") for _ in range(rng.randint(1, 6))]
        s = [unit[i % len(unit)] for i in range(w)]
        for _ in range(rng.randint(0, 3)):
            s[rng.randrange(w)] = rng.choice("ACGT")
    else:                                        # two copies of a block at a random distance
        s = [rng.choice("ACGT") for _ in range(w)]
        b = rng.randint(3, max(3, w // 3))
        a0 = rng.randrange(0, w - b)
        a1 = rng.randrange(0, w - b)
        s[a1:a1 + b] = s[a0:a0 + b]
    for _ in range(rng.choice([0, 0, 1, 2, 5])):
        s[rng.randrange(w)] = rng.choice("N-RY")
    return "".join(s)


def test_plane_arithmetic_equals_find_kmers():
    rng = random.Random(20251019)
    n_dups = 0
    for trial in range(4000):
        w = rng.randint(5, 64)
        k = rng.randint(2, min(16, w))
        window = _random_window(rng, w, rng.choice(["random", "repeat", "copy"]))
        got = packed_window_words(window, k)
        want = _oracle_words(window, k)
        assert [x for _, x in got] == want, (window, k)
        assert [q for q, _ in got] == sorted(q for q, _ in got)
        valid = sum(all(c in "ACGT" for c in window[q:q + k]) for q in range(w - k + 1))
        n_dups += valid - len(want)
    assert n_dups > 10000          # the repeat path was really exercised


def test_reverse_direction_is_the_reverse_complement_of_the_kept_words():
    rng = random.Random(7)
    for trial in range(500):
        w = rng.randint(13, 64)
        k = rng.randint(5, 13)
        window = _random_window(rng, w, rng.choice(["random", "repeat", "copy"]))
        got = [ko.reverse_complement(x) for _, x in packed_window_words(window, k)]
        want = [ko.reverse_complement(x) for x in ko.find_kmers(window, k)]      # main.rs:148-161, kmer_oracle.get_segment_manager
        assert got == want
        assert len(set(got)) == len(got)


def test_run_test_edge_cases():
    for k in range(1, 33):
        for n in (k - 1, k, k + 1, 40):
            if n < 0:
                continue
            m = (1 << n) - 1
            r = run_k(m, k)
            assert r == ((1 << max(0, n - k + 1)) - 1), (k, n)
