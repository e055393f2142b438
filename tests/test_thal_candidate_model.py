"""Executable specification of the branch-free loop candidate of the thal dimer kernels (csrc/thal.cu, flat_candidate) on the
CPU: the four kinds of a bulge / internal loop (one bulged base, longer bulge, 1x1 mismatch, general interior loop) as ONE
arithmetic sequence ((A + B) + C) + D over a single (S,H) table, zeros where a kind has fewer terms, against the four-way
branch the scalar algorithm uses (Primer3 thal calc_bulge_internal, restated in oracle/thal_oracle.c and in
loop_candidate).  Python floats are IEEE doubles, so bit equality here is the claim the kernels rely on: adding +0.0 changes
no finite or infinite value, and the association of the sums is the scalar code's."""
import math
import random
import struct

INF = math.inf
ILAS, ILAH = (-300 / 310.15), 0.0


def bits(x):
    return struct.pack("<d", x)


def branchy(t, inner, a_in, a_nx, b_in, b_nx, a_cl, a_pv, b_cl, b_pv, l1, l2):
    """loop_candidate (thal.cu) / calc_bulge_internal: tables are dicts of (S,H) by index tuple."""
    ls = l1 + l2 - 1
    iS, iH = inner
    if l1 == 0 or l2 == 0:
        if l1 + l2 == 1:
            S = t["bulge"][ls][0] + t["stack"][(a_in, a_cl, b_in, b_cl)][0]
            H = t["bulge"][ls][1] + t["stack"][(a_in, a_cl, b_in, b_cl)][1]
            if H > 0 or S > 0:
                H, S = INF, -1.0
            H += iH
            S += iS
            if not math.isfinite(H):
                H, S = INF, -1.0
        else:
            H = t["bulge"][ls][1] + t["atp"][(a_in, b_in)][1] + t["atp"][(a_cl, b_cl)][1]
            H += iH
            S = t["bulge"][ls][0] + t["atp"][(a_in, b_in)][0] + t["atp"][(a_cl, b_cl)][0]
            S += iS
            if not math.isfinite(H):
                H, S = INF, -1.0
            if H > 0 and S > 0:
                H, S = INF, -1.0
    else:
        xin, xcl = (a_in, a_nx, b_in, b_nx), (b_cl, b_pv, a_cl, a_pv)
        if l1 == 1 and l2 == 1:
            S = t["int2"][xin][0] + t["int2"][xcl][0]
            S += iS
            H = t["int2"][xin][1] + t["int2"][xcl][1]
            H += iH
        else:
            asym = abs(l1 - l2)
            H = t["interior"][ls][1] + t["tst"][xin][1] + t["tst"][xcl][1] + (ILAH * asym)
            H += iH
            S = t["interior"][ls][0] + t["tst"][xin][0] + t["tst"][xcl][0] + (ILAS * asym)
            S += iS
        if not math.isfinite(H):
            H, S = INF, -1.0
        if H > 0 and S > 0:
            H, S = INF, -1.0
    return S, H


def unified(t, inner, a_in, a_nx, b_in, b_nx, a_cl, a_pv, b_cl, b_pv, l1, l2):
    """flat_candidate: three table entries picked by integer selects, one sequence of additions."""
    ls = l1 + l2 - 1
    bulge = l1 == 0 or l2 == 0
    b1 = bulge and ls == 0
    one = l1 == 1 and l2 == 1
    xin, xcl = (a_in, a_nx, b_in, b_nx), (b_cl, b_pv, a_cl, a_pv)
    zero = (0.0, 0.0)
    A = t["bulge"][ls] if bulge else (t["int2"][xin] if one else t["interior"][ls])
    B = t["stack"][(a_in, a_cl, b_in, b_cl)] if b1 else (t["atp"][(a_in, b_in)] if bulge else (t["int2"][xcl] if one else t["tst"][xin]))
    C = zero if (b1 or one) else (t["atp"][(a_cl, b_cl)] if bulge else t["tst"][xcl])
    gen = not bulge and not one
    asym = abs(l1 - l2)
    dH = (ILAH * asym) if gen else 0.0
    dS = (ILAS * asym) if gen else 0.0
    H = A[1] + B[1] + C[1] + dH
    S = A[0] + B[0] + C[0] + dS
    if b1 and (H > 0 or S > 0):
        H, S = INF, -1.0
    H += inner[1]
    S += inner[0]
    if not math.isfinite(H):
        H, S = INF, -1.0
    if not b1 and H > 0 and S > 0:
        H, S = INF, -1.0
    return S, H


def _tables(rng):
    def entry():
        r = rng.random()
        if r < 0.08:
            return (-1.0, INF)                       # the joint-infinity rule of the parameter files
        if r < 0.12:
            return (rng.uniform(0, 5), rng.uniform(0, 900))   # positive terms: the "H > 0 and S > 0" rejections
        return (rng.uniform(-40, 2), rng.uniform(-12000, 400))
    quad = [(a, b, c, d) for a in range(4) for b in range(4) for c in range(4) for d in range(4)]
    return {"stack": {q: entry() for q in quad}, "int2": {q: entry() for q in quad}, "tst": {q: entry() for q in quad},
            "atp": {(a, b): (rng.choice([0.0, 6.9]), rng.choice([0.0, 2200.0])) for a in range(4) for b in range(4)},
            "bulge": [entry() for _ in range(30)], "interior": [entry() for _ in range(30)]}


def test_unified_candidate_is_bit_identical_to_the_four_way_branch():
    rng = random.Random(31)
    kinds = set()
    for trial in range(200):
        t = _tables(rng)
        for _ in range(400):
            l1, l2 = rng.randint(0, 14), rng.randint(0, 14)
            if l1 + l2 == 0 or l1 + l2 > 30:
                continue
            inner = (-1.0, INF) if rng.random() < 0.05 else (rng.uniform(-300, 5), rng.uniform(-90000, 900))
            args = [rng.randrange(4) for _ in range(8)]
            want = branchy(t, inner, *args, l1, l2)
            got = unified(t, inner, *args, l1, l2)
            assert bits(got[0]) == bits(want[0]) and bits(got[1]) == bits(want[1]), (l1, l2, inner, args)
            kinds.add((l1 == 0 or l2 == 0, l1 + l2 == 1, l1 == 1 and l2 == 1))
    assert len(kinds) == 4
