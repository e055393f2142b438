"""tools/a64emu - the AArch64 interpreter + Mach-O loader that executes the reference's own Primer3 2.6.1 executables
(od-msspe/bin/ntthal, od-msspe/bin/primer3_core) to produce the `*_emulated.json` goldens.  Two layers:
  * instruction-level checks on hand-encoded A64 words (no reference needed): arithmetic and flags, bit fields, bitmask
    immediates, division, load/store pairs with write-back, conversions, and the exactness of the fused multiply-add;
  * when /root/reference is present (this container; never the GPU box): the executables under the interpreter reproduce the
    reference's own known answers (delta_g.rs:197-230, primer.rs:238-250) and the committed fixtures are what they print."""
import json
import os
import struct
import sys

import pytest

from conftest import GOLDEN, ROOT

sys.path.insert(0, os.path.join(ROOT, "tools", "a64emu"))
import cpu as A   # noqa: E402

REF_BIN = "/root/reference/od-msspe/bin"
needs_reference = pytest.mark.skipif(not os.path.isfile(os.path.join(REF_BIN, "ntthal")), reason="reference executables not present")
BASE, STOP = 0x10000, 0xF0000000
RET = 0xD65F03C0


def run(words, x=None, d=None, mem_size=0x4000):
    mem = bytearray(mem_size)
    for i, w in enumerate(words + [RET]):
        struct.pack_into("<I", mem, 4 * i, w)
    c = A.CPU(mem, BASE)
    for k, v in (x or {}).items():
        c.x[k] = v & A.M64
    for k, v in (d or {}).items():
        c.v[k] = A.d2b(v)
    c.x[30], c.x[31] = STOP, BASE + mem_size - 0x100
    c.run(BASE, STOP, limit=10000)
    return c


def test_loop_with_flags_and_conditional_branch():
    # x0 = 0; x2 = 10; do { x0 += x2; } while (--x2 != 0)
    c = run([0xD2800000, 0xD2800142, 0x8B020000, 0xF1000442, 0x54FFFFC1])
    assert c.x[0] == 55 and c.x[2] == 0 and (c.z, c.c) == (1, 1)


def test_subs_sets_nzcv_like_the_architecture():
    c = run([0x6B040062], x={3: 5, 4: 7})                 # subs w2, w3, w4
    assert c.x[2] == 0xFFFFFFFE and (c.n, c.z, c.c, c.vf) == (1, 0, 0, 0)
    c = run([0x6B040062], x={3: 0x80000000, 4: 1})        # signed overflow
    assert c.x[2] == 0x7FFFFFFF and (c.n, c.z, c.c, c.vf) == (0, 0, 1, 1)


def test_bitmask_immediates_and_bit_fields():
    assert A.decode_bitmasks(0, 0b000111, 0, 32)[0] == 0xFF
    assert A.decode_bitmasks(0, 0b111100, 0, 64)[0] == 0x5555555555555555
    assert A.decode_bitmasks(1, 0b000111, 8, 64)[0] == 0xFF00000000000000
    assert run([0x12001C20], x={1: 0x12345678}).x[0] == 0x78                 # and w0, w1, #0xff
    assert run([0xD3482C20], x={1: 0xABCD}).x[0] == 0xB                      # ubfx x0, x1, #8, #4
    assert run([0x13041C20], x={1: 0xF0}).x[0] == 0xFFFFFFFF                 # sbfx w0, w1, #4, #4
    assert run([0xD344FC20], x={1: 0xF00}).x[0] == 0xF0                      # lsr x0, x1, #4
    assert run([0xD2A24680]).x[0] == 0x12340000                              # movz x0, #0x1234, lsl #16


def test_division_and_multiply_subtract():
    assert A.sx(run([0x9AC20C20], x={1: -7, 2: 2}).x[0], 64) == -3           # sdiv truncates toward zero
    assert run([0x9AC20C20], x={1: 5, 2: 0}).x[0] == 0                       # division by zero gives zero
    assert run([0x9B028C20], x={1: 6, 2: 7, 3: 100}).x[0] == 58              # msub x0, x1, x2, x3


def test_store_and_load_pair_with_write_back():
    c = run([0xA9BF07E0, 0xA8C10FE2], x={0: 0x1111, 1: 0x2222})              # stp x0, x1, [sp, #-16]! ; ldp x2, x3, [sp], #16
    assert (c.x[2], c.x[3]) == (0x1111, 0x2222) and c.x[31] == BASE + 0x4000 - 0x100


def test_float_conversions_and_exact_fused_multiply_add():
    assert A.sx(run([0x9E780020], d={1: -2.7}).x[0], 64) == -2               # fcvtzs x0, d1
    assert A.b2d(run([0x9E620020], x={1: -5}).v[0]) == -5.0                  # scvtf d0, x1
    a = 1.0 + 2.0 ** -30
    c = run([0x1F420C20], d={1: a, 2: a, 3: -(1.0 + 2.0 ** -29)})            # fmadd d0, d1, d2, d3
    assert A.b2d(c.v[0]) == 2.0 ** -60 and a * a - (1.0 + 2.0 ** -29) == 0.0  # one rounding, not two


def test_an_encoding_that_is_not_handled_is_refused_not_guessed():
    with pytest.raises(A.Unknown):
        run([0x00000000])


@needs_reference
def test_ntthal_under_the_interpreter_reproduces_delta_g_rs():
    from emu import run_ntthal
    gold = json.load(open(os.path.join(GOLDEN, "ntthal_delta_g_rs.json")))
    for g in gold:
        c = g["cond"]
        cond = ["-mv", "%.2f" % c["mv"], "-dv", "%.2f" % c["dv"], "-n", "%.2f" % c["dntp"], "-d", "%.2f" % c["dna"], "-t", "%.2f" % c["t"]]
        out, err, code, _ = run_ntthal(["-a", "ANY"] + cond + ["-s1", g["a"], "-s2", g["b"]])
        lines = out.split("\n")
        v = g["values"]
        assert code == 0 and err == ""
        assert lines[0] == "Calculated thermodynamical parameters for dimer:\tdS = %s\tdH = %s\tdG = %s\tt = %s" % (v["dS"], v["dH"], v["dG"], v["t"])
        assert [l.rstrip(" ") for l in lines[1:5]] == [t + "\t" + b for t, b in g["lines"]]
    # the reference's own protocol: -path <dir> -i, pairs on stdin (delta_g.rs:93-110)
    c = gold[2]["cond"]
    cond = ["-mv", "%.2f" % c["mv"], "-dv", "%.2f" % c["dv"], "-n", "%.2f" % c["dntp"], "-d", "%.2f" % c["dna"], "-t", "%.2f" % c["t"]]
    out, _, code, _ = run_ntthal(["-a", "ANY"] + cond + ["-path", "primer3_config/", "-i"],
                                 stdin="".join("%s,%s\n" % (g["a"], g["b"]) for g in gold[2:]).encode(),
                                 file_root="/root/reference/od-msspe/primer3_config")
    assert code == 0 and [l.split()[13] for l in out.split("\n")[0::5][:3]] == [g["values"]["dG"] for g in gold[2:]]


@needs_reference
def test_committed_ntthal_fixture_is_what_the_executable_prints():
    from emu import run_ntthal
    cases = json.load(open(os.path.join(GOLDEN, "ntthal_emulated.json")))["cases"]
    for c in cases[0:3] + cases[130:132] + cases[260:262] + cases[420:424]:
        out, _, code, _ = run_ntthal(c["args"], stdin=c["stdin"].encode())
        assert code == 0 and out == c["stdout"], c["args"]


@needs_reference
def test_primer3_core_under_the_interpreter_reproduces_primer_rs():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_primer3_core_emulated_golden as G
    fixture = json.load(open(os.path.join(GOLDEN, "primer3_core_emulated.json")))["primers"]
    pick = [fixture[0], fixture[150], fixture[230]]
    got, _ = G.run([p["primer"] for p in pick])
    assert got == pick
    assert (got[0]["TM"], got[0]["GC_PERCENT"], got[0]["SELF_ANY_TH"], got[0]["SELF_END_TH"], got[0]["HAIRPIN_TH"]) == \
        ("43.727", "53.846", "0.00", "0.00", "0.00")      # primer.rs:238-250
