#!/usr/bin/env python3
"""Reference-held ground truth for the nearest-neighbour tables: the parameter arrays COMPILED INTO the reference's own
Primer3 2.6.1 executables (od-msspe/bin/primer3_core and od-msspe/bin/ntthal, Mach-O arm64, which primer.rs:125-140 and
delta_g.rs:90-108 spawn).  They cannot be executed in this image, but their symbol tables are intact: thal.c's globals
(_stackEntropies, _tstack2Enthalpies, _defaultTetraloopEntropies, ...) are initialised data, i.e. the tables exactly as
Primer3's loader leaves them (5-symbol index A,C,G,T,N; the joint-infinity rule; -1.0 / +inf and 1e-11 / 0 sentinels;
the dangle3 index transposition; tri/tetraloop keys as base indices, sorted).

This script reads the arrays BY SYMBOL NAME (a small Mach-O 64 reader: LC_SEGMENT_64 sections + LC_SYMTAB), checks that
both executables hold bit-identical tables, and writes tests/golden/primer3_2_6_1_compiled_in_tables.json.  The fixture pins
  * csrc/thal_params_data.inc (generated from od-msspe/primer3_config/*.ds,*.dh) and
  * the expansion rules of the oracle (oracle/thal_oracle.c) and of the engine (csrc/thal_params.cu)
to something neither of them was written from (tests/test_oracle_thermo.py, tests/test_abi.py).

Run here (the GPU box has no /root/reference):  python tools/extract_primer3_compiled_in_tables.py
"""
import json
import os
import struct
import sys

REF_BIN = "/root/reference/od-msspe/bin"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden",
                   "primer3_2_6_1_compiled_in_tables.json")

# thal.c globals: name -> number of doubles
DOUBLE_TABLES = {
    "stackEntropies": 625, "stackEnthalpies": 625, "stackint2Entropies": 625, "stackint2Enthalpies": 625,
    "tstackEntropies": 625, "tstackEnthalpies": 625, "tstack2Entropies": 625, "tstack2Enthalpies": 625,
    "dangleEntropies3": 125, "dangleEnthalpies3": 125, "dangleEntropies5": 125, "dangleEnthalpies5": 125,
    "hairpinLoopEntropies": 30, "interiorLoopEntropies": 30, "bulgeLoopEntropies": 30,
    "hairpinLoopEnthalpies": 30, "interiorLoopEnthalpies": 30, "bulgeLoopEnthalpies": 30,
    "atpS": 25, "atpH": 25,
}
# struct { char loop[5 or 6]; double value; } = 16 bytes per entry; counts in numTriloops / numTetraloops
LOOP_TABLES = {
    "defaultTriloopEntropies": ("numTriloops", 5), "defaultTriloopEnthalpies": ("numTriloops", 5),
    "defaultTetraloopEntropies": ("numTetraloops", 6), "defaultTetraloopEnthalpies": ("numTetraloops", 6),
}


class MachO:
    def __init__(self, path):
        self.b = b = open(path, "rb").read()
        magic, = struct.unpack_from("<I", b, 0)
        if magic != 0xFEEDFACF:
            raise SystemExit("%s: not a thin little-endian Mach-O 64 (magic %#x)" % (path, magic))
        ncmds, = struct.unpack_from("<I", b, 16)
        off, self.sections, symtab = 32, [], None
        for _ in range(ncmds):
            cmd, size = struct.unpack_from("<II", b, off)
            if cmd == 0x19:  # LC_SEGMENT_64
                nsects, = struct.unpack_from("<I", b, off + 64)
                for i in range(nsects):
                    so = off + 72 + i * 80
                    addr, sz = struct.unpack_from("<QQ", b, so + 32)
                    foff, = struct.unpack_from("<I", b, so + 48)
                    flags, = struct.unpack_from("<I", b, so + 64)
                    self.sections.append((addr, sz, foff, (flags & 0xFF) == 1))  # S_ZEROFILL
            elif cmd == 0x2:  # LC_SYMTAB
                symtab = struct.unpack_from("<IIII", b, off + 8)
            off += size
        symoff, nsyms, stroff, _ = symtab
        self.syms = {}
        for i in range(nsyms):
            strx, _typ, _sect, _desc, val = struct.unpack_from("<IBBHQ", b, symoff + 16 * i)
            name = b[stroff + strx: b.index(b"\0", stroff + strx)].decode(errors="replace")
            self.syms[name] = val

    def read(self, name, nbytes):
        addr = self.syms["_" + name]
        for a, sz, foff, zerofill in self.sections:
            if a <= addr and addr + nbytes <= a + sz:
                return bytes(nbytes) if zerofill else self.b[foff + addr - a: foff + addr - a + nbytes]
        raise SystemExit("symbol %s not inside a section" % name)


def extract(path):
    m = MachO(path)
    out = {"doubles": {}, "loops": {}}
    for name, n in DOUBLE_TABLES.items():
        out["doubles"][name] = list(struct.unpack("<%dd" % n, m.read(name, 8 * n)))
    for name, (count_sym, ln) in LOOP_TABLES.items():
        n, = struct.unpack("<i", m.read(count_sym, 4))
        raw = m.read(name, 16 * n)
        ents = []
        for i in range(n):
            key = raw[16 * i: 16 * i + ln]
            assert all(c < 4 for c in key), (name, i, key)
            val, = struct.unpack_from("<d", raw, 16 * i + 8)
            ents.append(["".join("ACGT"[c] for c in key), val])
        out["loops"][name] = ents
    return out, m


def main():
    core, m = extract(os.path.join(REF_BIN, "primer3_core"))
    ntthal, _ = extract(os.path.join(REF_BIN, "ntthal"))
    same = json.dumps(core, sort_keys=True) == json.dumps(ntthal, sort_keys=True)
    if not same:
        raise SystemExit("primer3_core and ntthal hold different compiled-in tables")
    rel = b"libprimer3 release 2.6.1" in m.b
    enc = lambda v: "inf" if v == float("inf") else "-inf" if v == float("-inf") else repr(v)
    doc = {
        "source": "od-msspe/bin/primer3_core and od-msspe/bin/ntthal (Mach-O arm64), initialised data read by symbol name; "
                  "both executables bit-identical",
        "primer3_release_string_found": rel,
        "doubles": {k: [enc(x) for x in v] for k, v in core["doubles"].items()},
        "loops": {k: [[s, enc(x)] for s, x in v] for k, v in core["loops"].items()},
    }
    with open(OUT, "w") as f:
        json.dump(doc, f, separators=(",", ":"))
        f.write("\n")
    print("wrote", os.path.relpath(OUT), os.path.getsize(OUT), "bytes;",
          sum(len(v) for v in core["doubles"].values()), "doubles,",
          sum(len(v) for v in core["loops"].values()), "loop entries; release 2.6.1 string:", rel)


if __name__ == "__main__":
    sys.exit(main())
