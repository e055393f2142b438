"""The nearest-neighbour tables against REFERENCE-HELD ground truth: the arrays compiled into the Primer3 2.6.1 executables
the reference spawns (od-msspe/bin/primer3_core, primer.rs:125-140; od-msspe/bin/ntthal, delta_g.rs:90-108), read out of
their Mach-O data by symbol name (tools/extract_primer3_compiled_in_tables.py -> tests/golden/primer3_2_6_1_compiled_in_tables.json).

Those arrays are libprimer3's tables AFTER its loader ran: 5-symbol index (A,C,G,T,N), the joint-infinity rule, the -1.0 / +inf
and 1e-11 / 0 sentinels, the dangle3 index transposition, sorted tri/tetraloop keys.  Neither csrc/thal_params_data.inc (generated
from od-msspe/primer3_config/*) nor the two expansions (oracle/thal_oracle.c, csrc/thal_params.cu) were written from them, so a
transcription error in the shared .inc or a misread load rule can no longer be common-mode.  It also settles what
`primer3_core` computes with when the reference gives it no parameter path: exactly these tables."""
import ctypes as C
import json
import os
import struct

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

PLAIN = ["stackEntropies", "stackEnthalpies", "stackint2Entropies", "stackint2Enthalpies", "tstackEntropies", "tstackEnthalpies",
         "tstack2Entropies", "tstack2Enthalpies", "dangleEntropies3", "dangleEnthalpies3", "dangleEntropies5", "dangleEnthalpies5",
         "hairpinLoopEntropies", "interiorLoopEntropies", "bulgeLoopEntropies", "hairpinLoopEnthalpies", "interiorLoopEnthalpies",
         "bulgeLoopEnthalpies", "atpS", "atpH"]
KEYED = [("defaultTriloopEntropies", 5), ("defaultTriloopEnthalpies", 5), ("defaultTetraloopEntropies", 6),
         ("defaultTetraloopEnthalpies", 6)]


@pytest.fixture(scope="module")
def builtin():
    with open(os.path.join(GOLDEN, "primer3_2_6_1_compiled_in_tables.json")) as f:
        d = json.load(f)
    assert d["primer3_release_string_found"] is True
    dbl = {k: np.array([float(x) for x in v], dtype=np.float64) for k, v in d["doubles"].items()}
    loops = {k: [(s, float(x)) for s, x in v] for k, v in d["loops"].items()}
    return dbl, loops


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).tobytes()


def test_fixture_shape(builtin):
    dbl, loops = builtin
    assert sorted(dbl) == sorted(PLAIN) and sorted(loops) == sorted(k for k, _ in KEYED)
    assert sum(v.size for v in dbl.values()) == 5730
    assert [len(loops[k]) for k, _ in KEYED] == [16, 16, 77, 77]
    for k, ln in KEYED:
        keys = [s for s, _ in loops[k]]
        assert all(len(s) == ln for s in keys) and keys == sorted(keys)  # libprimer3 keeps them sorted for bsearch
    # the AT-closing penalty thal.c hard-codes (atpS / atpH)
    assert dbl["atpS"][3] == 6.9 == dbl["atpS"][15] and dbl["atpH"][3] == 2200.0 == dbl["atpH"][15]


def test_oracle_tables_equal_primer3_compiled_in(oracle_lib, builtin):
    dbl, loops = builtin
    L = oracle_lib.lib()
    L.oracle_thal_table.argtypes = [C.c_char_p, C.POINTER(C.c_double), C.c_int]
    L.oracle_thal_loop_table.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_double), C.c_int]
    for name in PLAIN:
        buf = (C.c_double * 625)()
        n = L.oracle_thal_table(name.encode(), buf, 625)
        assert n == dbl[name].size, name
        assert _bits(np.frombuffer(buf, dtype=np.float64, count=n)) == _bits(dbl[name]), name
    for name, ln in KEYED:
        keys = C.create_string_buffer(8 * 128)
        vals = (C.c_double * 128)()
        n = L.oracle_thal_loop_table(name.encode(), keys, vals, 128)
        got = [("".join("ACGT"[b] for b in keys.raw[8 * i: 8 * i + ln]), vals[i]) for i in range(n)]
        assert [(s, struct.pack("<d", v)) for s, v in got] == [(s, struct.pack("<d", v)) for s, v in loops[name]], name


def _engine_lib():
    import msspe_b200 as m
    if not os.path.exists(m.LIB_PATH):
        import subprocess
        subprocess.run([os.path.join(ROOT, "open-msspe-design_b200", "build.sh")], check=True)
    return m


def _engine_tables(m, raw):
    L = m.load_library()
    out = {}
    for name in PLAIN:
        buf = (C.c_double * 625)()
        n = L.msspe_thal_expanded_table(raw, name.encode(), buf, 625)
        assert n > 0, (name, n)
        out[name] = np.frombuffer(buf, dtype=np.float64, count=n).copy()
    for name, ln in KEYED:
        buf = (C.c_double * 256)()
        n = L.msspe_thal_expanded_table(raw, name.encode(), buf, 256)
        assert n >= 0 and n % 2 == 0, (name, n)
        ents = []
        for i in range(n // 2):
            key, s = int(buf[2 * i]), ""
            for _ in range(ln):
                s = "ACGTN"[key % 5] + s
                key //= 5
            ents.append((s, buf[2 * i + 1]))
        out[name] = ents
    return out


def test_engine_tables_equal_primer3_compiled_in(builtin):
    """The arrays the CUDA kernels index (host-side expansion, no device needed) == Primer3 2.6.1's compiled-in arrays."""
    dbl, loops = builtin
    m = _engine_lib()
    L = m.load_library()
    raw = C.create_string_buffer(m.RAW_PARAMS_BYTES)
    assert L.msspe_thal_params_default(raw) == 0
    got = _engine_tables(m, raw)
    for name in PLAIN:
        assert _bits(got[name]) == _bits(dbl[name]), name
    for name, _ in KEYED:
        assert [(s, struct.pack("<d", v)) for s, v in got[name]] == [(s, struct.pack("<d", v)) for s, v in loops[name]], name


def test_engine_expanded_table_errors_and_directory_source(builtin, tmp_path):
    """Unknown name / short buffer are errors; a primer3_config directory written from the embedded tables expands to the
    same arrays (the `ntthal -path` route of delta_g.rs:90), a perturbed one does not."""
    from test_abi import write_param_dir
    dbl, _ = builtin
    m = _engine_lib()
    L = m.load_library()
    raw = C.create_string_buffer(m.RAW_PARAMS_BYTES)
    assert L.msspe_thal_params_default(raw) == 0
    buf = (C.c_double * 625)()
    assert L.msspe_thal_expanded_table(raw, b"noSuchTable", buf, 625) == m.ERR_INVALID
    assert L.msspe_thal_expanded_table(raw, b"stackEntropies", buf, 624) == m.ERR_CAPACITY
    assert L.msspe_thal_expanded_table(None, b"stackEntropies", buf, 625) == m.ERR_INVALID
    write_param_dir(m, tmp_path / "same")
    write_param_dir(m, tmp_path / "shifted", stack_ds_shift=0.5)
    for sub, equal in (("same", True), ("shifted", False)):
        r = C.create_string_buffer(m.RAW_PARAMS_BYTES)
        err = C.create_string_buffer(256)
        assert L.msspe_thal_params_from_dir(str(tmp_path / sub).encode(), r, err, 256) == 0, err.value
        got = _engine_tables(m, r)
        assert (_bits(got["stackEntropies"]) == _bits(dbl["stackEntropies"])) is equal
        assert _bits(got["tstack2Enthalpies"]) == _bits(dbl["tstack2Enthalpies"])
