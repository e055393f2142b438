"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, exports every declared symbol,
and refuses to compute without a GPU (no fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import msspe_b200 as m
    if not os.path.exists(m.LIB_PATH):
        subprocess.run([os.path.join(ROOT, "open-msspe-design_b200", "build.sh")], check=True)
    return m


def test_header_symbols_all_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "od_msspe_b200.h")).read()
    declared = set(re.findall(r"\b(msspe_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(lib.ABI_SYMBOLS)
    L = lib.load_library()
    for s in declared:
        assert hasattr(L, s), s
    assert L.msspe_abi_version() == 2


def test_struct_layouts_match_header(lib, tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "od_msspe_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   "sizeof(msspe_thal_raw_params),sizeof(msspe_timing),sizeof(msspe_candidate),sizeof(msspe_thal_out),"
                   "sizeof(msspe_config),sizeof(msspe_thal_cond),sizeof(msspe_dimer_edge),sizeof(msspe_kmer_stat),sizeof(msspe_filter_cfg));}\n")
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [lib.RAW_PARAMS_BYTES, C.sizeof(lib.Timing), C.sizeof(lib.Candidate), C.sizeof(lib.ThalOut), C.sizeof(lib.Config),
            C.sizeof(lib.ThalCond), C.sizeof(lib.DimerEdge), lib.KMER_STAT_DTYPE.itemsize, C.sizeof(lib.FilterCfg)]
    assert got == want
    assert lib.CANDIDATE_DTYPE.itemsize == got[2] and lib.THAL_OUT_DTYPE.itemsize == got[3] and lib.EDGE_DTYPE.itemsize == got[6]


def test_bad_config_is_an_error_not_a_panic(lib):
    L = lib.load_library()
    h = C.c_void_p()
    # overlap (step) < search window: main.rs:201-203 panics; the ABI returns MSSPE_ERR_INVALID
    cfg = lib.Config(13, 500, 40, 50, 0, 0)
    assert L.msspe_create(C.byref(cfg), C.byref(h)) == lib.ERR_INVALID
    assert b"Overlap windows size" in L.msspe_last_error(None)
    cfg = lib.Config(40, 500, 250, 50, 0, 0)
    assert L.msspe_create(C.byref(cfg), C.byref(h)) == lib.ERR_INVALID


def test_no_cpu_fallback_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(lib.MsspeError) as e:
        lib.Engine()
    assert e.value.code == lib.ERR_CUDA and "no CPU fallback" in str(e.value)


def test_product_never_references_the_oracle():
    pkg = os.path.join(ROOT, "open-msspe-design_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".cu", ".cuh", ".cpp", ".h", ".py", ".sh")):
                txt = open(os.path.join(dp, fn), errors="ignore").read()
                assert "oracle" not in txt.lower() or fn == "thal.cu" and "oracle" not in txt, (dp, fn)


def write_param_dir(lib, path, stack_ds_shift=0.0):
    """The embedded tables written out as a primer3_config directory (stack.ds optionally shifted: a directory that
    differs from Primer3's compiled-in tables).  Returns the raw embedded block."""
    L = lib.load_library()
    a = C.create_string_buffer(lib.RAW_PARAMS_BYTES)
    assert L.msspe_thal_params_default(a) == 0
    import numpy as np
    d = np.frombuffer(a.raw[:8 * 2484], dtype=np.float64)
    names = [("stack.ds", 256), ("stack.dh", 256), ("stackmm.ds", 256), ("stackmm.dh", 256), ("dangle.ds", 128),
             ("dangle.dh", 128), ("loops.ds", 90), ("loops.dh", 90), ("tstack_tm_inf.ds", 256), ("tstack.dh", 256),
             ("tstack2.ds", 256), ("tstack2.dh", 256)]
    pos = 0
    fmt = lambda v: "inf" if np.isinf(v) else repr(float(v))
    os.makedirs(str(path), exist_ok=True)
    for fn, n in names:
        vals = d[pos:pos + n].copy()
        pos += n
        if fn == "stack.ds" and stack_ds_shift:
            vals[np.isfinite(vals)] += stack_ds_shift
        with open(os.path.join(str(path), fn), "w") as f:
            if fn.startswith("loops"):
                for r in range(30):
                    f.write("%d\t%s\n" % (r + 1, "\t".join(fmt(v) for v in vals[3 * r:3 * r + 3])))
            else:
                f.write("\n".join(fmt(v) for v in vals) + "\n")
    raw = a.raw
    off = 8 * 2484
    for base, cap, ln in (("triloop", 32, 5), ("tetraloop", 128, 6)):
        for ext in ("ds", "dh"):
            n = int.from_bytes(raw[off:off + 4], "little")
            seqs = raw[off + 4: off + 4 + cap * 8]
            vals = np.frombuffer(raw[off + 8 + cap * 8: off + 8 + cap * 16], dtype=np.float64)
            with open(os.path.join(str(path), "%s.%s" % (base, ext)), "w") as f:
                for i in range(n):
                    f.write("%s\t%s\n" % (seqs[8 * i:8 * i + ln].decode(), fmt(vals[i])))
            off += 8 + cap * 16
    return a


def test_embedded_params_equal_a_directory_load(lib, tmp_path):
    """msspe_thal_params_from_dir on files regenerated from the embedded tables must round-trip."""
    L = lib.load_library()
    a = write_param_dir(lib, tmp_path)
    b = C.create_string_buffer(lib.RAW_PARAMS_BYTES)
    err = C.create_string_buffer(256)
    assert L.msspe_thal_params_from_dir(str(tmp_path).encode(), b, err, 256) == 0, err.value
    assert a.raw == b.raw
