"""oracle.py -- ctypes access to oracle/liboracle.so.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class ThalCond(C.Structure):  # msspe_thal_cond
    _fields_ = [("mv", C.c_double), ("dv", C.c_double), ("dntp", C.c_double), ("dna_conc", C.c_double),
                ("temp_c", C.c_double), ("max_loop", C.c_int32), ("reserved", C.c_int32)]


class ThalOut(C.Structure):  # msspe_thal_out
    _fields_ = [("ds", C.c_double), ("dh", C.c_double), ("dg", C.c_double), ("tm", C.c_double),
                ("no_structure", C.c_int32), ("n_bp", C.c_int32)]


class OracleConfig(C.Structure):  # kmer_oracle.cpp Config == the od-msspe CLI flags (config.rs:11-148)
    _fields_ = [("kmer_size", C.c_uint64), ("window_size", C.c_uint64), ("overlap_size", C.c_uint64),
                ("max_mismatch_segments", C.c_uint64), ("max_iterations", C.c_uint64),
                ("search_windows_size", C.c_uint64),
                ("mv_conc", C.c_float), ("dv_conc", C.c_float), ("dntp_conc", C.c_float), ("dna_conc", C.c_float),
                ("annealing_temp", C.c_float), ("min_tm", C.c_float), ("max_tm", C.c_float),
                ("max_self_dimer_any_tm", C.c_float), ("max_self_dimer_end_tm", C.c_float),
                ("max_hairpin_tm", C.c_float), ("delta_g_threshold", C.c_float), ("tm_stddev", C.c_float),
                ("keep_all", C.c_int), ("check_cross_dimers", C.c_int), ("check_self_dimers", C.c_int),
                ("check_hairpin", C.c_int), ("disable_tm_stddev", C.c_int), ("disable_min_max_tm", C.c_int)]


def default_config(**kw) -> OracleConfig:
    """constants.rs:1-26 / config.rs defaults.  max_mismatch_segments 0 = auto (main.rs:658-660)."""
    c = OracleConfig(13, 500, 250, 0, 1000, 50, 50.0, 3.0, 0.0, 250.0, 25.0, 30.0, 60.0, 47.0, 47.0, 24.0,
                     -9000.0, 2.0, 0, 1, 1, 1, 0, 0)
    for k, v in kw.items():
        assert hasattr(c, k), k
        setattr(c, k, v)
    return c


def build() -> str:
    subprocess.run(["make", "-s", "-C", _HERE], check=True)
    return os.path.join(_HERE, "liboracle.so")


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.oracle_oligotm.restype = C.c_double
        L.oracle_oligotm.argtypes = [C.c_char_p] + [C.c_double] * 4
        L.oracle_gc_percent.restype = C.c_double
        L.oracle_gc_percent.argtypes = [C.c_char_p]
        L.oracle_thal.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(ThalCond), C.POINTER(ThalOut)]
        L.oracle_thal_stats.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(ThalCond), C.POINTER(C.c_long)]
        L.oracle_thal_load_dir.argtypes = [C.c_char_p]
        L.oracle_pipeline_run.restype = C.c_void_p
        L.oracle_pipeline_run.argtypes = [C.c_char_p, C.c_uint64, C.POINTER(OracleConfig), C.c_int]
        L.oracle_pipeline_free.argtypes = [C.c_void_p]
        for f in ("oracle_pipeline_csv", "oracle_pipeline_report"):
            getattr(L, f).restype = C.c_char_p
            getattr(L, f).argtypes = [C.c_void_p]
        L.oracle_pipeline_count.restype = C.c_uint64
        L.oracle_pipeline_count.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.oracle_pipeline_word.restype = C.c_char_p
        L.oracle_pipeline_word.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint64]
        L.oracle_pipeline_candidate.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.POINTER(C.c_uint64),
                                                C.POINTER(C.c_uint64), C.POINTER(C.c_float)]
        L.oracle_pipeline_stat.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.POINTER(C.c_float), C.POINTER(C.c_int)]
        L.oracle_pipeline_timing.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
        L.oracle_segment_slots.restype = C.c_uint64
        L.oracle_segment_slots.argtypes = [C.c_char_p, C.c_uint64] + [C.c_uint64] * 4 + [C.c_int, C.c_void_p,
                                                                                     C.c_uint64, C.c_void_p]
        L.oracle_select.restype = C.c_uint64
        L.oracle_select.argtypes = [C.c_char_p, C.c_uint64] + [C.c_uint64] * 4 + [C.c_int, C.c_uint64, C.c_uint64,
                                                                              C.c_void_p, C.c_void_p, C.c_void_p,
                                                                              C.c_void_p, C.c_uint64,
                                                                              C.POINTER(C.c_uint64),
                                                                              C.POINTER(C.c_double)]
        L.oracle_manager_create.restype = C.c_void_p
        L.oracle_manager_create.argtypes = [C.c_char_p, C.c_uint64] + [C.c_uint64] * 4
        L.oracle_manager_segments.restype = C.c_uint64
        L.oracle_manager_segments.argtypes = [C.c_void_p]
        L.oracle_manager_select.restype = C.c_uint64
        L.oracle_manager_select.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64,
                                            C.POINTER(C.c_uint64), C.POINTER(C.c_double)]
        L.oracle_manager_free.argtypes = [C.c_void_p]
        L.oracle_thal_load_embedded()
        _LIB = L
    return _LIB


def thal(o1: str, o2: str, ttype: int, cond: ThalCond) -> ThalOut:
    out = ThalOut()
    rc = lib().oracle_thal(o1.encode(), (o2 or o1).encode(), ttype, C.byref(cond), C.byref(out))
    assert rc == 0, rc
    return out


def thal_last_pairing(n: int):
    """Partner (1-based, in the reversed second oligo; 0 = unpaired) of every base of the first oligo in the duplex the most
    recent dimer `thal` call traced; after a HAIRPIN call, the partner position (1-based, same oligo) in the traced fold: what
    ntthal draws."""
    L = lib()
    buf = (C.c_int * 64)()
    L.oracle_thal_last_pairing(buf, 64)
    return list(buf)[:n]


def oligotm(s: str, mv=50.0, dv=1.5, dntp=0.6, dna=50.0) -> float:
    return lib().oracle_oligotm(s.encode(), mv, dv, dntp, dna)


def gc_percent(s: str) -> float:
    return lib().oracle_gc_percent(s.encode())


class PipelineResult:
    def __init__(self, handle):
        self._h = handle
        L = lib()
        self.csv = L.oracle_pipeline_csv(handle).decode()
        self.report = L.oracle_pipeline_report(handle).decode()
        self.candidates, self.filtered, self.final, self.stats = [[], []], [[], []], [[], []], [[], []]
        for d in (0, 1):
            n = L.oracle_pipeline_count(handle, d, 0)
            for i in range(n):
                f, t, s = C.c_uint64(), C.c_uint64(), C.c_float()
                L.oracle_pipeline_candidate(handle, d, i, C.byref(f), C.byref(t), C.byref(s))
                self.candidates[d].append((L.oracle_pipeline_word(handle, d, 0, i).decode(), f.value, t.value, s.value))
            for stage, dst in ((1, self.filtered), (2, self.final)):
                dst[d] = [L.oracle_pipeline_word(handle, d, stage, i).decode()
                          for i in range(L.oracle_pipeline_count(handle, d, stage))]
        t = (C.c_double * 5)()
        c = (C.c_uint64 * 4)()
        L.oracle_pipeline_timing(handle, t, c)
        self.seconds = dict(segments=t[0], select_fwd=t[1], select_rev=t[2], thermo=t[3], dimer=t[4])
        self.evals = (c[0], c[1])
        self.n_pairs = c[2]
        self.n_segments = c[3]

    def load_stats(self):
        L = lib()
        for d in (0, 1):
            self.stats[d] = []
            for i in range(len(self.candidates[d])):
                v = (C.c_float * 7)()
                fl = (C.c_int * 2)()
                L.oracle_pipeline_stat(self._h, d, i, v, fl)
                self.stats[d].append(dict(tm=v[0], gc=v[1], self_any=v[2], self_end=v[3], hairpin=v[4], mean=v[5],
                                          std=v[6], tm_ok=bool(fl[0]), runs=bool(fl[1])))
        return self.stats

    def close(self):
        if self._h:
            lib().oracle_pipeline_free(self._h)
            self._h = None


def run_pipeline(fasta: bytes, cfg: OracleConfig | None = None, stop_after: int = 0) -> PipelineResult:
    cfg = cfg or default_config()
    h = lib().oracle_pipeline_run(fasta, len(fasta), C.byref(cfg), stop_after)
    if not h:
        raise ValueError("No sequences found in the input file")
    r = PipelineResult(h)
    if stop_after == 0:
        r.load_stats()
    return r


def segment_slots(fasta: bytes, W, S, w, k, direction):
    import numpy as np
    L = lib()
    n = L.oracle_segment_slots(fasta, len(fasta), W, S, w, k, direction, None, 0, None)
    slots = max(0, w - k + 1)
    codes = np.empty(max(1, n * slots), dtype=np.uint64)
    part = np.empty(max(1, n), dtype=np.uint16)
    L.oracle_segment_slots(fasta, len(fasta), W, S, w, k, direction, codes.ctypes.data, n * slots, part.ctypes.data)
    return codes[:n * slots].reshape(n, slots), part[:n]


def select(fasta: bytes, W, S, w, k, direction, max_iter, mms):
    import numpy as np
    L = lib()
    codes = np.zeros(max_iter, dtype=np.uint64)
    freqs = np.zeros(max_iter, dtype=np.uint32)
    tied = np.zeros(max_iter, dtype=np.uint32)
    scores = np.zeros(max_iter, dtype=np.float32)
    ev, sec = C.c_uint64(), C.c_double()
    n = L.oracle_select(fasta, len(fasta), W, S, w, k, direction, max_iter, mms, codes.ctypes.data, freqs.ctypes.data,
                        tied.ctypes.data, scores.ctypes.data, max_iter, C.byref(ev), C.byref(sec))
    return dict(codes=codes[:n], freqs=freqs[:n], n_tied=tied[:n], scores=scores[:n], evals=ev.value, seconds=sec.value)


class Manager:
    """get_segment_manager (main.rs:196-235) built once and kept: repeated find_candidates_kmers calls on the same input."""

    def __init__(self, fasta: bytes, W, S, w, k):
        self._h = lib().oracle_manager_create(fasta, len(fasta), W, S, w, k)
        self.n_segments = lib().oracle_manager_segments(self._h)

    def select(self, direction, max_iter, mms):
        import numpy as np
        codes = np.zeros(max(1, max_iter), dtype=np.uint64)
        freqs = np.zeros(max(1, max_iter), dtype=np.uint32)
        ev, sec = C.c_uint64(), C.c_double()
        n = lib().oracle_manager_select(self._h, direction, max_iter, mms, codes.ctypes.data, freqs.ctypes.data, max_iter,
                                        C.byref(ev), C.byref(sec))
        return dict(codes=codes[:n], freqs=freqs[:n], evals=ev.value, seconds=sec.value)

    def close(self):
        if self._h:
            lib().oracle_manager_free(self._h)
            self._h = None
