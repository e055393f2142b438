"""msspe_b200 -- ctypes binding of libodmsspe_b200.so (the C ABI in include/od_msspe_b200.h).

This is the Python-side mirror of the reference's call sites (od-msspe/src/main.rs:693-752); the product is
the shared library, this module only marshals buffers.  There is no CPU fallback: if the CUDA library is
missing or no device is usable, every call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MSSPE_LIB") or os.path.join(os.path.dirname(_PKG), "libodmsspe_b200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_NOMEM, ERR_STATE, ERR_IO, ERR_CAPACITY = 0, -1, -2, -3, -4, -5, -6
DIR_FWD, DIR_REV = 0, 1
NO_KMER = np.uint64(0xFFFFFFFFFFFFFFFF)
SELECT_RECOUNT, SELECT_INCREMENTAL, SELECT_AUTO, SELECT_PARTITIONED, SELECT_BATCHED = 0, 1, 2, 3, 0x100
THAL_ANY, THAL_END1, THAL_HAIRPIN = 1, 2, 4

# every symbol include/od_msspe_b200.h declares
ABI_SYMBOLS = [
    "msspe_abi_version", "msspe_create", "msspe_destroy", "msspe_last_error", "msspe_set_stream", "msspe_reserve_pool", "msspe_synchronize",
    "msspe_get_timing", "msspe_reset_timing", "msspe_set_profiling", "msspe_load_genomes", "msspe_load_genomes_device",
    "msspe_build_index", "msspe_segment_info", "msspe_get_segment_kmers", "msspe_get_index", "msspe_select",
    "msspe_select_both", "msspe_coverage", "msspe_thal_params_default", "msspe_thal_params_from_dir",
    "msspe_set_thal_params", "msspe_primer_thermo", "msspe_thal_pairs", "msspe_thal_pairs_aligned", "msspe_cross_dimer",
    "msspe_fasta_open", "msspe_fasta_close", "msspe_fasta_records", "msspe_fasta_name", "msspe_fasta_bases",
    "msspe_fasta_offsets", "msspe_load_fasta", "msspe_kmer_stats", "msspe_kmer_stats_both", "msspe_coverage_summary", "msspe_vertex_cover", "msspe_shard_begin", "msspe_shard_buffers", "msspe_shard_count", "msspe_shard_firstpos", "msspe_shard_apply",
    "msspe_get_kernel_profile", "msspe_thal_expanded_table", "msspe_cross_dimer_device", "msspe_dist_unique_id", "msspe_dist_init", "msspe_select_both_dist",
]


class Config(C.Structure):
    _fields_ = [("kmer_size", C.c_uint32), ("window_size", C.c_uint32), ("overlap_size", C.c_uint32),
                ("search_windows_size", C.c_uint32), ("device", C.c_int32), ("flags", C.c_uint32)]


class Candidate(C.Structure):
    _fields_ = [("code", C.c_uint64), ("freq", C.c_uint32), ("n_tied", C.c_uint32), ("tie_score", C.c_float),
                ("reserved", C.c_uint32)]


class ThalCond(C.Structure):
    _fields_ = [("mv", C.c_double), ("dv", C.c_double), ("dntp", C.c_double), ("dna_conc", C.c_double),
                ("temp_c", C.c_double), ("max_loop", C.c_int32), ("reserved", C.c_int32)]


class ThalOut(C.Structure):
    _fields_ = [("ds", C.c_double), ("dh", C.c_double), ("dg", C.c_double), ("tm", C.c_double),
                ("no_structure", C.c_int32), ("n_bp", C.c_int32)]


class DimerEdge(C.Structure):
    _fields_ = [("pair", C.c_uint64), ("dg", C.c_double)]


class FilterCfg(C.Structure):
    _fields_ = [("min_tm", C.c_float), ("max_tm", C.c_float), ("max_self_dimer_any_tm", C.c_float),
                ("max_self_dimer_end_tm", C.c_float), ("max_hairpin_tm", C.c_float), ("tm_stddev", C.c_float),
                ("check_self_dimers", C.c_uint8), ("check_hairpin", C.c_uint8), ("disable_tm_stddev", C.c_uint8),
                ("disable_min_max_tm", C.c_uint8)]


def default_filter_cfg(**kw) -> "FilterCfg":
    """constants.rs:15-20 and the config.rs defaults."""
    f = FilterCfg(30.0, 60.0, 47.0, 47.0, 24.0, 2.0, 1, 1, 0, 0)
    for k, v in kw.items():
        assert hasattr(f, k), k
        setattr(f, k, v)
    return f


KMER_STAT_DTYPE = np.dtype([("code", "<u8"), ("tm", "<f4"), ("gc_percent", "<f4"), ("self_any_th", "<f4"), ("self_end_th", "<f4"),
                            ("hairpin_th", "<f4"), ("mean", "<f4"), ("std", "<f4"), ("tm_ok", "u1"), ("runs", "u1"), ("keep", "u1"),
                            ("reserved", "u1")])


class Timing(C.Structure):
    _fields_ = [("h2d_ms", C.c_float), ("encode_ms", C.c_float), ("index_ms", C.c_float), ("select_ms", C.c_float * 2),
                ("thermo_ms", C.c_float), ("dimer_ms", C.c_float), ("select_evals", C.c_uint64 * 2),
                ("select_postings_read", C.c_uint64 * 2), ("select_iterations", C.c_uint32 * 2),
                ("kernel_launches", C.c_uint32), ("count_kernel_ms", C.c_float * 2),
                ("count_kernel_launches", C.c_uint32 * 2)]


KERNEL_PROF_DTYPE = np.dtype([("name", "S56"), ("ms", "<f4"), ("launches", "<u4"), ("alg_bytes", "<u8")])
CANDIDATE_DTYPE = np.dtype([("code", "<u8"), ("freq", "<u4"), ("n_tied", "<u4"), ("tie_score", "<f4"), ("reserved", "<u4")])
THAL_OUT_DTYPE = np.dtype([("ds", "<f8"), ("dh", "<f8"), ("dg", "<f8"), ("tm", "<f8"), ("no_structure", "<i4"), ("n_bp", "<i4")])
EDGE_DTYPE = np.dtype([("pair", "<u8"), ("dg", "<f8")])

_lib = None


class MsspeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("msspe error %d: %s" % (code, msg))
        self.code = code


def load_library():
    """dlopen the CUDA engine; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MsspeError(ERR_IO, "%s not built -- run open-msspe-design_b200/build.sh (no CPU fallback exists)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.msspe_last_error.restype = C.c_char_p
    L.msspe_last_error.argtypes = [C.c_void_p]
    L.msspe_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
    L.msspe_destroy.argtypes = [C.c_void_p]
    L.msspe_destroy.restype = None
    L.msspe_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    L.msspe_reserve_pool.argtypes = [C.c_void_p, C.c_uint64]
    L.msspe_synchronize.argtypes = [C.c_void_p]
    L.msspe_get_timing.argtypes = [C.c_void_p, C.POINTER(Timing)]
    L.msspe_reset_timing.argtypes = [C.c_void_p]
    L.msspe_set_profiling.argtypes = [C.c_void_p, C.c_int]
    L.msspe_load_genomes.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
    L.msspe_load_genomes_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
    L.msspe_build_index.argtypes = [C.c_void_p]
    L.msspe_segment_info.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.msspe_get_segment_kmers.argtypes = [C.c_void_p, C.c_uint8, C.c_void_p, C.c_uint64]
    L.msspe_get_index.argtypes = [C.c_void_p, C.c_uint8, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_void_p,
                                  C.c_void_p, C.c_void_p]
    L.msspe_select.argtypes = [C.c_void_p, C.c_uint8, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint32)]
    L.msspe_select_both.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint32),
                                    C.c_void_p, C.POINTER(C.c_uint32)]
    L.msspe_fasta_open.argtypes = [C.c_char_p, C.c_uint32, C.POINTER(C.c_void_p), C.c_char_p, C.c_size_t]
    L.msspe_fasta_close.argtypes = [C.c_void_p]
    L.msspe_fasta_close.restype = None
    L.msspe_fasta_records.argtypes = [C.c_void_p]
    L.msspe_fasta_records.restype = C.c_uint32
    L.msspe_fasta_name.argtypes = [C.c_void_p, C.c_uint32]
    L.msspe_fasta_name.restype = C.c_char_p
    L.msspe_fasta_bases.argtypes = [C.c_void_p]
    L.msspe_fasta_bases.restype = C.c_void_p
    L.msspe_fasta_offsets.argtypes = [C.c_void_p]
    L.msspe_fasta_offsets.restype = C.c_void_p
    L.msspe_load_fasta.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, C.POINTER(C.c_void_p)]
    L.msspe_coverage_summary.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p,
                                         C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint64)]
    L.msspe_vertex_cover.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p,
                                     C.POINTER(C.c_uint32)]
    L.msspe_coverage.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_uint64]
    L.msspe_thal_params_default.argtypes = [C.c_void_p]
    L.msspe_thal_expanded_table.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_double), C.c_uint32]
    L.msspe_thal_params_from_dir.argtypes = [C.c_char_p, C.c_void_p, C.c_char_p, C.c_size_t]
    L.msspe_set_thal_params.argtypes = [C.c_void_p, C.c_void_p]
    L.msspe_primer_thermo.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32] + [C.c_void_p] * 5
    L.msspe_thal_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_int32,
                                   C.POINTER(ThalCond), C.c_void_p]
    L.msspe_cross_dimer.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(ThalCond), C.c_uint32,
                                    C.c_uint32, C.c_double, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_void_p,
                                    C.c_uint64, C.POINTER(C.c_uint64)]
    L.msspe_cross_dimer_device.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(ThalCond), C.c_uint32, C.c_uint32,
                                           C.c_double, C.c_uint64, C.c_uint64, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64),
                                           C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    L.msspe_dist_unique_id.argtypes = [C.c_void_p]
    L.msspe_dist_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    L.msspe_select_both_dist.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint32), C.c_void_p, C.POINTER(C.c_uint32)]
    L.msspe_get_kernel_profile.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
    L.msspe_kmer_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(FilterCfg), C.c_void_p]
    L.msspe_kmer_stats_both.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(FilterCfg), C.c_void_p, C.c_void_p]
    L.msspe_shard_begin.argtypes = [C.c_void_p, C.c_uint8]
    L.msspe_shard_buffers.argtypes = [C.c_void_p, C.c_uint8, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    L.msspe_shard_count.argtypes = [C.c_void_p, C.c_uint8, C.POINTER(C.c_uint64)]
    L.msspe_shard_firstpos.argtypes = [C.c_void_p, C.c_uint8, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
    L.msspe_shard_apply.argtypes = [C.c_void_p, C.c_uint8, C.c_uint32, C.c_uint32, C.c_void_p]
    _lib = L
    return L


RAW_PARAMS_BYTES = 25024  # sizeof(msspe_thal_raw_params); checked by tests/test_abi.py


def encode_word(word: str) -> int:
    v = 0
    for ch in word:
        v = (v << 2) | "ACGT".index(ch)
    return v


def decode_word(code: int, k: int) -> str:
    return "".join("ACGT"[(int(code) >> (2 * (k - 1 - i))) & 3] for i in range(k))


def pack_records(seqs) -> tuple[np.ndarray, np.ndarray]:
    """Concatenate record sequences (bytes) and build the offsets array msspe_load_genomes expects."""
    offs = np.zeros(len(seqs) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
    bases = np.frombuffer(b"".join(seqs), dtype=np.uint8).copy() if seqs else np.zeros(0, dtype=np.uint8)
    return bases, offs


class Fasta:
    """A parsed FASTA file (to_records, main.rs:108-122): names, offsets[n+1] and the normalised bases."""

    def __init__(self, L, handle):
        self.L, self.h = L, handle
        n = L.msspe_fasta_records(handle)
        self.names = [L.msspe_fasta_name(handle, i).decode() for i in range(n)]
        self.offsets = np.ctypeslib.as_array(C.cast(L.msspe_fasta_offsets(handle), C.POINTER(C.c_uint64)), shape=(n + 1,)).copy()
        nb = int(self.offsets[-1])
        self.bases = (np.ctypeslib.as_array(C.cast(L.msspe_fasta_bases(handle), C.POINTER(C.c_uint8)), shape=(nb,)).copy()
                      if nb else np.zeros(0, np.uint8))

    def sequences(self):
        return [self.bases[int(self.offsets[i]):int(self.offsets[i + 1])].tobytes().decode("latin-1") for i in range(len(self.names))]

    def close(self):
        if self.h:
            self.L.msspe_fasta_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def fasta_open(path: str, n_threads: int = 0) -> Fasta:
    """Parse only (host; no device needed)."""
    L = load_library()
    h = C.c_void_p()
    err = C.create_string_buffer(256)
    rc = L.msspe_fasta_open(os.fsencode(path), n_threads, C.byref(h), err, 256)
    if rc != OK:
        raise MsspeError(rc, err.value.decode())
    return Fasta(L, h)


class Engine:
    """One msspe_ctx.  Method names follow the reference functions they replace."""

    def __init__(self, kmer_size=13, window_size=500, overlap_size=250, search_windows_size=50, device=0):
        self.L = load_library()
        self.k = kmer_size
        cfg = Config(kmer_size, window_size, overlap_size, search_windows_size, device, 0)
        h = C.c_void_p()
        rc = self.L.msspe_create(C.byref(cfg), C.byref(h))
        if rc != OK:
            raise MsspeError(rc, self.L.msspe_last_error(None).decode())
        self.h = h
        self._keep = None

    def close(self):
        if getattr(self, "h", None):
            self.L.msspe_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != OK:
            raise MsspeError(rc, self.L.msspe_last_error(self.h).decode())

    # -- get_segment_manager (main.rs:196-235) --
    def load_genomes(self, bases: np.ndarray, offsets: np.ndarray):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        self._check(self.L.msspe_load_genomes(self.h, bases.ctypes.data, offsets.ctypes.data, len(offsets) - 1))

    def load_fasta(self, path: str, n_threads: int = 0) -> Fasta:
        """Parse and load in one call (chunked host-to-device copies overlap the parse)."""
        h = C.c_void_p()
        self._check(self.L.msspe_load_fasta(self.h, os.fsencode(path), n_threads, C.byref(h)))
        return Fasta(self.L, h)

    def load_genomes_device(self, device_ptr: int, offsets: np.ndarray, keepalive=None):
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        self._keep = keepalive
        self._check(self.L.msspe_load_genomes_device(self.h, C.c_void_p(device_ptr), offsets.ctypes.data, len(offsets) - 1))

    def reserve_pool(self, nbytes: int):
        """Background first touch of the stream-ordered pool (cold start); the first build_index waits for it."""
        self._check(self.L.msspe_reserve_pool(self.h, C.c_uint64(int(nbytes))))

    def set_stream(self, cuda_stream: int):
        self._check(self.L.msspe_set_stream(self.h, C.c_void_p(cuda_stream)))

    def build_index(self):
        self._check(self.L.msspe_build_index(self.h))

    def segment_info(self):
        g, mp, s = C.c_uint64(), C.c_uint32(), C.c_uint32()
        self._check(self.L.msspe_segment_info(self.h, C.byref(g), C.byref(mp), C.byref(s)))
        return g.value, mp.value, s.value

    def segment_kmers(self, direction: int) -> np.ndarray:
        g, _, s = self.segment_info()
        out = np.empty(max(1, g * s), dtype=np.uint64)
        self._check(self.L.msspe_get_segment_kmers(self.h, direction, out.ctypes.data, g * s))
        return out[:g * s].reshape(g, s) if s else out[:0].reshape(g, 0)

    # -- make_kmer_segments_windows_mapping (main.rs:237-255) --
    def index(self, direction: int):
        nc, npost = C.c_uint64(), C.c_uint64()
        self._check(self.L.msspe_get_index(self.h, direction, C.byref(nc), C.byref(npost), None, None, None))
        codes = np.empty(max(1, nc.value), dtype=np.uint64)
        offs = np.empty(nc.value + 1, dtype=np.uint64)
        post = np.empty(max(1, npost.value), dtype=np.uint32)
        self._check(self.L.msspe_get_index(self.h, direction, C.byref(nc), C.byref(npost), codes.ctypes.data,
                                           offs.ctypes.data, post.ctypes.data))
        return codes[:nc.value], offs, post[:npost.value]

    # -- find_candidates_kmers (main.rs:331-406) --
    def select(self, direction: int, max_iterations: int, max_mismatch_segments: int, mode=SELECT_RECOUNT) -> np.ndarray:
        out = np.zeros(max(1, max_iterations), dtype=CANDIDATE_DTYPE)
        n = C.c_uint32()
        self._check(self.L.msspe_select(self.h, direction, max_iterations, max_mismatch_segments, mode, out.ctypes.data, C.byref(n)))
        return out[:n.value]

    def select_both(self, max_iterations: int, max_mismatch_segments: int, mode=SELECT_RECOUNT):
        a = np.zeros(max(1, max_iterations), dtype=CANDIDATE_DTYPE)
        b = np.zeros(max(1, max_iterations), dtype=CANDIDATE_DTYPE)
        na, nb = C.c_uint32(), C.c_uint32()
        self._check(self.L.msspe_select_both(self.h, max_iterations, max_mismatch_segments, mode, a.ctypes.data, C.byref(na),
                                             b.ctypes.data, C.byref(nb)))
        return a[:na.value], b[:nb.value]

    def coverage(self, fwd_codes, rev_codes):
        g, _, _ = self.segment_info()
        f = np.ascontiguousarray(fwd_codes, dtype=np.uint64)
        r = np.ascontiguousarray(rev_codes, dtype=np.uint64)
        cov = np.zeros(max(1, g), dtype=np.uint8)
        part = np.zeros(max(1, g), dtype=np.uint16)
        rec = np.zeros(max(1, g), dtype=np.uint32)
        self._check(self.L.msspe_coverage(self.h, f.ctypes.data, len(f), r.ctypes.data, len(r), cov.ctypes.data,
                                          part.ctypes.data, rec.ctypes.data, g))
        return cov[:g], part[:g], rec[:g]

    def coverage_summary(self, fwd_codes, rev_codes, n_records: int):
        """print_coverage_report's aggregation (main.rs:518-574) reduced on the device: per-record and per-partition
        (covered, total) segment counts and the number of covered segments."""
        _, maxp, _ = self.segment_info()
        f = np.ascontiguousarray(fwd_codes, dtype=np.uint64)
        r = np.ascontiguousarray(rev_codes, dtype=np.uint64)
        npart = int(maxp) + 1
        rc = np.zeros(max(1, n_records), dtype=np.uint32); rt = np.zeros(max(1, n_records), dtype=np.uint32)
        pc = np.zeros(npart, dtype=np.uint32); pt = np.zeros(npart, dtype=np.uint32)
        ncov = C.c_uint64(0)
        self._check(self.L.msspe_coverage_summary(self.h, f.ctypes.data, len(f), r.ctypes.data, len(r), rc.ctypes.data,
                                                  rt.ctypes.data, n_records, pc.ctypes.data, pt.ctypes.data, npart, C.byref(ncov)))
        return rc[:n_records], rt[:n_records], pc, pt, int(ncov.value)

    def vertex_cover(self, codes, edge_a, edge_b) -> np.ndarray:
        """main.rs:754-798 on the device: codes = distinct primer words, (edge_a[e], edge_b[e]) = conflict edges as
        node indices; returns deleted[n] (uint8)."""
        codes = np.ascontiguousarray(codes, dtype=np.uint64)
        ea = np.ascontiguousarray(edge_a, dtype=np.uint32); eb = np.ascontiguousarray(edge_b, dtype=np.uint32)
        assert len(ea) == len(eb)
        deleted = np.zeros(max(1, len(codes)), dtype=np.uint8)
        nd = C.c_uint32(0)
        self._check(self.L.msspe_vertex_cover(self.h, codes.ctypes.data, len(codes), ea.ctypes.data, eb.ctypes.data, len(ea),
                                              deleted.ctypes.data, C.byref(nd)))
        assert int(deleted[:len(codes)].sum()) == nd.value
        return deleted[:len(codes)]

    # -- check_primers (primer.rs:143-166) --
    def primer_thermo(self, codes, oligo_len=None):
        codes = np.ascontiguousarray(codes, dtype=np.uint64)
        n = len(codes)
        outs = [np.zeros(max(1, n), dtype=np.float64) for _ in range(5)]
        self._check(self.L.msspe_primer_thermo(self.h, codes.ctypes.data, n, oligo_len or self.k, *[o.ctypes.data for o in outs]))
        return dict(zip(("tm", "gc", "self_any", "self_end", "hairpin"), [o[:n] for o in outs]))

    # -- get_kmer_stats + filter_kmers (main.rs:408-516) --
    def kmer_stats(self, codes, cfg: "FilterCfg" = None, oligo_len=None) -> np.ndarray:
        codes = np.ascontiguousarray(codes, dtype=np.uint64)
        out = np.zeros(max(1, len(codes)), dtype=KMER_STAT_DTYPE)
        cfg = cfg or default_filter_cfg()
        self._check(self.L.msspe_kmer_stats(self.h, codes.ctypes.data, len(codes), oligo_len or self.k, C.byref(cfg), out.ctypes.data))
        return out[:len(codes)]

    def kmer_stats_both(self, fwd_codes, rev_codes, cfg: "FilterCfg" = None, oligo_len=None):
        """get_kmer_stats + filter_kmers for both directions with one device batch (main.rs:723-724)."""
        f = np.ascontiguousarray(fwd_codes, dtype=np.uint64); r = np.ascontiguousarray(rev_codes, dtype=np.uint64)
        of = np.zeros(max(1, len(f)), dtype=KMER_STAT_DTYPE); orv = np.zeros(max(1, len(r)), dtype=KMER_STAT_DTYPE)
        cfg = cfg or default_filter_cfg()
        self._check(self.L.msspe_kmer_stats_both(self.h, f.ctypes.data, len(f), r.ctypes.data, len(r), oligo_len or self.k, C.byref(cfg),
                                                 of.ctypes.data, orv.ctypes.data))
        return of[:len(f)], orv[:len(r)]

    def thal_pairs(self, a, b, ttype, cond: ThalCond, oligo_len=None) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.uint64)
        b = np.ascontiguousarray(b if b is not None else a, dtype=np.uint64)
        out = np.zeros(max(1, len(a)), dtype=THAL_OUT_DTYPE)
        self._check(self.L.msspe_thal_pairs(self.h, a.ctypes.data, b.ctypes.data, len(a), oligo_len or self.k, ttype,
                                            C.byref(cond), out.ctypes.data))
        return out[:len(a)]

    # -- run_ntthal (delta_g.rs:83-153) --
    def cross_dimer(self, codes, cond: ThalCond, dg_limit: float, row_begin=0, row_end=None, oligo_len=None,
                    edge_capacity=None, nostruct_capacity=None):
        codes = np.ascontiguousarray(codes, dtype=np.uint64)
        n = len(codes)
        row_end = n if row_end is None else row_end
        rows = row_end - row_begin
        ecap = edge_capacity if edge_capacity is not None else max(1024, rows * n)
        ncap = nostruct_capacity if nostruct_capacity is not None else max(1024, rows * n)
        edges = np.zeros(ecap, dtype=EDGE_DTYPE)
        nos = np.zeros(ncap, dtype=np.uint64)
        ne, nn = C.c_uint64(), C.c_uint64()
        self._check(self.L.msspe_cross_dimer(self.h, codes.ctypes.data, n, oligo_len or self.k, C.byref(cond), row_begin,
                                             row_end, dg_limit, edges.ctypes.data, ecap, C.byref(ne), nos.ctypes.data,
                                             ncap, C.byref(nn)))
        return edges[:ne.value], nos[:nn.value]

    def cross_dimer_device(self, codes, cond: ThalCond, dg_limit: float, row_begin=0, row_end=None, oligo_len=None,
                           edge_capacity=None, nostruct_capacity=None):
        """msspe_cross_dimer with the compacted lists left on the device: returns (edges, nostruct) as torch tensors viewing
        ctx-owned buffers (int64 [n, 2] = (pair, dG bits) and int64 [m]); valid until the next cross-dimer call."""
        import torch
        codes = np.ascontiguousarray(codes, dtype=np.uint64)
        n = len(codes)
        row_end = n if row_end is None else row_end
        rows = row_end - row_begin
        ecap = edge_capacity if edge_capacity is not None else max(1024, rows * n)
        ncap = nostruct_capacity if nostruct_capacity is not None else max(1024, rows * n)
        pe, pn, ne, nn = C.c_void_p(), C.c_void_p(), C.c_uint64(), C.c_uint64()
        self._check(self.L.msspe_cross_dimer_device(self.h, codes.ctypes.data, n, oligo_len or self.k, C.byref(cond), row_begin, row_end,
                                                    dg_limit, ecap, ncap, C.byref(pe), C.byref(ne), C.byref(pn), C.byref(nn)))
        e = (torch.as_tensor(Engine._DevArray(pe.value, 2 * ne.value, "<i8"), device="cuda").view(-1, 2) if ne.value
             else torch.zeros((0, 2), dtype=torch.int64, device="cuda"))
        s = (torch.as_tensor(Engine._DevArray(pn.value, nn.value, "<i8"), device="cuda") if nn.value
             else torch.zeros(0, dtype=torch.int64, device="cuda"))
        return e, s

    def dist_init(self, dist, device):
        """Communicator of the multi-GPU loop: rank 0 draws the NCCL id, torch.distributed carries its 128 bytes."""
        import torch
        idt = torch.zeros(128, dtype=torch.uint8)
        if dist.get_rank() == 0:
            buf = C.create_string_buffer(128)
            rc = self.L.msspe_dist_unique_id(buf)
            if rc != OK:
                raise MsspeError(rc, "msspe_dist_unique_id failed (libnccl.so.2 not loadable?)")
            idt = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        idt = idt.to(device)
        dist.broadcast(idt, 0)
        raw = bytes(idt.cpu().numpy().tobytes())
        self._check(self.L.msspe_dist_init(self.h, raw, dist.get_rank(), dist.get_world_size()))

    def select_both_dist(self, max_iterations: int, max_mismatch_segments: int):
        """find_candidates_kmers of the WHOLE job (both directions) from this rank's column shard; collective."""
        a = np.zeros(max(1, max_iterations), dtype=CANDIDATE_DTYPE)
        b = np.zeros(max(1, max_iterations), dtype=CANDIDATE_DTYPE)
        na, nb = C.c_uint32(), C.c_uint32()
        self._check(self.L.msspe_select_both_dist(self.h, max_iterations, max_mismatch_segments, a.ctypes.data, C.byref(na),
                                                  b.ctypes.data, C.byref(nb)))
        return a[:na.value], b[:nb.value]

    def kernel_profile(self) -> np.ndarray:
        """Per kernel class since the last reset_timing: launches, algorithmic bytes, device ms (while profiling is on)."""
        out = np.zeros(16, dtype=KERNEL_PROF_DTYPE)
        n = C.c_uint32()
        self._check(self.L.msspe_get_kernel_profile(self.h, out.ctypes.data, 16, C.byref(n)))
        return out[:n.value]

    def set_profiling(self, on: bool):
        self._check(self.L.msspe_set_profiling(self.h, 1 if on else 0))

    # -- genome-sharded selection: per-rank primitives (msspe_b200.distributed.select_sharded drives them) --
    class _DevArray:
        """Zero-copy view of a device buffer for torch.as_tensor (CUDA array interface)."""

        def __init__(self, ptr, n, typestr):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}

    def shard_begin(self, direction: int):
        self._check(self.L.msspe_shard_begin(self.h, direction))

    def _shard_buffers(self, direction: int):
        pc, pf, n = C.c_void_p(), C.c_void_p(), C.c_uint64()
        self._check(self.L.msspe_shard_buffers(self.h, direction, C.byref(pc), C.byref(pf), C.byref(n)))
        return pc.value, pf.value, n.value

    def shard_codes(self, direction: int):
        import torch
        pc, _, n = self._shard_buffers(direction)
        if n == 0:
            return torch.zeros(0, dtype=torch.int64, device="cuda")
        return torch.as_tensor(Engine._DevArray(pc, n, "<i8"), device="cuda")  # codes < 2^62: order-preserving as int64

    def shard_count(self, direction: int):
        import torch
        live = C.c_uint64()
        self._check(self.L.msspe_shard_count(self.h, direction, C.byref(live)))
        _, pf, n = self._shard_buffers(direction)
        f = torch.as_tensor(Engine._DevArray(pf, n, "<i4"), device="cuda") if n else torch.zeros(0, dtype=torch.int32, device="cuda")
        return f, live.value

    def shard_n_part(self) -> int:
        g, maxp, _ = self.segment_info()
        return (maxp + 1) if g else 0

    def shard_firstpos(self, direction: int, local_ids: np.ndarray, n_part: int) -> np.ndarray:
        ids = np.ascontiguousarray(local_ids, dtype=np.uint32)
        out = np.full((len(ids), n_part), 0xFFFFFFFF, dtype=np.uint32)
        if len(ids) and n_part:
            self._check(self.L.msspe_shard_firstpos(self.h, direction, ids.ctypes.data, len(ids), n_part, out.ctypes.data))
        return out

    def shard_apply(self, direction: int, local_id: int, n_part: int) -> np.ndarray:
        out = np.zeros(max(1, n_part), dtype=np.uint8)
        self._check(self.L.msspe_shard_apply(self.h, direction, local_id & 0xFFFFFFFF, n_part, out.ctypes.data))
        return out[:n_part]

    def set_thal_params_dir(self, path: str):
        buf = C.create_string_buffer(RAW_PARAMS_BYTES + 64)
        err = C.create_string_buffer(512)
        rc = self.L.msspe_thal_params_from_dir(path.encode(), buf, err, 512)
        if rc != OK:
            raise MsspeError(rc, err.value.decode())
        self._check(self.L.msspe_set_thal_params(self.h, buf))

    def timing(self) -> Timing:
        t = Timing()
        self._check(self.L.msspe_get_timing(self.h, C.byref(t)))
        return t

    def reset_timing(self):
        self._check(self.L.msspe_reset_timing(self.h))

    def synchronize(self):
        self._check(self.L.msspe_synchronize(self.h))
